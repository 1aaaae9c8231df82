#!/usr/bin/env python
"""Static instruction census of one kernel in a cubin / shared library (no GPU needed).

    python tools/sass_count.py mara3_b200/libmara3_b200.so 'stage_strip<4, 64, true, 2, false, false>'

Prints the SASS instruction count per class (fp64 pipe, MUFU, shared / global memory, shuffles,
integer / select, control), which is what the per-phase instruction budget in DESIGN.md is checked
against before GPU time is spent (straight-line kernels: static count ~ dynamic count per CTA).
"""
import collections
import re
import subprocess
import sys


def classify(op):
    base = op.split(".")[0]
    if base in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"):
        return "fp64"
    if base == "MUFU":
        return "mufu"
    if base in ("F2F", "I2F", "F2I", "I2FP", "F2FP"):
        return "convert"
    if base in ("LDS", "STS", "LDSM"):
        return "shared"
    if base in ("LDG", "STG", "LD", "ST", "LDC", "LDCU", "ULDC", "CCTL", "ATOMG", "RED", "ATOM"):
        return "global/const"
    if base in ("LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "UTMAPF", "LDGDEPBAR", "DEPBAR"):
        return "async-copy"
    if base in ("SHFL", "VOTE", "VOTEU", "MATCH", "REDUX"):
        return "shuffle/vote"
    if base in ("BAR", "BRA", "EXIT", "BSSY", "BSYNC", "CALL", "RET", "WARPSYNC", "NOP", "YIELD", "BREAK", "MEMBAR", "ERRBAR", "FENCE", "ELECT", "NANOSLEEP"):
        return "control"
    if base in ("FSEL", "SEL", "ISETP", "LOP3", "IADD3", "IMAD", "LEA", "SHF", "MOV", "PRMT", "IABS", "IMNMX", "VIADD", "VIMNMX", "PLOP3", "P2R", "R2P",
                "UMOV", "UIADD3", "ULOP3", "UIMAD", "USHF", "ULEA", "UISETP", "USEL", "UPLOP3", "S2R", "S2UR", "CS2R", "R2UR", "FSETP", "FADD", "FMUL", "FFMA",
                "UPRMT", "IADD", "UIADD", "HFMA2", "FMNMX", "POPC", "FLO", "BREV", "SGXT", "UFLO", "UPOPC", "BMSK", "UBMSK", "LEPC"):
        return "int/select/move"
    return "other:" + base


def main():
    lib, pattern = sys.argv[1], sys.argv[2]
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    mangled = re.findall(r"Function : (\S+)", sass)
    want = [m for m, d in zip(mangled, names) if pattern in d]
    if len(want) != 1:
        print("matches:", [d for d in names if pattern in d])
        sys.exit(1)
    body = sass.split("Function : " + want[0])[1].split("Function : ")[0]
    counts = collections.Counter()
    detail = collections.Counter()
    for line in body.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
        if m:
            op = m.group(1)
            counts[classify(op)] += 1
            detail[op.split(".")[0]] += 1
    total = sum(counts.values())
    print(f"{pattern}: {total} SASS instructions")
    for k, v in counts.most_common():
        print(f"  {k:18s} {v:6d}  {100.0 * v / total:5.1f} %")
    print("  top opcodes:", ", ".join(f"{k} {v}" for k, v in detail.most_common(24)))


if __name__ == "__main__":
    main()

"""Distance (in SASS instructions) from each global load of a kernel to the first instruction that reads its destination.
A short distance in an in-order warp means a long-scoreboard stall.  usage: python tools/sass_load_distance.py <cubin|so> <kernel substring>"""
import re, subprocess, sys
sass = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
mangled = re.findall(r"Function : (\S+)", sass)
names = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.splitlines()
want = [m for m, d in zip(mangled, names) if sys.argv[2] in d]
assert len(want) == 1, [d for d in names if sys.argv[2] in d]
body = sass.split("Function : " + want[0])[1].split("Function : ")[0]
ins = [m.group(1).strip() for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(.*?);", body)]
for i, s in enumerate(ins):
    m = re.match(r"(@!?U?P\d\s+)?LDG\S*\s+(R\d+),", s)
    if not m: continue
    r = int(m.group(2)[1:]); regs = {f"R{r}", f"R{r+1}"} if ".64" in s else {f"R{r}"}
    if ".128" in s: regs = {f"R{r+k}" for k in range(4)}
    for j in range(i + 1, len(ins)):
        ops = ins[j].split(",", 1)[1] if "," in ins[j] else ""
        if any(re.search(rf"\b{x}\b", ops) for x in regs) or (ins[j].startswith(("ST", "@")) and any(re.search(rf"\b{x}\b", ins[j]) for x in regs)):
            print(f"{i:5d} {s[:50]:50s} -> +{j - i:4d}  {ins[j][:60]}"); break

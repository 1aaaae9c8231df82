"""Top stalled SASS instructions of an ncu source page export.  usage: python tools/ncu_top_stalls.py <src.csv> [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sec = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
hdr = rows[sec[0]]
body = [x for x in rows[sec[0]+1:(sec[1]-1 if len(sec) > 1 else len(rows))] if len(x) == len(hdr)]
isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(x[ist]) for x in body)
print("total samples", tot)
by_reason = {}
for x in body:
    for i in stall_cols:
        if x[i] not in ("", "0"): by_reason[hdr[i]] = by_reason.get(hdr[i], 0) + int(x[i])
print({k: round(v / tot, 3) for k, v in sorted(by_reason.items(), key=lambda kv: -kv[1])})
for idx, x in sorted(enumerate(body), key=lambda ix: -int(ix[1][ist]))[:n]:
    reasons = sorted([(int(x[i]), hdr[i][6:]) for i in stall_cols if x[i] not in ("", "0")], reverse=True)[:2]
    print(f"{idx:5d} {int(x[ist]):6d} {x[iex]:>9s}  {x[isrc].strip()[:64]:64s} {reasons}")

"""Summarise an ncu capture of one kernel: headline counters, stall reasons, per-phase opcode mix.
usage: python tools/ncu_summary.py <raw.csv> <sass.csv> <cell-warps per launch>"""
import csv, collections, re, sys
raw, sass, W = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "smsp__cycles_elapsed.avg.per_second"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w: print(f"{w:66s} {r[i]:>18s} {units[i]}")
items = []
for i, h in enumerate(hdr):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try: items.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError: pass
tot = sum(v for v, _ in items)
print("stalls:", ", ".join(f"{h} {v/tot:.2f}" for v, h in sorted(items, reverse=True)[:9]))
rows = list(csv.reader(open(sass)))
sections = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
hdr = rows[sections[0]]
body = rows[sections[0] + 1:(sections[1] - 1 if len(sections) > 1 else len(rows))]
isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
phase, acc, stall, tot = 0, collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
for x in body:
    if len(x) <= iex: continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", x[isrc].strip())
    op = (m.group(2) if m else x[isrc]).split(".")[0]
    n = int(x[iex]); acc[phase][op] += n; tot[phase] += n; stall[phase] += int(x[ist])
    if op == "BAR": phase += 1
print(f"total warp-instr per cell-warp: {sum(tot.values())/W:.1f}")
for p in sorted(acc):
    dp = sum(acc[p][o] for o in ("DADD", "DMUL", "DFMA", "DSETP"))
    print(f"phase {p}: total {tot[p]/W:7.1f}  DP {dp/W:6.1f}  stall samples {stall[p]:5d}  " + ", ".join(f"{o} {n/W:.1f}" for o, n in acc[p].most_common(14)))

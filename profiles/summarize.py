#!/usr/bin/env python
"""Turn an ncu report (gpurun_out/*.ncu-rep, captured with `ncu --set full --clock-control none --import-source on`)
into the short text summary that is committed under profiles/.

    python profiles/summarize.py gpurun_out/prof_x.ncu-rep <units per launch> [unit name] > profiles/r01_x.txt
"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
uname = sys.argv[3] if len(sys.argv) > 3 else "unit"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, un, vals = rows[0], rows[1], rows[2]
get = {h: (vals[i], un[i]) for i, h in enumerate(hdr)}
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print(f"# {rep}")
for w in want:
    if w in get:
        print(f"{w:78s} {get[w][0]:>24s} {get[w][1]}")
print("warp stall reasons (warps per issue-active cycle, > 0.15):")
for h in hdr:
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(get[h][0] or 0) > 0.15:
        print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {float(get[h][0]):.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]; data = rows[2:]
iA, iE, iS = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
tot = sum(int(r[iE]) for r in data)
print(f"SASS opcode mix: {tot} warp instructions, {tot / units:.2f} per {uname} ({units:.0f} {uname}s per launch)")
c, s = Counter(), Counter()
for r in data:
    t = r[iA].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[iE]); s[op] += int(r[iS])
ts = max(1, sum(s.values()))
for op, n in c.most_common(18):
    print(f"  {op:10s} {n / tot * 100:6.2f}%  {n / units:8.2f} per {uname}   stall samples {s[op] / ts * 100:5.1f}%")

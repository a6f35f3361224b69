#!/usr/bin/env python
"""Per-source-line instruction and stall-sample counts of one profiled kernel.

ncu's SASS page (per-instruction counters, in address order) is joined with `nvdisasm -g` of the object file the
kernel came from (same instruction order, `//## File ... line N` markers from -lineinfo).

    python profiles/by_line.py gpurun_out/prof_x.ncu-rep maximilian_b200/build/spectral.o stft_kernelILi16 [units] [top]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, obj, sym = sys.argv[1], sys.argv[2], sys.argv[3]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
by_stall = os.environ.get("BY_STALL") == "1"      # order by stall samples instead of executed instructions

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
prof = [(r[iS].strip(), int(r[iE]), int(r[iN])) for r in rows[2:]]

with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
lines, cur, inside = [], None, False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        inside = sym in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(1).strip()))
assert len(lines) == len(prof), (len(lines), len(prof))
agg = defaultdict(lambda: [0, 0])
tot = sum(p[1] for p in prof); ts = max(1, sum(p[2] for p in prof))
for (loc, _), (_, e, s) in zip(lines, prof):
    agg[loc][0] += e; agg[loc][1] += s
srcs = {}
print(f"{tot / units:.1f} warp instructions per unit; top {top} source lines")
for loc, (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_stall else 0])[:top]:
    f, n = loc if loc else ("?", 0)
    if f not in srcs:
        p = os.path.join(os.environ.get("CSRC", os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc")), f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][n - 1].strip()[:100] if 0 < n <= len(srcs[f]) else ""
    print(f"{e / units:9.1f} ({e / tot * 100:5.1f}%)  stall {s / ts * 100:5.1f}%  {f}:{n}  {text}")

#!/usr/bin/env python3
"""Turns an ncu launch list (ncu --metrics gpu__time_duration.sum --csv) into the per-kernel summary kept under profiles/.
    python tools/summarize_launches.py gpurun_out/r01d_launches.csv profiles/r01 [launches_per_proof]
writes <prefix>_launches_full.csv (the launches of the LAST proof in the capture) and <prefix>_launches_summary.csv"""
import csv
import re
import sys
from collections import OrderedDict

src, prefix = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
# the capture holds warm-up launches, identical proofs and possibly other work after them: one steady-state proof's worth of launches
# is the window between the last two constraint kernels (stages 4-9 of one proof + stages 1-3 of the next: same kernels, rotated)
idx = [i for i, r in enumerate(rows) if "constraint_eval" in r[ki]]
if len(sys.argv) > 3:
    per_proof = int(sys.argv[3])
    last = rows[-per_proof:]
else:
    last = rows[idx[-2] + 1: idx[-1] + 1]
short = lambda k: re.sub(r"\(.*", "", k).replace("void ", "").replace("dg::", "")
with open(prefix + "_launches_full.csv", "w") as f:
    f.write(f"# launches of one steady-state proof (window between the last two constraint kernels of the capture), in order; source: {src}\n# index,kernel,grid,block,ns\n")
    for i, r in enumerate(last):
        f.write(f"{i},{short(r[ki])},{r[gi].replace(',', ' ')},{r[bi].replace(',', ' ')},{r[vi]}\n")
tot = OrderedDict()
for r in last:
    k = short(r[ki])
    t = tot.setdefault(k, [0, 0.0])
    t[0] += 1
    t[1] += float(r[vi]) / 1e6
total = sum(t[1] for t in tot.values())
with open(prefix + "_launches_summary.csv", "w") as f:
    f.write("# ncu launch list summary: one proof of the 2^20-step collatz trace (w=26, blowup 32) -- gpu__time_duration.sum, --clock-control none\n")
    f.write("# command: " + (sys.argv[4] if len(sys.argv) > 4 else "ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline") + "\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's stage_ms, not absolutes\n")
    f.write(f"# total {total:.2f} ms over {len(last)} launches\nkernel,launches,total_ms,share\n")
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{ms:.3f},{ms / total:.4f}\n")
print(open(prefix + "_launches_summary.csv").read())

#!/usr/bin/env python3
"""Per-stage device times of one workload: python tools/stage_times.py [log_n] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import distaff_b200 as dg
from distaff_b200 import backend
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tr, name = bench.build_trace(log_n)
buf = backend.DeviceBuffer(tr.registers.nbytes).upload(tr.registers)
acc = []
for i in range(reps + 2):
    p = dg.prove_device(buf, tr.width, tr.length, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
    if i >= 2:
        acc.append([p.stats["total_ms"]] + p.stats["stage_ms"])
a = np.mean(np.array(acc), axis=0)
print(os.environ.get("DG_AIR_LB", "-"), name, "total %.1f ms | " % a[0] + " ".join("%.1f" % x for x in a[1:]), flush=True)

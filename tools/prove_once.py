#!/usr/bin/env python3
"""Runs a few proofs of the benchmark workload (for ncu): python tools/prove_once.py [log_n] [count]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                      # noqa: E402
import distaff_b200 as dg         # noqa: E402
from distaff_b200 import backend  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
count = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tr, name = bench.build_trace(log_n)
buf = backend.DeviceBuffer(tr.registers.nbytes).upload(tr.registers)
for i in range(count):
    p = dg.prove_device(buf, tr.width, tr.length, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
    print(i, name, len(p.bytes), p.stats["total_ms"], p.stats["kernel_launches"], flush=True)

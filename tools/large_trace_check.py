#!/usr/bin/env python3
"""Traces beyond the headline size (2^21, 2^22 steps: three-pass transforms, > 4 GiB intermediates): the proof must be accepted by the
restated reference verifier and by the GPU verifier, and be reproducible.   python tools/large_trace_check.py [log_n ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                      # noqa: E402
import distaff_b200 as dg                         # noqa: E402
from oracle import pyoracle as po                 # noqa: E402

ok = True
for log_n in [int(a) for a in sys.argv[1:]] or [21]:
    t0 = time.time()
    tr, name = bench.build_trace(log_n)
    t1 = time.time()
    p = dg.prove(tr)
    p2 = dg.prove(tr)
    v_cpu = po.verify(tr.program_hash, tr.public_inputs, tr.outputs, p.bytes)
    v_gpu = dg.verify(tr.program_hash, tr.public_inputs, tr.outputs, p.bytes)
    bad = dg.verify(tr.program_hash, tr.public_inputs, [tr.outputs[0] + 1], p.bytes)
    good = v_cpu is None and v_gpu is None and p.bytes == p2.bytes and bad is not None
    ok &= good
    print(f"2^{log_n} steps x {tr.width} registers ({name}, trace built in {t1 - t0:.0f} s): device {p2.stats['total_ms']:.1f} ms, "
          f"{len(p.bytes)} proof bytes, oracle verifier: {v_cpu}, GPU verifier: {v_gpu}, reproducible: {p.bytes == p2.bytes}, tampered outputs: {bad}", flush=True)
print("LARGE_TRACE_CHECK", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)

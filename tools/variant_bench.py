#!/usr/bin/env python3
"""One kernel-variant measurement per process (the variant switches are read once per process from the environment):
   DG_AIR_CFG / DG_NTT_RMAX / DG_NTT_BT / DG_NTT_INLINE ...  python tools/variant_bench.py [log_n] [reps]
Prints per-stage device times of the 2^log_n-step collatz proof and whether the proof equals the committed oracle digest."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                     # noqa: E402
import bench                           # noqa: E402
import distaff_b200 as dg              # noqa: E402
from distaff_b200 import backend, hostvm   # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cache = "/tmp/dg_trace_%d.npz" % log_n
if os.path.exists(cache):
    z = np.load(cache, allow_pickle=True)
    tr = hostvm.ExecutionTrace(z["regs"], int(z["cd"]), int(z["ld"]), int(z["sd"]), bytes(z["ph"]), [int(x) for x in z["pi"]], [int(x) for x in z["out"]])
else:
    tr, _ = bench.build_trace(log_n)
    np.savez(cache, regs=tr.registers, cd=tr.ctx_depth, ld=tr.loop_depth, sd=tr.stack_depth, ph=np.frombuffer(tr.program_hash, dtype=np.uint8),
             pi=np.array(tr.public_inputs, dtype=object), out=np.array(tr.outputs, dtype=object))
buf = backend.DeviceBuffer(tr.registers.nbytes).upload(tr.registers)
acc = []
for i in range(reps + 2):
    p = dg.prove_device(buf, tr.width, tr.length, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
    if i >= 2:
        acc.append([p.stats["total_ms"]] + p.stats["stage_ms"])
a = np.mean(np.array(acc), axis=0)
sha = hashlib.sha256(p.bytes).hexdigest()
gold = bench.golden_for(log_n)
ok = (gold["proof_sha256"] == sha) if gold else None
tag = " ".join("%s=%s" % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith("DG_") and k != "DG_DEVICE")
print("VARIANT [%s] 2^%d total %.2f ms | %s | golden=%s launches=%d" % (tag, log_n, a[0], " ".join("%.2f" % x for x in a[1:]), ok, p.stats["kernel_launches"]), flush=True)

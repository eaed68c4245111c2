#!/usr/bin/env python3
"""Regenerates tests/golden/collatz_2_20.json: the CPU oracle proves bench.py's 2^20-step collatz trace (single thread: ~12 min and
~25 GB here; ORACLE_THREADS=N to use N threads -- the bytes do not depend on it) and the SHA-256 of trace and proof are recorded."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                         # noqa: E402
from oracle import pyoracle as po    # noqa: E402

po.set_threads(int(os.environ.get("ORACLE_THREADS", "1")))
tr, name = bench.build_trace(20)
regs = np.ascontiguousarray(tr.registers)
t0 = time.time()
r = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
dt = time.time() - t0
assert r.error is None, r.error
assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, r.proof) is None
out = {"name": name, "trace_sha256": hashlib.sha256(regs.tobytes()).hexdigest(), "proof_sha256": hashlib.sha256(r.proof).hexdigest(),
       "proof_len": len(r.proof), "oracle_prove_s": dt}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "collatz_2_20.json"), "w"))
print(out)

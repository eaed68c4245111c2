#!/usr/bin/env python3
"""Differential debugging: run the CUDA prover with DG_DEBUG_DUMP and compare every dumped stage against the CPU oracle.
    python tools/debug_stages.py [program-name]      (names from tests/programs.py)"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
d = tempfile.mkdtemp()
os.environ["DG_DEBUG_DUMP"] = d
import distaff_b200 as dg              # noqa: E402
from oracle import pyoracle as po      # noqa: E402
from tests import programs             # noqa: E402

names = sys.argv[1:] or ["fib13"]
P = programs.small_programs()
for name in names:
    tr = P[name]
    ref = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs, keep_large=True)
    print(name, "oracle error:", ref.error, "n=", tr.length, "w=", tr.width)
    try:
        proof = dg.prove(tr)
    except Exception as e:
        print("  GPU prove failed:", e)
        proof = None
    for vec in ("i_evals", "f_evals", "t_evals", "constraint_poly", "composition_poly"):
        path = os.path.join(d, vec + ".bin")
        if not os.path.exists(path):
            print("  ", vec, "not dumped")
            continue
        got = np.fromfile(path, dtype=np.uint64).reshape(-1, 2)
        want = ref.vector(vec)
        same = got.shape == want.shape and np.array_equal(got, want)
        bad = None if same else (np.nonzero((got != want).any(axis=1))[0][:8] if got.shape == want.shape else "shape %s vs %s" % (got.shape, want.shape))
        print("  ", vec, "OK" if same else f"MISMATCH at {bad}")
    if proof:
        print("   trace_root", proof.trace_root == ref.digest("trace_root"), "constraint_root", proof.constraint_root == ref.digest("constraint_root"),
              "nonce", proof.pow_nonce == ref.u64s("pow_nonce")[0], "bytes", proof.bytes == ref.proof, len(proof.bytes), len(ref.proof))
        print("   stats", proof.stats)

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
EXT, NQ, GR = [int(x) for x in os.environ.get("DG_OPTS", "32,50,20").split(",")]
P = programs.small_programs()
for name in names:
    tr = P[name]
    ref = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs, ext=EXT, num_queries=NQ, grinding=GR, keep_large=True)
    print(name, "oracle error:", ref.error, "n=", tr.length, "w=", tr.width)
    try:
        proof = dg.prove(tr, dg.ProofOptions(EXT, NQ, GR))
    except Exception as e:
        print("  GPU prove failed:", e)
        proof = None
    # the GPU path builds the boundary numerators directly in coefficient form: compare them with the interpolated oracle evaluations
    for vec, src in (("i_evals", "i_coeffs"), ("f_evals", "f_coeffs"), ("t_evals", "t_evals"), ("constraint_poly", "constraint_poly"),
                     ("composition_poly", "composition_poly")):
        path = os.path.join(d, src + ".bin")
        if not os.path.exists(path):
            print("  ", vec, "not dumped")
            continue
        got = np.fromfile(path, dtype=np.uint64).reshape(-1, 2)
        want = ref.vector(vec)
        if src.endswith("_coeffs"):
            want = po.fft(want, inverse=True)
        same = got.shape == want.shape and np.array_equal(got, want)
        bad = None if same else (np.nonzero((got != want).any(axis=1))[0][:8] if got.shape == want.shape else "shape %s vs %s" % (got.shape, want.shape))
        print("  ", vec, "OK" if same else f"MISMATCH at {bad}")
    fr = os.path.join(d, "fri_roots.bin")
    if os.path.exists(fr):
        got = open(fr, "rb").read()
        want = b"".join(ref.digests("fri_roots"))
        print("   fri_roots", [got[i:i + 32] == want[i:i + 32] for i in range(0, max(len(got), len(want)), 32)])
    ps = os.path.join(d, "positions.bin")
    if os.path.exists(ps):
        print("   positions", list(np.fromfile(ps, dtype=np.uint64)) == ref.u64s("positions"))
    if proof:
        print("   trace_root", proof.trace_root == ref.digest("trace_root"), "constraint_root", proof.constraint_root == ref.digest("constraint_root"),
              "nonce", proof.pow_nonce == ref.u64s("pow_nonce")[0], "bytes", proof.bytes == ref.proof, len(proof.bytes), len(ref.proof))
        if proof.bytes != ref.proof and len(proof.bytes) == len(ref.proof):
            diff = [i for i in range(len(ref.proof)) if proof.bytes[i] != ref.proof[i]]
            print("   first differing byte offsets:", diff[:10], "count", len(diff))
        print("   stats", proof.stats)

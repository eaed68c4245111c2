#!/usr/bin/env python3
"""Regenerates tests/golden/oracle_proofs.json.

The reference cannot be built or imported in this environment (Rust, no toolchain) and its own tests pin no proof bytes
(SURVEY.md F5), so these fixtures are produced by the CPU oracle (oracle/, a restatement whose primitives are pinned by
the reference's KATs, see tests/test_oracle_*.py, and whose proofs are accepted by the restated verifier).  They freeze
the oracle's behaviour so that (a) oracle regressions are caught and (b) the -m gpu tests can check the CUDA prover
against committed values as well as against the live oracle.   Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po   # noqa: E402
from tests import programs          # noqa: E402

out = {}
for name, tr in programs.small_programs().items():
    r = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
    assert r.error is None, (name, r.error)
    assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, r.proof) is None
    out[name] = {
        "n": tr.length, "w": tr.width, "ctx": tr.ctx_depth, "loop": tr.loop_depth, "stack": tr.stack_depth,
        "inputs": [str(v) for v in tr.public_inputs], "outputs": [str(v) for v in tr.outputs],
        "program_hash": tr.program_hash.hex(),
        "trace_sha256": hashlib.sha256(tr.registers.tobytes()).hexdigest(),
        "proof_len": len(r.proof), "proof_sha256": hashlib.sha256(r.proof).hexdigest(),
        "trace_root": r.digest("trace_root").hex(), "constraint_root": r.digest("constraint_root").hex(),
        "fri_roots": [d.hex() for d in r.digests("fri_roots")],
        "pow_nonce": r.u64s("pow_nonce")[0], "positions5": r.u64s("positions")[:5],
    }
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_proofs.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print("wrote", path, len(out), "programs")

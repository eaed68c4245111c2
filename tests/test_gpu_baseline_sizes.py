"""-m gpu parity at the sizes BASELINE.json names: fibonacci 2^16 steps, Merkle-path verification (2^14 steps, 6-wide
Rescue constraints), the 65-register trace (two-chunk BLAKE3 rows inside a full proof) -- byte-equal to the CPU oracle -- and the
headline 2^20-step collatz proof against the digest of the oracle's proof committed in tests/golden/collatz_2_20.json (the oracle
needs ~12 minutes for that one, so only its SHA-256 travels; tests/golden/make_collatz_2_20.py regenerates it).
Reference end-to-end tests these follow: /root/reference/src/tests/mod.rs:12-63, /root/reference/src/examples/{fibonacci,merkle,collatz}.rs."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import programs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)


@pytest.fixture(scope="module")
def dg():
    import distaff_b200
    from distaff_b200 import backend
    backend.device_info()
    return distaff_b200


@pytest.fixture(scope="module")
def fast_oracle(po):
    """the oracle with all host threads (identical bytes for any thread count: tests/test_oracle_stark.py::test_threads_do_not_change_the_proof)"""
    po.set_threads(min(32, os.cpu_count() or 1))      # the restatement is fastest at ~32 threads on the 128-thread hosts (profiles/r02_oracle_threads.txt)
    yield po
    po.set_threads(1)


def _check(dg, po, tr, n_expected=None):
    if n_expected is not None:
        assert tr.length == n_expected
    proof = dg.prove(tr)
    ref = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
    assert ref.error is None, ref.error
    assert proof.trace_root == ref.digest("trace_root")
    assert proof.constraint_root == ref.digest("constraint_root")
    assert proof.pow_nonce == ref.u64s("pow_nonce")[0]
    assert proof.bytes == ref.proof
    assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof.bytes) is None


def test_fibonacci_2_16(dg, fast_oracle):
    """BASELINE configs[1]: fibonacci, 2^16 trace steps (16 operations per term, examples/fibonacci.rs)"""
    from distaff_b200 import hostvm
    _check(dg, fast_oracle, hostvm.fibonacci((1 << 16) // 16 - 6), 1 << 16)


def test_merkle_paths_2_14(dg, fast_oracle):
    """BASELINE configs[2]: Merkle-path verification, 2^14 steps, Rescue rounds (enforce_rescr) on most cycles.  SURVEY.md quotes it as
    `merkle 256`, which the reference example cannot execute (see programs.merkle_paths); four depth-64 paths give the same shape."""
    _check(dg, fast_oracle, programs.merkle_paths(64, 4, fast_oracle), 1 << 14)
    _check(dg, fast_oracle, programs.merkle_example(64, fast_oracle), 1 << 12)      # the example at its largest depth


def test_wide_trace_two_chunk_rows(dg, fast_oracle):
    """65 registers: each committed row is 1040 bytes = two BLAKE3 chunks + parent (trace_table.rs:174-185, lib.rs:83)"""
    tr = programs.wide_program()
    assert tr.width == 65
    _check(dg, fast_oracle, tr)


def test_headline_collatz_2_20_matches_the_oracle_digest(dg, po):
    """BASELINE configs[3] / bench.py's workload: the 2^20-step collatz trace.  The GPU proof must hash to the digest of the CPU
    oracle's proof of the same trace, and the restated reference verifier must accept it."""
    import bench
    gold = json.load(open(os.path.join(HERE, "golden", "collatz_2_20.json")))
    tr, name = bench.build_trace(20)
    assert name == gold["name"] and tr.length == 1 << 20
    assert hashlib.sha256(np.ascontiguousarray(tr.registers).tobytes()).hexdigest() == gold["trace_sha256"]
    proof = dg.prove(tr)
    assert len(proof.bytes) == gold["proof_len"]
    assert hashlib.sha256(proof.bytes).hexdigest() == gold["proof_sha256"]
    assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof.bytes) is None
    # a second proof (arena path instead of the measuring pool path) is the same
    assert dg.prove(tr).bytes == proof.bytes

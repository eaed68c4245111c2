"""End-to-end checks of the oracle prover/verifier pair, after the reference's integration tests
(/root/reference/src/tests/mod.rs:12-63): execute -> prove -> verify == Ok, and the three tampering cases with the
reference's exact error string.  Also pins the oracle against the committed golden digests (tests/golden/)."""
import hashlib
import json
import os

import pytest

from tests import programs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_proofs.json")
FRI_ERR = "verification of low-degree proof failed: evaluations did not match column value at depth 0"


@pytest.fixture(scope="module")
def proven(po):
    out = {}
    for name, tr in programs.small_programs().items():
        r = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
        assert r.error is None, (name, r.error)
        out[name] = (tr, r)
    return out


def test_program_results():
    P = programs.small_programs()
    assert P["fib13"].outputs == [233] and P["fib_span"].outputs == [3]
    assert P["cmp"].outputs == [1] and P["collatz3"].outputs == [7]
    assert P["if_else"].outputs == [8] and P["if_else0"].outputs == [15]
    assert P["fib13"].length == 256 and P["fib13"].width == 20


def test_prove_then_verify(po, proven):
    for name, (tr, r) in proven.items():
        assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, r.proof) is None, name


def test_tampering_is_rejected_like_the_reference(po, proven):
    tr, r = proven["fib_span"]
    assert po.verify(tr.program_hash, [1, 1], tr.outputs, r.proof) == FRI_ERR
    assert po.verify(tr.program_hash, tr.public_inputs, [5], r.proof) == FRI_ERR
    bad = bytes([1]) + tr.program_hash[1:]
    assert po.verify(bad, tr.public_inputs, tr.outputs, r.proof) == FRI_ERR
    corrupt = bytearray(r.proof)
    corrupt[40] ^= 1          # inside trace_info/trace nodes
    assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, bytes(corrupt)) is not None


def test_proof_layout_fields(proven):
    # bincode layout: root[32] | domain_depth, ctx, loop, stack u8 | op_count u32 | ...  ... | nonce u64 | options[4]
    tr, r = proven["fib13"]
    p = r.proof
    assert p[32] == 13 and p[33] == tr.ctx_depth and p[34] == tr.loop_depth and p[35] == tr.stack_depth
    assert p[-4:] == bytes([5, 50, 20, 0])
    assert 60_000 < len(p) < 70_000          # README.md:151 quotes 62 KB for 2^8 operations


def test_invalid_trace_is_rejected(po, fib13):
    regs = fib13.registers.copy()
    regs[fib13.width - 1, 100, 0] += 1
    r = po.prove(regs, fib13.ctx_depth, fib13.loop_depth, fib13.public_inputs, fib13.outputs)
    assert r.error is not None and "not satisfied" in r.error


def test_matches_golden_digests(proven):
    golden = json.load(open(GOLDEN))
    for name, (tr, r) in proven.items():
        g = golden[name]
        assert hashlib.sha256(r.proof).hexdigest() == g["proof_sha256"], name
        assert r.digest("trace_root").hex() == g["trace_root"]
        assert r.digest("constraint_root").hex() == g["constraint_root"]
        assert r.u64s("pow_nonce")[0] == g["pow_nonce"]
        assert r.u64s("positions")[:5] == g["positions5"]


def test_wide_trace_two_chunk_rows(po):
    """w = 65 > 64 registers: leaves are multi-chunk BLAKE3 hashes; the restated prover and verifier still agree"""
    tr = programs.wide_program()
    assert (tr.width, tr.ctx_depth, tr.loop_depth, tr.stack_depth) == (65, 15, 3, 32)
    r = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs, num_queries=20, grinding=8)
    assert r.error is None
    assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, r.proof) is None
    bad = bytearray(r.proof)
    bad[40] ^= 1
    assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, bytes(bad)) is not None


def test_fri_round_trip_reference_tests(po):
    """fri/mod.rs mod tests prove_verify / verify_fail: a random polynomial of degree 63 (64) over a 512-point domain, default options,
    with the reference's exact error strings"""
    import ctypes
    import numpy as np
    from distaff_b200 import felt

    def evaluations(domain_size, degree, seed):
        coeffs = felt.to_ints(felt.random_elements(degree + 1, seed)) + [0] * (domain_size - degree - 1)
        return po.fft(felt.from_ints(coeffs))

    def run(ev, max_degree, drop=0):
        msg = ctypes.create_string_buffer(256)
        rc = po.lib().or_fri_roundtrip(np.ascontiguousarray(ev).ctypes.data, ev.shape[0], max_degree, drop, msg, 256)
        return rc, msg.value.decode()

    ev = evaluations(512, 63, 11)
    assert run(ev, 63) == (0, "")
    assert run(ev, 62) == (1, "remainder is not a valid degree 14 polynomial")
    ev2 = evaluations(512, 64, 12)
    assert run(ev2, 63) == (1, "remainder is not a valid degree 15 polynomial")
    assert run(ev2, 63, drop=1) == (1, "evaluations did not match column value at depth 0")


def test_threads_do_not_change_the_proof(po):
    """the optional host threading of the oracle (or_set_threads) only re-partitions exact field arithmetic: identical bytes"""
    from distaff_b200 import hostvm
    tr = hostvm.collatz(7)                      # 4096 steps: large enough for the threaded FFT / Merkle / batch-inversion paths
    po.set_threads(1)
    a = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
    try:
        po.set_threads(4)
        b = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
    finally:
        po.set_threads(1)
    assert a.error is None and b.error is None
    assert a.proof == b.proof

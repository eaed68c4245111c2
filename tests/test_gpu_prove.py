"""-m gpu parity tests of the whole prove pipeline through the C-ABI (dg_prove): proof bytes must be identical to the CPU
oracle's, accepted by the (restated) reference verifier, and equal to the committed golden digests."""
import hashlib
import json
import os

import pytest

from tests import programs

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_proofs.json")


@pytest.fixture(scope="module")
def dg():
    import distaff_b200
    from distaff_b200 import backend
    backend.device_info()
    return distaff_b200


def _same(proof, ref):
    """byte equality with a readable diagnosis (which commitment diverged first)"""
    if proof.bytes == ref.proof:
        return
    diff = [i for i in range(min(len(proof.bytes), len(ref.proof))) if proof.bytes[i] != ref.proof[i]]
    raise AssertionError("proof mismatch: len %d vs %d, trace_root %s, constraint_root %s, nonce %s vs %s, %d differing bytes from offset %s" % (
        len(proof.bytes), len(ref.proof), proof.trace_root == ref.digest("trace_root"), proof.constraint_root == ref.digest("constraint_root"),
        proof.pow_nonce, ref.u64s("pow_nonce")[0], len(diff), diff[:5]))


@pytest.fixture(scope="module")
def small():
    return programs.small_programs()


def test_proofs_are_bit_identical_to_the_oracle(dg, po, small):
    golden = json.load(open(GOLDEN))
    for name, tr in small.items():
        proof = dg.prove(tr)
        ref = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
        assert ref.error is None
        assert proof.trace_root == ref.digest("trace_root"), name
        assert proof.constraint_root == ref.digest("constraint_root"), name
        assert proof.pow_nonce == ref.u64s("pow_nonce")[0], name
        assert proof.bytes == ref.proof, name
        assert hashlib.sha256(proof.bytes).hexdigest() == golden[name]["proof_sha256"], name
        assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof.bytes) is None, name
        assert proof.stats["kernel_launches"] > 0


@pytest.mark.parametrize("prog,ext,queries,grinding", [("collatz3", 16, 30, 8), ("collatz3", 64, 20, 12), ("collatz3", 128, 10, 0),
                                                        ("fib13", 256, 5, 4), ("collatz3", 256, 5, 4), ("hash", 128, 128, 1)])
def test_other_proof_options(dg, po, small, prog, ext, queries, grinding):
    tr = small[prog]
    proof = dg.prove(tr, dg.ProofOptions(ext, queries, grinding))
    ref = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs, ext=ext, num_queries=queries, grinding=grinding)
    assert ref.error is None
    _same(proof, ref)
    verdict = po.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof.bytes)
    if prog == "collatz3" and ext == 256:
        # reference quirk, reproduced by the restated verifier: fri/verifier.rs:86 divides max_degree_plus_1 by 4 per layer with
        # truncation, so for blowup 256 and odd log2(trace length) (here 2^11) it demands a remainder of degree 2 where the
        # honest remainder has degree 3.  The prover output is still byte-identical to the reference prover's.
        assert verdict == "verification of low-degree proof failed: remainder is not a valid degree 2 polynomial"
    else:
        assert verdict is None


def test_medium_traces(dg, po):
    from distaff_b200 import hostvm
    for tr in (hostvm.fibonacci(250), programs.merkle_example(20, po), hostvm.collatz(7)):     # n = 4096, 2048, 4096
        proof = dg.prove(tr)
        ref = po.prove(tr.registers, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
        assert ref.error is None
        assert proof.bytes == ref.proof
        assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof.bytes) is None


def test_device_resident_entry_point(dg, po, small):
    from distaff_b200 import backend
    tr = small["fib13"]
    buf = backend.DeviceBuffer(tr.registers.nbytes).upload(tr.registers)
    proof = dg.prove_device(buf, tr.width, tr.length, tr.ctx_depth, tr.loop_depth, tr.public_inputs, tr.outputs)
    assert proof.bytes == dg.prove(tr).bytes


def test_invalid_trace_is_reported_not_proven(dg, small):
    from distaff_b200 import backend, hostvm
    tr = small["fib13"]
    regs = tr.registers.copy()
    regs[tr.width - 1, 100, 0] += 1
    bad = hostvm.ExecutionTrace(regs, tr.ctx_depth, tr.loop_depth, tr.stack_depth, tr.program_hash, tr.public_inputs, tr.outputs)
    with pytest.raises(backend.DgError) as e:
        dg.prove(bad)
    assert e.value.code == -5                    # DG_ERR_UNSATISFIED  (evaluator.rs:152-157 panics in the reference)


def test_argument_validation(dg, small):
    from distaff_b200 import backend, hostvm
    tr = small["fib13"]
    short = hostvm.ExecutionTrace(tr.registers[:, :8].copy(), tr.ctx_depth, tr.loop_depth, tr.stack_depth, tr.program_hash, [1, 0], [1])
    with pytest.raises(backend.DgError):
        dg.prove(short)                          # trace shorter than MIN_TRACE_LENGTH
    with pytest.raises(AssertionError):
        dg.ProofOptions(8, 50, 20)               # options.rs:36


def test_host_rng_callbacks_give_the_same_proof(dg, po, small):
    """dg_set_rng_callbacks: all Fiat-Shamir draws supplied by the host -- here by the oracle's restatement of
    field::prng_vector (field.rs:264-275) and compute_query_positions (stark/utils/mod.rs:25-44) through C function pointers -- must give
    the byte-identical proof, and the callbacks must really be on the path (call counts, failure propagation)."""
    from distaff_b200 import backend, felt
    calls = {"field": 0, "positions": 0}

    def draw_field(seed, count):
        calls["field"] += 1
        return felt.from_ints(po.prng_vector(seed, count)).tobytes()

    def draw_positions(seed, domain, ext, nq):
        calls["positions"] += 1
        return po.query_positions(seed, domain, ext, nq)

    try:
        for name in ("fib13", "collatz3", "hash"):
            tr = small[name]
            backend.set_rng_callbacks()
            want = dg.prove(tr).bytes
            backend.set_rng_callbacks(draw_field, draw_positions)
            before = dict(calls)
            got = dg.prove(tr).bytes
            assert got == want, name
            assert calls["positions"] == before["positions"] + 1
            assert calls["field"] >= before["field"] + 3          # constraint coefficients, composition coefficients, >= 1 FRI layer
        # a failing position callback surfaces as the reference's "needed more query positions" condition (DG_ERR_EXHAUSTED)
        backend.set_rng_callbacks(draw_field, lambda *a: (_ for _ in ()).throw(RuntimeError("no")))
        with pytest.raises(backend.DgError) as e:
            dg.prove(small["fib13"])
        assert e.value.code == -4
        # positions that violate compute_query_positions' invariants are rejected, not proven
        backend.set_rng_callbacks(draw_field, lambda seed, domain, ext, nq: [ext * (i + 1) for i in range(nq)])
        with pytest.raises(backend.DgError):
            dg.prove(small["fib13"])
    finally:
        backend.set_rng_callbacks()
    assert dg.prove(small["fib13"]).bytes == dg.prove(small["fib13"]).bytes

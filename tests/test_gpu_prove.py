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

"""-m gpu parity of the GPU verifier (dg_verify, distaff_b200/csrc/verifier.cu) with the restated reference verifier of the oracle
(/root/reference/src/stark/verifier.rs:11-75, fri/verifier.rs:11-131, crypto/merkle.rs:154-263): same accept / reject decision and the
same error string on honest proofs, on the three tampering cases of the reference's own test (src/tests/mod.rs:32-63) and on proofs
with single corrupted bytes in every section of the encoding."""
import pytest

from tests import programs

pytestmark = pytest.mark.gpu
LDP = "verification of low-degree proof failed: evaluations did not match column value at depth 0"


@pytest.fixture(scope="module")
def dg():
    import distaff_b200
    from distaff_b200 import backend
    backend.device_info()
    return distaff_b200


@pytest.fixture(scope="module")
def proofs(dg):
    out = {}
    for name, tr in programs.small_programs().items():
        out[name] = (tr, dg.prove(tr).bytes)
    return out


def test_honest_proofs_are_accepted(dg, po, proofs):
    for name, (tr, proof) in proofs.items():
        assert po.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof) is None, name
        assert dg.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof) is None, name


def test_reference_tampering_cases(dg, po, proofs):
    """src/tests/mod.rs:32-63: wrong inputs, wrong outputs, wrong program hash -- all die in the FRI consistency check"""
    tr, proof = proofs["fib_span"]
    assert tr.public_inputs == [1, 0] and tr.outputs == [3]
    bad_hash = bytes([1]) + tr.program_hash[1:]
    for args in ((tr.program_hash, [1, 1], tr.outputs), (tr.program_hash, tr.public_inputs, [5]), (bad_hash, tr.public_inputs, tr.outputs)):
        assert po.verify(*args, proof) == LDP
        assert dg.verify(*args, proof) == LDP


def test_other_options_including_the_blowup_256_quirk(dg, po):
    small = programs.small_programs()
    for prog, ext, queries, grinding in (("collatz3", 16, 30, 8), ("collatz3", 64, 20, 12), ("fib13", 256, 5, 4), ("collatz3", 256, 5, 4), ("hash", 128, 128, 1)):
        tr = small[prog]
        proof = dg.prove(tr, dg.ProofOptions(ext, queries, grinding)).bytes
        want = po.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof)
        assert dg.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof) == want, (prog, ext)
    # fri/verifier.rs:86 rejects the honest blowup-256 proof of a 2^11-step trace (truncating division): reproduced on both sides
    assert want is None or "remainder" in want


def test_corrupted_bytes_get_the_same_verdict(dg, po, proofs):
    from distaff_b200 import backend
    import random
    rng = random.Random(7)
    checked = rejected = 0
    for name in ("fib13", "collatz3", "hash", "deep_stack"):
        tr, proof = proofs[name]
        # offsets spread over the whole encoding: roots, node lists, opened rows, constraint leaves, deep values, FRI layers, remainder, nonce
        offsets = [0, 31, 36, 40] + [rng.randrange(44, len(proof) - 16) for _ in range(40)] + [len(proof) - 12, len(proof) - 5]
        for off in offsets:
            bad = bytearray(proof)
            bad[off] ^= 1 << rng.randrange(8)
            bad = bytes(bad)
            want = po.verify(tr.program_hash, tr.public_inputs, tr.outputs, bad)
            try:
                got = dg.verify(tr.program_hash, tr.public_inputs, tr.outputs, bad)
            except backend.DgError:
                got = "malformed"
            if want is not None and (want.startswith("exception") or "too short" in want):
                assert got is not None, (name, off)          # bytes that no longer parse into a well-formed proof: both refuse, each in its own words
            else:
                assert got == want, (name, off, got, want)
            checked += 1
            rejected += want is not None
    assert checked > 150 and rejected > 100


def test_truncated_bytes_are_refused(dg, proofs):
    from distaff_b200 import backend
    tr, proof = proofs["fib13"]
    for cut in (0, 10, 100, len(proof) - 1):
        with pytest.raises(backend.DgError):
            dg.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof[:cut])
    with pytest.raises(backend.DgError):
        dg.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof + b"\0")


def test_headline_proof_verifies_on_the_gpu(dg, po):
    import bench
    tr, _ = bench.build_trace(16)
    proof = dg.prove(tr).bytes
    assert dg.verify(tr.program_hash, tr.public_inputs, tr.outputs, proof) is None
    assert dg.verify(tr.program_hash, tr.public_inputs, [tr.outputs[0] + 1], proof) == LDP

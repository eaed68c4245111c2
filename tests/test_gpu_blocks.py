"""-m gpu parity tests of the CUDA building blocks against the CPU oracle, through the C-ABI (distaff_b200.api).
Bit-exact: everything here is integer / byte work."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M = 2**128 - 45 * 2**40 + 1


@pytest.fixture(scope="module")
def dg():
    import distaff_b200
    from distaff_b200 import backend
    info = backend.device_info()          # fails loudly without a GPU / without the built .so
    assert info["sm_count"] > 0
    return distaff_b200


def _rand(n, seed):
    from distaff_b200 import felt
    return felt.random_elements(n, seed)


EDGE = [0, 1, 2, M - 1, M - 2, 2**64 - 1, 2**64, 2**64 + 1, 2**127, (M + 1) // 2, 45 * 2**40 - 1, 45 * 2**40, 2**128 - 2**88, M - 2**64]


def test_field_ops_ptx_and_portable_match_oracle(dg, po):
    from distaff_b200 import felt
    a_int = [x for x in EDGE for _ in EDGE] + felt.to_ints(_rand(4000, 1))
    b_int = [y for _ in EDGE for y in EDGE] + felt.to_ints(_rand(4000, 2))
    a, b = felt.from_ints(a_int), felt.from_ints(b_int)
    for op, f in (("add", lambda x, y: (x + y) % M), ("sub", lambda x, y: (x - y) % M), ("mul", lambda x, y: x * y % M)):
        want = [f(x, y) for x, y in zip(a_int, b_int)]
        for impl in (0, 1):
            assert felt.to_ints(dg.field_op(op, a, b, impl=impl)) == want, (op, impl)
    sub = felt.from_ints(a_int[:300])
    want = [pow(x, M - 2, M) if x else 0 for x in a_int[:300]]
    assert felt.to_ints(dg.field_op("inv", sub, impl=0)) == want
    e = felt.from_ints(b_int[:300])
    want = [pow(x, y, M) if x else 0 for x, y in zip(a_int[:300], b_int[:300])]
    assert felt.to_ints(dg.field_op("exp", sub, e, impl=0)) == want
    # spot-check against the oracle's limb-for-limb restatement of field.rs::mul
    for x, y in list(zip(a_int, b_int))[:400]:
        assert po.field_op("mul", x, y) == x * y % M


def test_mul_crafted_reduction_paths(dg):
    """operand pairs whose product is congruent to a tiny residue: the Solinas fold overflows 2^128 / ends with an all-ones top limb,
    paths random operands reach with probability ~2^-35 (tools/gen_mul_vectors.py); checked against Python integers for every multiply"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gen_mul_vectors
    from distaff_b200 import felt
    pairs = gen_mul_vectors.pairs()
    a = felt.from_ints([p[0] for p in pairs])
    b = felt.from_ints([p[1] for p in pairs])
    want = [(x * y) % M for x, y in pairs]
    for impl in (0, 1, 2, 3):
        assert felt.to_ints(dg.field_op("mul", a, b, impl=impl)) == want, impl
    # sums in [M, 2^128) (all-ones top limb without overflow) and differences that borrow
    near = [M - 1, M - 2, M - 2**32, M - 2**64, M - 2**96, 2**128 - 2**96 - 1 - (2**128 - M), 2**127, 2**127 - 1, 1, 2, 2**96, 2**96 - 1, 45 * 2**40, 45 * 2**40 - 1]
    xs = [x % M for x in near for _ in near]
    ys = [y % M for _ in near for y in near]
    fa, fb = felt.from_ints(xs), felt.from_ints(ys)
    for impl in (0, 1):
        assert felt.to_ints(dg.field_op("add", fa, fb, impl=impl)) == [(x + y) % M for x, y in zip(xs, ys)], impl
        assert felt.to_ints(dg.field_op("sub", fa, fb, impl=impl)) == [(x - y) % M for x, y in zip(xs, ys)], impl


def test_unreduced_dot_products(dg):
    """288-bit accumulation of 6 products reduced once (fe_dot, used by the constraint kernel's mat-vecs) == Python integers;
    operands include the crafted pairs and all-(M-1) rows (largest accumulator)"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gen_mul_vectors
    from distaff_b200 import felt
    pairs = gen_mul_vectors.pairs()
    xs = [p[0] for p in pairs] + [M - 1] * 60 + felt.to_ints(_rand(6000, 31))
    ys = [p[1] for p in pairs] + [M - 1] * 60 + felt.to_ints(_rand(6000, 32))
    n = len(xs) // 6 * 6
    xs, ys = xs[:n], ys[:n]
    want = [sum(xs[6 * i + j] * ys[6 * i + j] for j in range(6)) % M for i in range(n // 6)]
    for impl in (0, 1):
        assert felt.to_ints(dg.field_op("dot6", felt.from_ints(xs), felt.from_ints(ys), impl=impl)) == want, impl


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 5, 8, 10, 11, 12, 13, 16, 19, 20, 21, 22])
def test_ntt_matches_oracle(dg, po, log_n):
    n = 1 << log_n
    batch = 3 if log_n <= 16 else 1
    x = _rand(n * batch, 100 + log_n).reshape(batch, n, 2)
    y = dg.ntt(x)
    for b in range(batch):
        assert np.array_equal(y[b], po.fft(x[b])), (log_n, b)
    back = dg.intt(y)
    assert np.array_equal(back, x)


def test_ntt_edge_inputs(dg, po):
    from distaff_b200 import felt
    n = 1 << 12
    for vec in ([0] * n, [M - 1] * n, [1] + [0] * (n - 1), list(range(n))):
        x = felt.from_ints(vec)
        assert np.array_equal(dg.ntt(x), po.fft(x))
        assert np.array_equal(dg.intt(x), po.fft(x, inverse=True))


@pytest.mark.parametrize("log_n,blowup", [(4, 16), (4, 32), (8, 32), (10, 32), (11, 32), (13, 32), (16, 32), (12, 64)])
def test_lde_matches_reference_extension(dg, po, log_n, blowup):
    # TraceTable::extend: interpolate (iNTT n), zero-pad to N, evaluate (NTT N)   trace_table.rs:143-169
    n = 1 << log_n
    batch = 3 if log_n <= 13 else 2
    x = _rand(n * batch, 200 + log_n).reshape(batch, n, 2)
    got = dg.lde(x, blowup)
    for b in range(batch):
        poly = po.fft(x[b], inverse=True)
        padded = np.zeros((n * blowup, 2), dtype=np.uint64)
        padded[:n] = poly
        want = po.fft(padded)
        assert np.array_equal(got[b], want), (log_n, b)
        assert np.array_equal(got[b, ::blowup], x[b])        # the extension agrees with the trace on the trace domain


@pytest.mark.parametrize("w", [1, 3, 4, 5, 16, 17, 20, 26, 63, 64, 65, 100, 127])
def test_row_hashing_matches_blake3(dg, po, w):
    rows = 257
    cols = _rand(w * rows, 300 + w).reshape(w, rows, 2)
    got = dg.hash_rows(cols)
    for r in (0, 1, 17, 255, 256):
        row_bytes = cols[:, r, :].tobytes()
        assert got[32 * r:32 * r + 32] == po.hash("blake3", row_bytes), (w, r)


@pytest.mark.parametrize("log_l", [1, 2, 3, 5, 10, 11, 12, 16])
def test_merkle_nodes_match_oracle(dg, po, log_l):
    n = 1 << log_l
    leaves = np.random.Generator(np.random.PCG64(log_l)).integers(0, 256, size=n * 32, dtype=np.uint8).tobytes()
    assert dg.merkle_build(leaves) == po.merkle_nodes("blake3", leaves)


def _le16(values):
    return b"".join(int(v).to_bytes(16, "little") for v in values)


def test_algebraic_hash_reference_vectors(dg):
    """known answers of the reference's own tests (hash.rs:264-297): poseidon / rescue of [1, 2, 3, 4]"""
    msg = _le16([1, 2, 3, 4])
    assert list(dg.hash64(msg, "poseidon")) == [
        224, 9, 85, 92, 75, 117, 136, 23, 142, 67, 249, 199, 39, 177, 97, 129,
        93, 192, 153, 131, 76, 160, 94, 162, 200, 192, 187, 5, 159, 69, 48, 165]
    assert list(dg.hash64(msg, "rescue")) == [
        148, 191, 96, 185, 107, 196, 170, 28, 161, 214, 196, 211, 158, 111, 135, 32,
        122, 173, 195, 37, 123, 60, 246, 104, 176, 53, 127, 67, 38, 208, 69, 54]


@pytest.mark.parametrize("name", ["blake3", "rescue", "poseidon"])
def test_hash64_matches_oracle(dg, po, name):
    n = 300                                              # not a multiple of the block size
    msgs = _rand(4 * n, 77).reshape(n, 4, 2)             # valid field elements
    msgs[0] = 0                                          # all-zero message
    msgs[1, :, 0] = M & 0xFFFFFFFFFFFFFFFF               # M - 1 in every lane
    msgs[1, :, 1] = M >> 64
    msgs[1, :, 0] -= 1
    raw = msgs.tobytes()
    got = dg.hash64(raw, name)
    assert len(got) == 32 * n
    for i in range(n):
        assert got[32 * i:32 * i + 32] == po.hash(name, raw[64 * i:64 * i + 64]), (name, i)
    assert dg.hash64(b"", name) == b""                   # empty batch


@pytest.mark.parametrize("name", ["rescue", "poseidon"])
@pytest.mark.parametrize("log_l", [1, 2, 3, 6, 9])
def test_algebraic_merkle_nodes_match_oracle(dg, po, name, log_l):
    n = 1 << log_l
    leaves = _rand(2 * n, 900 + log_l).tobytes()         # leaves are digests = pairs of field elements (merkle.rs:321-338)
    assert dg.merkle_build(leaves, name) == po.merkle_nodes(name, leaves)


def test_pow_nonce_is_the_smallest(dg, po):
    import ctypes
    for g, seed_byte in ((0, 1), (8, 2), (12, 3), (16, 4), (20, 5)):
        seed = bytes([seed_byte] * 32)
        new_seed, nonce = dg.find_pow_nonce(seed, g)
        out = ctypes.create_string_buffer(32)
        want = po.lib().or_find_pow_nonce(seed, g, out)
        assert nonce == want and new_seed == out.raw, g


def test_invalid_arguments_return_errors(dg):
    from distaff_b200 import backend
    with pytest.raises(backend.DgError):
        dg.merkle_build(b"\0" * 96)                      # 3 leaves: not a power of two (merkle.rs:26)
    with pytest.raises(backend.DgError):
        dg.find_pow_nonce(b"\0" * 32, 33)                # options.rs:43

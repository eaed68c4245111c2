"""Pins oracle/math.h against the reference's own unit tests (re-stated with Python big ints):
/root/reference/src/math/field.rs:346-447, fft.rs:117-157, polynom.rs:287-515, quartic.rs:170-227."""
import random

import numpy as np

M = 2**128 - 45 * 2**40 + 1
G = 23953097886125630542083529559205016746


def rnd(rng):
    return rng.randrange(M)


def test_field_known_answers(po):
    # field.rs:346-402
    assert po.field_op("add", 2, 3) == 5
    assert po.field_op("add", M - 1, 1) == 0
    assert po.field_op("add", M - 1, 2) == 1
    assert po.field_op("sub", 5, 3) == 2
    assert po.field_op("sub", 3, 5) == M - 2
    assert po.field_op("mul", 5, 3) == 15
    t = M - 1
    assert po.field_op("mul", t, t) == 1
    assert po.field_op("mul", t, 2) == M - 2
    assert po.field_op("mul", t, 4) == M - 4
    assert po.field_op("mul", (M + 1) // 2, 2) == 1
    assert po.field_op("inv", 1) == 1
    assert po.field_op("inv", 0) == 0
    assert po.field_op("exp", 0, 5) == 0
    assert po.field_op("exp", 7, 0) == 1


def test_field_random_vs_bigint(po):
    rng = random.Random(1)
    edge = [0, 1, 2, M - 1, M - 2, 2**64 - 1, 2**64, 2**127, (M + 1) // 2, 45 * 2**40 - 1]
    pairs = [(a, b) for a in edge for b in edge] + [(rnd(rng), rnd(rng)) for _ in range(2000)]
    for a, b in pairs:
        assert po.field_op("add", a, b) == (a + b) % M
        assert po.field_op("sub", a, b) == (a - b) % M
        assert po.field_op("mul", a, b) == (a * b) % M
    for _ in range(200):
        x = rnd(rng)
        assert po.field_op("mul", x, po.field_op("inv", x)) == (1 if x else 0)
        p = rnd(rng)
        assert po.field_op("exp", x, p) == pow(x, p, M)


def test_roots_of_unity(po):
    # field.rs:438-447
    assert po.root_of_unity(2**40) == G
    assert pow(G, 2**40, M) == 1 and pow(G, 2**39, M) != 1
    assert po.root_of_unity(2**39) == pow(G, 2, M)
    for k in (1, 4, 10, 25):
        assert po.root_of_unity(2**k) == pow(G, 2**(40 - k), M)


def test_inv_many_maps_zero_to_zero(po):
    rng = random.Random(2)
    vals = [rnd(rng) for _ in range(50)]
    vals[3] = 0
    vals[17] = 0
    a = po.fvec(vals)
    out = np.zeros_like(a)
    po.lib().or_inv_many(a.ctypes.data, out.ctypes.data, len(vals))
    for v, r in zip(vals, po.ints(out)):
        assert r == (pow(v, M - 2, M) if v else 0)


def test_fft_matches_naive_evaluation(po):
    # fft.rs:117-157: fft_in_place + permute == polynom::eval at all domain points
    rng = random.Random(3)
    for n in (4, 8, 16, 256, 1024):
        p = [rnd(rng) for _ in range(n)]
        w = pow(G, 2**40 // n, M)
        got = po.ints(po.fft(po.fvec(p)))
        pts = range(n) if n <= 256 else list(range(0, n, 97)) + [n - 1]
        for i in pts:
            x = pow(w, i, M)
            assert got[i] == sum(c * pow(x, k, M) for k, c in enumerate(p)) % M
        back = po.ints(po.fft(po.fvec(got), inverse=True))
        assert back == p


def test_fft_in_place_with_reference_twiddles(po):
    # same contract through the raw entry point with bit-reversed twiddles (fft.rs:58-63)
    rng = random.Random(4)
    n = 64
    p = po.fvec([rnd(rng) for _ in range(n)])
    root = po.fvec([po.root_of_unity(n)])
    tw = np.zeros((n // 2, 2), dtype=np.uint64)
    po.lib().or_get_twiddles(root.ctypes.data, n, 0, tw.ctypes.data)
    a = p.copy()
    po.lib().or_fft_in_place(a.ctypes.data, n, tw.ctypes.data, 1)
    assert po.ints(a) == po.ints(po.fft(p))


def _poly_divmod(a, b):
    a = a[:]
    out = [0] * (len(a) - len(b) + 1)
    inv = pow(b[-1], M - 2, M)
    for i in range(len(out) - 1, -1, -1):
        q = a[i + len(b) - 1] * inv % M
        out[i] = q
        for j, c in enumerate(b):
            a[i + j] = (a[i + j] - q * c) % M
    return out, a


def test_syn_div_equals_long_division(po):
    # polynom.rs:456-463 and the remainder-dropping behaviour (polynom.rs:180-197)
    rng = random.Random(5)
    for n in (2, 3, 8, 33):
        a = [rnd(rng) for _ in range(n)]
        b = rnd(rng)
        arr = po.fvec(a)
        fb = po.fvec([b])
        po.lib().or_syn_div(arr.ctypes.data, n, fb.ctypes.data)
        q, _ = _poly_divmod(a, [(-b) % M, 1])
        assert po.ints(arr) == q + [0]


def test_syn_div_expanded_matches_reference_test(po):
    # polynom.rs:466-490
    ys = [0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 13, 14, 15]
    poly = po.ints(po.fft(po.fvec(ys), inverse=True))
    root = po.root_of_unity(16)
    d12 = pow(root, 12, M)
    z_poly, rem = _poly_divmod([M - 1, 0, 0, 0, 1], [(-d12) % M, 1])
    assert all(r == 0 for r in rem[:1])
    arr = po.fvec(poly)
    exc = po.fvec([d12])
    po.lib().or_syn_div_expanded(arr.ctypes.data, 16, 4, exc.ctypes.data, 1)
    expected, rem = _poly_divmod(poly, z_poly)
    got = po.ints(arr)
    assert got[:len(expected)] == expected
    assert all(v == 0 for v in got[len(expected):])


def test_quartic_interpolate_batch_equals_lagrange(po):
    # quartic.rs:178-191
    rng = random.Random(6)
    n = 7
    xs = [[rnd(rng) for _ in range(4)] for _ in range(n)]
    ys = [[rnd(rng) for _ in range(4)] for _ in range(n)]
    fx = po.fvec([v for r in xs for v in r])
    fy = po.fvec([v for r in ys for v in r])
    out = np.zeros_like(fx)
    po.lib().or_quartic_interpolate_batch(fx.ctypes.data, fy.ctypes.data, n, out.ctypes.data)
    got = po.ints(out)
    for r in range(n):
        coeffs = got[4 * r:4 * r + 4]
        for x, y in zip(xs[r], ys[r]):
            assert sum(c * pow(x, k, M) for k, c in enumerate(coeffs)) % M == y
        lx, ly = po.fvec(xs[r]), po.fvec(ys[r])
        lo = np.zeros_like(lx)
        po.lib().or_lagrange(lx.ctypes.data, ly.ctypes.data, 4, lo.ctypes.data)
        assert po.ints(lo) == coeffs


def test_quartic_transpose_layout(po):
    # quartic.rs:220-227
    v = po.fvec(list(range(16)))
    out = np.zeros_like(v)
    po.lib().or_quartic_transpose(v.ctypes.data, 16, 1, out.ctypes.data)
    assert po.ints(out) == [0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15]
    out2 = np.zeros((8, 2), dtype=np.uint64)
    po.lib().or_quartic_transpose(v.ctypes.data, 16, 2, out2.ctypes.data)
    assert po.ints(out2) == [0, 4, 8, 12, 2, 6, 10, 14]

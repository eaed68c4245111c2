#!/usr/bin/env python3
"""Limb-level model of fe_reduce_wide (fp128.cuh): reduction of a 9-limb (288-bit) accumulator of up to 128 products modulo
M = 2^128 - C, C = K*2^32 - 1, K = 11520.  Every multiply-add is checked to stay inside 64 bits, every borrow chain inside its width;
the result is compared with Python integers.  python tools/model_reduce9.py"""
import random

M = 2**128 - 45 * 2**40 + 1
C = 2**128 - M
K = 11520
MASK32 = 2**32 - 1


def wmad(a, b, c):
    assert 0 <= a < 2**32 and 0 <= b < 2**32 and 0 <= c < 2**64
    d = a * b + c
    assert d < 2**64, "mad.wide overflow"
    return d


def reduce9(r):
    assert len(r) == 9 and all(0 <= x < 2**32 for x in r)
    # step 1: V = lo + ((H*K) << 32) - H, H = r[4..8]
    t0 = wmad(r[4], K, r[1])
    t1 = wmad(r[2], 1, wmad(r[5], K, t0 >> 32))
    t2 = wmad(r[3], 1, wmad(r[6], K, t1 >> 32))
    t3 = wmad(r[7], K, t2 >> 32)
    t4 = wmad(r[8], K, t3 >> 32)
    v = [r[0], t0 & MASK32, t1 & MASK32, t2 & MASK32, t3 & MASK32, t4 & MASK32, t4 >> 32]
    V = sum(x << (32 * i) for i, x in enumerate(v))
    H = sum(r[4 + i] << (32 * i) for i in range(5))
    V -= H
    assert V >= 0
    assert V == sum(r[i] << (32 * i) for i in range(4)) + H * C
    v = [(V >> (32 * i)) & MASK32 for i in range(7)]
    assert v[6] == 0 and v[5] < 2**22, (v[5], v[6])
    # step 2: fold T = v5:v4
    u1 = wmad(v[4], K, v[1])
    u2 = wmad(v[2], 1, wmad(v[5], K, u1 >> 32))
    u3 = wmad(v[3], 1, u2 >> 32)
    w = [v[0], u1 & MASK32, u2 & MASK32, u3 & MASK32, u3 >> 32]
    W = sum(x << (32 * i) for i, x in enumerate(w)) - (v[4] + (v[5] << 32))
    assert 0 <= W < 2**128 + 2**99
    cy = W >> 128
    assert cy in (0, 1)
    lo = W & (2**128 - 1)
    # canonical form: rare path when cy or top limb all ones
    if cy or (lo >> 96) == MASK32:
        q = lo + C
        g = q >> 128
        if cy or g:
            lo = q & (2**128 - 1)
    assert lo < M
    return lo


def main():
    rnd = random.Random(7)
    n = 0
    for trial in range(200000):
        kind = trial % 5
        if kind == 0:
            val = rnd.getrandbits(288) % (128 * M * M)
        elif kind == 1:
            val = sum((rnd.randrange(M) * rnd.randrange(M)) for _ in range(rnd.randrange(1, 129)))
        elif kind == 2:
            val = 128 * (M - 1) * (M - 1) - rnd.getrandbits(rnd.randrange(1, 200))
        elif kind == 3:
            val = rnd.getrandbits(rnd.randrange(1, 262))
        else:
            t = rnd.randrange(1, 2**52)
            val = (t << 128) + 2**128 - rnd.randrange(1, t * C + 1)      # drives the final overflow
            val = min(val, 128 * M * M - 1)
        r = [(val >> (32 * i)) & MASK32 for i in range(9)]
        assert val < 2**288
        if r[8] >= 128:
            continue
        assert reduce9(r) == val % M
        n += 1
    print(n, "accumulators reduced correctly")


if __name__ == "__main__":
    main()

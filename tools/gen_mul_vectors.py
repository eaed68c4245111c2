#!/usr/bin/env python3
"""Crafted operand pairs for the modular multiplication: canonical a, b < M whose product is congruent to a tiny residue rho, so that
the partially reduced value of the Solinas fold overflows 2^128 (rho >= C - j) or ends with an all-ones top limb.  Random operands hit
these paths with probability ~2^-35; tests/test_gpu_blocks.py and tools/bench_modmul.cu read the pairs.
    python tools/gen_mul_vectors.py            -> tools/_bin/mul_vectors.bin (pairs of 16-byte little-endian elements)"""
import os
import random

M = 2**128 - 45 * 2**40 + 1
C = 45 * 2**40 - 1


def pairs(seed=1234, per=6):
    rnd = random.Random(seed)
    out = []
    rhos = [0, 1, 2, C - 2, C - 1, C, C + 1, C + 2, 2 * C - 3, 2 * C - 2, 2 * C - 1, 2 * C, 2**46, 2**60 + 12345, 2**64 - 1, 2**64, 2**92, 2**93 - 1,
            M - 1, M - 2, M - C, M - C - 1, M - 2**32, M - 2**64, M - 2**96, M - 2**96 - 1, M - 2**96 + 1, 2**128 - 2**96 - 1 - (2**128 - M)]
    for rho in rhos:
        rho %= M
        for bits in (47, 64, 90, 100, 110, 120, 127, 128)[:per + 2]:
            b = rnd.getrandbits(bits) % M
            if b == 0:
                b = 3
            a = rho * pow(b, M - 2, M) % M
            out.append((a, b))
            out.append((b, a))
    return out


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "tools", "_bin"), exist_ok=True)
    with open(os.path.join(root, "tools", "_bin", "mul_vectors.bin"), "wb") as f:
        for a, b in pairs():
            f.write(a.to_bytes(16, "little") + b.to_bytes(16, "little"))
    print(len(pairs()), "pairs")


if __name__ == "__main__":
    main()

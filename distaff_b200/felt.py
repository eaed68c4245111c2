"""Field-element array helpers shared by the host-side Python code.

A vector of field elements of F_M (M = 2^128 - 45*2^40 + 1, /root/reference/src/math/field.rs:11) is a
C-contiguous numpy array of dtype uint64 and shape (..., 2): [..., 0] = low 64 bits, [..., 1] = high 64 bits,
i.e. exactly the 16 little-endian bytes of a Rust u128 (/root/reference/src/utils/mod.rs:35-41).
"""
import numpy as np

M = 2**128 - 45 * 2**40 + 1
G = 23953097886125630542083529559205016746   # 2^40-th root of unity, field.rs:14
_MASK = 2**64 - 1


def from_ints(values):
    """list/iterable of python ints -> (n, 2) uint64 array"""
    vals = list(values)
    a = np.empty((len(vals), 2), dtype=np.uint64)
    for i, v in enumerate(vals):
        a[i, 0] = v & _MASK
        a[i, 1] = v >> 64
    return a


def to_ints(a):
    """(..., 2) uint64 array -> flat list of python ints"""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in a]


def as_bytes(a):
    return np.ascontiguousarray(a, dtype=np.uint64).tobytes()


def from_bytes(b):
    return np.frombuffer(bytes(b), dtype=np.uint64).reshape(-1, 2).copy()


def root_of_unity(order):
    """field::get_root_of_unity (field.rs:228-234)"""
    assert order > 0 and order & (order - 1) == 0 and order <= 2**40
    return pow(G, 2**40 // order, M)


def random_elements(n, seed):
    """n pseudo-random canonical field elements (deterministic; SplitMix64 -> 128 bits -> mod M)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    raw = rng.integers(0, 2**64, size=(n, 2), dtype=np.uint64)
    # reduce values >= M (probability ~2^-82 each, handled anyway)
    hi_max = np.uint64(M >> 64)
    lo_m = np.uint64(M & _MASK)
    bad = (raw[:, 1] == hi_max) & (raw[:, 0] >= lo_m)
    raw[bad, 0] -= lo_m
    raw[bad, 1] = 0
    return raw

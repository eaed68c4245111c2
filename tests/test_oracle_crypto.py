"""Pins oracle/crypto.h: hash KATs of the reference (/root/reference/src/crypto/hash.rs:264-297), BLAKE3 against the
official implementation (pip `blake3`), Merkle tree structure tests (/root/reference/src/crypto/merkle.rs:339-518),
ChaCha20 keystream against `cryptography`, and the StdRng/Uniform restatement cross-check values."""
import struct

import numpy as np
import pytest


def _le(values):
    return b"".join(int(v).to_bytes(16, "little") for v in values)


def test_algebraic_hash_kats(po):
    msg = _le([1, 2, 3, 4])
    assert list(po.hash("poseidon", msg)) == [
        224, 9, 85, 92, 75, 117, 136, 23, 142, 67, 249, 199, 39, 177, 97, 129,
        93, 192, 153, 131, 76, 160, 94, 162, 200, 192, 187, 5, 159, 69, 48, 165]
    assert list(po.hash("rescue", msg)) == [
        148, 191, 96, 185, 107, 196, 170, 28, 161, 214, 196, 211, 158, 111, 135, 32,
        122, 173, 195, 37, 123, 60, 246, 104, 176, 53, 127, 67, 38, 208, 69, 54]
    assert list(po.hash("gmimc", msg)) == [
        115, 208, 64, 41, 162, 43, 134, 243, 236, 80, 161, 106, 195, 234, 30, 26,
        71, 74, 255, 77, 41, 125, 25, 152, 162, 106, 65, 108, 84, 216, 37, 37]
    with pytest.raises(ValueError):
        po.hash("rescue", b"\0" * 65)      # hash.rs:152


def test_blake3_matches_official_implementation(po):
    blake3 = pytest.importorskip("blake3")
    data = bytes((i * 131 + 7) % 251 for i in range(5000))
    lengths = list(range(0, 130)) + [255, 256, 320, 400, 448, 1023, 1024, 1025, 1040, 2031, 2032, 2048, 2049, 3072, 4097, 5000]
    for n in lengths:
        assert po.hash("blake3", data[:n]) == blake3.blake3(data[:n]).digest(), n


def _h2(po, a, b):
    return po.hash("poseidon", a + b)


LEAVES8 = [bytes([(37 * i + 11 * j + 5) % 256 for j in range(32)]) for i in range(8)]


def test_merkle_tree_structure(po):
    # merkle.rs:340-362: nodes[1] is the root, heap layout
    for n in (4, 8):
        leaves = LEAVES8[:n]
        nodes = po.merkle_nodes("poseidon", b"".join(leaves))
        level = leaves
        while len(level) > 1:
            level = [_h2(po, level[i], level[i + 1]) for i in range(0, len(level), 2)]
        assert nodes[32:64] == level[0]
        assert nodes[:32] == b"\0" * 32
    nodes = po.merkle_nodes("poseidon", b"".join(LEAVES8))
    assert nodes[4 * 32:5 * 32] == _h2(po, LEAVES8[0], LEAVES8[1])
    assert nodes[2 * 32:3 * 32] == _h2(po, _h2(po, LEAVES8[0], LEAVES8[1]), _h2(po, LEAVES8[2], LEAVES8[3]))


def _parse_batch(raw):
    off = 0

    def u64():
        nonlocal off
        v = struct.unpack_from("<Q", raw, off)[0]
        off += 8
        return v

    def dvec():
        nonlocal off
        n = u64()
        out = [raw[off + 32 * i: off + 32 * i + 32] for i in range(n)]
        off += 32 * n
        return out
    values = dvec()
    nodes = [dvec() for _ in range(u64())]
    depth = raw[off]
    return values, nodes, depth


def test_merkle_prove_batch_reference_cases(po):
    # merkle.rs:427-493 (exact value / node lists)
    L = LEAVES8
    h = lambda a, b: _h2(po, a, b)
    leaves = b"".join(L)
    v, nodes, depth = _parse_batch(po.merkle_prove_batch("poseidon", leaves, [1]))
    assert (v, depth) == ([L[1]], 3)
    assert nodes == [[L[0], h(L[2], L[3]), h(h(L[4], L[5]), h(L[6], L[7]))]]
    v, nodes, depth = _parse_batch(po.merkle_prove_batch("poseidon", leaves, [1, 2]))
    assert v == [L[1], L[2]]
    assert nodes == [[L[0], h(h(L[4], L[5]), h(L[6], L[7]))], [L[3]]]
    v, nodes, depth = _parse_batch(po.merkle_prove_batch("poseidon", leaves, [1, 6]))
    assert v == [L[1], L[6]]
    assert nodes == [[L[0], h(L[2], L[3])], [L[7], h(L[4], L[5])]]
    v, nodes, depth = _parse_batch(po.merkle_prove_batch("poseidon", leaves, list(range(8))))
    assert v == L and nodes == [[], [], [], []]


def test_merkle_verify_batch_reference_cases(po):
    # merkle.rs:495-518
    leaves = b"".join(LEAVES8)
    root = po.merkle_nodes("poseidon", leaves)[32:64]
    vb = lambda idx, proof: po.merkle_verify_batch("poseidon", root, idx, proof)
    p = po.merkle_prove_batch("poseidon", leaves, [1])
    assert vb([1], p) == 1 and vb([2], p) == 0
    p = po.merkle_prove_batch("poseidon", leaves, [1, 2])
    assert vb([1, 2], p) == 1 and vb([1], p) == 0 and vb([1, 3], p) == 0 and vb([1, 2, 3], p) == 0
    for idx in ([1, 6], [1, 3, 6], list(range(8))):
        assert vb(idx, po.merkle_prove_batch("poseidon", leaves, idx)) == 1


def test_chacha20_keystream(po):
    ciphers = pytest.importorskip("cryptography.hazmat.primitives.ciphers")
    seed = bytes(range(32))
    words = np.zeros(16 * 8, dtype=np.uint32)
    po.lib().or_chacha_words(seed, len(words), words.ctypes.data)
    algo = ciphers.algorithms.ChaCha20(seed, b"\0" * 16)
    ks = ciphers.Cipher(algo, mode=None).encryptor().update(b"\0" * (4 * len(words)))
    assert words.tobytes() == ks


def test_stdrng_uniform_cross_check_values(po):
    # SURVEY.md appendix C cross-check values (computed by an independent Python restatement of rand 0.7.3)
    seed = bytes(range(32))
    assert po.prng_vector(seed, 1)[0] == 97422350404758380040720986772812566605
    assert po.query_positions(seed, 2**13, 2**14, 4) == [3395, 2345, 6541, 4676]  # ext > domain: nothing rejected


def test_query_positions_rules(po):
    # stark/utils/mod.rs:25-44: no multiples of the extension factor, no duplicates
    seed = bytes([9] * 32)
    pos = po.query_positions(seed, 2**13, 32, 50)
    assert len(pos) == 50 and len(set(pos)) == 50
    assert all(p % 32 != 0 and p < 2**13 for p in pos)
    with pytest.raises(ValueError):
        po.query_positions(seed, 64, 32, 64)

"""CPU-side tests of the product's host logic (distaff_b200/csrc/host_fs.cu, exported as dg_host_*) against the oracle:
Fiat-Shamir draws, query positions, seed hashing, batch-proof planning and the periodic constant tables.  No GPU needed."""
import ctypes
import os
import struct

import numpy as np
import pytest

M = 2**128 - 45 * 2**40 + 1


@pytest.fixture(scope="module")
def L():
    from distaff_b200 import backend
    if not os.path.exists(backend.LIB_PATH):
        pytest.skip("libdistaff_gpu.so not built")
    return backend.lib()


def test_prng_vector_matches_oracle(L, po):
    from distaff_b200 import felt
    for seed in (bytes(range(32)), bytes([7] * 32), bytes(32)):
        out = np.zeros((600, 2), dtype=np.uint64)
        assert L.dg_host_prng_vector(seed, 600, out.ctypes.data) == 0
        assert felt.to_ints(out) == po.prng_vector(seed, 600)
    assert felt.to_ints(out[:1])[0] < M


def test_query_positions_match_oracle(L, po):
    for seed, domain, ext, nq in ((bytes(range(32)), 2**13, 32, 50), (bytes([3] * 32), 2**25, 32, 50), (bytes([9] * 32), 2**10, 16, 128)):
        out = np.zeros(nq, dtype=np.uint64)
        assert L.dg_host_query_positions(seed, domain, ext, nq, out.ctypes.data) == 0
        assert [int(x) for x in out] == po.query_positions(seed, domain, ext, nq)
    out = np.zeros(64, dtype=np.uint64)
    assert L.dg_host_query_positions(bytes(32), 64, 32, 64, out.ctypes.data) == -4      # cannot find 64 positions: utils/mod.rs:39-41


def test_short_blake3_matches_oracle(L, po):
    data = bytes((i * 7 + 1) % 256 for i in range(1024))
    for n in (0, 1, 32, 63, 64, 65, 320, 352, 1023, 1024):
        out = ctypes.create_string_buffer(32)
        assert L.dg_host_blake3(data[:n], n, out) == 0
        assert out.raw == po.hash("blake3", data[:n])
    assert L.dg_host_blake3(data + b"x", 1025, out) != 0


def _oracle_plan(po, leaves, indexes):
    """structure of the oracle's prove_batch: per slot the list of 32-byte nodes"""
    raw = po.merkle_prove_batch("blake3", leaves, indexes)
    off = 0

    def u64():
        nonlocal off
        v = struct.unpack_from("<Q", raw, off)[0]
        off += 8
        return v
    nv = u64()
    off += 32 * nv
    slots = []
    for _ in range(u64()):
        k = u64()
        slots.append([raw[off + 32 * i: off + 32 * i + 32] for i in range(k)])
        off += 32 * k
    return slots, raw[off]


def test_batch_proof_plan_matches_oracle_structure(L, po):
    rng = np.random.Generator(np.random.PCG64(5))
    for log_l in (3, 6, 10):
        n = 1 << log_l
        leaves = rng.integers(0, 256, size=n * 32, dtype=np.uint8).tobytes()
        nodes = po.merkle_nodes("blake3", leaves)
        for trial in range(6):
            k = int(rng.integers(1, min(n, 40) + 1))
            idx = [int(x) for x in rng.choice(n, size=k, replace=False)]
            want, depth = _oracle_plan(po, leaves, idx)
            arr = np.array(idx, dtype=np.uint64)
            out = np.zeros(4096, dtype=np.uint64)
            written = ctypes.c_size_t(0)
            assert L.dg_host_plan_batch(arr.ctypes.data, k, n, out.ctypes.data, 4096, ctypes.byref(written)) == 0
            flat = [int(x) for x in out[:written.value]]
            n_slots, d = flat[0], flat[1]
            assert d == depth and n_slots == len(want)
            pos = 2
            for s in range(n_slots):
                cnt = flat[pos]
                pos += 1
                got = []
                for _ in range(cnt):
                    is_leaf, index = flat[pos], flat[pos + 1]
                    pos += 2
                    src = leaves if is_leaf else nodes
                    got.append(src[32 * index: 32 * index + 32])
                assert got == want[s], (log_l, idx, s)
    arr = np.array([1, 1], dtype=np.uint64)
    assert L.dg_host_plan_batch(arr.ctypes.data, 2, 8, out.ctypes.data, 4096, ctypes.byref(written)) != 0     # repeating indexes (merkle.rs:302)


def test_periodic_tables_interpolate_the_round_constants(L):
    # rows 0, 8, 16, ... of the 8x-extended cycle are the original 16 round constants; masks are 0/1 there
    from distaff_b200 import felt
    out = np.zeros((128 * 23, 2), dtype=np.uint64)
    assert L.dg_host_periodic_tables(out.ctypes.data) == 0
    t = np.array(felt.to_ints(out), dtype=object).reshape(128, 23)
    masks = [[0] + [1] * 15, [1] * 15 + [0], [0] + [1] * 7 + [0] + [1] * 7]
    for m in range(3):
        assert [int(t[8 * k, 8 + m]) for k in range(16)] == masks[m]
    assert all(0 <= int(v) < M for v in t.reshape(-1))
    # the extension is a degree < 16 polynomial: check one column against Lagrange evaluation through Python big ints
    G = 23953097886125630542083529559205016746
    w128 = pow(G, 2**40 // 128, M)
    xs = [pow(w128, 8 * k, M) for k in range(16)]
    col = 3
    ys = [int(t[8 * k, col]) for k in range(16)]
    for s in (1, 5, 77, 127):
        x = pow(w128, s, M)
        acc = 0
        for i in range(16):
            num, den = 1, 1
            for j in range(16):
                if i != j:
                    num = num * (x - xs[j]) % M
                    den = den * (xs[i] - xs[j]) % M
            acc = (acc + ys[i] * num * pow(den, M - 2, M)) % M
        assert acc == int(t[s, col])


def test_sharded_tree_index_algebra(L):
    """ShardGeom (distaff_b200/csrc/shard.cu) against a brute-force model: build the global heap of item ids, the per-rank local heaps,
    the per-rank mid heaps (subtree over the rank's k-range of the level with n*G nodes) and the replicated top heap, then check that
    every global node / item is found at the location the product code computes."""
    import ctypes
    out = (ctypes.c_int64 * 3)()
    LOCAL, TOP, MID = 0, 1, 2
    for n, log_blk, log_g in ((4, 2, 1), (8, 3, 2), (8, 2, 3), (16, 0, 3), (8, 5, 0), (4, 1, 2), (64, 2, 3)):
        blk, G = 1 << log_blk, 1 << log_g
        total = n * blk * G
        # global tree: node = frozenset of the level-0 items below it
        level = [frozenset([i]) for i in range(total)]
        glob = {}
        size = total
        while size > 1:
            level = [level[2 * i] | level[2 * i + 1] for i in range(size // 2)]
            size //= 2
            for o, s_ in enumerate(level):
                glob[size + o] = s_
        # local trees: rank g holds items i = (k*G + g)*blk + j at local index k*blk + j
        local = []
        for g in range(G):
            its = [frozenset([(k * G + g) * blk + j]) for k in range(n) for j in range(blk)]
            heap = {}
            lv, sz = its, n * blk
            while sz > n:
                lv = [lv[2 * i] | lv[2 * i + 1] for i in range(sz // 2)]
                sz //= 2
                for o, s_ in enumerate(lv):
                    heap[sz + o] = s_
            local.append((its, heap, lv if blk > 1 else its))       # [2] = the n subtree roots, by k
        # mid heaps: rank g gets, from every rank g', the roots of k in [g n/G, (g+1) n/G); node k'*G + g' of its level with n nodes
        mid = []
        for g in range(G):
            chunk = n // G
            lv = [None] * n
            for g2 in range(G):
                for k2 in range(chunk):
                    lv[k2 * G + g2] = local[g2][2][g * chunk + k2]
            heap = {n + o: s_ for o, s_ in enumerate(lv)}
            sz = n
            while sz > 1:
                lv = [lv[2 * i] | lv[2 * i + 1] for i in range(sz // 2)]
                sz //= 2
                for o, s_ in enumerate(lv):
                    heap[sz + o] = s_
            mid.append(heap)
        top = {G + g: mid[g][1] for g in range(G)}
        lv, sz = [mid[g][1] for g in range(G)], G
        while sz > 1:
            lv = [lv[2 * i] | lv[2 * i + 1] for i in range(sz // 2)]
            sz //= 2
            for o, s_ in enumerate(lv):
                top[sz + o] = s_
        for i in range(total):
            assert L.dg_host_shard_locate(n, log_blk, log_g, 0, i, out) == 0
            owner, kind, idx = out[0], out[1], out[2]
            assert kind == LOCAL and local[owner][0][idx] == frozenset([i])
        for h in range(1, total):
            assert L.dg_host_shard_locate(n, log_blk, log_g, 1, h, out) == 0
            owner, kind, idx = out[0], out[1], out[2]
            if kind == TOP:
                assert owner == -1 and idx == h and top[idx] == glob[h]
            elif kind == MID:
                assert mid[owner][idx] == glob[h], (n, log_blk, log_g, h)
            else:
                assert local[owner][1][idx] == glob[h], (n, log_blk, log_g, h)


def test_crafted_multiplication_operands_are_what_they_claim():
    """tools/gen_mul_vectors.py (used by the GPU arithmetic tests): canonical operands whose product is a tiny / near-M residue"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gen_mul_vectors as g
    pairs = g.pairs()
    assert len(pairs) >= 400
    small = 0
    for a, b in pairs:
        assert 0 <= a < g.M and 0 <= b < g.M
        r = a * b % g.M
        if r < 2**94 or r > g.M - 2**97:
            small += 1
    assert small == len(pairs)
    assert g.pairs() == pairs                                   # deterministic


def test_generator_against_the_published_chacha20_vector(L):
    """External pin (no code of this repo involved in the expected values): the ChaCha20 keystream for the all-zero key and nonce is
    the published test vector 76b8e0ad a0f13d90 405d6ae5 5386bd28 bdd219b8 a08ded1a a836efcc 8b770dc7 ...  StdRng::from_seed([0; 32])
    (rand 0.7.3 = rand_chacha ChaCha20, 64-bit counter 0, stream 0) emits those bytes; next_u64 = two words, low first;
    Uniform(0..2^32) for usize = the top 32 bits of a u64; compute_query_positions drops multiples of the extension factor
    (stark/utils/mod.rs:25-44).  So the first draw 0x903df1a0 (a multiple of 32) is rejected and the positions start with 0x28bd8653."""
    import ctypes
    ks = bytes.fromhex("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
    words = [int.from_bytes(ks[4 * i:4 * i + 4], "little") for i in range(16)]
    u64s = [words[2 * i] | (words[2 * i + 1] << 32) for i in range(8)]
    expect = []
    for v in u64s:
        p = v >> 32                                   # Uniform(0..2^32): widening multiply by 2^32 = top 32 bits, zone = all
        if p % 32 != 0 and p not in expect:
            expect.append(p)
    assert expect[0] == 0x28bd8653 and (u64s[0] >> 32) == 0x903df1a0
    out = (ctypes.c_uint64 * 5)()
    assert L.dg_host_query_positions(bytes(32), 1 << 32, 32, 5, out) == 0
    assert list(out) == expect[:5]
    # and field::prng_vector's first element for the zero seed: floor(v * M / 2^128) with v = u64s[0] | u64s[1] << 64 (Standard u128: low word first)
    M = 2**128 - 45 * 2**40 + 1
    v = u64s[0] | (u64s[1] << 64)
    assert (v * M) % 2**128 <= M - 1                  # accepted on the first draw
    buf = (ctypes.c_uint8 * 16)()
    assert L.dg_host_prng_vector(bytes(32), 1, buf) == 0
    assert int.from_bytes(bytes(buf), "little") == (v * M) >> 128


def test_merkle_verification_plan_equals_verify_batch(L):
    """The hashing plan dg_verify runs on the device (verifier.cu: plan_verify_batch, exported as dg_host_merkle_verify_plan) executed on the
    CPU with blake3 must accept exactly what MerkleTree::verify_batch (merkle.rs:154-263, restated in the oracle) accepts: honest batch
    proofs for random index sets, and the same proofs with a corrupted value / node / index set / node-list shape."""
    import ctypes
    import random
    import struct
    import blake3 as b3
    from oracle import pyoracle as po
    rng = random.Random(11)

    def parse(proof):
        off = 0
        (nv,) = struct.unpack_from("<Q", proof, off); off += 8
        values = [proof[off + 32 * i: off + 32 * i + 32] for i in range(nv)]; off += 32 * nv
        (ns,) = struct.unpack_from("<Q", proof, off); off += 8
        nodes = []
        for _ in range(ns):
            (k,) = struct.unpack_from("<Q", proof, off); off += 8
            nodes.append([proof[off + 32 * i: off + 32 * i + 32] for i in range(k)]); off += 32 * k
        return values, nodes, proof[off]

    def encode(values, nodes, depth):
        out = struct.pack("<Q", len(values)) + b"".join(values) + struct.pack("<Q", len(nodes))
        for slot in nodes:
            out += struct.pack("<Q", len(slot)) + b"".join(slot)
        return out + bytes([depth])

    def run_plan(root, indexes, values, nodes, depth):
        counts = (ctypes.c_uint32 * max(1, len(nodes)))(*[len(s) for s in nodes])
        idx = (ctypes.c_uint64 * max(1, len(indexes)))(*indexes)
        ops = (ctypes.c_uint32 * 30000)()
        ls = (ctypes.c_uint32 * 64)()
        n_ops, n_levels, root_slot = ctypes.c_uint32(0), ctypes.c_uint32(0), ctypes.c_uint32(0)
        rc = L.dg_host_merkle_verify_plan(idx, len(indexes), depth, len(values), counts, len(nodes), ops, 30000, ctypes.byref(n_ops), ls, 64,
                                          ctypes.byref(n_levels), ctypes.byref(root_slot))
        if rc != 0:
            return False
        pool = list(values) + [d for s in nodes for d in s]
        pool += [None] * n_ops.value
        for lvl in range(n_levels.value):
            for o in range(ls[lvl], ls[lvl + 1]):
                l, r, out = ops[3 * o], ops[3 * o + 1], ops[3 * o + 2]
                pool[out] = b3.blake3(pool[l] + pool[r]).digest()
        return pool[root_slot.value] == root

    accepted = rejected = 0
    for log_l in (3, 5, 8, 11):
        n = 1 << log_l
        leaves = bytes(rng.getrandbits(8) for _ in range(32 * n))
        root = po.merkle_nodes("blake3", leaves)[32:64]
        for trial in range(12):
            k = rng.randrange(1, min(n, 40) + 1)
            indexes = rng.sample(range(n), k)
            proof = po.merkle_prove_batch("blake3", leaves, indexes)
            values, nodes, depth = parse(proof)
            cases = [(indexes, values, nodes, depth)]
            v2 = list(values); v2[rng.randrange(len(v2))] = bytes(32); cases.append((indexes, v2, nodes, depth))
            slots = [i for i, s in enumerate(nodes) if s]
            if slots:
                n2 = [list(s) for s in nodes]; s = rng.choice(slots); n2[s][rng.randrange(len(n2[s]))] = bytes(range(32)); cases.append((indexes, values, n2, depth))
                n3 = [list(s) for s in nodes]; n3[rng.choice(slots)].pop(); cases.append((indexes, values, n3, depth))
            other = [i ^ 1 for i in indexes]
            if sorted(other) != sorted(indexes) and len(set(other)) == len(other):
                cases.append((other, values, nodes, depth))
            cases.append((indexes[:-1], values, nodes, depth) if k > 1 else (indexes, values, nodes[:-1], depth))
            for idx, vals, nds, dep in cases:
                want = po.merkle_verify_batch("blake3", root, idx, encode(vals, nds, dep))
                got = run_plan(root, idx, vals, nds, dep)
                assert want in (0, 1) and got == (want == 1), (log_l, trial, idx[:5])
                accepted += got
                rejected += not got
    assert accepted >= 48 and rejected >= 100

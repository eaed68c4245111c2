// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or called by the product path.
//
// CPU restatement of the reference's crypto layer:
//   hash::{rescue, poseidon, gmimc}  /root/reference/src/crypto/hash.rs:119-203, helpers :222-254
//   hash::blake3                     /root/reference/src/crypto/hash.rs:205-209 -> crate blake3 0.3.5 (NOT in the
//                                    reference tree; restated here from the published BLAKE3 spec and pinned in
//                                    tests against the `blake3` pip package = official implementation)
//   MerkleTree / BatchMerkleProof    /root/reference/src/crypto/merkle.rs:25-124, 154-312
//   StdRng (rand 0.7.3 = ChaCha20) + Uniform sampling, used by field::prng* (/root/reference/src/math/field.rs:264-275)
//                                    and utils::compute_query_positions (/root/reference/src/stark/utils/mod.rs:25-44);
//                                    third-party, restated from the crates' documented algorithm: PARITY UNPINNED by the
//                                    reference (no test vectors in-tree); ChaCha20 keystream itself is checked against
//                                    the `cryptography` package in tests.
#ifndef ORACLE_CRYPTO_H
#define ORACLE_CRYPTO_H

#include "math.h"
#include "ref_constants.h"
#include <array>
#include <map>
#include <set>

namespace oracle {

typedef std::array<uint8_t, 32> Digest;
typedef void (*HashFn)(const uint8_t *in, size_t len, uint8_t *out32);

static inline u128 cst(const unsigned long long c[2]) { return mk128(c[0], c[1]); }

// ---------------------------------------------------------------------------------------------
// BLAKE3 (spec restatement; hash mode, 32-byte output)
// ---------------------------------------------------------------------------------------------
namespace b3 {
static const uint32_t IV[8] = { 0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19 };
static const int PERM[16] = { 2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8 };
enum { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline void g(uint32_t *s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx; s[d] = rotr(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my; s[d] = rotr(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 7);
}
static inline void compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter, uint32_t block_len,
                            uint32_t flags, uint32_t out[8]) {
    uint32_t s[16], m[16], t[16];
    for (int i = 0; i < 8; i++) s[i] = cv[i];
    for (int i = 0; i < 4; i++) s[8 + i] = IV[i];
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = block_len; s[15] = flags;
    for (int i = 0; i < 16; i++) m[i] = block[i];
    for (int r = 0; r < 7; r++) {
        g(s, 0, 4, 8, 12, m[0], m[1]);   g(s, 1, 5, 9, 13, m[2], m[3]);
        g(s, 2, 6, 10, 14, m[4], m[5]);  g(s, 3, 7, 11, 15, m[6], m[7]);
        g(s, 0, 5, 10, 15, m[8], m[9]);  g(s, 1, 6, 11, 12, m[10], m[11]);
        g(s, 2, 7, 8, 13, m[12], m[13]); g(s, 3, 4, 9, 14, m[14], m[15]);
        for (int i = 0; i < 16; i++) t[i] = m[PERM[i]];
        for (int i = 0; i < 16; i++) m[i] = t[i];
    }
    for (int i = 0; i < 8; i++) out[i] = s[i] ^ s[i + 8];
}
// chaining value of one chunk (<= 1024 bytes); `root` = this chunk is the whole message
static inline void chunk_cv(const uint8_t *data, size_t len, uint64_t chunk_counter, bool root, uint32_t out[8]) {
    uint32_t cv[8];
    for (int i = 0; i < 8; i++) cv[i] = IV[i];
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    for (size_t b = 0; b < nblocks; b++) {
        uint8_t buf[64];
        memset(buf, 0, 64);
        size_t take = std::min((size_t)64, len - b * 64);
        if (len == 0) take = 0;
        memcpy(buf, data + b * 64, take);
        uint32_t words[16];
        for (int i = 0; i < 16; i++)
            words[i] = (uint32_t)buf[4 * i] | ((uint32_t)buf[4 * i + 1] << 8) | ((uint32_t)buf[4 * i + 2] << 16) | ((uint32_t)buf[4 * i + 3] << 24);
        uint32_t flags = 0;
        if (b == 0) flags |= CHUNK_START;
        if (b == nblocks - 1) { flags |= CHUNK_END; if (root) flags |= ROOT; }
        uint32_t o[8];
        compress(cv, words, chunk_counter, (uint32_t)take, flags, o);
        for (int i = 0; i < 8; i++) cv[i] = o[i];
    }
    for (int i = 0; i < 8; i++) out[i] = cv[i];
}
static inline void parent_cv(const uint32_t l[8], const uint32_t r[8], bool root, uint32_t out[8]) {
    uint32_t block[16];
    for (int i = 0; i < 8; i++) { block[i] = l[i]; block[8 + i] = r[i]; }
    compress(IV, block, 0, 64, PARENT | (root ? ROOT : 0), out);
}
// recursive tree: left subtree = largest power-of-two number of chunks strictly less than the total
static void subtree(const uint8_t *data, size_t len, uint64_t chunk0, bool root, uint32_t out[8]) {
    if (len <= 1024) { chunk_cv(data, len, chunk0, root, out); return; }
    size_t chunks = (len + 1023) / 1024;
    size_t left = 1;
    while (left * 2 < chunks) left *= 2;
    uint32_t l[8], r[8];
    subtree(data, left * 1024, chunk0, false, l);
    subtree(data + left * 1024, len - left * 1024, chunk0 + left, false, r);
    parent_cv(l, r, root, out);
}
} // namespace b3

static inline void blake3(const uint8_t *in, size_t len, uint8_t *out32) {
    uint32_t o[8];
    b3::subtree(in, len, 0, true, o);
    for (int i = 0; i < 8; i++) { out32[4 * i] = o[i]; out32[4 * i + 1] = o[i] >> 8; out32[4 * i + 2] = o[i] >> 16; out32[4 * i + 3] = o[i] >> 24; }
}

// ---------------------------------------------------------------------------------------------
// Algebraic hashes over the field (hash.rs:119-203); <= 64 input bytes, 32 output bytes
// ---------------------------------------------------------------------------------------------
namespace alg {
static inline void load_state(const uint8_t *in, size_t len, u128 st[6]) {
    uint8_t buf[96];
    memset(buf, 0, sizeof buf);
    if (len > 64) throw std::runtime_error("expected 64 or fewer input bytes");
    memcpy(buf, in, len);
    memcpy(st, buf, 96);
}
static inline void add_constants(u128 st[6], size_t off) { for (int i = 0; i < 6; i++) st[i] = field::add(st[i], cst(REF_HASH_ARK[off + i])); }
static inline void apply_sbox(u128 st[6]) { for (int i = 0; i < 6; i++) st[i] = field::exp(st[i], 3); }
static inline void apply_inv_sbox(u128 st[6]) { u128 ia = cst(REF_INV_ALPHA); for (int i = 0; i < 6; i++) st[i] = field::exp(st[i], ia); }
static inline void apply_mds(u128 st[6]) {
    u128 r[6];
    for (int i = 0; i < 6; i++) {
        r[i] = 0;
        for (int j = 0; j < 6; j++) r[i] = field::add(r[i], field::mul(cst(REF_HASH_MDS[i * 6 + j]), st[j]));
    }
    memcpy(st, r, sizeof r);
}
} // namespace alg

static inline void rescue(const uint8_t *in, size_t len, uint8_t *out32) {  // hash.rs:151-177
    u128 st[6];
    alg::load_state(in, len, st);
    alg::add_constants(st, 0);
    for (int i = 0; i < 10; i++) {
        alg::apply_inv_sbox(st); alg::apply_mds(st); alg::add_constants(st, (i * 2 + 1) * 6);
        alg::apply_sbox(st);     alg::apply_mds(st); alg::add_constants(st, (i * 2 + 2) * 6);
    }
    memcpy(out32, st, 32);
}
static inline void poseidon(const uint8_t *in, size_t len, uint8_t *out32) {  // hash.rs:119-147
    u128 st[6];
    alg::load_state(in, len, st);
    for (int i = 0; i < 91; i++) {
        alg::add_constants(st, i * 6);
        if (i < 4 || i >= 87) alg::apply_sbox(st);
        else st[5] = field::exp(st[5], 3);
        alg::apply_mds(st);
    }
    memcpy(out32, st, 32);
}
static inline void gmimc(const uint8_t *in, size_t len, uint8_t *out32) {  // hash.rs:181-201
    u128 st[6];
    alg::load_state(in, len, st);
    for (int i = 0; i < 166; i++) {
        u128 s0 = st[0];
        u128 mask = field::exp(field::add(s0, cst(REF_HASH_ARK[i])), 3);
        for (int j = 1; j < 6; j++) st[j - 1] = field::add(mask, st[j]);
        st[5] = s0;
    }
    memcpy(out32, st, 32);
}

// ---------------------------------------------------------------------------------------------
// Merkle tree (merkle.rs)
// ---------------------------------------------------------------------------------------------
struct BatchMerkleProof {
    std::vector<Digest> values;
    std::vector<std::vector<Digest>> nodes;
    uint8_t depth;
};

// merkle.rs:269-294
static inline std::vector<Digest> build_merkle_nodes(const std::vector<Digest> &leaves, HashFn hash) {
    size_t n = leaves.size() / 2;
    std::vector<Digest> nodes(2 * n);
    nodes[0].fill(0);
    const int T = fft::host_threads();
    #pragma omp parallel for num_threads(T) if (T > 1 && n >= 4096)
    for (size_t i = 0; i < n; i++) hash(leaves[2 * i].data(), 64, nodes[n + i].data());  // leaves are contiguous 32-byte arrays
    if (T <= 1) {
        for (size_t i = n - 1; i >= 1; i--) hash(nodes[2 * i].data(), 64, nodes[i].data());
    } else {                                                 // same nodes, level by level (a level only reads the one below it)
        for (size_t m = n / 2; m >= 1; m >>= 1) {
            #pragma omp parallel for num_threads(T) if (m >= 4096)
            for (size_t i = m; i < 2 * m; i++) hash(nodes[2 * i].data(), 64, nodes[i].data());
        }
    }
    return nodes;
}

struct MerkleTree {
    std::vector<Digest> nodes, values;
    MerkleTree() {}
    MerkleTree(std::vector<Digest> leaves, HashFn hash) {  // merkle.rs:25-34
        if (leaves.size() < 2 || (leaves.size() & (leaves.size() - 1))) throw std::runtime_error("number of leaves must be a power of 2 and >= 2");
        nodes = build_merkle_nodes(leaves, hash);
        values = std::move(leaves);
    }
    const Digest &root() const { return nodes[1]; }

    // merkle.rs:46-61
    std::vector<Digest> prove(size_t index) const {
        std::vector<Digest> proof;
        proof.push_back(values[index]);
        proof.push_back(values[index ^ 1]);
        index = (index + nodes.size()) >> 1;
        while (index > 1) { proof.push_back(nodes[index ^ 1]); index >>= 1; }
        return proof;
    }

    // merkle.rs:64-124  (including the `nodes[i]` slot-reuse behaviour of the level loop)
    BatchMerkleProof prove_batch(const std::vector<size_t> &indexes_in) const {
        size_t n = values.size();
        std::map<size_t, size_t> index_map;
        for (size_t i = 0; i < indexes_in.size(); i++) {
            if (indexes_in[i] > n) throw std::runtime_error("invalid index");
            index_map[indexes_in[i]] = i;
        }
        if (index_map.size() != indexes_in.size()) throw std::runtime_error("repeating indexes detected");
        std::set<size_t> norm;
        for (size_t idx : indexes_in) norm.insert(idx - (idx & 1));
        std::vector<size_t> indexes(norm.begin(), norm.end());

        BatchMerkleProof p;
        p.values.assign(index_map.size(), Digest());
        std::vector<size_t> next;
        for (size_t index : indexes) {
            const Digest &v1 = values[index], &v2 = values[index + 1];
            auto i1 = index_map.find(index), i2 = index_map.find(index + 1);
            if (i1 != index_map.end()) {
                if (i2 != index_map.end()) { p.values[i1->second] = v1; p.values[i2->second] = v2; p.nodes.push_back({}); }
                else { p.values[i1->second] = v1; p.nodes.push_back({ v2 }); }
            } else { p.values[i2->second] = v2; p.nodes.push_back({ v1 }); }
            next.push_back((index + n) >> 1);
        }
        uint8_t depth = (uint8_t)__builtin_ctzll(n);
        for (uint8_t d = 1; d < depth; d++) {
            std::vector<size_t> cur = next;
            next.clear();
            size_t i = 0;
            while (i < cur.size()) {
                size_t sibling = cur[i] ^ 1;
                if (i + 1 < cur.size() && cur[i + 1] == sibling) i += 1;
                else p.nodes[i].push_back(nodes[sibling]);
                next.push_back(sibling >> 1);
                i += 1;
            }
        }
        p.depth = depth;
        return p;
    }

    // merkle.rs:127-151
    static bool verify(const Digest &root, size_t index, const std::vector<Digest> &proof, HashFn hash) {
        uint8_t buf[64]; Digest v;
        size_t r = index & 1;
        memcpy(buf, proof[r].data(), 32); memcpy(buf + 32, proof[1 - r].data(), 32);
        hash(buf, 64, v.data());
        index = (index + ((size_t)1 << (proof.size() - 1))) >> 1;
        for (size_t i = 2; i < proof.size(); i++) {
            if ((index & 1) == 0) { memcpy(buf, v.data(), 32); memcpy(buf + 32, proof[i].data(), 32); }
            else { memcpy(buf, proof[i].data(), 32); memcpy(buf + 32, v.data(), 32); }
            hash(buf, 64, v.data());
            index >>= 1;
        }
        return v == root;
    }

    // merkle.rs:154-263
    static bool verify_batch(const Digest &root, const std::vector<size_t> &indexes_in, const BatchMerkleProof &proof, HashFn hash) {
        uint8_t buf[64];
        std::map<size_t, Digest> v;
        size_t offset = (size_t)1 << proof.depth;
        std::map<size_t, size_t> index_map;
        for (size_t i = 0; i < indexes_in.size(); i++) {
            if (indexes_in[i] > offset - 1) return false;
            index_map[indexes_in[i]] = i;
        }
        if (index_map.size() != indexes_in.size()) return false;
        std::set<size_t> norm;
        for (size_t idx : indexes_in) norm.insert(idx - (idx & 1));
        std::vector<size_t> indexes(norm.begin(), norm.end());
        if (indexes.size() != proof.nodes.size()) return false;

        std::vector<size_t> next, ptrs;
        for (size_t i = 0; i < indexes.size(); i++) {
            size_t index = indexes[i];
            auto i1 = index_map.find(index), i2 = index_map.find(index + 1);
            if (i1 != index_map.end()) {
                if (proof.values.size() <= i1->second) return false;
                memcpy(buf, proof.values[i1->second].data(), 32);
                if (i2 != index_map.end()) {
                    if (proof.values.size() <= i2->second) return false;
                    memcpy(buf + 32, proof.values[i2->second].data(), 32);
                    ptrs.push_back(0);
                } else {
                    if (proof.nodes[i].size() < 1) return false;
                    memcpy(buf + 32, proof.nodes[i][0].data(), 32);
                    ptrs.push_back(1);
                }
            } else {
                if (proof.nodes[i].size() < 1) return false;
                memcpy(buf, proof.nodes[i][0].data(), 32);
                if (i2 == index_map.end()) return false;
                if (proof.values.size() <= i2->second) return false;
                memcpy(buf + 32, proof.values[i2->second].data(), 32);
                ptrs.push_back(1);
            }
            Digest parent;
            hash(buf, 64, parent.data());
            size_t parent_index = (offset + index) >> 1;
            v[parent_index] = parent;
            next.push_back(parent_index);
        }
        for (uint8_t d = 1; d < proof.depth; d++) {
            std::vector<size_t> cur = next;
            next.clear();
            size_t i = 0;
            while (i < cur.size()) {
                size_t node_index = cur[i], sibling_index = node_index ^ 1;
                Digest sibling;
                if (i + 1 < cur.size() && cur[i + 1] == sibling_index) {
                    auto it = v.find(sibling_index);
                    if (it == v.end()) return false;
                    sibling = it->second;
                    i += 1;
                } else {
                    size_t pointer = ptrs[i];
                    if (proof.nodes[i].size() <= pointer) return false;
                    sibling = proof.nodes[i][pointer];
                    ptrs[i] += 1;
                }
                auto itn = v.find(node_index);
                if (itn == v.end()) return false;
                if (node_index & 1) { memcpy(buf, sibling.data(), 32); memcpy(buf + 32, itn->second.data(), 32); }
                else { memcpy(buf, itn->second.data(), 32); memcpy(buf + 32, sibling.data(), 32); }
                Digest parent;
                hash(buf, 64, parent.data());
                size_t parent_index = node_index >> 1;
                v[parent_index] = parent;
                next.push_back(parent_index);
                i += 1;
            }
        }
        auto it = v.find(1);
        return it != v.end() && it->second == root;
    }
};

// ---------------------------------------------------------------------------------------------
// rand 0.7.3 StdRng = ChaCha20Rng (rand_chacha 0.2): key = seed, 64-bit block counter from 0,
// stream id 0, 20 rounds; next_u64 = two consecutive keystream words (low word first).
// ---------------------------------------------------------------------------------------------
struct ChaChaRng {
    uint32_t key[8];
    uint64_t counter;
    uint32_t buf[16];
    int pos;
    explicit ChaChaRng(const uint8_t seed[32]) : counter(0), pos(16) {
        for (int i = 0; i < 8; i++)
            key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
    }
    static inline uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
    static inline void qr(uint32_t *s, int a, int b, int c, int d) {
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 16);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 12);
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 8);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 7);
    }
    void refill() {
        uint32_t in[16] = { 0x61707865, 0x3320646e, 0x79622d32, 0x6b206574 };
        for (int i = 0; i < 8; i++) in[4 + i] = key[i];
        in[12] = (uint32_t)counter; in[13] = (uint32_t)(counter >> 32); in[14] = 0; in[15] = 0;
        uint32_t s[16];
        memcpy(s, in, sizeof s);
        for (int r = 0; r < 10; r++) {
            qr(s, 0, 4, 8, 12); qr(s, 1, 5, 9, 13); qr(s, 2, 6, 10, 14); qr(s, 3, 7, 11, 15);
            qr(s, 0, 5, 10, 15); qr(s, 1, 6, 11, 12); qr(s, 2, 7, 8, 13); qr(s, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) buf[i] = s[i] + in[i];
        counter++;
        pos = 0;
    }
    uint32_t next_u32() { if (pos >= 16) refill(); return buf[pos++]; }
    uint64_t next_u64() { uint64_t lo = next_u32(); uint64_t hi = next_u32(); return lo | (hi << 32); }
    // Standard distribution for u128: low 64 bits first, then high
    u128 next_u128() { u64 lo = next_u64(); u64 hi = next_u64(); return mk128(lo, hi); }
};

// Uniform::from(0..M) over u128  (UniformInt::sample, widening multiply + zone rejection)
static inline u128 sample_field(ChaChaRng &rng) {
    // range = M, ints_to_reject = (2^128 - M) % M = 2^128 - M ; zone = MAX - ints_to_reject = M - 1
    const u128 zone = M - 1;
    for (;;) {
        u128 v = rng.next_u128();
        // 256-bit product v * M -> (hi, lo)
        u64 v0 = (u64)v, v1 = (u64)(v >> 64), m0 = (u64)M, m1 = (u64)(M >> 64);
        u128 p00 = (u128)v0 * m0, p01 = (u128)v0 * m1, p10 = (u128)v1 * m0, p11 = (u128)v1 * m1;
        u128 mid = (p00 >> 64) + (u64)p01 + (u64)p10;
        u128 lo = ((u128)(u64)mid << 64) | (u64)p00;
        u128 hi = p11 + (p01 >> 64) + (p10 >> 64) + (mid >> 64);
        if (lo <= zone) return hi;
    }
}
// field.rs:264-275
static inline u128 prng(const uint8_t seed[32]) { ChaChaRng g(seed); return sample_field(g); }
static inline std::vector<u128> prng_vector(const uint8_t seed[32], size_t length) {
    ChaChaRng g(seed);
    std::vector<u128> r(length);
    for (size_t i = 0; i < length; i++) r[i] = sample_field(g);
    return r;
}
// Uniform::from(0..range) over usize (64-bit): widening multiply of a u64 draw, zone rejection
static inline uint64_t sample_usize(ChaChaRng &rng, uint64_t range) {
    uint64_t ints_to_reject = (0 - range) % range;  // (2^64 - range) % range
    uint64_t zone = ~(uint64_t)0 - ints_to_reject;
    for (;;) {
        uint64_t v = rng.next_u64();
        u128 p = (u128)v * range;
        if ((uint64_t)p <= zone) return (uint64_t)(p >> 64);
    }
}

} // namespace oracle
#endif

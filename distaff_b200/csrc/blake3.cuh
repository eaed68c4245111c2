// BLAKE3 compression function for device code (32-byte digests, hash mode).
// The reference commits with `hash::blake3` = crate blake3 0.3.5 (/root/reference/src/crypto/hash.rs:205-209,
// /root/reference/src/stark/options.rs:89); the crate is not vendored, so this follows the published BLAKE3 spec
// (IV, message permutation, 7 rounds, flags).  Call sites on the prove path only ever hash
//   * 64-byte inputs (Merkle nodes, FRI rows, proof-of-work): one compression, flags CHUNK_START|CHUNK_END|ROOT
//   * trace rows of 16*w bytes, w < 128 => at most two 1024-byte chunks + one parent compression.
#pragma once
#include <cstdint>

namespace dg {
namespace b3 {

enum { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

#define DG_B3_IV0 0x6A09E667u
#define DG_B3_IV1 0xBB67AE85u
#define DG_B3_IV2 0x3C6EF372u
#define DG_B3_IV3 0xA54FF53Au
#define DG_B3_IV4 0x510E527Fu
#define DG_B3_IV5 0x9B05688Cu
#define DG_B3_IV6 0x1F83D9ABu
#define DG_B3_IV7 0x5BE0CD19u

__host__ __device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

#define DG_B3_G(a, b, c, d, mx, my)            \
    a = a + b + (mx); d = rotr32(d ^ a, 16);   \
    c = c + d;        b = rotr32(b ^ c, 12);   \
    a = a + b + (my); d = rotr32(d ^ a, 8);    \
    c = c + d;        b = rotr32(b ^ c, 7);

// one round with the message schedule given as compile-time word indices
#define DG_B3_ROUND(m, i0, i1, i2, i3, i4, i5, i6, i7, i8, i9, i10, i11, i12, i13, i14, i15) \
    DG_B3_G(s0, s4, s8, s12, m[i0], m[i1])   DG_B3_G(s1, s5, s9, s13, m[i2], m[i3])          \
    DG_B3_G(s2, s6, s10, s14, m[i4], m[i5])  DG_B3_G(s3, s7, s11, s15, m[i6], m[i7])         \
    DG_B3_G(s0, s5, s10, s15, m[i8], m[i9])  DG_B3_G(s1, s6, s11, s12, m[i10], m[i11])       \
    DG_B3_G(s2, s7, s8, s13, m[i12], m[i13]) DG_B3_G(s3, s4, s9, s14, m[i14], m[i15])

// cv (in/out, 8 words) <- compress(cv, m[16], counter, block_len, flags), truncated to the chaining value
__host__ __device__ __forceinline__ void compress(uint32_t cv[8], const uint32_t m[16], uint64_t counter, uint32_t block_len, uint32_t flags) {
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = DG_B3_IV0, s9 = DG_B3_IV1, s10 = DG_B3_IV2, s11 = DG_B3_IV3;
    uint32_t s12 = (uint32_t)counter, s13 = (uint32_t)(counter >> 32), s14 = block_len, s15 = flags;
    // message schedule: round r uses m[perm^r(i)]; the seven permutations are spelled out so that every index is a constant
    DG_B3_ROUND(m, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    DG_B3_ROUND(m, 2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    DG_B3_ROUND(m, 3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    DG_B3_ROUND(m, 10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    DG_B3_ROUND(m, 12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    DG_B3_ROUND(m, 9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    DG_B3_ROUND(m, 11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
    cv[0] = s0 ^ s8;  cv[1] = s1 ^ s9;  cv[2] = s2 ^ s10; cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12; cv[5] = s5 ^ s13; cv[6] = s6 ^ s14; cv[7] = s7 ^ s15;
}

// Variant for the ALU-bound row-hashing kernel: the first addition of every "a + b + m" is issued as a multiply-add a*one + b on the
// FMA pipe.  `one` must be a run-time 1 (a kernel parameter): ptxas folds a literal multiplier back into an IADD3.
__device__ __forceinline__ uint32_t madd1(uint32_t a, uint32_t one, uint32_t b) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}
#define DG_B3_GF(a, b, c, d, mx, my)                          \
    a = madd1(a, one, b) + (mx); d = rotr32(d ^ a, 16);       \
    c = c + d;                   b = rotr32(b ^ c, 12);       \
    a = madd1(a, one, b) + (my); d = rotr32(d ^ a, 8);        \
    c = c + d;                   b = rotr32(b ^ c, 7);
#define DG_B3_ROUNDF(m, i0, i1, i2, i3, i4, i5, i6, i7, i8, i9, i10, i11, i12, i13, i14, i15) \
    DG_B3_GF(s0, s4, s8, s12, m[i0], m[i1])   DG_B3_GF(s1, s5, s9, s13, m[i2], m[i3])          \
    DG_B3_GF(s2, s6, s10, s14, m[i4], m[i5])  DG_B3_GF(s3, s7, s11, s15, m[i6], m[i7])         \
    DG_B3_GF(s0, s5, s10, s15, m[i8], m[i9])  DG_B3_GF(s1, s6, s11, s12, m[i10], m[i11])       \
    DG_B3_GF(s2, s7, s8, s13, m[i12], m[i13]) DG_B3_GF(s3, s4, s9, s14, m[i14], m[i15])
__device__ __forceinline__ void compress_fma(uint32_t cv[8], const uint32_t m[16], uint64_t counter, uint32_t block_len, uint32_t flags, uint32_t one) {
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = DG_B3_IV0, s9 = DG_B3_IV1, s10 = DG_B3_IV2, s11 = DG_B3_IV3;
    uint32_t s12 = (uint32_t)counter, s13 = (uint32_t)(counter >> 32), s14 = block_len, s15 = flags;
    DG_B3_ROUNDF(m, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    DG_B3_ROUNDF(m, 2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    DG_B3_ROUNDF(m, 3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    DG_B3_ROUNDF(m, 10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    DG_B3_ROUNDF(m, 12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    DG_B3_ROUNDF(m, 9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    DG_B3_ROUNDF(m, 11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
    cv[0] = s0 ^ s8;  cv[1] = s1 ^ s9;  cv[2] = s2 ^ s10; cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12; cv[5] = s5 ^ s13; cv[6] = s6 ^ s14; cv[7] = s7 ^ s15;
}

__host__ __device__ __forceinline__ void iv(uint32_t cv[8]) {
    cv[0] = DG_B3_IV0; cv[1] = DG_B3_IV1; cv[2] = DG_B3_IV2; cv[3] = DG_B3_IV3;
    cv[4] = DG_B3_IV4; cv[5] = DG_B3_IV5; cv[6] = DG_B3_IV6; cv[7] = DG_B3_IV7;
}

// digest of exactly 64 bytes given as 16 little-endian words
__host__ __device__ __forceinline__ void hash64(const uint32_t m[16], uint32_t out[8]) {
    iv(out);
    compress(out, m, 0, 64, CHUNK_START | CHUNK_END | ROOT);
}

}  // namespace b3
}  // namespace dg

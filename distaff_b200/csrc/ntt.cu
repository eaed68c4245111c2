// Batched radix-2 NTT / inverse NTT and coset low-degree extension over F_M for sm_100a.
//
// Replaces, for the prove hot path, the reference's recursive in-place FFT + bit-reversal permutation
//   /root/reference/src/math/fft.rs:16-79, /root/reference/src/math/polynom.rs:34-41,93-103
// and the zero-padded extension loop of /root/reference/src/stark/trace/trace_table.rs:143-169.
// Contract (pinned by fft.rs:117-157): natural-order input -> natural-order DFT  X[k] = sum_j x[j] w^(jk).
//
// Design: a transform of size n = 2^log_n is split into at most three passes of <= 1024-point sub-transforms that run
// entirely in shared memory (decimation in frequency, twiddles for the in-block stages staged in shared memory).  Every
// pass streams HBM once with 16-byte vector accesses; a block owns a tile of T neighbouring "lanes" (independent
// sub-transforms whose elements are adjacent in memory) so that global reads and writes are T*16-byte contiguous
// segments.  Inter-pass twiddles w^(lane*k) come from a two-level power table (2 loads + 1 multiply).
//
// Low-degree extension does not zero-pad: evaluating P (n coefficients) on the LDE domain of size N = b*n is done as b
// independent size-n transforms of the coset-scaled coefficients p[m] * w_N^(c*m); the result is stored coset-major
// ([c][k] <-> LDE index b*k + c), which is the layout every later kernel (leaf hashing, constraint evaluation, FRI)
// consumes with unit-stride reads.  This removes log2(b) of the log2(N) butterfly levels and all work on zeros.
// The fully unrolled rounds are far larger than the 32 KB instruction cache; with the field multiplication out of line
// (one shared 80-instruction body) the pass kernels shrink from 17.6k to 10k instructions and run ~4% faster (B200, 2^20 x 32 LDE).
#define DG_MUL_CALL 1
#include "common.cuh"

namespace dg {

struct PassGeom {
    int log_t;                          // lanes per block (power of two)
    unsigned num_tiles;                 // blockIdx.x = outer * num_tiles + tile
    long long in_outer, in_lane, in_point;
    long long out_outer, out_lane, out_point;
    long long in_batch_y, out_batch_y, in_batch_z, out_batch_z;
    int lane_major;                     // shared-memory layout: 0 = [point][lane], 1 = [lane][point] (padded)
    int tw_on;                          // multiply output k of lane (tile*T+lane) by tw^((tile*T+lane)*k)
    TwiddleRef tw;
    int has_scale;
    fe scale;
    int coset_on;                       // input transform of the LDE: sum_f src[j + f*fold_stride] * cw^(c*(j + f*fold_stride))
    int fold;
    long long fold_stride;
    TwiddleRef cw;
    int coset_fast;                     // fold == 1: input factor from a single-level table, lane factor merged into the output twiddle
    const fe *cw_point;                 // cw_point[e] = (w_N^in_point)^e, e < cw_point_mask + 1
    unsigned cw_point_mask;
    const fe *tw_full;                  // coset_fast: tw_full[coset][k * out_point + lane] = cw^(lane * (k * blowup + coset)), or null
    long long tw_full_stride;           // elements per coset (= transform size n)
    int log_blowup;
    unsigned coset0;                    // first coset handled by this launch (blockIdx.y = coset - coset0)
    const fe *roots;                    // per-stage twiddle tables of the L-point transform: W_st[j] = w_L^(j << st), back to back
};

__device__ __forceinline__ fe tw_lookup(const TwiddleRef &t, unsigned long long e) {
    unsigned ee = (unsigned)e & t.mask;
    fe a = t.lo[ee & ((1u << t.lo_bits) - 1u)];
    fe b = t.hi[ee >> t.lo_bits];
    return fe_mul(a, b);
}

// ---- pass kernel -------------------------------------------------------------------------------------------------------
// One block transforms a tile of T lanes x L points.  The log2(L) decimation-in-frequency stages are grouped into rounds of
// up to 4 stages that run entirely in registers on 2^rho elements per thread ("unit"); shared memory is touched only
// between rounds.  The first round reads its operands straight from global memory and the last one writes straight back,
// so a 1024-point sub-transform costs 2 shared-memory round trips and 2 barriers instead of 10.  Stage twiddles come from
// per-stage compact tables W_st[j] = w_L^(j << st) (unit-stride, conflict-free) staged in shared memory.
template <int LOG_L> struct Rounds {
    static constexpr int R1 = LOG_L <= 4 ? LOG_L : 4;
    static constexpr int REM = LOG_L - R1;
    static constexpr int R2 = REM == 0 ? 0 : (REM <= 4 ? REM : (REM + 1) / 2);
    static constexpr int R3 = REM - R2;
};

template <int LOG_L, int S0, int RHO>
__device__ __forceinline__ void dif_regs(fe *x, const fe *s_tw, int g_lo) {
    constexpr int L = 1 << LOG_L, R = 1 << RHO;
    constexpr int LOG_SP = LOG_L - S0 - RHO;
#pragma unroll
    for (int u = 0; u < RHO; u++) {
        const int hr = R >> (u + 1);
        const int st = S0 + u;
        const fe *W = s_tw + (L - (L >> st));
#pragma unroll
        for (int i = 0; i < R; i++) {
            if ((i & hr) == 0) {
                fe a = x[i], b = x[i + hr];
                x[i] = fe_add(a, b);
                fe d = fe_sub(a, b);
                // in the last round (LOG_SP == 0, g_lo == 0) the twiddle index is a compile-time constant: index 0 is w^0 = 1
                if (st != LOG_L - 1 && !(LOG_SP == 0 && (i & (hr - 1)) == 0)) d = fe_mul(d, W[g_lo + ((i & (hr - 1)) << LOG_SP)]);
                x[i + hr] = d;
            }
        }
    }
}

template <int LOG_L, bool LANE_MAJOR>
__device__ __forceinline__ int sidx(int pos, int t, int T) {
    constexpr int L = 1 << LOG_L;
    constexpr int LS = L + (L >> 3) + 1;                 // padded lane stride, one pad element per 8 points
    return LANE_MAJOR ? (t * LS + pos + (pos >> 3)) : (pos * T + t);
}

template <int LOG_L, int S0, int RHO, bool FIRST, bool LAST, bool LANE_MAJOR>
__device__ __forceinline__ void ntt_round(const fe *__restrict__ src, fe *__restrict__ dst, fe *s_data, const fe *s_tw, const PassGeom &g,
                                          unsigned tile, long long in_base) {
    constexpr int L = 1 << LOG_L, R = 1 << RHO;
    constexpr int LOG_B = LOG_L - S0, LOG_SP = LOG_B - RHO;
    constexpr int N_GLO = 1 << LOG_SP, N_GHI = 1 << S0;
    const int T = 1 << g.log_t;
    const int units = (L >> RHO) * T;
    for (int u = threadIdx.x; u < units; u += blockDim.x) {
        int t, g_lo, g_hi;
        if (!LANE_MAJOR || LAST) {             // lanes fastest: global accesses of neighbouring threads are contiguous across lanes
            t = u & (T - 1);
            const int rest = u >> g.log_t;
            g_lo = rest & (N_GLO - 1);
            g_hi = rest >> LOG_SP;
        } else {                               // points fastest: contiguous rows of the last pass / conflict-free shared accesses
            g_lo = u & (N_GLO - 1);
            const int rest = u >> LOG_SP;
            g_hi = rest & (N_GHI - 1);
            t = rest >> S0;
        }
        const int gbase = (g_hi << LOG_B) + g_lo;
        fe x[R];
#pragma unroll
        for (int m = 0; m < R; m++) {
            const int pos = gbase + (m << LOG_SP);
            if (FIRST) {
                const long long j = in_base + (long long)t * g.in_lane + (long long)pos * g.in_point;
                if (g.coset_fast) {
                    // p[j] * w_N^(c*pos*in_point); the lane part w_N^(c*lane) rides on the output twiddle
                    x[m] = fe_mul(src[j], g.cw_point[((g.coset0 + (unsigned)blockIdx.y) * (unsigned)pos) & g.cw_point_mask]);
                } else if (g.coset_on) {
                    // sum_f src[j + f n] w_N^(c (j + f n)) = w_N^(c j) * Horner_f(src[j + f n]; u),  u = w_N^(c n) (constant per coset)
                    const unsigned long long c = g.coset0 + blockIdx.y;
                    const fe u = tw_lookup(g.cw, c * (unsigned long long)g.fold_stride);
                    fe v = src[j + (long long)(g.fold - 1) * g.fold_stride];
                    for (int f = g.fold - 2; f >= 0; f--) v = fe_add(fe_mul(v, u), src[j + (long long)f * g.fold_stride]);
                    x[m] = fe_mul(v, tw_lookup(g.cw, c * (unsigned long long)j));
                } else {
                    x[m] = src[j];
                }
            } else {
                x[m] = s_data[sidx<LOG_L, LANE_MAJOR>(pos, t, T)];
            }
        }
        dif_regs<LOG_L, S0, RHO>(x, s_tw, g_lo);
#pragma unroll
        for (int m = 0; m < R; m++) {
            const int pos = gbase + (m << LOG_SP);
            if (LAST) {                        // position q holds X[bitrev(q)]
                const unsigned k = __brev((unsigned)pos) >> (32 - LOG_L);
                fe v = x[m];
                if (g.coset_fast && g.tw_on) {
                    if (g.tw_full)          // streamed table in the layout of the output: one 16-byte load instead of two loads and a multiplication
                        v = fe_mul(v, g.tw_full[(long long)(g.coset0 + blockIdx.y) * g.tw_full_stride + (long long)(tile * T + t) * g.out_lane + (long long)k * g.out_point]);
                    else
                        v = fe_mul(v, tw_lookup(g.cw, (unsigned long long)(tile * T + t) * (((unsigned long long)k << g.log_blowup) + g.coset0 + blockIdx.y)));
                }
                else if (g.tw_on) v = fe_mul(v, tw_lookup(g.tw, (unsigned long long)(tile * T + t) * k));
                if (g.has_scale) v = fe_mul(v, g.scale);
                dst[(long long)t * g.out_lane + (long long)k * g.out_point] = v;
            } else {
                s_data[sidx<LOG_L, LANE_MAJOR>(pos, t, T)] = x[m];
            }
        }
    }
}

template <int LOG_L, bool LANE_MAJOR>
__global__ void __launch_bounds__(256, 2) ntt_pass_kernel(const fe *__restrict__ src, fe *__restrict__ dst, const PassGeom g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int L = 1 << LOG_L;
    typedef Rounds<LOG_L> RD;
    fe *s_tw = reinterpret_cast<fe *>(smem_raw);          // L entries: per-stage tables back to back
    fe *s_data = s_tw + L;
    const int T = 1 << g.log_t;
    const unsigned tile = blockIdx.x % g.num_tiles, outer = blockIdx.x / g.num_tiles;
    const long long in_base = (long long)outer * g.in_outer + (long long)tile * T * g.in_lane;   // index inside the vector
    src += (long long)blockIdx.y * g.in_batch_y + (long long)blockIdx.z * g.in_batch_z;
    dst += (long long)blockIdx.y * g.out_batch_y + (long long)blockIdx.z * g.out_batch_z + (long long)outer * g.out_outer +
           (long long)tile * T * g.out_lane;
    for (int i = threadIdx.x; i < L - 1; i += blockDim.x) s_tw[i] = g.roots[i];
    if (g.tw_full) {
        // the streamed twiddles are consumed in the last round: start pulling this block's T*16-byte segments (one per output k) into L2 now
        const fe *tb = g.tw_full + (long long)(g.coset0 + blockIdx.y) * g.tw_full_stride + (long long)tile * T * g.out_lane;
        for (int k = threadIdx.x; k < L; k += blockDim.x)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(tb + (long long)k * g.out_point));
    }
    __syncthreads();
    ntt_round<LOG_L, 0, RD::R1, true, RD::R2 == 0, LANE_MAJOR>(src, dst, s_data, s_tw, g, tile, in_base);
    if constexpr (RD::R2 > 0) {
        __syncthreads();
        ntt_round<LOG_L, RD::R1, RD::R2, false, RD::R3 == 0, LANE_MAJOR>(src, dst, s_data, s_tw, g, tile, in_base);
    }
    if constexpr (RD::R3 > 0) {
        __syncthreads();
        ntt_round<LOG_L, RD::R1 + RD::R2, RD::R3, false, true, LANE_MAJOR>(src, dst, s_data, s_tw, g, tile, in_base);
    }
}

typedef void (*PassKernel)(const fe *, fe *, const PassGeom);
template <bool LM> static PassKernel pass_kernel_t(int log_l) {
    switch (log_l) {
        case 1: return ntt_pass_kernel<1, LM>;  case 2: return ntt_pass_kernel<2, LM>;  case 3: return ntt_pass_kernel<3, LM>;
        case 4: return ntt_pass_kernel<4, LM>;  case 5: return ntt_pass_kernel<5, LM>;  case 6: return ntt_pass_kernel<6, LM>;
        case 7: return ntt_pass_kernel<7, LM>;  case 8: return ntt_pass_kernel<8, LM>;  case 9: return ntt_pass_kernel<9, LM>;
        case 10: return ntt_pass_kernel<10, LM>;
    }
    throw Error(-1, "unsupported sub-transform size");
}

static void launch_pass(Context &c, int log_l, PassGeom g, const fe *src, fe *dst, unsigned blocks_x, unsigned by, unsigned bz) {
    const int L = 1 << log_l, T = 1 << g.log_t;
    const size_t data = g.lane_major ? (size_t)T * (L + (L >> 3) + 1) : (size_t)L * T;
    const size_t smem = ((size_t)L + data) * sizeof(fe);
    int threads = L * T / 16;
    if (threads > 256) threads = 256;
    if (threads < 32) threads = 32;
    PassKernel k = g.lane_major ? pass_kernel_t<true>(log_l) : pass_kernel_t<false>(log_l);
    static bool attr_set[2][MAX_LOG_L + 1] = {{false}};
    if (!attr_set[g.lane_major][log_l]) {
        DG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set[g.lane_major][log_l] = true;
    }
    DG_REQUIRE(by <= 65535 && bz <= 65535, "batch too large for one launch");
    k<<<dim3(blocks_x, by, bz), threads, smem, c.stream>>>(src, dst, g); c.launches++;
    DG_CUDA(cudaGetLastError());
}

static int lanes_log(int log_l, long long available) {
    int lt = 4;                                   // 16 lanes
    while ((1 << (log_l + lt)) > 4096 && lt > 0) lt--;    // keep tiles at 4096 elements = 64 KB
    while ((1LL << lt) > available && lt > 0) lt--;
    return lt;
}

// split log_n into 1..3 pass sizes (outermost first)
static int split_passes(int log_n, int l[3]) {
    if (log_n <= MAX_LOG_L) { l[0] = log_n; return 1; }
    if (log_n <= 2 * MAX_LOG_L) { l[0] = (log_n + 1) / 2; l[1] = log_n / 2; return 2; }
    DG_REQUIRE(log_n <= 3 * MAX_LOG_L, "transform too large");
    l[0] = (log_n + 2) / 3; l[1] = (log_n + 1) / 3; l[2] = log_n / 3;
    return 3;
}

struct CosetSpec { bool on; int log_blowup; int fold; unsigned coset0; };

// table[c][k * R1 + lane] = w_N^(lane * (k * b + c)): the twiddle the first LDE pass applies to output k of lane `lane` on coset c.
// It depends on the shape only (n, b, first pass size), so it is built once per shape and kept: N elements (512 MB for 2^20 x 32).
__global__ void lde_twiddle_fill_kernel(fe *table, TwiddleRef cw, int log_n, int log_r1, int log_b) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> (log_n + log_b)) return;
    const unsigned long long c = i >> log_n, pos = i & ((1ULL << log_n) - 1);
    const unsigned long long lane = pos & ((1ULL << log_r1) - 1), k = pos >> log_r1;
    table[i] = tw_lookup(cw, lane * ((k << log_b) + c));
}
static const fe *lde_twiddle_table(Context &c, int log_n, int log_b, int l0) {
    static long long cap = -1;
    if (cap < 0) { const char *e = getenv("DG_LDE_TW_MB"); cap = (e ? atoll(e) : 1024) << 20; }    // 0 disables the table
    const size_t bytes = ((size_t)16 << (log_n + log_b));
    if ((long long)bytes > cap) return nullptr;
    const long long key = ((long long)log_n << 16) | (log_b << 8) | l0;
    auto it = c.lde_twiddles.find(key);
    if (it == c.lde_twiddles.end()) {
        DevBuf t;
        t.alloc(bytes, true);
        const unsigned long long cnt = 1ULL << (log_n + log_b);
        lde_twiddle_fill_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, c.stream>>>(t.as<fe>(), c.twiddle(log_n + log_b, false), log_n, log_n - l0, log_b);
        c.launches++;
        DG_CUDA(cudaGetLastError());
        it = c.lde_twiddles.emplace(key, std::move(t)).first;
    }
    return it->second.as<fe>();
}

// Runs the passes of one batched transform.  `by` = number of y-batches (cosets for the LDE, else 1), `bz` = vectors.
// src strides: vector stride src_stride (z), y stride 0 for the LDE (every coset reads the same coefficients).
static void run_transform(Context &c, const fe *src, fe *dst, int log_n, bool inverse, unsigned by, unsigned bz, long long src_stride_z,
                          long long dst_stride_y, long long dst_stride_z, CosetSpec cs) {
    int l[3];
    const int np = split_passes(log_n, l);
    const long long n = 1LL << log_n;
    fe scale = fe_make(1, 0);
    if (inverse) scale = host_inv(fe_make((unsigned long long)n, 0));

    PassGeom base;
    memset(&base, 0, sizeof base);
    if (cs.on) {
        base.coset_on = 1;
        base.fold = cs.fold;
        base.fold_stride = n;
        base.cw = c.twiddle(log_n + cs.log_blowup, false);
        base.log_blowup = cs.log_blowup;
        base.coset0 = cs.coset0;
        if (cs.fold == 1) {
            // order of w_N^in_point where in_point = n / N1 (first pass) : N / in_point = N1 << log_blowup
            const int log_order = l[0] + cs.log_blowup;
            base.coset_fast = 1;
            base.cw_point = c.single_table(log_order);
            base.cw_point_mask = (1u << log_order) - 1u;
        }
    }

    fe *tmp = nullptr;
    long long tmp_stride_y = n, tmp_stride_z = n * by;
    if (np > 1) {
        c.ntt_tmp.ensure((size_t)n * by * bz * sizeof(fe), true);
        tmp = c.ntt_tmp.as<fe>();
    }

    if (np == 1) {
        PassGeom g = base;
        g.log_t = 0; g.num_tiles = 1;
        g.in_point = 1; g.out_point = 1; g.in_lane = 0; g.out_lane = 0;
        g.in_batch_y = 0; g.in_batch_z = src_stride_z; g.out_batch_y = dst_stride_y; g.out_batch_z = dst_stride_z;
        g.lane_major = 1; g.tw_on = 0;
        g.has_scale = inverse; g.scale = scale;
        g.roots = c.roots(l[0], inverse);
        launch_pass(c, l[0], g, src, dst, 1, by, bz);
        return;
    }

    const long long N1 = 1LL << l[0];
    const long long R1 = n >> l[0];
    {   // pass 1: N1-point transforms over j1 (stride R1), twiddle w_n^(j' * k1)
        PassGeom g = base;
        g.log_t = lanes_log(l[0], R1);
        g.num_tiles = (unsigned)(R1 >> g.log_t);
        g.in_point = R1; g.in_lane = 1; g.out_point = R1; g.out_lane = 1;
        g.in_batch_y = 0; g.in_batch_z = src_stride_z; g.out_batch_y = tmp_stride_y; g.out_batch_z = tmp_stride_z;
        g.lane_major = 0;
        g.tw_on = 1; g.tw = c.twiddle(log_n, inverse);
        if (g.coset_fast) { g.tw_full = lde_twiddle_table(c, log_n, cs.log_blowup, l[0]); g.tw_full_stride = n; }
        g.roots = c.roots(l[0], inverse);
        launch_pass(c, l[0], g, src, tmp, g.num_tiles, by, bz);
    }
    long long N2 = 1;
    if (np == 3) {   // pass 2: within every row k1, N2-point transforms over ja (stride N3), twiddle w_R1^(jb * ka)
        N2 = 1LL << l[1];
        const long long N3 = 1LL << l[2];
        PassGeom g;
        memset(&g, 0, sizeof g);
        g.log_t = lanes_log(l[1], N3);
        g.num_tiles = (unsigned)(N3 >> g.log_t);
        g.in_point = N3; g.in_lane = 1; g.in_outer = R1; g.out_point = N3; g.out_lane = 1; g.out_outer = R1;
        g.in_batch_y = tmp_stride_y; g.in_batch_z = tmp_stride_z; g.out_batch_y = tmp_stride_y; g.out_batch_z = tmp_stride_z;
        g.lane_major = 0;
        g.tw_on = 1; g.tw = c.twiddle(log_n - l[0], inverse);
        g.roots = c.roots(l[1], inverse);
        launch_pass(c, l[1], g, tmp, tmp, (unsigned)(g.num_tiles * N1), by, bz);
    }
    {   // last pass: contiguous NL-point transforms; output index k1 + N1*ka + N1*N2*kb
        const int ll = l[np - 1];
        PassGeom g;
        memset(&g, 0, sizeof g);
        g.log_t = lanes_log(ll, N1);
        g.num_tiles = (unsigned)(N1 >> g.log_t);
        g.in_point = 1; g.in_lane = R1; g.in_outer = (np == 3) ? (1LL << ll) : 0;
        g.out_lane = 1; g.out_outer = (np == 3) ? N1 : 0; g.out_point = N1 * N2;
        g.in_batch_y = tmp_stride_y; g.in_batch_z = tmp_stride_z; g.out_batch_y = dst_stride_y; g.out_batch_z = dst_stride_z;
        g.lane_major = 1; g.tw_on = 0;
        g.has_scale = inverse; g.scale = scale;
        g.roots = c.roots(ll, inverse);
        launch_pass(c, ll, g, tmp, dst, (unsigned)(g.num_tiles * N2), by, bz);
    }
}

void ntt_batch(Context &c, const fe *src, fe *dst, int log_n, int batch, size_t src_stride, size_t dst_stride, bool inverse) {
    DG_REQUIRE(log_n >= 1 && log_n <= 30, "log_n out of range");
    const size_t n = (size_t)1 << log_n;
    // bound the scratch: process vectors in chunks of at most ~1 GiB of scratch
    size_t max_chunk = std::max<size_t>(1, ((size_t)1 << 30) / (n * sizeof(fe)));
    if (max_chunk > 65535) max_chunk = 65535;
    for (size_t b0 = 0; b0 < (size_t)batch; b0 += max_chunk) {
        size_t nb = std::min(max_chunk, (size_t)batch - b0);
        run_transform(c, src + b0 * src_stride, dst + b0 * dst_stride, log_n, inverse, 1, (unsigned)nb, (long long)src_stride, 0,
                      (long long)dst_stride, CosetSpec{false, 0, 1, 0});
    }
}

void lde_batch(Context &c, const fe *src, fe *dst, int log_n, int log_blowup, int fold, int batch, size_t src_stride, size_t dst_stride,
               unsigned coset0, unsigned ncosets) {
    DG_REQUIRE(log_n >= 1 && log_n + log_blowup <= 30, "LDE domain too large");
    const size_t n = (size_t)1 << log_n;
    const unsigned cosets = ncosets ? ncosets : (1u << log_blowup);
    DG_REQUIRE(coset0 + cosets <= (1u << log_blowup), "coset range out of bounds");
    size_t max_chunk = std::max<size_t>(1, ((size_t)1 << 30) / (n * cosets * sizeof(fe)));
    if (max_chunk > 65535) max_chunk = 65535;
    for (size_t b0 = 0; b0 < (size_t)batch; b0 += max_chunk) {
        size_t nb = std::min(max_chunk, (size_t)batch - b0);
        run_transform(c, src + b0 * src_stride, dst + b0 * dst_stride, log_n, false, cosets, (unsigned)nb, (long long)src_stride,
                      (long long)n, (long long)dst_stride, CosetSpec{true, log_blowup, fold, coset0});
    }
}

}  // namespace dg

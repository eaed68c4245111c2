// Pass kernels of the batched NTT / coset LDE (included by ntt.cu and ntt_inl.cu, which differ in how fe_mul is emitted:
// out of line -- one shared body, small code -- or inlined).  DG_NTT_TAG keeps the kernel symbols of the two translation units apart.
#pragma once
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
// up to RMAX stages that run entirely in registers on 2^rho elements per thread ("unit"); shared memory is touched only
// between rounds.  The first round reads its operands straight from global memory and the last one writes straight back.
// RMAX = 4: 16 elements per unit, 3 rounds for 1024 points (fewest shared-memory round trips, ~110 registers);
// RMAX = 3 / 2: 8 / 4 elements per unit, 4 / 5 rounds (smaller unrolled bodies, fewer registers, more resident warps).
// Stage twiddles come from per-stage compact tables W_st[j] = w_L^(j << st) (unit-stride, conflict-free) staged in shared memory.
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

template <int LOG_L, int RMAX, int S0, bool LANE_MAJOR>
__device__ __forceinline__ void ntt_rounds(const fe *__restrict__ src, fe *__restrict__ dst, fe *s_data, const fe *s_tw, const PassGeom &g,
                                           unsigned tile, long long in_base) {
    constexpr int REM = LOG_L - S0, LEFT = (REM + RMAX - 1) / RMAX, RHO = (REM + LEFT - 1) / LEFT;   // even split, largest round first
    ntt_round<LOG_L, S0, RHO, S0 == 0, S0 + RHO == LOG_L, LANE_MAJOR>(src, dst, s_data, s_tw, g, tile, in_base);
    if constexpr (S0 + RHO < LOG_L) {
        __syncthreads();
        ntt_rounds<LOG_L, RMAX, S0 + RHO, LANE_MAJOR>(src, dst, s_data, s_tw, g, tile, in_base);
    }
}

template <int LOG_L, bool LANE_MAJOR, int RMAX, int BT, int MINB, int TAG>
__global__ void __launch_bounds__(BT, MINB) ntt_pass_kernel(const fe *__restrict__ src, fe *__restrict__ dst, const PassGeom g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int L = 1 << LOG_L;
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
    ntt_rounds<LOG_L, RMAX, 0, LANE_MAJOR>(src, dst, s_data, s_tw, g, tile, in_base);
}

typedef void (*PassKernel)(const fe *, fe *, const PassGeom);

}  // namespace dg

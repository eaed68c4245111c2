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
#include "ntt_pass.cuh"

namespace dg {

// kernel variants (ntt_pass.cuh): RMAX stages per register round, BT threads per block; inline-multiply instantiations live in ntt_inl.cu
PassKernel pass_kernel_inline(bool lane_major, int log_l, int rmax, int bt);       // nullptr when that combination is not instantiated
template <bool LM, int RMAX, int BT, int MINB> static PassKernel pass_kernel_t(int log_l) {
    switch (log_l) {
        case 1: return ntt_pass_kernel<1, LM, RMAX, BT, MINB, 0>;  case 2: return ntt_pass_kernel<2, LM, RMAX, BT, MINB, 0>;
        case 3: return ntt_pass_kernel<3, LM, RMAX, BT, MINB, 0>;  case 4: return ntt_pass_kernel<4, LM, RMAX, BT, MINB, 0>;
        case 5: return ntt_pass_kernel<5, LM, RMAX, BT, MINB, 0>;  case 6: return ntt_pass_kernel<6, LM, RMAX, BT, MINB, 0>;
        case 7: return ntt_pass_kernel<7, LM, RMAX, BT, MINB, 0>;  case 8: return ntt_pass_kernel<8, LM, RMAX, BT, MINB, 0>;
        case 9: return ntt_pass_kernel<9, LM, RMAX, BT, MINB, 0>;  case 10: return ntt_pass_kernel<10, LM, RMAX, BT, MINB, 0>;
    }
    throw Error(-1, "unsupported sub-transform size");
}

// Kernel configuration per pass kind: A = strided passes (first / middle), B = the contiguous last pass; "f" variants are used by the
// transforms whose first pass folds 8 coefficient blocks (constraint / composition LDE).  B200, 2^20 x 32 LDE of 26 columns (r02):
//   out-of-line multiply, 16-element units, 256 threads (r01)   57.8 ms
//   out-of-line, 4-element units, 512 threads                   53.3 ms      (32 instead of 16 resident warps per SM)
//   inline multiply, 8-element units, 512 threads               51.8 ms      inline, 4-element units, 512 threads: 52.7 ms
//   inline, 256 threads (8 / 4-element units)                   61.2 / 58.7 ms
// Environment overrides (read once): DG_NTT_A / DG_NTT_B / DG_NTT_AF = "rmax,threads,inline", DG_NTT_TILE = log2 of the tile size.
struct PassCfg { int rmax, bt, inl; };
struct NttConfig { PassCfg a, b, af; int tile_log; };
static PassCfg parse_cfg(const char *env, PassCfg d) {
    const char *e = getenv(env);
    if (e) { int r = 0, t = 0, i = 0; if (sscanf(e, "%d,%d,%d", &r, &t, &i) == 3) d = PassCfg{r, t, i}; }
    if (d.rmax < 2 || d.rmax > 4) d.rmax = 4;
    if (d.bt != 512 && d.bt != 1024) d.bt = 256;
    return d;
}
static const NttConfig &ntt_config() {
    static NttConfig cfg;
    static bool init = false;
    if (!init) {
        cfg.a = parse_cfg("DG_NTT_A", PassCfg{3, 512, 1});
        cfg.b = parse_cfg("DG_NTT_B", PassCfg{3, 512, 1});
        cfg.af = parse_cfg("DG_NTT_AF", PassCfg{2, 512, 1});
        const char *e = getenv("DG_NTT_TILE");
        cfg.tile_log = e ? atoi(e) : 12;
        if (cfg.tile_log < 12 || cfg.tile_log > 13) cfg.tile_log = 12;
        init = true;
    }
    return cfg;
}

static void launch_pass(Context &c, int log_l, PassGeom g, const fe *src, fe *dst, unsigned blocks_x, unsigned by, unsigned bz) {
    const int L = 1 << log_l, T = 1 << g.log_t;
    const size_t data = g.lane_major ? (size_t)T * (L + (L >> 3) + 1) : (size_t)L * T;
    const size_t smem = ((size_t)L + data) * sizeof(fe);
    const NttConfig &nc = ntt_config();
    const PassCfg cfg = g.lane_major ? nc.b : ((g.coset_on && !g.coset_fast) ? nc.af : nc.a);
    int rmax = cfg.rmax, bt = cfg.bt;
    PassKernel k = nullptr;
    if (cfg.inl) k = pass_kernel_inline(g.lane_major != 0, log_l, rmax, bt);
    if (!k) {
        const bool lm = g.lane_major != 0;
        if (bt == 1024) bt = 512;
        if (rmax == 4) { bt = 256; k = lm ? pass_kernel_t<true, 4, 256, 2>(log_l) : pass_kernel_t<false, 4, 256, 2>(log_l); }
        else if (rmax == 3 && bt == 512) k = lm ? pass_kernel_t<true, 3, 512, 2>(log_l) : pass_kernel_t<false, 3, 512, 2>(log_l);
        else if (rmax == 3) k = lm ? pass_kernel_t<true, 3, 256, 3>(log_l) : pass_kernel_t<false, 3, 256, 3>(log_l);
        else if (bt == 512) k = lm ? pass_kernel_t<true, 2, 512, 2>(log_l) : pass_kernel_t<false, 2, 512, 2>(log_l);
        else k = lm ? pass_kernel_t<true, 2, 256, 4>(log_l) : pass_kernel_t<false, 2, 256, 4>(log_l);
    }
    int threads = (L * T) >> (log_l < rmax ? log_l : rmax);           // one unit per thread in the largest round
    if (threads > bt) threads = bt;
    if (threads < 32) threads = 32;
    set_func_smem(c, (const void *)k, 200 * 1024);
    DG_REQUIRE(by <= 65535 && bz <= 65535, "batch too large for one launch");
    // (r02: folding the vector index into blockIdx.x so that all columns of a (tile, coset) share the streamed twiddles in L2 cost 120 bytes
    //  of spills at the 64-register cap and made the trace LDE 1% slower: rejected)
    k<<<dim3(blocks_x, by, bz), threads, smem, c.stream>>>(src, dst, g); c.launches++;
    DG_CUDA(cudaGetLastError());
}

static int lanes_log(int log_l, long long available) {
    int lt = 4;                                   // 16 lanes
    const int tile = 1 << ntt_config().tile_log;
    while ((1 << (log_l + lt)) > tile && lt > 0) lt--;    // keep tiles at 4096 elements = 64 KB (8192 with DG_NTT_TILE=13)
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

// ---- fold-8 input transform for all cosets at once -----------------------------------------------------------------------------------
// Extending a polynomial with 8n coefficients (constraint / composition polynomial) over all b cosets needs, for every coefficient
// position j < n and coset c, the value  w_N^(c j) * sum_{f<8} p[j + f n] * zeta^(c f),  zeta = w_N^n = w_b.  Inside the pass kernel that
// is a Horner evaluation per (coset, position): 9 multiplications per input element, as much as the whole transform that follows.
// For all cosets together the inner sums are a b-point DFT of the zero-padded 8-vector: with c = (b/8) q + r it is, per residue r, a
// twist by zeta^(r f) (7 multiplications) and an 8-point DFT over q (5 multiplications); the outer factor is a geometric progression in
// q.  One thread per (r, j): 29 multiplications for 8 outputs instead of 72.  The result feeds b plain size-n transforms.
__global__ void __launch_bounds__(256) prefold8_kernel(const fe *__restrict__ src, fe *__restrict__ dst, int log_n, int log_b, const fe *__restrict__ zeta,
                                                       TwiddleRef twN, fe w8, fe w8_2, fe w8_3) {
    const unsigned long long n = 1ULL << log_n;
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >> (log_n + log_b - 3)) return;
    const unsigned long long j = t & (n - 1);
    const unsigned r = (unsigned)(t >> log_n);                 // residue of the coset index modulo b/8
    const unsigned bmask = (1u << log_b) - 1u, step = 1u << (log_b - 3);
    fe x[8];
#pragma unroll
    for (int f = 0; f < 8; f++) x[f] = src[j + (unsigned long long)f * n];
    if (r) {
#pragma unroll
        for (int f = 1; f < 8; f++) x[f] = fe_mul(x[f], zeta[(r * (unsigned)f) & bmask]);
    }
    // 8-point DFT X[q] = sum_f x[f] w8^(q f), decimation in frequency, outputs in natural order
    fe u[4], v[4];
    const fe wp[4] = {fe_make(1, 0), w8, w8_2, w8_3};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        u[i] = fe_add(x[i], x[i + 4]);
        v[i] = fe_sub(x[i], x[i + 4]);
        if (i) v[i] = fe_mul(v[i], wp[i]);
    }
    fe X[8];
    {
        fe p0 = fe_add(u[0], u[2]), p1 = fe_add(u[1], u[3]), q0 = fe_sub(u[0], u[2]), q1 = fe_mul(fe_sub(u[1], u[3]), w8_2);
        X[0] = fe_add(p0, p1); X[4] = fe_sub(p0, p1); X[2] = fe_add(q0, q1); X[6] = fe_sub(q0, q1);
        p0 = fe_add(v[0], v[2]); p1 = fe_add(v[1], v[3]); q0 = fe_sub(v[0], v[2]); q1 = fe_mul(fe_sub(v[1], v[3]), w8_2);
        X[1] = fe_add(p0, p1); X[5] = fe_sub(p0, p1); X[3] = fe_add(q0, q1); X[7] = fe_sub(q0, q1);
    }
    // outer factor w_N^(c j), c = step q + r: w_N^(r j) * (w_N^(step j))^q
    fe cur = tw_lookup(twN, (unsigned long long)r * j);
    const fe rho = tw_lookup(twN, (unsigned long long)step * j);
#pragma unroll
    for (int q = 0; q < 8; q++) {
        dst[((unsigned long long)(step * (unsigned)q + r) << log_n) + j] = fe_mul(X[q], cur);
        if (q < 7) cur = fe_mul(cur, rho);
    }
}

void lde_batch(Context &c, const fe *src, fe *dst, int log_n, int log_blowup, int fold, int batch, size_t src_stride, size_t dst_stride,
               unsigned coset0, unsigned ncosets) {
    DG_REQUIRE(log_n >= 1 && log_n + log_blowup <= 30, "LDE domain too large");
    const size_t n = (size_t)1 << log_n;
    const unsigned cosets = ncosets ? ncosets : (1u << log_blowup);
    DG_REQUIRE(coset0 + cosets <= (1u << log_blowup), "coset range out of bounds");
    static int prefold = -1;
    if (prefold < 0) { const char *e = getenv("DG_LDE_PREFOLD"); prefold = e ? atoi(e) : 1; }
    if (prefold && fold == 8 && batch == 1 && cosets == (1u << log_blowup) && log_blowup >= 3 && log_blowup <= 8) {
        // all cosets: input transform for every coset in one kernel, then b plain transforms (r02, 2^20 x 32: 3.7 -> 2.5 ms per polynomial)
        DevBuf pre((size_t)n * cosets * sizeof(fe));
        const fe w8 = host_root_of_unity(3), w8_2 = fe_mul(w8, w8);
        const unsigned long long threads = (unsigned long long)n << (log_blowup - 3);
        prefold8_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, c.stream>>>(src, pre.as<fe>(), log_n, log_blowup, c.single_table(log_blowup),
                                                                                 c.twiddle(log_n + log_blowup, false), w8, w8_2, fe_mul(w8_2, w8)); c.launches++;
        DG_CUDA(cudaGetLastError());
        ntt_batch(c, pre.as<fe>(), dst, log_n, (int)cosets, n, n, false);
        return;
    }
    // scratch of the two-pass transforms: one intermediate of n * cosets elements per vector; more vectors per launch = fewer passes over
    // the streamed first-pass twiddles (DG_NTT_SCRATCH_MB, default 4096)
    static size_t scratch_cap = 0;
    if (!scratch_cap) { const char *e = getenv("DG_NTT_SCRATCH_MB"); scratch_cap = (size_t)(e ? atoll(e) : 4096) << 20; }
    size_t max_chunk = std::max<size_t>(1, scratch_cap / (n * cosets * sizeof(fe)));
    if (max_chunk > 65535) max_chunk = 65535;
    for (size_t b0 = 0; b0 < (size_t)batch; b0 += max_chunk) {
        size_t nb = std::min(max_chunk, (size_t)batch - b0);
        run_transform(c, src + b0 * src_stride, dst + b0 * dst_stride, log_n, false, cosets, (unsigned)nb, (long long)src_stride,
                      (long long)n, (long long)dst_stride, CosetSpec{true, log_blowup, fold, coset0});
    }
}

}  // namespace dg

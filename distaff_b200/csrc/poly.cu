// Polynomial kernels of the prove pipeline (coefficient-domain steps 4 and 6 of /root/reference/src/stark/prover.rs):
//   polynom::syn_div_in_place            /root/reference/src/math/polynom.rs:190-197   -> weighted suffix scan
//   polynom::syn_div_expanded_in_place   /root/reference/src/math/polynom.rs:202-236   -> strided suffix sums + 2-tap stencil
//   polynom::eval (Horner)               /root/reference/src/math/polynom.rs:9-17       -> dot product with a power table
//   parallel::mul_acc / add_in_place     /root/reference/src/math/parallel.rs           -> fused element-wise kernels
//
// The sequential recurrences of the reference are re-expressed as parallel scans.  With q[i] = sum_{j>i} a[j] b^(j-i-1)
// (exactly what the synthetic-division loop leaves in a[i], remainder dropped) one has
//      q[i] = b^-(i+1) * sum_{j>i} a[j] b^j
// i.e. an element-wise scaling, a plain (addition-only) exclusive suffix sum, and another scaling.  Field arithmetic is
// exact, so the coefficients are identical to the reference's.
#include "poly.h"

namespace dg {

__device__ __forceinline__ fe pw(const PowRef &t, unsigned long long e) {
    return fe_mul(t.lo[e & ((1ULL << t.lo_bits) - 1ULL)], t.hi[e >> t.lo_bits]);
}

__global__ void pow_fill_kernel(fe *out, fe base, unsigned long long count) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = fe_pow_u64(base, i);
}

PowTable::PowTable(Context &c, fe base, unsigned long long len) {
    lo_bits = 1;
    while ((1ULL << (2 * lo_bits)) < len) lo_bits++;
    const unsigned long long lo_n = 1ULL << lo_bits;
    const unsigned long long hi_n = (len + lo_n - 1) / lo_n + 1;
    lo.alloc(lo_n * sizeof(fe));
    hi.alloc(hi_n * sizeof(fe));
    fe step = fe_pow_u64(base, lo_n);
    pow_fill_kernel<<<(unsigned)((lo_n + 127) / 128), 128, 0, c.stream>>>(lo.as<fe>(), base, lo_n); c.launches++;
    pow_fill_kernel<<<(unsigned)((hi_n + 127) / 128), 128, 0, c.stream>>>(hi.as<fe>(), step, hi_n); c.launches++;
    DG_CUDA(cudaGetLastError());
}

// ---- exclusive suffix sum (plain additions) -----------------------------------------------------------------------------------------
static const int SCAN_THREADS = 256, SCAN_PER_THREAD = 4, SCAN_BLOCK = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ fe shfl_down_fe(fe v, int d) {
    fe r;
    r.lo = __shfl_down_sync(0xffffffffu, v.lo, d);
    r.hi = __shfl_down_sync(0xffffffffu, v.hi, d);
    return r;
}

// block-wide inclusive suffix sum of one value per thread; returns (inclusive suffix over threads >= tid); total in *block_total
__device__ __forceinline__ fe block_suffix_inclusive(fe v, fe *s_warp, fe *block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        fe o = shfl_down_fe(v, d);
        if (lane + d < 32) v = fe_add(v, o);
    }
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    fe above = fe_make(0, 0);
    for (int w2 = warp + 1; w2 < nwarps; w2++) above = fe_add(above, s_warp[w2]);
    if (block_total) {
        fe tot = fe_make(0, 0);
        for (int w2 = 0; w2 < nwarps; w2++) tot = fe_add(tot, s_warp[w2]);
        *block_total = tot;
    }
    __syncthreads();
    return fe_add(v, above);
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const fe *__restrict__ data, fe *__restrict__ sums, unsigned long long len) {
    __shared__ fe s_warp[SCAN_THREADS / 32];
    const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_BLOCK + (unsigned long long)threadIdx.x * SCAN_PER_THREAD;
    fe v = fe_make(0, 0);
#pragma unroll
    for (int u = 0; u < SCAN_PER_THREAD; u++)
        if (base + u < len) v = fe_add(v, data[base + u]);
    fe total;
    block_suffix_inclusive(v, s_warp, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// data[i] <- sum_{j>i, j in block} data[j] + carry[block]     (carry may be null)
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(fe *__restrict__ data, const fe *__restrict__ carry, unsigned long long len) {
    __shared__ fe s_warp[SCAN_THREADS / 32];
    const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_BLOCK + (unsigned long long)threadIdx.x * SCAN_PER_THREAD;
    fe x[SCAN_PER_THREAD];
    fe v = fe_make(0, 0);
#pragma unroll
    for (int u = 0; u < SCAN_PER_THREAD; u++) {
        x[u] = (base + u < len) ? data[base + u] : fe_make(0, 0);
        v = fe_add(v, x[u]);
    }
    fe incl = block_suffix_inclusive(v, s_warp, nullptr);
    fe run = fe_sub(incl, v);                           // sum over threads strictly above
    if (carry) run = fe_add(run, carry[blockIdx.x]);
#pragma unroll
    for (int u = SCAN_PER_THREAD - 1; u >= 0; u--) {
        if (base + u < len) data[base + u] = run;
        run = fe_add(run, x[u]);
    }
}

void suffix_scan_exclusive(Context &c, fe *data, unsigned long long len) {
    const unsigned long long nblk = (len + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nblk == 1) {
        scan_apply_kernel<<<1, SCAN_THREADS, 0, c.stream>>>(data, nullptr, len); c.launches++;
        DG_CUDA(cudaGetLastError());
        return;
    }
    DevBuf sums(nblk * sizeof(fe));
    scan_block_sums_kernel<<<(unsigned)nblk, SCAN_THREADS, 0, c.stream>>>(data, sums.as<fe>(), len); c.launches++;
    DG_CUDA(cudaGetLastError());
    suffix_scan_exclusive(c, sums.as<fe>(), nblk);
    scan_apply_kernel<<<(unsigned)nblk, SCAN_THREADS, 0, c.stream>>>(data, sums.as<fe>(), len); c.launches++;
    DG_CUDA(cudaGetLastError());
}

// ---- synthetic division by (x - b) ----------------------------------------------------------------------------------------------------
__global__ void scale_by_pow_kernel(const fe *__restrict__ in, fe *__restrict__ out, PowRef t, unsigned long long offset, unsigned long long len,
                                    fe sub0) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    fe v = in[i];
    if (i == 0) v = fe_sub(v, sub0);
    out[i] = fe_mul(v, pw(t, i + offset));
}

// out[i] = sum_{j>i} (in[j] - [j==0] sub0) b^(j-i-1);   `scratch` holds len elements.  in may equal out.
void syn_div(Context &c, const fe *in, fe *out, fe *scratch, unsigned long long len, const PowRef &b_pows, const PowRef &binv_pows, fe sub0) {
    const unsigned blocks = (unsigned)((len + 255) / 256);
    scale_by_pow_kernel<<<blocks, 256, 0, c.stream>>>(in, scratch, b_pows, 0, len, sub0); c.launches++;
    DG_CUDA(cudaGetLastError());
    suffix_scan_exclusive(c, scratch, len);
    scale_by_pow_kernel<<<blocks, 256, 0, c.stream>>>(scratch, out, binv_pows, 1, len, fe_make(0, 0)); c.launches++;
    DG_CUDA(cudaGetLastError());
}

// ---- division by (x^n - 1) / (x - e) -----------------------------------------------------------------------------------------------------
// s[i + m n] = sum_{m' >= m} a[i + m' n]
__global__ void strided_suffix_kernel(const fe *__restrict__ a, fe *__restrict__ s, unsigned long long n, int blocks_m) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe run = fe_make(0, 0);
    for (int m = blocks_m - 1; m >= 0; m--) {
        run = fe_add(run, a[i + (unsigned long long)m * n]);
        s[i + (unsigned long long)m * n] = run;
    }
}
// out[idx] = s[idx+n-1] - e * s[idx+n]  for idx <= len-n (with s[len] = 0), else 0 ; optionally accumulated: out = base0 + base1 + that
__global__ void expanded_stencil_kernel(const fe *__restrict__ s, const fe *__restrict__ add0, const fe *__restrict__ add1, fe *__restrict__ out,
                                        unsigned long long n, unsigned long long len, fe e) {
    unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= len) return;
    fe v = fe_make(0, 0);
    if (idx <= len - n) {
        v = s[idx + n - 1];
        if (idx + n < len) v = fe_sub(v, fe_mul(e, s[idx + n]));
    }
    if (add0) v = fe_add(v, add0[idx]);
    if (add1) v = fe_add(v, add1[idx]);
    out[idx] = v;
}
void syn_div_expanded_sum(Context &c, const fe *a, fe *scratch, const fe *add0, const fe *add1, fe *out, unsigned long long n, unsigned long long len, fe e) {
    strided_suffix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(a, scratch, n, (int)(len / n)); c.launches++;
    expanded_stencil_kernel<<<(unsigned)((len + 255) / 256), 256, 0, c.stream>>>(scratch, add0, add1, out, n, len, e); c.launches++;
    DG_CUDA(cudaGetLastError());
}

// ---- evaluation of many polynomials at two points (DEEP values) -----------------------------------------------------------------------------
// partial[(col*2 + p) * chunks + chunk] = sum_{k in chunk} poly[col][k] * x_p^k,  x_0 = z (table zt), x_1 = z*g (zt * gt)
static const int EVAL_CHUNK = 4096;
__global__ void __launch_bounds__(256) eval2_partial_kernel(const fe *__restrict__ polys, unsigned long long n, PowRef zt, TwiddleRef gt, fe *__restrict__ partial,
                                                            int two_points) {
    __shared__ fe s_warp[2][8];
    const unsigned long long col = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
    const fe *p = polys + col * n;
    fe a0 = fe_make(0, 0), a1 = fe_make(0, 0);
    for (int u = 0; u < EVAL_CHUNK / 256; u++) {
        unsigned long long k = chunk * EVAL_CHUNK + (unsigned long long)u * 256 + threadIdx.x;
        if (k < n) {
            fe v = p[k];
            fe zk = pw(zt, k);
            fe t = fe_mul(v, zk);
            a0 = fe_add(a0, t);
            if (two_points) {
                unsigned ee = (unsigned)k & gt.mask;
                fe gk = fe_mul(gt.lo[ee & ((1u << gt.lo_bits) - 1u)], gt.hi[ee >> gt.lo_bits]);
                a1 = fe_add(a1, fe_mul(t, gk));
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { a0 = fe_add(a0, shfl_down_fe(a0, d)); a1 = fe_add(a1, shfl_down_fe(a1, d)); }
    if (lane == 0) { s_warp[0][warp] = a0; s_warp[1][warp] = a1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        fe t0 = fe_make(0, 0), t1 = fe_make(0, 0);
        for (int w2 = 0; w2 < 8; w2++) { t0 = fe_add(t0, s_warp[0][w2]); t1 = fe_add(t1, s_warp[1][w2]); }
        partial[(col * 2 + 0) * chunks + chunk] = t0;
        partial[(col * 2 + 1) * chunks + chunk] = t1;
    }
}
__global__ void reduce_partials_kernel(const fe *__restrict__ partial, fe *__restrict__ out, unsigned long long chunks) {
    // one block of 32 threads per output
    const unsigned long long o = blockIdx.x;
    fe a = fe_make(0, 0);
    for (unsigned long long i = threadIdx.x; i < chunks; i += 32) a = fe_add(a, partial[o * chunks + i]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) a = fe_add(a, shfl_down_fe(a, d));
    if (threadIdx.x == 0) out[o] = a;
}
// out[col*2 + p] = poly_col(x_p)
void eval_polys_at(Context &c, const fe *polys, unsigned long long n, int cols, const PowRef &zt, const TwiddleRef &gt, bool two_points, fe *out) {
    const unsigned chunks = (unsigned)((n + EVAL_CHUNK - 1) / EVAL_CHUNK);
    DevBuf partial((size_t)cols * 2 * chunks * sizeof(fe));
    eval2_partial_kernel<<<dim3(chunks, cols), 256, 0, c.stream>>>(polys, n, zt, gt, partial.as<fe>(), two_points ? 1 : 0); c.launches++;
    reduce_partials_kernel<<<cols * 2, 32, 0, c.stream>>>(partial.as<fe>(), out, chunks); c.launches++;
    DG_CUDA(cudaGetLastError());
}

// ---- linear combinations ------------------------------------------------------------------------------------------------------------------
// t1[k] = sum_i cc1[i] P_i[k],  t2[k] = sum_i cc2[i] P_i[k]
__global__ void lincomb2_kernel(const fe *__restrict__ polys, unsigned long long n, int w, const fe *__restrict__ cc1, const fe *__restrict__ cc2,
                                fe *__restrict__ t1, fe *__restrict__ t2) {
    unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    fe a = fe_make(0, 0), b = fe_make(0, 0);
    for (int i = 0; i < w; i++) {
        fe v = polys[(unsigned long long)i * n + k];
        a = fe_add(a, fe_mul(v, cc1[i]));
        b = fe_add(b, fe_mul(v, cc2[i]));
    }
    t1[k] = a; t2[k] = b;
}
void lincomb2(Context &c, const fe *polys, unsigned long long n, int w, const fe *cc1, const fe *cc2, fe *t1, fe *t2) {
    lincomb2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(polys, n, w, cc1, cc2, t1, t2); c.launches++;
    DG_CUDA(cudaGetLastError());
}

// Boundary-constraint numerators in coefficient form.  The reference evaluates, at every point x of the 8n-point constraint domain,
//   I(x) = sum_j (T_j(x) - in_j) (a_j + b_j x^adj)            (evaluator.rs:181-326, adj = 6n + 2)
// and interpolates the evaluations afterwards (constraint_poly.rs).  I has degree n - 1 + adj < 8n, so the interpolant IS the polynomial
//   I = (sum_j a_j T_j - Ka) + x^adj (sum_j b_j T_j - Kb),     Ka = sum_j a_j in_j,  Kb = sum_j b_j in_j
// whose coefficients are two linear combinations of the trace polynomials' coefficients: n*nb multiplications instead of 8n*nb.
// coef = [a_init | b_init | a_final | b_final], nb entries each; ic / fc receive the 8n coefficients of the first / last step numerators.
__global__ void boundary_coeffs_kernel(const fe *__restrict__ polys, unsigned long long n, int nb, const fe *__restrict__ coef, fe KiA, fe KiB,
                                       fe KfA, fe KfB, unsigned long long adj, fe *__restrict__ ic, fe *__restrict__ fc) {
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    fe ia = fe_make(0, 0), ib = ia, fa = ia, fb = ia;
    for (int j = 0; j < nb; j++) {
        const fe v = polys[(unsigned long long)j * n + k];
        ia = fe_add(ia, fe_mul(v, coef[j]));
        ib = fe_add(ib, fe_mul(v, coef[nb + j]));
        fa = fe_add(fa, fe_mul(v, coef[2 * nb + j]));
        fb = fe_add(fb, fe_mul(v, coef[3 * nb + j]));
    }
    if (k == 0) { ia = fe_sub(ia, KiA); ib = fe_sub(ib, KiB); fa = fe_sub(fa, KfA); fb = fe_sub(fb, KfB); }
    const fe zero = fe_make(0, 0);
    ic[k] = ia; fc[k] = fa;
    ic[adj + k] = ib; fc[adj + k] = fb;
    // the gaps [n, adj) and [adj + n, 8n): 5n + 2 + (n - 2) = 6n entries, six per thread
    for (unsigned long long q = n + k; q < 8 * n; q += n)
        if (q < adj || q >= adj + n) { ic[q] = zero; fc[q] = zero; }
}
void boundary_coeffs(Context &c, const fe *polys, unsigned long long n, int nb, const fe *coef, fe KiA, fe KiB, fe KfA, fe KfB, fe *ic, fe *fc) {
    boundary_coeffs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(polys, n, nb, coef, KiA, KiB, KfA, KfB, 6 * n + 2, ic, fc); c.launches++;
    DG_CUDA(cudaGetLastError());
}

// composition polynomial (trace_table.rs:241-258, constraint_poly.rs:49):
//   comp[k] = cq[k]*kc + [k < n] (t1q[k]+t2q[k])*k1 + [inc <= k < inc+n] (t1q[k-inc]+t2q[k-inc])*k2
__global__ void compose_kernel(const fe *__restrict__ t1q, const fe *__restrict__ t2q, const fe *__restrict__ cq, fe *__restrict__ comp,
                               unsigned long long n, unsigned long long len, unsigned long long inc, fe k1, fe k2, fe kc) {
    unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= len) return;
    fe v = fe_mul(cq[k], kc);
    if (k < n) v = fe_add(v, fe_mul(fe_add(t1q[k], t2q[k]), k1));
    if (k >= inc && k < inc + n) v = fe_add(v, fe_mul(fe_add(t1q[k - inc], t2q[k - inc]), k2));
    comp[k] = v;
}
void compose(Context &c, const fe *t1q, const fe *t2q, const fe *cq, fe *comp, unsigned long long n, unsigned long long len, unsigned long long inc,
             fe k1, fe k2, fe kc) {
    compose_kernel<<<(unsigned)((len + 255) / 256), 256, 0, c.stream>>>(t1q, t2q, cq, comp, n, len, inc, k1, k2, kc); c.launches++;
    DG_CUDA(cudaGetLastError());
}

// ---- gathers for the query openings -----------------------------------------------------------------------------------------------------------
// out[q*w + j] = ext[j * col_stride + phys_q]   (phys_q = position inside the rank's coset-major slab, computed on the host)
__global__ void gather_rows_kernel(const fe *__restrict__ ext, int w, unsigned long long col_stride, const unsigned long long *__restrict__ phys,
                                   int nq, fe *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * w) return;
    int q = t / w, j = t % w;
    out[t] = ext[(unsigned long long)j * col_stride + phys[q]];
}
void gather_rows(Context &c, const fe *ext, int w, unsigned long long col_stride, const unsigned long long *d_phys, int nq, fe *d_out) {
    gather_rows_kernel<<<(nq * w + 127) / 128, 128, 0, c.stream>>>(ext, w, col_stride, d_phys, nq, d_out); c.launches++;
    DG_CUDA(cudaGetLastError());
}
// out[t] = src[idx[t]] for 32-byte items
__global__ void gather32_kernel(const uint4 *__restrict__ src, const unsigned long long *__restrict__ idx, int count, uint4 *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    out[2 * t] = src[2 * idx[t]];
    out[2 * t + 1] = src[2 * idx[t] + 1];
}
void gather32(Context &c, const void *src, const unsigned long long *d_idx, int count, void *d_out) {
    if (count == 0) return;
    gather32_kernel<<<(count + 127) / 128, 128, 0, c.stream>>>((const uint4 *)src, d_idx, count, (uint4 *)d_out); c.launches++;
    DG_CUDA(cudaGetLastError());
}
// out[t] = src[idx[t]] for 16-byte items
__global__ void gather16_kernel(const fe *__restrict__ src, const unsigned long long *__restrict__ idx, int count, fe *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) out[t] = src[idx[t]];
}
void gather16(Context &c, const fe *src, const unsigned long long *d_idx, int count, fe *d_out) {
    if (count == 0) return;
    gather16_kernel<<<(count + 127) / 128, 128, 0, c.stream>>>(src, d_idx, count, d_out); c.launches++;
    DG_CUDA(cudaGetLastError());
}

}  // namespace dg

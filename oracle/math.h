// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or called by the product path.
//
// CPU restatement (single thread, unsigned __int128) of the reference's math layer:
//   field   : /root/reference/src/math/field.rs:27-244   (add, sub, mul, inv, inv_many, exp, roots, power series)
//   fft     : /root/reference/src/math/fft.rs:16-108      (recursive in-place radix-2 FFT, permuted output)
//   polynom : /root/reference/src/math/polynom.rs:9-256   (eval, fft eval/interpolate, syn_div*, lagrange)
//   quartic : /root/reference/src/math/quartic.rs:6-162   (degree-3 batch interpolate / evaluate, transpose)
// The algorithms deliberately mirror the reference's sequential structure (this file is also the
// "restated reference (C++)" CPU baseline); the GPU path computes the same values differently.
#ifndef ORACLE_MATH_H
#define ORACLE_MATH_H

#include <cstdint>
#include <cstring>
#include <cassert>
#include <vector>
#include <stdexcept>
#include <algorithm>

namespace oracle {

typedef unsigned __int128 u128;
typedef uint64_t u64;

static inline u128 mk128(u64 lo, u64 hi) { return ((u128)hi << 64) | lo; }

// field.rs:11   M = 2^128 - 45 * 2^40 + 1
static const u128 M = mk128(0xffffd30000000001ULL, 0xffffffffffffffffULL);
// field.rs:14   2^40-th root of unity  (23953097886125630542083529559205016746)
static const u128 G = mk128(0x86b8723e1920f4aaULL, 0x120532e7b364080aULL);

namespace field {

// field.rs:27-31
static inline u128 add(u128 a, u128 b) {
    u128 z = M - b;
    return a < z ? M - z + a : a - z;
}
// field.rs:33-36
static inline u128 sub(u128 a, u128 b) { return a < b ? M - b + a : a - b; }

// --- helpers of field.rs:287-336, kept limb-for-limb so the multiplication follows the reference ---
struct u192 { u64 w0, w1, w2; };

static inline u192 mul_128x64(u128 a, u64 b) {
    u128 z_lo = (u128)(u64)a * (u128)b;
    u128 z_hi = (a >> 64) * (u128)b;
    z_hi = z_hi + (z_lo >> 64);
    return { (u64)z_lo, (u64)z_hi, (u64)(z_hi >> 64) };
}
static inline u192 sub_192x192(u64 a0, u64 a1, u64 a2, u64 b0, u64 b1, u64 b2) {
    u128 z0 = (u128)a0 - (u128)b0;
    u128 z1 = (u128)a1 - ((u128)b1 + (z0 >> 127));
    u128 z2 = (u128)a2 - ((u128)b2 + (z1 >> 127));
    return { (u64)z0, (u64)z1, (u64)z2 };
}
static inline u192 add_192x192(u64 a0, u64 a1, u64 a2, u64 b0, u64 b1, u64 b2) {
    u128 z0 = (u128)a0 + (u128)b0;
    u128 z1 = (u128)a1 + (u128)b1 + (z0 >> 64);
    u128 z2 = (u128)a2 + (u128)b2 + (z1 >> 64);
    return { (u64)z0, (u64)z1, (u64)z2 };
}
static inline u192 mul_by_modulus(u64 a) {
    u128 a_lo = (u128)a * M;  // wrapping
    u64 a_hi = a == 0 ? 0 : a - 1;
    return { (u64)a_lo, (u64)(a_lo >> 64), a_hi };
}
static inline u192 mul_reduce(u64 z0, u64 z1, u64 z2) {
    u192 q = mul_by_modulus(z2);
    return sub_192x192(z0, z1, z2, q.w0, q.w1, q.w2);
}
static inline void sub_modulus(u64 &lo, u64 &hi) {
    u128 z = (u128)0 - M;
    z += (u128)lo;
    z += (u128)hi << 64;
    lo = (u64)z; hi = (u64)(z >> 64);
}

// field.rs:38-73
static inline u128 mul(u128 a, u128 b) {
    u192 x = mul_128x64(a, (u64)(b >> 64));
    x = mul_reduce(x.w0, x.w1, x.w2);
    if (x.w2 == 1) sub_modulus(x.w0, x.w1);

    u192 y = mul_128x64(a, (u64)b);
    u128 t = (u128)y.w1 + x.w0;
    u64 y1 = (u64)t;
    t = (u128)y.w2 + x.w1 + (u64)(t >> 64);
    u64 y2 = (u64)t;
    u64 y3 = (u64)(t >> 64);
    if (y3 == 1) sub_modulus(y1, y2);

    u192 z = mul_reduce(y.w0, y1, y2);
    if (z.w2 == 1 || (z.w1 == (u64)(M >> 64) && z.w0 >= (u64)M)) sub_modulus(z.w0, z.w1);
    return mk128(z.w0, z.w1);
}

static inline u128 neg(u128 x) { return sub(0, x); }

// field.rs:201-219
static inline u128 exp(u128 b, u128 p) {
    if (b == 0) return 0;
    if (p == 0) return 1;
    u128 r = 1;
    while (p > 0) {
        if (p & 1) r = mul(r, b);
        p >>= 1;
        b = mul(b, b);
    }
    return r;
}

// field.rs:83-162  binary extended-GCD inverse; inv(0) = 0
static inline u128 inv(u128 x) {
    if (x == 0) return 0;
    const u64 M0 = (u64)M, M1 = (u64)(M >> 64);
    u128 v = M;
    u192 a = {0, 0, 0};
    u192 u = (x & 1) ? u192{ (u64)x, (u64)(x >> 64), 0 }
                     : add_192x192((u64)x, (u64)(x >> 64), 0, M0, M1, 0);
    u192 d = { M0 - 1, M1, 0 };
    while (v != 1) {
        while (u.w2 > 0 || mk128(u.w0, u.w1) > v) {
            u = sub_192x192(u.w0, u.w1, u.w2, (u64)v, (u64)(v >> 64), 0);
            d = add_192x192(d.w0, d.w1, d.w2, a.w0, a.w1, a.w2);
            while ((u.w0 & 1) == 0) {
                if (d.w0 & 1) d = add_192x192(d.w0, d.w1, d.w2, M0, M1, 0);
                u.w0 = (u.w0 >> 1) | ((u.w1 & 1) << 63);
                u.w1 = (u.w1 >> 1) | ((u.w2 & 1) << 63);
                u.w2 >>= 1;
                d.w0 = (d.w0 >> 1) | ((d.w1 & 1) << 63);
                d.w1 = (d.w1 >> 1) | ((d.w2 & 1) << 63);
                d.w2 >>= 1;
            }
        }
        v = v - mk128(u.w0, u.w1);
        a = add_192x192(a.w0, a.w1, a.w2, d.w0, d.w1, d.w2);
        while ((v & 1) == 0) {
            if (a.w0 & 1) a = add_192x192(a.w0, a.w1, a.w2, M0, M1, 0);
            v >>= 1;
            a.w0 = (a.w0 >> 1) | ((a.w1 & 1) << 63);
            a.w1 = (a.w1 >> 1) | ((a.w2 & 1) << 63);
            a.w2 >>= 1;
        }
    }
    u128 r = mk128(a.w0, a.w1);
    while (a.w2 > 0 || r >= M) {
        a = sub_192x192(a.w0, a.w1, a.w2, M0, M1, 0);
        r = mk128(a.w0, a.w1);
    }
    return r;
}

static inline u128 div(u128 a, u128 b) { return mul(a, inv(b)); }

// field.rs:173-192  Montgomery-trick batch inversion; zeros map to zeros
static inline void inv_many_fill(const u128 *values, u128 *result, size_t n) {
    u128 last = 1;
    for (size_t i = 0; i < n; i++) {
        result[i] = last;
        if (values[i] != 0) last = mul(last, values[i]);
    }
    last = inv(last);
    for (size_t i = n; i-- > 0;) {
        if (values[i] == 0) result[i] = 0;
        else {
            result[i] = mul(last, result[i]);
            last = mul(last, values[i]);
        }
    }
}
static inline std::vector<u128> inv_many(const std::vector<u128> &v) {
    std::vector<u128> r(v.size());
    inv_many_fill(v.data(), r.data(), v.size());
    return r;
}

// field.rs:228-234
static inline u128 get_root_of_unity(size_t order) {
    if (order == 0 || (order & (order - 1))) throw std::runtime_error("order must be a power of 2");
    unsigned tz = __builtin_ctzll(order);
    if (tz > 40) throw std::runtime_error("order cannot exceed 2^40");
    return exp(G, (u128)1 << (40 - tz));
}
// field.rs:237-244
static inline std::vector<u128> get_power_series(u128 b, size_t length) {
    std::vector<u128> r(length);
    r[0] = 1;
    for (size_t i = 1; i < length; i++) r[i] = mul(r[i - 1], b);
    return r;
}
// field.rs:69-73
static inline void mul_acc(u128 *a, const u128 *b, u128 c, size_t n) {
    for (size_t i = 0; i < n; i++) a[i] = add(a[i], mul(b[i], c));
}

} // namespace field

namespace fft {

static const size_t MAX_LOOP = 256;  // fft.rs:7

// Optional host threading (or_set_threads).  The reference's prover passes num_threads = 1 to every FFT / batch call, so 1 is the
// faithful default; with T > 1 the same butterflies run in OpenMP tasks (independent sub-transforms, chunked twiddle loops) and
// the column / row / step loops of the prover are split over threads.  Every value is identical for any T.
inline int &host_threads() { static int t = 1; return t; }

static inline void butterfly(u128 *v, size_t offset, size_t stride) {
    size_t i = offset, j = offset + stride;
    u128 t = v[i];
    v[i] = field::add(t, v[j]);
    v[j] = field::sub(t, v[j]);
}
static inline void butterfly_twiddle(u128 *v, u128 tw, size_t offset, size_t stride) {
    size_t i = offset, j = offset + stride;
    u128 t = v[i];
    v[j] = field::mul(v[j], tw);
    v[i] = field::add(t, v[j]);
    v[j] = field::sub(t, v[j]);
}

// fft.rs:16-56 (num_threads fixed to 1, as every call site in the prover does)
static void fft_in_place(u128 *values, size_t len, const u128 *twiddles, size_t count, size_t stride, size_t offset) {
    size_t size = len / stride;
    if (size > 2) {
        if (stride == count && count < MAX_LOOP) {
            fft_in_place(values, len, twiddles, 2 * count, 2 * stride, offset);
        } else {
            fft_in_place(values, len, twiddles, count, 2 * stride, offset);
            fft_in_place(values, len, twiddles, count, 2 * stride, offset + stride);
        }
    }
    for (size_t o = offset; o < offset + count; o++) butterfly(values, o, stride);
    size_t last_offset = offset + size * stride;
    size_t i = 0;
    for (size_t o = offset; o < last_offset; o += 2 * stride, i++) {
        if (i == 0) continue;
        for (size_t j = o; j < o + count; j++) butterfly_twiddle(values, twiddles[i], j, stride);
    }
}

// task-parallel form of fft_in_place: the two half-size sub-transforms of the `else` branch are independent tasks, the twiddle loop
// of a large level is cut into chunks.  Must be called inside an OpenMP parallel region (see fft_in_place_mt).
static void fft_in_place_tasks(u128 *values, size_t len, const u128 *twiddles, size_t count, size_t stride, size_t offset) {
    size_t size = len / stride;
    if (size * count <= ((size_t)1 << 14)) { fft_in_place(values, len, twiddles, count, stride, offset); return; }   // elements below this call
    if (stride == count && count < MAX_LOOP) {
        fft_in_place_tasks(values, len, twiddles, 2 * count, 2 * stride, offset);
    } else {
        #pragma omp task default(shared)
        fft_in_place_tasks(values, len, twiddles, count, 2 * stride, offset);
        #pragma omp task default(shared)
        fft_in_place_tasks(values, len, twiddles, count, 2 * stride, offset + stride);
        #pragma omp taskwait
    }
    for (size_t o = offset; o < offset + count; o++) butterfly(values, o, stride);
    const size_t groups = size / 2;                      // group i covers offsets [offset + 2 i stride, .. + count)
    const size_t chunk = std::max<size_t>(1, 4096 / count);
    #pragma omp taskloop default(shared) grainsize(1)
    for (size_t c0 = 1; c0 < groups; c0 += chunk) {
        const size_t c1 = std::min(groups, c0 + chunk);
        for (size_t i = c0; i < c1; i++) {
            const size_t o = offset + 2 * i * stride;
            for (size_t j = o; j < o + count; j++) butterfly_twiddle(values, twiddles[i], j, stride);
        }
    }
}
static inline void fft_in_place_mt(u128 *values, size_t len, const u128 *twiddles) {
    const int T = host_threads();
    if (T <= 1 || len < ((size_t)1 << 15)) { fft_in_place(values, len, twiddles, 1, 1, 0); return; }
    #pragma omp parallel num_threads(T)
    #pragma omp single
    fft_in_place_tasks(values, len, twiddles, 1, 1, 0);
}

static inline size_t permute_index(size_t size, size_t index) {
    if (size == 1) return 0;
    unsigned bits = __builtin_ctzll(size);
    size_t r = 0;
    for (unsigned b = 0; b < bits; b++) r |= ((index >> b) & 1) << (bits - 1 - b);
    return r;
}
// fft.rs:71-79
static inline void permute(u128 *v, size_t n) {
    for (size_t i = 0; i < n; i++) {
        size_t j = permute_index(n, i);
        if (j > i) std::swap(v[i], v[j]);
    }
}
// fft.rs:58-69
static inline std::vector<u128> get_twiddles(u128 root, size_t size) {
    std::vector<u128> tw = field::get_power_series(root, size / 2);
    permute(tw.data(), tw.size());
    return tw;
}
static inline std::vector<u128> get_inv_twiddles(u128 root, size_t size) {
    u128 inv_root = field::exp(root, (u128)(size - 1));
    return get_twiddles(inv_root, size);
}

} // namespace fft

namespace polynom {

// polynom.rs:9-17
static inline u128 eval(const u128 *p, size_t n, u128 x) {
    u128 y = 0, pw = 1;
    for (size_t i = 0; i < n; i++) {
        y = field::add(y, field::mul(p[i], pw));
        pw = field::mul(pw, x);
    }
    return y;
}
static inline u128 eval(const std::vector<u128> &p, u128 x) { return eval(p.data(), p.size(), x); }

// polynom.rs:34-41
static inline void eval_fft_twiddles(u128 *p, size_t n, const u128 *twiddles, bool unpermute) {
    fft::fft_in_place_mt(p, n, twiddles);
    if (unpermute) fft::permute(p, n);
}
// polynom.rs:93-103
static inline void interpolate_fft_twiddles(u128 *v, size_t n, const u128 *inv_twiddles, bool unpermute) {
    fft::fft_in_place_mt(v, n, inv_twiddles);
    u128 inv_len = field::inv((u128)n);
    for (size_t i = 0; i < n; i++) v[i] = field::mul(v[i], inv_len);
    if (unpermute) fft::permute(v, n);
}
static inline void eval_fft(std::vector<u128> &p) {
    u128 g = field::get_root_of_unity(p.size());
    std::vector<u128> tw = fft::get_twiddles(g, p.size());
    eval_fft_twiddles(p.data(), p.size(), tw.data(), true);
}
static inline void interpolate_fft(std::vector<u128> &v) {
    u128 g = field::get_root_of_unity(v.size());
    std::vector<u128> tw = fft::get_inv_twiddles(g, v.size());
    interpolate_fft_twiddles(v.data(), v.size(), tw.data(), true);
}

// polynom.rs:190-197
static inline void syn_div_in_place(u128 *a, size_t n, u128 b) {
    u128 c = 0;
    for (size_t i = n; i-- > 0;) {
        u128 t = field::add(a[i], field::mul(b, c));
        a[i] = c;
        c = t;
    }
}
// polynom.rs:202-236
static inline void syn_div_expanded_in_place(u128 *a, size_t len, size_t degree, const u128 *exceptions, size_t n_exc) {
    std::vector<u128> result(a, a + len);
    result.reserve(len + n_exc);
    size_t degree_offset = len - degree;
    for (size_t i = degree_offset; i-- > 0;) result[i] = field::add(result[i], result[i + degree]);
    for (size_t e = 0; e < n_exc; e++) {
        u128 exception = field::neg(exceptions[e]);
        result.push_back(0);
        u128 next_term = result[0];
        result[0] = 0;
        for (size_t i = 0; i < result.size() - 1; i++) {
            result[i] = field::add(result[i], field::mul(next_term, exception));
            std::swap(next_term, result[i + 1]);
        }
    }
    for (size_t i = 0; i < degree_offset + n_exc; i++) a[i] = result[degree + i];
    for (size_t i = degree_offset + n_exc; i < len; i++) a[i] = 0;
}

// polynom.rs:240-245
static inline size_t degree_of(const u128 *p, size_t n) {
    for (size_t i = n; i-- > 0;) if (p[i] != 0) return i;
    return 0;
}
static inline size_t infer_degree(const std::vector<u128> &evaluations) {
    std::vector<u128> p(evaluations);
    interpolate_fft(p);
    return degree_of(p.data(), p.size());
}

// polynom.rs:156-178  long division (remainder ignored)
static inline std::vector<u128> div(const std::vector<u128> &a_in, const std::vector<u128> &b) {
    size_t apos = degree_of(a_in.data(), a_in.size());
    std::vector<u128> a(a_in);
    size_t bpos = degree_of(b.data(), b.size());
    if (apos < bpos) throw std::runtime_error("cannot divide by polynomial of higher degree");
    std::vector<u128> result(apos - bpos + 1, 0);
    for (size_t i = result.size(); i-- > 0;) {
        u128 quot = field::div(a[apos], b[bpos]);
        result[i] = quot;
        for (size_t j = bpos; j-- > 0;) a[i + j] = field::sub(a[i + j], field::mul(b[j], quot));
        apos--;
    }
    return result;
}
// polynom.rs:262-279
static inline std::vector<u128> get_zero_roots(const std::vector<u128> &xs) {
    size_t n = xs.size() + 1;
    std::vector<u128> result(n);
    n -= 1;
    result[n] = 1;
    for (size_t i = 0; i < xs.size(); i++) {
        n -= 1;
        result[n] = 0;
        for (size_t j = n; j < xs.size(); j++)
            result[j] = field::sub(result[j], field::mul(result[j + 1], xs[i]));
    }
    return result;
}
// polynom.rs:47-76  Lagrange interpolation
static inline std::vector<u128> interpolate(const std::vector<u128> &xs, const std::vector<u128> &ys) {
    std::vector<u128> roots = get_zero_roots(xs);
    std::vector<u128> divisor = {0, 1};
    std::vector<std::vector<u128>> numerators;
    for (size_t i = 0; i < xs.size(); i++) {
        divisor[0] = field::neg(xs[i]);
        numerators.push_back(div(roots, divisor));
    }
    std::vector<u128> denominators;
    for (size_t i = 0; i < xs.size(); i++) denominators.push_back(eval(numerators[i], xs[i]));
    denominators = field::inv_many(denominators);
    std::vector<u128> result(xs.size(), 0);
    for (size_t i = 0; i < xs.size(); i++) {
        u128 y_slice = field::mul(ys[i], denominators[i]);
        for (size_t j = 0; j < xs.size(); j++)
            if (numerators[i][j] != 0 && ys[i] != 0)
                result[j] = field::add(result[j], field::mul(numerators[i][j], y_slice));
    }
    return result;
}

} // namespace polynom

namespace quartic {

struct Q { u128 v[4]; };

// quartic.rs:6-17
static inline u128 eval(const u128 *p, u128 x) {
    u128 y = field::add(p[0], field::mul(p[1], x));
    u128 x2 = field::mul(x, x);
    y = field::add(y, field::mul(p[2], x2));
    u128 x3 = field::mul(x2, x);
    y = field::add(y, field::mul(p[3], x3));
    return y;
}
// quartic.rs:20-31
static inline std::vector<u128> evaluate_batch(const std::vector<Q> &polys, u128 x) {
    std::vector<u128> r(polys.size());
    const int T = fft::host_threads();
    #pragma omp parallel for num_threads(T) if (T > 1 && polys.size() >= 4096)
    for (size_t i = 0; i < polys.size(); i++) r[i] = eval(polys[i].v, x);
    return r;
}
// quartic.rs:37-135
static inline std::vector<Q> interpolate_batch(const std::vector<Q> &xs_all, const std::vector<Q> &ys_all) {
    using namespace field;
    size_t n = xs_all.size();
    std::vector<Q> equations(n * 4);
    std::vector<u128> inverses(n * 4);
    const int T = fft::host_threads();
    #pragma omp parallel for num_threads(T) if (T > 1 && n >= 4096)
    for (size_t i = 0; i < n; i++) {
        const size_t j = 4 * i;
        const u128 *xs = xs_all[i].v;
        u128 x01 = mul(xs[0], xs[1]), x02 = mul(xs[0], xs[2]), x03 = mul(xs[0], xs[3]);
        u128 x12 = mul(xs[1], xs[2]), x13 = mul(xs[1], xs[3]), x23 = mul(xs[2], xs[3]);
        equations[j]     = {{ mul(neg(x12), xs[3]), add(add(x12, x13), x23), sub(sub(neg(xs[1]), xs[2]), xs[3]), 1 }};
        inverses[j]      = eval(equations[j].v, xs[0]);
        equations[j + 1] = {{ mul(neg(x02), xs[3]), add(add(x02, x03), x23), sub(sub(neg(xs[0]), xs[2]), xs[3]), 1 }};
        inverses[j + 1]  = eval(equations[j + 1].v, xs[1]);
        equations[j + 2] = {{ mul(neg(x01), xs[3]), add(add(x01, x03), x13), sub(sub(neg(xs[0]), xs[1]), xs[3]), 1 }};
        inverses[j + 2]  = eval(equations[j + 2].v, xs[2]);
        equations[j + 3] = {{ mul(neg(x01), xs[2]), add(add(x01, x02), x12), sub(sub(neg(xs[0]), xs[1]), xs[2]), 1 }};
        inverses[j + 3]  = eval(equations[j + 3].v, xs[3]);
    }
    if (T > 1 && n >= 4096) {                            // Montgomery's trick per chunk: the same (exact) inverses as one global batch
        const size_t total = n * 4, chunk = (total + (size_t)T * 4 - 1) / ((size_t)T * 4);
        std::vector<u128> out(total);
        #pragma omp parallel for num_threads(T)
        for (size_t c0 = 0; c0 < total; c0 += chunk) inv_many_fill(inverses.data() + c0, out.data() + c0, std::min(chunk, total - c0));
        inverses.swap(out);
    } else {
        inverses = inv_many(inverses);
    }
    std::vector<Q> result(n);
    #pragma omp parallel for num_threads(T) if (T > 1 && n >= 4096)
    for (size_t i = 0; i < n; i++) {
        const size_t j = 4 * i;
        const u128 *ys = ys_all[i].v;
        u128 inv_y = mul(ys[0], inverses[j]);
        for (int k = 0; k < 4; k++) result[i].v[k] = mul(inv_y, equations[j].v[k]);
        for (int e = 1; e < 4; e++) {
            inv_y = mul(ys[e], inverses[j + e]);
            for (int k = 0; k < 4; k++)
                result[i].v[k] = add(result[i].v[k], mul(inv_y, equations[j + e].v[k]));
        }
    }
    return result;
}
// quartic.rs:137-152
static inline std::vector<Q> transpose(const u128 *vec, size_t len, size_t stride) {
    if (len % (4 * stride) != 0) throw std::runtime_error("vector length must be divisible by 4*stride");
    size_t rows = len / (4 * stride);
    std::vector<Q> r(rows);
    for (size_t i = 0; i < rows; i++)
        r[i] = {{ vec[i * stride], vec[(i + rows) * stride], vec[(i + 2 * rows) * stride], vec[(i + 3 * rows) * stride] }};
    return r;
}

} // namespace quartic

} // namespace oracle
#endif

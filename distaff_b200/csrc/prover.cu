// Host orchestration of the CUDA prove pipeline: the body that replaces /root/reference/src/stark/prover.rs:17-169.
// Stage numbering and names follow the reference's nine debug!() sections (prover.rs:19-167) so that per-stage timings of
// the CPU prover and of this backend line up.  Only 32-byte roots, a handful of challenges and the final openings cross
// the PCIe bus after the register traces have been uploaded.
#include "prover.h"
#include "air.h"
#include "host_fs.h"
#include "poly.h"
#include "shard.h"
#include <array>
#include <memory>

namespace dg {

namespace {

struct StageClock {
    cudaStream_t s;
    cudaEvent_t ev[10];
    explicit StageClock(cudaStream_t stream) : s(stream) { for (auto &e : ev) DG_CUDA(cudaEventCreate(&e)); }
    ~StageClock() { for (auto &e : ev) cudaEventDestroy(e); }
    void mark(int i) { DG_CUDA(cudaEventRecord(ev[i], s)); }
    float between(int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, ev[a], ev[b]); return ms; }
};

void d2h(Context &c, void *dst, const void *src, size_t bytes) {
    DG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.stream));
    DG_CUDA(cudaStreamSynchronize(c.stream));
}
void h2d(Context &c, void *dst, const void *src, size_t bytes) {
    DG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.stream));
}

// optional dump of intermediate device buffers (differential debugging against the oracle): DG_DEBUG_DUMP=<dir>
void debug_dump(Context &c, const char *name, const void *dev, size_t bytes) {
    const char *dir = getenv("DG_DEBUG_DUMP");
    if (!dir) return;
    std::vector<uint8_t> host(bytes);
    d2h(c, host.data(), dev, bytes);
    std::string path = std::string(dir) + "/" + name + ".bin";
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return;
    fwrite(host.data(), 1, bytes, f);
    fclose(f);
}

void debug_dump_host(const char *name, const void *host, size_t bytes) {
    const char *dir = getenv("DG_DEBUG_DUMP");
    if (!dir) return;
    std::string path = std::string(dir) + "/" + name + ".bin";
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return;
    fwrite(host, 1, bytes, f);
    fclose(f);
}


// fetches 32-byte items src[idx[i]] to the host
std::vector<Digest> fetch32(Context &c, const void *src, const std::vector<uint64_t> &idx) {
    std::vector<Digest> out(idx.size());
    if (idx.empty()) return out;
    DevBuf d_idx(idx.size() * 8), d_out(idx.size() * 32);
    h2d(c, d_idx.p, idx.data(), idx.size() * 8);
    gather32(c, src, d_idx.as<unsigned long long>(), (int)idx.size(), d_out.p);
    d2h(c, out.data(), d_out.p, idx.size() * 32);
    return out;
}
std::vector<fe> fetch16(Context &c, const fe *src, const std::vector<uint64_t> &idx) {
    std::vector<fe> out(idx.size());
    if (idx.empty()) return out;
    DevBuf d_idx(idx.size() * 8), d_out(idx.size() * 16);
    h2d(c, d_idx.p, idx.data(), idx.size() * 8);
    gather16(c, src, d_idx.as<unsigned long long>(), (int)idx.size(), d_out.as<fe>());
    d2h(c, out.data(), d_out.p, idx.size() * 16);
    return out;
}

// materialises the node lists of a batch proof; `leaf_fetch` maps leaf indices and `node_fetch` heap indices to 32-byte values
template <typename LeafFetch, typename NodeFetch>
std::vector<std::vector<Digest>> resolve_plan(const fs::BatchPlan &plan, LeafFetch leaf_fetch, NodeFetch node_fetch) {
    std::vector<uint64_t> leaf_idx, node_idx;
    for (auto &slot : plan.nodes)
        for (auto &r : slot) (r.leaf ? leaf_idx : node_idx).push_back(r.index);
    std::vector<Digest> leaves = leaf_fetch(leaf_idx), nodes = node_fetch(node_idx);
    std::vector<std::vector<Digest>> out(plan.nodes.size());
    size_t li = 0, ni = 0;
    for (size_t s = 0; s < plan.nodes.size(); s++)
        for (auto &r : plan.nodes[s]) out[s].push_back(r.leaf ? leaves[li++] : nodes[ni++]);
    return out;
}

void write_digest_vec(fs::ByteWriter &w, const std::vector<Digest> &v) {
    w.u64(v.size());
    for (auto &d : v) w.raw(d.data(), 32);
}
void write_digest_vv(fs::ByteWriter &w, const std::vector<std::vector<Digest>> &v) {
    w.u64(v.size());
    for (auto &x : v) write_digest_vec(w, x);
}
void write_felt_vec(fs::ByteWriter &w, const std::vector<fe> &v) {
    w.u64(v.size());
    for (auto &x : v) w.felt(x);
}

int ilog2(uint64_t v) { int l = 0; while ((1ULL << l) < v) l++; return l; }

struct FriLayerDev {
    DevBuf leaves, nodes, folded;     // row hashes, tree, and the folded values (= values of the next layer)
    const fe *vals;
    Layout layout;                    // layout of `vals` (domain size 2^layout.log_d)
    Digest root;
};

}  // namespace

// d_regs: register traces in device memory; when `host_cols` is given they are not there yet: column chunks are uploaded on the
// copy stream while the previous chunk is being interpolated and extended (the upload hides behind the LDE)
static Proof *prove_core(Context &c, fe *d_regs, const uint8_t *const *host_cols, uint32_t width, uint64_t length, uint32_t ctx_depth,
                         uint32_t loop_depth, const uint8_t *inputs16, uint32_t n_inputs, const uint8_t *outputs16, uint32_t n_outputs,
                         const dg_options_t &opt, dg_prove_stats_t *stats) {
    // ---- argument checks (trace_table.rs:23-58, options.rs:29-50, lib.rs:33-34) -----------------------------------------------
    const uint64_t n = length, b = opt.extension_factor;
    DG_REQUIRE(opt.hash_id == 0, "unsupported hash function (only blake3 is serialisable, options.rs:107)");
    DG_REQUIRE(b >= 16 && b <= 256 && (b & (b - 1)) == 0, "extension_factor must be a power of 2 between 16 and 256");
    DG_REQUIRE(opt.num_queries > 0 && opt.num_queries <= 128, "num_queries must be in 1..128");
    DG_REQUIRE(opt.grinding_factor <= 32, "grinding factor cannot be greater than 32");
    DG_REQUIRE(n >= 16 && (n & (n - 1)) == 0, "execution trace length must be a power of 2 and at least 16");
    DG_REQUIRE(ctx_depth <= 16, "context depth cannot be greater than 16");
    DG_REQUIRE(loop_depth <= 8, "loop depth cannot be greater than 8");
    DG_REQUIRE(width < 128, "execution trace cannot have more than 128 registers");
    DG_REQUIRE(width > 15 + ctx_depth + loop_depth, "user stack must consist of at least one register");
    DG_REQUIRE(n_inputs <= 8 && n_outputs <= 8, "cannot have more than 8 public inputs / outputs");
    const int w = (int)width, log_n = ilog2(n), log_b = ilog2(b), log_N = log_n + log_b;
    DG_REQUIRE(log_N <= 30, "LDE domain too large");
    const uint64_t N = n * b, E = n * 8;
    const int stack_depth = w - 15 - (int)ctx_depth - (int)loop_depth;
    DG_REQUIRE(stack_depth <= 32, "stack depth cannot be greater than 32");
    std::vector<fe> inputs(n_inputs), outputs(n_outputs);
    if (n_inputs) memcpy(inputs.data(), inputs16, n_inputs * 16);
    if (n_outputs) memcpy(outputs.data(), outputs16, n_outputs * 16);

    ArenaScope arena_scope;                   // all DevBufs below come from the per-proof arena (no driver allocation inside a proof)
    StageClock clk(c.stream);
    const unsigned long long launches0 = c.launches;
    Proof *proof = new Proof();
    std::unique_ptr<Proof> guard(proof);

    // ---- sharding: rank g owns the LDE cosets [c0, c0 + nc) of every column (world == 1: all of them) ------------------------------
    const int G = c.world, g = c.rank;
    int log_g = 0;
    while ((1 << log_g) < G) log_g++;
    DG_REQUIRE((b >> log_g) >= 4 && G <= 8, "extension factor too small for this many GPUs (need >= 4 cosets per rank)");
    const int log_nc = log_b - log_g;
    const uint64_t nc = 1ULL << log_nc, N_loc = n * nc;
    const unsigned c0 = (unsigned)(g * nc);

    // ---- 1: extend execution trace ---------------------------------------------------------------------------------------------------
    clk.mark(0);
    DevBuf polys((size_t)w * n * 16), ext((size_t)w * N_loc * 16);
    if (!host_cols) {
        ntt_batch(c, d_regs, polys.as<fe>(), log_n, w, n, n, true);
        lde_batch(c, polys.as<fe>(), ext.as<fe>(), log_n, log_b, 1, w, n, N_loc, c0, (unsigned)nc);
    } else {
        const int chunk = (int)std::max<uint64_t>(1, std::min<uint64_t>(w, ((uint64_t)1 << 25) / (n * 16)));   // ~32 MB per upload
        std::vector<cudaEvent_t> done((w + chunk - 1) / chunk);
        {   // the destination comes from the stream-ordered pool of the compute stream: order the copy stream after it
            cudaEvent_t ready;
            DG_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
            DG_CUDA(cudaEventRecord(ready, c.stream));
            DG_CUDA(cudaStreamWaitEvent(c.copy_stream, ready, 0));
            cudaEventDestroy(ready);
        }
        for (size_t i = 0; i < done.size(); i++) {           // enqueue every upload first: the copy engine runs ahead of the compute stream
            DG_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
            for (int j = (int)i * chunk; j < std::min(w, (int)(i + 1) * chunk); j++)
                DG_CUDA(cudaMemcpyAsync(d_regs + (size_t)j * n, host_cols[j], n * 16, cudaMemcpyHostToDevice, c.copy_stream));
            DG_CUDA(cudaEventRecord(done[i], c.copy_stream));
        }
        for (size_t i = 0; i < done.size(); i++) {
            const int j0 = (int)i * chunk, cols = std::min(w, j0 + chunk) - j0;
            DG_CUDA(cudaStreamWaitEvent(c.stream, done[i], 0));
            ntt_batch(c, d_regs + (size_t)j0 * n, polys.as<fe>() + (size_t)j0 * n, log_n, cols, n, n, true);
            lde_batch(c, polys.as<fe>() + (size_t)j0 * n, ext.as<fe>() + (size_t)j0 * N_loc, log_n, log_b, 1, cols, n, N_loc, c0, (unsigned)nc);
        }
        for (auto &e : done) cudaEventDestroy(e);
    }

    // ---- 2: trace Merkle tree ----------------------------------------------------------------------------------------------------------
    clk.mark(1);
    DevBuf t_leaves(N_loc * 32);
    hash_trace_rows(c, ext.as<fe>(), t_leaves.p, w, log_n, log_nc);          // local rows, [k][c - c0]
    ShardedTree t_tree;
    t_tree.build(c, t_leaves.p, n, log_nc);
    memcpy(proof->trace_root, t_tree.root.data(), 32);

    // ---- 3: evaluate constraints --------------------------------------------------------------------------------------------------------
    clk.mark(2);
    fe last_row[3];     // op_counter and program hash of the last trace step (evaluator.rs:73-74)
    for (int j = 0; j < 3; j++) d2h(c, &last_row[j], d_regs + (size_t)j * n + (n - 1), 16);
    const fe op_count = last_row[0];
    const fe program_hash[2] = {last_row[1], last_row[2]};
    fs::ConstraintCoefficients cc = fs::draw_constraint_coefficients(proof->trace_root, ctx_depth, loop_depth, stack_depth, inputs, outputs,
                                                                      op_count, program_hash);
    static DevBuf d_periodic;
    if (!d_periodic.p) {
        std::vector<fe> per = fs::periodic_tables();
        d_periodic.alloc(per.size() * 16, true);
        h2d(c, d_periodic.p, per.data(), per.size() * 16);
    }
    const size_t T = cc.coefA.size(), nb = cc.bAi.size();
    DevBuf d_coef((2 * T + 4 * nb) * 16), d_violation(4);
    {
        std::vector<fe> pack;
        pack.insert(pack.end(), cc.coefA.begin(), cc.coefA.end());
        pack.insert(pack.end(), cc.coefB.begin(), cc.coefB.end());
        pack.insert(pack.end(), cc.bAi.begin(), cc.bAi.end());
        pack.insert(pack.end(), cc.bBi.begin(), cc.bBi.end());
        pack.insert(pack.end(), cc.bAf.begin(), cc.bAf.end());
        pack.insert(pack.end(), cc.bBf.begin(), cc.bBf.end());
        h2d(c, d_coef.p, pack.data(), pack.size() * 16);
        DG_CUDA(cudaStreamSynchronize(c.stream));
    }
    DG_CUDA(cudaMemsetAsync(d_violation.p, 0, 4, c.stream));
    DevBuf evals(3 * E * 16);                 // [boundary numerator, first step | boundary numerator, last step | transition combination]
    {
        const int num_c8 = 8 >> log_g;
        const uint64_t E_loc = n * num_c8;
        DevBuf evals_loc(E_loc * 16), gathered(E * 16);
        AirParams P;
        memset(&P, 0, sizeof P);
        P.w = w; P.ctx_depth = ctx_depth; P.loop_depth = loop_depth; P.stack_depth = stack_depth;
        P.cl = std::max<int>(ctx_depth, 1); P.ll = std::max<int>(loop_depth, 1); P.sl = std::max(stack_depth, 8);
        P.log_n = log_n; P.log_blowup = log_b;
        P.ext = ext.as<fe>(); P.col_stride = N_loc;
        P.c8_base = g * num_c8; P.num_c8 = num_c8;
        P.t_ev = evals_loc.as<fe>();
        P.periodic = d_periodic.as<fe>();
        const fe *base = d_coef.as<fe>();
        P.coefA = base; P.coefB = base + T;
        P.twN = c.twiddle(log_N, false);
        static const int GROUP_DEG[6] = {2, 3, 4, 6, 7, 8};
        for (int gi = 0; gi < 6; gi++) P.inc[gi] = (8 * n - 1) - (n - 1) * GROUP_DEG[gi];
        P.violation = d_violation.as<unsigned>();
        launch_constraint_eval(c, P);
        comm_all_reduce_max_u32(c, d_violation.as<unsigned>(), 1);
        unsigned violation = 0;
        d2h(c, &violation, d_violation.p, 4);
        if (violation) throw Error(DG_ERR_UNSATISFIED, "transition constraints at step " + std::to_string(violation - 1) + " were not satisfied");
        // every rank interpolates the transition combination: gather the coset slabs, then go to natural step order
        comm_all_gather(c, evals_loc.as<fe>(), gathered.as<fe>(), E_loc * 16);
        transpose_cosets(c, gathered.as<fe>(), evals.as<fe>() + 2 * E, log_n, 3, 1);
        // boundary constraints (evaluator.rs:181-326), directly as the 8n coefficients the reference obtains by interpolation
        boundary_coeffs(c, polys.as<fe>(), n, (int)nb, base + 2 * T, cc.KiA, cc.KiB, cc.KfA, cc.KfB, evals.as<fe>(), evals.as<fe>() + E);
    }
    debug_dump(c, "t_evals", evals.as<fe>() + 2 * E, E * 16);

    // ---- 4: convert constraint evaluations into a polynomial -----------------------------------------------------------------------------
    clk.mark(3);
    const int log_E = log_n + 3;
    DevBuf combined(E * 16), scratch(E * 16), scratch2(E * 16);
    const fe root_n = host_root_of_unity(log_n);
    const fe x_last = host_inv(root_n);                        // w_n^(n-1)   (evaluator.rs:128-131)
    {
        ntt_batch(c, evals.as<fe>() + 2 * E, evals.as<fe>() + 2 * E, log_E, 1, E, E, true);
        debug_dump(c, "i_coeffs", evals.as<fe>(), E * 16);
        debug_dump(c, "f_coeffs", evals.as<fe>() + E, E * 16);
        fe *ic = evals.as<fe>(), *fc = evals.as<fe>() + E, *tc = evals.as<fe>() + 2 * E;
        PowTable one_t(c, fe_make(1, 0), E + 1), xl_t(c, x_last, E + 1), xli_t(c, root_n, E + 1);
        syn_div(c, ic, ic, scratch.as<fe>(), E, one_t.ref(), one_t.ref(), fe_make(0, 0));            // / (x - 1)
        syn_div(c, fc, fc, scratch.as<fe>(), E, xl_t.ref(), xli_t.ref(), fe_make(0, 0));             // / (x - x_last)
        syn_div_expanded_sum(c, tc, scratch.as<fe>(), ic, fc, combined.as<fe>(), n, E, x_last);      // / ((x^n - 1)/(x - x_last)), summed
    }
    debug_dump(c, "constraint_poly", combined.p, E * 16);

    // ---- 5: constraint evaluations over the LDE domain + their Merkle tree -----------------------------------------------------------------
    clk.mark(4);
    DevBuf c_ext(N_loc * 16), c_items((N_loc / 4) * 32);
    lde_batch(c, combined.as<fe>(), c_ext.as<fe>(), log_n, log_b, 8, 1, E, N_loc, c0, (unsigned)nc);
    constraint_items_local(c, c_ext.as<fe>(), log_n, log_nc, c_items.p);      // first tree level: H(4 evaluations), [k][c4 local]
    ShardedTree c_tree;
    c_tree.build(c, c_items.p, n, log_nc - 2);
    memcpy(proof->constraint_root, c_tree.root.data(), 32);

    // ---- 6: DEEP composition polynomial ---------------------------------------------------------------------------------------------------------
    clk.mark(5);
    fs::CompositionCoefficients dc = fs::draw_composition_coefficients(proof->constraint_root, w);
    const fe z = dc.z, zg = fe_mul(z, root_n);
    std::vector<fe> state1(w), state2(w);
    DevBuf comp(E * 16), comp_ext(N * 16);
    {
        PowTable z_t(c, z, E + 1), zi_t(c, host_inv(z), E + 1), zg_t(c, zg, n + 1), zgi_t(c, host_inv(zg), n + 1);
        TwiddleRef g_t = c.twiddle(log_n, false);
        DevBuf d_deep((size_t)(2 * w + 2) * 16);
        eval_polys_at(c, polys.as<fe>(), n, w, z_t.ref(), g_t, true, d_deep.as<fe>());
        eval_polys_at(c, combined.as<fe>(), E, 1, z_t.ref(), g_t, false, d_deep.as<fe>() + 2 * w);
        std::vector<fe> deep(2 * w + 2);
        d2h(c, deep.data(), d_deep.p, deep.size() * 16);
        fe sub1 = fe_make(0, 0), sub2 = fe_make(0, 0);
        for (int i = 0; i < w; i++) {
            state1[i] = deep[2 * i]; state2[i] = deep[2 * i + 1];
            sub1 = fe_add(sub1, fe_mul(state1[i], dc.trace1[i]));
            sub2 = fe_add(sub2, fe_mul(state2[i], dc.trace2[i]));
        }
        const fe c_at_z = deep[2 * w];
        DevBuf d_cc((size_t)2 * w * 16), t12(2 * n * 16);
        h2d(c, d_cc.p, dc.trace1.data(), (size_t)w * 16);
        h2d(c, d_cc.as<fe>() + w, dc.trace2.data(), (size_t)w * 16);
        fe *t1 = t12.as<fe>(), *t2 = t12.as<fe>() + n;
        lincomb2(c, polys.as<fe>(), n, w, d_cc.as<fe>(), d_cc.as<fe>() + w, t1, t2);
        syn_div(c, t1, t1, scratch.as<fe>(), n, z_t.ref(), zi_t.ref(), sub1);                          // (T1(x) - T1(z)) / (x - z)
        syn_div(c, t2, t2, scratch.as<fe>(), n, zg_t.ref(), zgi_t.ref(), sub2);                        // (T2(x) - T2(zg)) / (x - zg)
        syn_div(c, combined.as<fe>(), scratch2.as<fe>(), scratch.as<fe>(), E, z_t.ref(), zi_t.ref(), c_at_z);   // (C(x) - C(z)) / (x - z)
        compose(c, t1, t2, scratch2.as<fe>(), comp.as<fe>(), n, E, 6 * n + 1, dc.t1_degree, dc.t2_degree, dc.constraints);
        debug_dump(c, "composition_poly", comp.p, E * 16);
        if (G == 1) {
            lde_batch(c, comp.as<fe>(), comp_ext.as<fe>(), log_n, log_b, 8, 1, E, N);
        } else {   // extend the own cosets, then every rank gets the whole vector (rank-major == coset-major) to run FRI redundantly
            DevBuf comp_loc(N_loc * 16);
            lde_batch(c, comp.as<fe>(), comp_loc.as<fe>(), log_n, log_b, 8, 1, E, N_loc, c0, (unsigned)nc);
            comm_all_gather(c, comp_loc.p, comp_ext.p, N_loc * 16);
        }
    }

    // ---- 7: FRI layers ---------------------------------------------------------------------------------------------------------------------------
    clk.mark(6);
    std::vector<FriLayerDev> layers;
    {
        TwiddleRef inv_root = c.twiddle(log_N, true);
        const fe tau_inv = host_inv(host_root_of_unity(2));
        const fe inv4 = host_inv(fe_make(4, 0));
        const fe *cur = comp_ext.as<fe>();
        Layout lay{log_N, log_b};
        for (;;) {
            const int log_r = lay.log_d - 2;
            const uint64_t R = 1ULL << log_r;
            const Layout rows{log_r, (lay.log_b >= 0 && log_r >= lay.log_b) ? lay.log_b : -1};
            layers.emplace_back();
            FriLayerDev &L = layers.back();
            L.vals = cur; L.layout = lay;
            L.leaves.alloc(R * 32); L.nodes.alloc(R * 32);
            fri_hash_rows(c, cur, lay, rows, L.leaves.p);
            merkle_build(c, L.leaves.p, L.nodes.p, R);
            d2h(c, L.root.data(), (const uint8_t *)L.nodes.p + 32, 32);
            if (R * 4 <= 256) break;                              // MAX_REMAINDER_LENGTH (fri/mod.rs:13)
            fs::Rng rng(L.root.data());
            const fe alpha = rng.field();                          // special_x = prng(root)  (fri/prover.rs:29)
            L.folded.alloc(R * 16);                                // values of the next layer, owned by this one
            fri_fold(c, cur, lay, L.folded.as<fe>(), rows, alpha, inv_root, log_N, tau_inv, inv4);
            cur = L.folded.as<fe>();
            lay = rows;
        }
    }

    // ---- 8: query positions ------------------------------------------------------------------------------------------------------------------------
    clk.mark(7);
    std::vector<uint64_t> positions;
    {
        std::vector<uint8_t> roots;
        for (auto &L : layers) roots.insert(roots.end(), L.root.begin(), L.root.end());
        uint8_t seed[32];
        debug_dump_host("fri_roots", roots.data(), roots.size());
        fs::blake3_short(roots.data(), roots.size(), seed);
        proof->pow_nonce = pow_search(c, seed, opt.grinding_factor);
        pow_hash(seed, proof->pow_nonce, proof->pow_seed);
        try {
            positions = fs::query_positions(proof->pow_seed, N, b, opt.num_queries);
        } catch (const std::exception &e) { throw Error(DG_ERR_EXHAUSTED, e.what()); }
        debug_dump_host("positions", positions.data(), positions.size() * 8);
    }

    // ---- 9: build proof object -------------------------------------------------------------------------------------------------------------------------
    clk.mark(8);
    fs::ByteWriter out;
    {
        const int nq = (int)positions.size();
        // trace rows at the queried positions (trace_table.rs:127-134): the rank owning the position's coset reads the row
        std::vector<fe> rows((size_t)nq * w);
        {
            std::vector<uint64_t> phys(nq);
            std::vector<int> owners(nq);
            for (int q = 0; q < nq; q++) {
                const uint64_t cpos = positions[q] & (b - 1), k = positions[q] >> log_b;
                owners[q] = (int)(cpos >> log_nc);
                phys[q] = owners[q] == g ? (((cpos - c0) << log_n) + k) : 0;
            }
            DevBuf d_pos(nq * 8), d_rows((size_t)nq * w * 16);
            h2d(c, d_pos.p, phys.data(), nq * 8);
            gather_rows(c, ext.as<fe>(), w, N_loc, d_pos.as<unsigned long long>(), nq, d_rows.as<fe>());
            std::vector<uint8_t> got = exchange_owned(c, d_rows.p, nq, (size_t)w * 16, owners);
            memcpy(rows.data(), got.data(), got.size());
        }
        // trace tree openings: leaves are the row hashes
        fs::BatchPlan tplan = fs::plan_batch_proof(positions, N);
        auto trace_nodes = resolve_plan(tplan, [&](const std::vector<uint64_t> &idx) { return t_tree.fetch_items(c, idx); },
                                        [&](const std::vector<uint64_t> &idx) { return t_tree.fetch_nodes(c, idx); });

        // constraint tree openings: leaf j = evaluations (2j, 2j+1), unhashed (prover.rs:180-187)
        auto constraint_leaves = [&](const std::vector<uint64_t> &idx) {
            std::vector<uint64_t> phys;
            std::vector<int> owners;
            for (uint64_t j : idx) {
                const uint64_t i = 2 * j, cpos = i & (b - 1), k = i >> log_b;
                const int owner = (int)(cpos >> log_nc);
                owners.push_back(owner);
                const uint64_t p0 = owner == g ? (((cpos - c0) << log_n) + k) : 0;
                phys.push_back(p0);
                phys.push_back(owner == g ? p0 + n : 0);            // evaluation 2j+1 lives in the next coset, same k
            }
            std::vector<Digest> o(idx.size());
            if (idx.empty()) return o;
            DevBuf d_idx(phys.size() * 8), d_out(phys.size() * 16);
            h2d(c, d_idx.p, phys.data(), phys.size() * 8);
            gather16(c, c_ext.as<fe>(), d_idx.as<unsigned long long>(), (int)phys.size(), d_out.as<fe>());
            std::vector<uint8_t> got = exchange_owned(c, d_out.p, idx.size(), 32, owners);
            memcpy(o.data(), got.data(), got.size());
            return o;
        };
        auto constraint_nodes = [&](const std::vector<uint64_t> &idx) {
            // heap indices of the tree over N/2 leaves: [N/4, N/2) is the first hashed level (= level-0 items of c_tree)
            std::vector<uint64_t> items, inner;
            std::vector<size_t> ipos, npos;
            for (size_t q = 0; q < idx.size(); q++) {
                if (idx[q] >= N / 4) { items.push_back(idx[q] - N / 4); ipos.push_back(q); }
                else { inner.push_back(idx[q]); npos.push_back(q); }
            }
            std::vector<Digest> o(idx.size());
            std::vector<Digest> a = c_tree.fetch_items(c, items), bb = c_tree.fetch_nodes(c, inner);
            for (size_t i = 0; i < a.size(); i++) o[ipos[i]] = a[i];
            for (size_t i = 0; i < bb.size(); i++) o[npos[i]] = bb[i];
            return o;
        };
        std::vector<uint64_t> c_positions = fs::constraint_positions(positions);
        fs::BatchPlan cplan = fs::plan_batch_proof(c_positions, N / 2);
        std::vector<Digest> c_values = constraint_leaves(cplan.value_leaves);
        auto c_nodes_open = resolve_plan(cplan, constraint_leaves, constraint_nodes);

        // ---- serialise (proof.rs:10-37; bincode: u64 length prefixes, arrays raw, little endian)
        out.raw(proof->trace_root, 32);
        out.u8(tplan.depth); out.u8((uint8_t)ctx_depth); out.u8((uint8_t)loop_depth); out.u8((uint8_t)stack_depth);
        out.u32((uint32_t)op_count.lo);                                               // op_count as u32 (proof.rs:62)
        write_digest_vv(out, trace_nodes);
        out.u64(nq);
        for (int q = 0; q < nq; q++) {
            out.u64(w);
            for (int j = 0; j < w; j++) out.felt(rows[(size_t)q * w + j]);
        }
        out.raw(proof->constraint_root, 32);
        write_digest_vec(out, c_values);
        write_digest_vv(out, c_nodes_open);
        out.u8(cplan.depth);
        write_felt_vec(out, state1);
        write_felt_vec(out, state2);

        // FRI proof (fri/prover.rs:55-95)
        out.u64(layers.size() - 1);
        std::vector<uint64_t> fpos = positions;
        for (size_t d = 0; d + 1 < layers.size(); d++) {
            FriLayerDev &L = layers[d];
            const uint64_t D = 1ULL << L.layout.log_d, R = D / 4;
            fpos = fs::augmented_positions(fpos, D);
            fs::BatchPlan plan = fs::plan_batch_proof(fpos, R);
            std::vector<uint64_t> phys;
            for (uint64_t p : fpos)
                for (int j = 0; j < 4; j++) phys.push_back(L.layout.phys(p + j * R));
            std::vector<fe> vals = fetch16(c, L.vals, phys);
            auto nodes = resolve_plan(plan, [&](const std::vector<uint64_t> &idx) { return fetch32(c, L.leaves.p, idx); },
                                      [&](const std::vector<uint64_t> &idx) { return fetch32(c, L.nodes.p, idx); });
            out.raw(L.root.data(), 32);
            out.u64(fpos.size());
            for (auto &v : vals) out.felt(v);
            write_digest_vv(out, nodes);
            out.u8(plan.depth);
        }
        {
            FriLayerDev &L = layers.back();
            const uint64_t D = 1ULL << L.layout.log_d;
            std::vector<uint64_t> phys(D);
            for (uint64_t i = 0; i < D; i++) phys[i] = L.layout.phys(i);     // remainder in column-major row order == natural order
            std::vector<fe> rem = fetch16(c, L.vals, phys);
            out.raw(L.root.data(), 32);
            write_felt_vec(out, rem);
        }
        out.u64(proof->pow_nonce);
        out.u8((uint8_t)log_b); out.u8((uint8_t)opt.num_queries); out.u8((uint8_t)opt.grinding_factor); out.u8(0);
    }
    clk.mark(9);
    DG_CUDA(cudaStreamSynchronize(c.stream));
    proof->bytes = std::move(out.b);
    if (stats) {
        for (int i = 0; i < 9; i++) stats->stage_ms[i] = clk.between(i, i + 1);
        stats->h2d_ms = 0.0f;                        // host variant: uploads overlap stage 1 and are part of stage_ms[0]
        stats->total_ms = clk.between(0, 9);
        stats->kernel_launches = c.launches - launches0;
    }
    return guard.release();
}

Proof *prove_device(Context &c, const fe *d_regs, uint32_t width, uint64_t length, uint32_t ctx_depth, uint32_t loop_depth,
                    const uint8_t *inputs16, uint32_t n_inputs, const uint8_t *outputs16, uint32_t n_outputs, const dg_options_t &opt,
                    dg_prove_stats_t *stats, float) {
    return prove_core(c, const_cast<fe *>(d_regs), nullptr, width, length, ctx_depth, loop_depth, inputs16, n_inputs, outputs16, n_outputs, opt, stats);
}

Proof *prove_host(Context &c, const dg_trace_t &trace, const uint8_t *inputs16, uint32_t n_inputs, const uint8_t *outputs16,
                  uint32_t n_outputs, const dg_options_t &opt, dg_prove_stats_t *stats) {
    DG_REQUIRE(trace.columns && trace.width >= 16 && trace.width < 128, "invalid trace");
    DG_REQUIRE(trace.length >= 16 && (trace.length & (trace.length - 1)) == 0, "execution trace length must be a power of 2 and at least 16");
    for (uint32_t j = 0; j < trace.width; j++) DG_REQUIRE(trace.columns[j] != nullptr, "null register column");
    c.upload_buf.ensure((size_t)trace.length * 16 * trace.width, true);
    return prove_core(c, c.upload_buf.as<fe>(), trace.columns, trace.width, trace.length, trace.ctx_depth, trace.loop_depth, inputs16, n_inputs,
                      outputs16, n_outputs, opt, stats);
}

}  // namespace dg

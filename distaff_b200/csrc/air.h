// Parameters of the constraint-evaluation kernel (air.cu).
#pragma once
#include "common.cuh"

namespace dg {

struct AirParams {
    int w, ctx_depth, loop_depth, stack_depth;
    int cl, ll, sl;                     // padded stack lengths: max(depth, 1 / 1 / 8)   (trace_state.rs:58-60)
    int log_n, log_blowup;
    int n_boundary_regs;                // registers [0, n_boundary_regs) carry boundary coefficients
    const fe *ext;                      // extended trace slab of this rank: [w][local cosets][n], column stride col_stride
    unsigned long long col_stride;
    int c8_base, num_c8;                // evaluation-domain cosets handled here: c8 in [c8_base, c8_base + num_c8)  (c8 = step mod 8)
    fe *i_ev, *f_ev, *t_ev;             // outputs, coset-major: [c8 - c8_base][k]  (step s = 8k + c8)
    const fe *periodic;                 // [128][23] = sponge ARK (8) | masks (3) | hasher ARK (12), row = step % 128
    const fe *coefA, *coefB;            // per transition constraint (evaluation order): cc[2i], cc[2i+1] of its flattened slot
    const fe *bAi, *bBi, *bAf, *bBf;    // per register boundary coefficients (first step / last step)
    fe KiA, KiB, KfA, KfB;              // sum_j expected_j * coefficient_j
    TwiddleRef twN;                     // powers of the LDE root w_N
    unsigned long long b_adj;           // boundary degree adjustment 6n + 2   (evaluator.rs:408-412)
    unsigned long long inc[6];          // incremental degrees of the groups 2,3,4,6,7,8  (evaluator.rs:395-402)
    unsigned *violation;                // set to step+1 when a trace-domain point violates a transition constraint
};

void launch_constraint_eval(Context &c, const AirParams &P);

}  // namespace dg

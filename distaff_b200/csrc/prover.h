// Host-side orchestration of the prove pipeline (prover.cu) -- the body that replaces
// /root/reference/src/stark/prover.rs:17-169.
#pragma once
#include "../../include/distaff_gpu.h"
#include "common.cuh"

namespace dg {

struct Proof {
    std::vector<uint8_t> bytes;          // bincode encoding of StarkProof
    uint8_t trace_root[32], constraint_root[32], pow_seed[32];
    unsigned long long pow_nonce = 0;
};

Proof *prove_host(Context &c, const dg_trace_t &trace, const uint8_t *inputs16, uint32_t n_inputs, const uint8_t *outputs16,
                  uint32_t n_outputs, const dg_options_t &opt, dg_prove_stats_t *stats);
Proof *prove_device(Context &c, const fe *d_registers, uint32_t width, uint64_t length, uint32_t ctx_depth, uint32_t loop_depth,
                    const uint8_t *inputs16, uint32_t n_inputs, const uint8_t *outputs16, uint32_t n_outputs, const dg_options_t &opt,
                    dg_prove_stats_t *stats, float h2d_ms);

}  // namespace dg

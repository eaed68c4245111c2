#include "prover.h"
namespace dg {
Proof *prove_host(Context &, const dg_trace_t &, const uint8_t *, uint32_t, const uint8_t *, uint32_t, const dg_options_t &, dg_prove_stats_t *) {
    throw Error(-1, "dg_prove: not implemented yet");
}
Proof *prove_device(Context &, const fe *, uint32_t, uint64_t, uint32_t, uint32_t, const uint8_t *, uint32_t, const uint8_t *, uint32_t,
                    const dg_options_t &, dg_prove_stats_t *, float) {
    throw Error(-1, "dg_prove_device: not implemented yet");
}
}  // namespace dg

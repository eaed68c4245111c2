/* Minimal C host for the C-ABI of include/distaff_gpu.h: reads a register trace from a file, proves it on the GPU with dg_prove()
 * and writes the bincode bytes of the StarkProof.  This is the call sequence the Rust shim of INTEGRATION.md performs inside
 * stark::prove (/root/reference/src/stark/prover.rs:17); no Python, no torch.
 *
 *   gcc -O2 -Iinclude examples/prove_trace.c -Ldistaff_b200 -ldistaff_gpu -Wl,-rpath,$PWD/distaff_b200 -o prove_trace
 *   ./prove_trace trace.bin proof.bin [extension_factor num_queries grinding_factor]
 *
 * trace.bin (little endian): u32 width, ctx_depth, loop_depth, n_inputs, n_outputs, reserved; u64 length;
 *                            n_inputs x 16 bytes, n_outputs x 16 bytes, then width columns of length x 16 bytes. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "distaff_gpu.h"

static int fail(const char *what) {
    fprintf(stderr, "%s: %s\n", what, dg_last_error());
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s trace.bin proof.bin [ext queries grinding]\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    uint32_t hdr[6];
    uint64_t length;
    if (fread(hdr, 4, 6, f) != 6 || fread(&length, 8, 1, f) != 1) { fprintf(stderr, "short header\n"); return 2; }
    const uint32_t width = hdr[0], n_in = hdr[3], n_out = hdr[4];
    uint8_t inputs[8 * 16], outputs[8 * 16];
    if (n_in > 8 || n_out > 8 || fread(inputs, 16, n_in, f) != n_in || fread(outputs, 16, n_out, f) != n_out) { fprintf(stderr, "bad inputs\n"); return 2; }
    uint8_t **cols = (uint8_t **)malloc(sizeof(uint8_t *) * width);
    for (uint32_t j = 0; j < width; j++) {
        cols[j] = (uint8_t *)malloc((size_t)length * 16);
        if (fread(cols[j], 16, (size_t)length, f) != length) { fprintf(stderr, "short column %u\n", j); return 2; }
    }
    fclose(f);

    dg_trace_t trace = { (const uint8_t *const *)cols, width, length, hdr[1], hdr[2] };
    dg_options_t opt = { 32, 50, 20, 0 };                 /* ProofOptions::default(), options.rs:82-90 */
    if (argc >= 6) { opt.extension_factor = (uint32_t)atoi(argv[3]); opt.num_queries = (uint32_t)atoi(argv[4]); opt.grinding_factor = (uint32_t)atoi(argv[5]); }
    dg_proof_t *proof = NULL;
    dg_prove_stats_t stats;
    if (dg_prove(&trace, inputs, n_in, outputs, n_out, &opt, &proof, &stats) != DG_OK) return fail("dg_prove");
    size_t len = 0;
    if (dg_proof_serialized_len(proof, &len) != DG_OK) return fail("dg_proof_serialized_len");
    uint8_t *bytes = (uint8_t *)malloc(len);
    if (dg_proof_serialize(proof, bytes, len) != DG_OK) return fail("dg_proof_serialize");
    dg_proof_free(proof);
    f = fopen(argv[2], "wb");
    if (!f || fwrite(bytes, 1, len, f) != len) { perror(argv[2]); return 2; }
    fclose(f);
    printf("proof: %zu bytes, %.3f ms on the device, %llu kernel launches\n", len, stats.total_ms, (unsigned long long)stats.kernel_launches);
    return 0;
}

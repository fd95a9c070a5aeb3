// The reference's example (example/qkd_ldpc_example.cpp: Johnson, "Introducing LDPC codes", example 2.5) on the GPU
// library: the N = 6 code, Alice 0 0 1 0 1 1, Bob 1 0 1 0 1 1, QBER 0.2, all six decoders through the C ABI.
// The reference hard-codes a dense-matrix path; here the matrix is written out in place (4 checks x 6 bits).
//   build: make -C qkd_ldpc_v_b200/host example        run: qkd_ldpc_v_b200/qkdldpc_example
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../include/qkdldpc.h"

int main() {
    // H = 110100 / 011010 / 100011 / 001101 as CSR (check_nodes of the reference's H_matrix)
    const std::vector<int32_t> row_ptr{0, 3, 6, 9, 12}, col_idx{0, 1, 3, 1, 2, 4, 0, 4, 5, 2, 3, 5};
    qkdldpc_code *code = nullptr;
    if (qkdldpc_code_create(&code, 6, 4, 12, row_ptr.data(), col_idx.data(), 0, nullptr) != QKDLDPC_OK) {
        std::fprintf(stderr, "ERROR: %s\n", qkdldpc_last_error());
        return 1;
    }
    const uint32_t alice = 0b110100u, bob = 0b110101u;   // bit i of the word = key bit i
    const double qber = 0.2;
    const char *names[6] = {"SPA", "SPA-LIN-APPROX", "NMSA", "OMSA", "ANMSA", "AOMSA"};
    int status = 0;
    for (int alg = 0; alg < 6; ++alg) {
        qkdldpc_params p{};
        p.algorithm = alg;
        p.max_iterations = 100;
        p.primary = 0.8;
        p.secondary = 0.5;
        p.enable_threshold = 1;
        p.threshold = 100.;
        p.message_precision = 32;
        uint32_t out = 0;
        int32_t iters = 0;
        uint8_t flags = 0;
        if (qkdldpc_decode_batch(code, &p, 1, &alice, &bob, &qber, 1, nullptr, 0, nullptr, 0, &out, &iters, &flags, nullptr) != QKDLDPC_OK) {
            std::fprintf(stderr, "ERROR: %s\n", qkdldpc_last_error());
            status = 1;
            break;
        }
        std::printf("%-15s iterations %d, syndromes %s, keys %s, Bob's corrected key:", names[alg], iters,
                    (flags & QKDLDPC_FLAG_SYNDROMES_MATCH) ? "match" : "DIFFER", (flags & QKDLDPC_FLAG_KEYS_MATCH) ? "match" : "DIFFER");
        for (int i = 0; i < 6; ++i) std::printf(" %u", (out >> i) & 1u);
        std::printf("\n");
    }
    qkdldpc_code_destroy(code);
    return status;
}

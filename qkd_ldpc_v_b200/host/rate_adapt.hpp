// Rate adaptation by puncturing and shortening (https://arxiv.org/abs/1007.1616), untainted puncturing
// (https://arxiv.org/abs/1103.6149) and the privacy-maintenance position lists, as the reference computes them on
// the host (array_and_matrix_operations.cpp:140-256, 975-1223). Everything here runs once per (matrix, combination);
// the per-frame work (LLR classes, extended frames) is done by the GPU library from the position lists.
#pragma once
#include <filesystem>
#include <vector>

#include "matrix.hpp"
#include "xoshiro256pp.hpp"

namespace qkdldpc {

// H_matrix_params, array_and_matrix_operations.hpp:27-57
struct H_matrix_params {
    double delta{}, efficiency{}, punctured_fraction{}, shortened_fraction{}, adapted_code_rate{};
    std::vector<int> punctured_bits{}, shortened_bits{}, bits_to_remove{};
};

// adapt_code_rate, :1129-1223. Returns empty position lists when (QBER, delta, f_EC) is outside the achievable range
// (the combination is then skipped, simulation.cpp:413-415). `warn` receives the reference's warning text, if any.
H_matrix_params adapt_code_rate(Xoshiro256pp &prng, const H_matrix &matrix, double QBER, double delta, double efficiency,
                                bool untainted_puncturing, std::string *warn = nullptr);

// select_punctured_bits_untainted, :1002-1068 (selection order, maximal set).
std::vector<int> select_punctured_bits_untainted(Xoshiro256pp &prng, const H_matrix &matrix);

// get_punctured_bits_untainted, :1076-1123: reads `<matrix>.untp` next to the matrix; if it is missing/empty the list
// is generated and written -- to `cache_dir` when given (the matrix directory may be read-only), else next to the matrix.
std::vector<int> get_punctured_bits_untainted(const fs::path &matrix_path, Xoshiro256pp &prng, const H_matrix &matrix,
                                              const fs::path &cache_dir = {});

std::vector<int> get_bits_positions_to_remove(const H_matrix &matrix);                                   // :140-185
std::vector<int> get_bits_positions_to_remove_rate_adapt(const H_matrix &matrix, const H_matrix_params &p);  // :189-256

// Frame construction of QKD_LDPC_RATE_ADAPT (qkd_ldpc_algorithm.cpp:1136-1174): punctured positions take one random
// bit per party drawn from the trial's generator AFTER key generation, shortened positions are 0, the rest are the
// first n payload bits of the keys (quirk Q12).
void extend_frame(Xoshiro256pp &prng, const H_matrix_params &p, const std::vector<int> &alice, const std::vector<int> &bob,
                  std::vector<int> &alice_ext, std::vector<int> &bob_ext);

}  // namespace qkdldpc

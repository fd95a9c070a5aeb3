// Trial input generation, bit-compatible with the reference's run_trial (simulation.cpp:549-555):
//   fill_random_bits  (array_and_matrix_operations.cpp:889-901)  -- N draws of uniform_int_distribution<int>(0,1)
//   inject_errors     (array_and_matrix_operations.cpp:905-933)  -- exactly floor(N*QBER) flips at the first
//                                                                    positions of a std::shuffle'd index vector
// and the per-trial seeds of QKD_LDPC_batch_simulation (simulation.cpp:713-719).
// The streams are identical to the reference's because the same libstdc++ distributions / std::shuffle are driven
// by the same xoshiro256++ engine (host/xoshiro256pp.hpp); tests/test_host_rng.py pins this against golden keys
// produced by the compiled reference.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "xoshiro256pp.hpp"

namespace qkdldpc {

void fill_random_bits(Xoshiro256pp &prng, std::vector<int> &bit_array);

// Returns the accurate QBER = floor(N*QBER) / N.
double inject_errors(Xoshiro256pp &prng, const std::vector<int> &bit_array, double QBER,
                     std::vector<int> &bit_array_with_errors_out);

std::vector<std::uint64_t> trial_seeds(std::uint64_t simulation_seed, std::size_t trials_number);

// Packs 0/1 ints into the C ABI's frame layout (bit i -> word i>>5, position i&31).
void pack_frame(const std::vector<int> &bits, std::uint32_t *words_out);

}  // namespace qkdldpc

#include "keygen.hpp"

#include <algorithm>
#include <limits>
#include <numeric>
#include <random>

namespace qkdldpc {

void fill_random_bits(Xoshiro256pp &prng, std::vector<int> &bit_array) {
    std::uniform_int_distribution<int> coin(0, 1);
    for (auto &b : bit_array) b = coin(prng);
}

double inject_errors(Xoshiro256pp &prng, const std::vector<int> &bit_array, double QBER,
                     std::vector<int> &bit_array_with_errors_out) {
    const std::size_t n = bit_array.size();
    const std::size_t num_errors = static_cast<std::size_t>(static_cast<double>(n) * QBER);
    bit_array_with_errors_out = bit_array;
    if (num_errors > 0) {
        std::vector<std::size_t> pos(n);
        std::iota(pos.begin(), pos.end(), std::size_t{0});
        std::shuffle(pos.begin(), pos.end(), prng);
        for (std::size_t i = 0; i < num_errors; ++i) bit_array_with_errors_out[pos[i]] ^= 1;
    }
    return static_cast<double>(num_errors) / static_cast<double>(n);
}

std::vector<std::uint64_t> trial_seeds(std::uint64_t simulation_seed, std::size_t trials_number) {
    Xoshiro256pp prng(simulation_seed);
    std::uniform_int_distribution<std::size_t> any(0, std::numeric_limits<std::size_t>::max());
    std::vector<std::uint64_t> seeds(trials_number);
    for (auto &s : seeds) s = any(prng);
    return seeds;
}

void pack_frame(const std::vector<int> &bits, std::uint32_t *words_out) {
    const std::size_t words = (bits.size() + 31) / 32;
    std::fill(words_out, words_out + words, 0u);
    for (std::size_t i = 0; i < bits.size(); ++i)
        if (bits[i]) words_out[i >> 5] |= 1u << (i & 31);
}

}  // namespace qkdldpc

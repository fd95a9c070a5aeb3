// xoshiro256++ engine seeded through SplitMix64 (public-domain algorithm by
// Blackman & Vigna). Written from the published algorithm; it provides the
// part of Reputeless/Xoshiro-cpp v1.1's `XoshiroCpp::Xoshiro256PlusPlus`
// that the reference uses (reference pin: CMakeLists.txt:33-37; call sites
// src/array_and_matrix_operations.cpp:889-933, src/simulation.cpp:549,713-719).
//
// Known answers: state {1,2,3,4} -> 41943041, 58720359,
// 3588806011781223, 3591011842654386; seed 1 -> first output 1847458086238483744.
#pragma once
#include <array>
#include <cstdint>

namespace qkdldpc {

class Xoshiro256pp {
public:
    using result_type = std::uint64_t;
    using state_type = std::array<std::uint64_t, 4>;

    explicit constexpr Xoshiro256pp(std::uint64_t seed = 0x2545F4914F6CDD1DULL) noexcept : s_{} {
        // four successive SplitMix64 outputs fill the state
        std::uint64_t x = seed;
        for (auto &w : s_) {
            x += 0x9e3779b97f4a7c15ULL;
            std::uint64_t z = x;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            w = z ^ (z >> 31);
        }
    }
    explicit constexpr Xoshiro256pp(state_type st) noexcept : s_(st) {}

    constexpr result_type operator()() noexcept {
        const std::uint64_t r = rotl(s_[0] + s_[3], 23) + s_[0];
        const std::uint64_t t = s_[1] << 17;
        s_[2] ^= s_[0];
        s_[3] ^= s_[1];
        s_[1] ^= s_[2];
        s_[0] ^= s_[3];
        s_[2] ^= t;
        s_[3] = rotl(s_[3], 45);
        return r;
    }

    static constexpr result_type min() noexcept { return 0; }
    static constexpr result_type max() noexcept { return ~std::uint64_t{0}; }
    constexpr state_type state() const noexcept { return s_; }

private:
    static constexpr std::uint64_t rotl(std::uint64_t v, int k) noexcept {
        return (v << k) | (v >> (64 - k));
    }
    state_type s_;
};

}  // namespace qkdldpc

// extern "C" view of the host-side helpers, so that the Python tests / bench can drive the same C++ code the
// qkdldpc_sim binary uses (input generation, loaders, rate adaptation).
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "keygen.hpp"

#define HOST_API extern "C" __attribute__((visibility("default")))

// run_trial's key generation for a list of per-trial seeds; packed output [count][words]; returns accurate QBER.
HOST_API double qkdhost_gen_keys(const std::uint64_t *seeds, std::int64_t count, std::int64_t n, double qber,
                                 std::uint32_t *alice_packed, std::uint32_t *bob_packed, int threads) {
    const std::size_t words = static_cast<std::size_t>((n + 31) / 32);
    if (threads < 1) threads = 1;
    std::vector<double> acc(static_cast<std::size_t>(threads), 0.0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        const std::int64_t lo = count * t / threads, hi = count * (t + 1) / threads;
        pool.emplace_back([=, &acc] {
            std::vector<int> a(static_cast<std::size_t>(n)), b;
            for (std::int64_t i = lo; i < hi; ++i) {
                qkdldpc::Xoshiro256pp prng(seeds[i]);
                qkdldpc::fill_random_bits(prng, a);
                acc[static_cast<std::size_t>(t)] = qkdldpc::inject_errors(prng, a, qber, b);
                qkdldpc::pack_frame(a, alice_packed + static_cast<std::size_t>(i) * words);
                qkdldpc::pack_frame(b, bob_packed + static_cast<std::size_t>(i) * words);
            }
        });
    }
    for (auto &th : pool) th.join();
    return static_cast<double>(static_cast<std::size_t>(static_cast<double>(n) * qber)) / static_cast<double>(n);
}

HOST_API void qkdhost_trial_seeds(std::uint64_t simulation_seed, std::int64_t count, std::uint64_t *seeds_out) {
    auto s = qkdldpc::trial_seeds(simulation_seed, static_cast<std::size_t>(count));
    std::memcpy(seeds_out, s.data(), s.size() * sizeof(std::uint64_t));
}

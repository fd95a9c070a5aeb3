// Synthetic key generator (device RNG). Included by api.cu only.
#pragma once
#include "common.cuh"

namespace qk {

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------------------------
// Synthetic keys with run_trial's distribution (simulation.cpp:549-555): Alice iid uniform, Bob = Alice with
// exactly `n_err` = floor(n*qber) flips at uniformly random distinct positions (rejection of repeats is uniform).
__device__ __forceinline__ u64 mix64(u64 z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(128) gen_keys_kernel(int n, int words, int n_err, u64 seed, uint32_t *alice,
                                                       uint32_t *bob) {
    extern __shared__ uint32_t s_err[];   // [words]
    const long long f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const u64 fkey = mix64(seed ^ mix64((u64)f));
    for (int w = tid; w < words; w += blockDim.x) {
        uint32_t r = (uint32_t)(mix64(fkey + (u64)w) >> 32);
        const int rem = n - w * 32;
        if (rem < 32) r &= (rem <= 0) ? 0u : ((1u << rem) - 1u);
        alice[f * words + w] = r;
        s_err[w] = 0;
    }
    __syncthreads();
    if (tid < 32) {
        int cnt = 0;
        u64 ctr = 0;
        while (cnt < n_err) {
            const int want = min(32, n_err - cnt);
            bool fresh = false;
            if (lane < want) {
                const u64 r = mix64(fkey ^ mix64(0x5bd1e995ull + ctr * 32 + lane));
                const uint32_t pos = (uint32_t)(((r >> 32) * (u64)n) >> 32);
                const uint32_t bit = 1u << (pos & 31);
                fresh = !(atomicOr(&s_err[pos >> 5], bit) & bit);
            }
            cnt += __popc(__ballot_sync(0xffffffffu, fresh));
            ++ctr;
        }
    }
    __syncthreads();
    for (int w = tid; w < words; w += blockDim.x) bob[f * words + w] = alice[f * words + w] ^ s_err[w];
}

}  // namespace qk

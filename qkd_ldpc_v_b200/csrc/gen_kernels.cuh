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

// ---------------------------------------------------------------------------------------------------------------
// Reference-compatible trial inputs ON THE DEVICE: the frames this kernel writes are bit-identical to what the
// reference's run_trial builds on the CPU for the same per-trial seed (simulation.cpp:549-555, 743):
//   prng = xoshiro256++ seeded through SplitMix64 with seeds[n] + combination index (XoshiroCpp v1.1);
//   fill_random_bits   (array_and_matrix_operations.cpp:889-901): N draws, bit = draw >> 63 (libstdc++'s
//                      uniform_int_distribution<int>(0,1) on a 64-bit engine: Lemire's method with range 2);
//   inject_errors      (:905-933): std::shuffle of 0..N-1, flip the bits at the first floor(N*QBER) entries;
//   QKD_LDPC_RATE_ADAPT (qkd_ldpc_algorithm.cpp:1148-1174): punctured positions (ascending) take one more draw per
//                      party, shortened positions are 0, the others take the first payload bits in order.
// libstdc++ 13's std::shuffle (bits/stl_algo.h:3742-3805) walks i = 1..N-1 swapping a[i] with a[j], j uniform in
// [0, i], two positions per engine draw: x = uniform(0, (i+1)(i+2)-1) by Lemire's nearly-divisionless method
// (bits/uniform_int_dist.h:257-281, 128-bit product, rejection below (2^64 - R) % R), j_i = x / (i+2),
// j_{i+1} = x % (i+2); an even N first spends one draw on a[1] <-> a[draw >> 63].
// Only the first K = floor(N*QBER) entries of the shuffled array are used, and they can be had WITHOUT the array:
// before step i, a[i] still holds i (earlier steps only touched indices < i), so step i writes the value i into
// a[j]; a step with i >= K therefore matters only when j < K (prefix[j] = i), whatever it moves out of the prefix
// never comes back. Steps i < K are real swaps inside the prefix. One thread per frame, sequential like the CPU code.
struct Xoshiro256pp {
    u64 s0, s1, s2, s3;
    __device__ __forceinline__ explicit Xoshiro256pp(u64 seed) {
        u64 x = seed;
        auto next = [&x]() {
            x += 0x9e3779b97f4a7c15ull;
            u64 z = x;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            return z ^ (z >> 31);
        };
        s0 = next(); s1 = next(); s2 = next(); s3 = next();
    }
    static __device__ __forceinline__ u64 rotl(u64 v, int k) { return (v << k) | (v >> (64 - k)); }
    __device__ __forceinline__ u64 operator()() {
        const u64 r = rotl(s0 + s3, 23) + s0;
        const u64 t = s1 << 17;
        s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t;
        s3 = rotl(s3, 45);
        return r;
    }
    // std::uniform_int_distribution<uint64>{0, range - 1}: Lemire's nearly-divisionless method, as libstdc++ does it
    __device__ __forceinline__ u64 below(u64 range) {
        u64 d = (*this)();
        u64 lo = d * range;
        if (lo < range) {
            const u64 threshold = (0ull - range) % range;
            while (lo < threshold) {
                d = (*this)();
                lo = d * range;
            }
        }
        return __umul64hi(d, range);
    }
};

// One parameter combination of a sweep: frames [c * trials, (c+1) * trials) of a launch belong to combination c and use
// trial seeds seeds[0 .. trials) + seed_offset (the reference's seeds[n] + curr_sim, simulation.cpp:743).
struct RefKeygenCombo {
    u64 seed_offset;
    int n_err;                 // floor(n * QBER)
    int rate_adapt;            // punctured / shortened masks present
};

struct RefKeygenArgs {
    int n, words;
    long long n_frames, first_frame;   // frames of this launch; global index of its first frame
    long long trials;          // frames per combination
    const u64 *seeds;          // [trials] per-trial seeds (simulation.cpp:713-719)
    const RefKeygenCombo *combos;
    uint32_t *alice, *bob;     // [n_frames][words] output frames of this launch (extended frames with rate adaptation)
    const uint32_t *masks;     // [combo][2][words] packed punctured / shortened positions (shortened excludes punctured)
    int any_rate_adapt;        // the scratch holds raw-key words (some combination is rate-adapted)
    int max_err;               // prefix entries per thread in the scratch
    uint32_t *scratch;         // thread-interleaved: [(2*words if any_rate_adapt) + max_err][threads of the launch]
    long long scratch_stride;  // threads of the launch
};

__global__ void __launch_bounds__(128) ref_keygen_kernel(const RefKeygenArgs a) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.n_frames) return;
    const long long gf = a.first_frame + f, combo = gf / a.trials;
    const RefKeygenCombo cb = a.combos[combo];
    const int rate_adapt = cb.rate_adapt;
    const long long T = a.scratch_stride;
    uint32_t *raw_a = a.scratch + f;                                  // raw_a[w * T]  (rate adaptation only)
    uint32_t *raw_b = raw_a + (a.any_rate_adapt ? (long long)a.words * T : 0);
    uint32_t *prefix = a.scratch + (a.any_rate_adapt ? 2ll * a.words * T : 0) + f;   // prefix[j * T]
    uint32_t *out_a = a.alice + f * a.words, *out_b = a.bob + f * a.words;
    Xoshiro256pp g(a.seeds[gf % a.trials] + cb.seed_offset);

    // fill_random_bits
    for (int w = 0; w < a.words; ++w) {
        const int cnt = min(32, a.n - w * 32);
        uint32_t word = 0;
        for (int b = 0; b < cnt; ++b) word |= (uint32_t)(g() >> 63) << b;
        if (rate_adapt) { raw_a[(long long)w * T] = word; raw_b[(long long)w * T] = word; }
        else { out_a[w] = word; out_b[w] = word; }
    }

    // inject_errors: the first K entries of std::shuffle(0..N-1), without the array
    const int K = cb.n_err;
    if (K > 0) {
        for (int j = 0; j < K; ++j) prefix[(long long)j * T] = (uint32_t)j;
        auto step = [&](uint32_t i, uint32_t j) {   // std::iter_swap(first + i, first + j), j <= i
            if (i < (uint32_t)K) {
                const uint32_t vi = prefix[(long long)i * T], vj = prefix[(long long)j * T];
                prefix[(long long)i * T] = vj;
                prefix[(long long)j * T] = vi;
            } else if (j < (uint32_t)K) {
                prefix[(long long)j * T] = i;
            }
        };
        uint32_t i = 1;
        const uint32_t N = (uint32_t)a.n;
        if (N > 1 && (N % 2) == 0) {
            step(1, (uint32_t)(g() >> 63));
            i = 2;
        }
        for (; i < N; i += 2) {
            const u64 b1 = (u64)i + 2;                       // __swap_range + 1
            const u64 x = g.below(((u64)i + 1) * b1);
            uint32_t p0, p1;
            if (b1 * b1 <= 0xFFFFFFFFull) {                  // x < (i+1)(i+2) fits 32 bits: cheap division
                p0 = (uint32_t)x / (uint32_t)b1;
                p1 = (uint32_t)x % (uint32_t)b1;
            } else {
                p0 = (uint32_t)(x / b1);
                p1 = (uint32_t)(x % b1);
            }
            step(i, p0);
            step(i + 1, p1);
        }
        for (int j = 0; j < K; ++j) {
            const uint32_t p = prefix[(long long)j * T];
            if (rate_adapt) raw_b[(long long)(p >> 5) * T] ^= 1u << (p & 31);
            else out_b[p >> 5] ^= 1u << (p & 31);
        }
    }
    if (!rate_adapt) return;
    const uint32_t *punct_mask = a.masks + combo * 2 * a.words, *short_mask = punct_mask + a.words;

    // QKD_LDPC_RATE_ADAPT frame construction
    int k = 0;   // payload index n of the reference loop
    for (int w = 0; w < a.words; ++w) {
        const uint32_t pm = punct_mask[w], sm = short_mask[w];
        const int cnt = min(32, a.n - w * 32);
        uint32_t wa = 0, wb = 0;
        for (int b = 0; b < cnt; ++b) {
            if ((pm >> b) & 1u) {
                wa |= (uint32_t)(g() >> 63) << b;
                wb |= (uint32_t)(g() >> 63) << b;
            } else if (!((sm >> b) & 1u)) {
                wa |= ((raw_a[(long long)(k >> 5) * T] >> (k & 31)) & 1u) << b;
                wb |= ((raw_b[(long long)(k >> 5) * T] >> (k & 31)) & 1u) << b;
                ++k;
            }
        }
        out_a[w] = wa;
        out_b[w] = wb;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// remove_bits (array_and_matrix_operations.cpp:259-287): the step after the decoder in the protocol -- privacy
// maintenance / removal of the punctured and shortened positions leaves the final key. `kept` lists the surviving
// positions in ascending order; one warp packs 32 of them per ballot. grid = (frames, ceil(words_out / warps per CTA)).
__global__ void __launch_bounds__(256) remove_bits_kernel(int words_in, int n_keep, int words_out, const int *kept, const uint32_t *in,
                                                          uint32_t *out) {
    const long long f = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= words_out) return;
    const int k = w * 32 + lane;
    uint32_t bit = 0;
    if (k < n_keep) {
        const int p = __ldg(kept + k);
        bit = (in[f * words_in + (p >> 5)] >> (p & 31)) & 1u;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, bit != 0);
    if (lane == 0) out[f * words_out + w] = word;
}

// The same for the frames of a multi-combination launch (frame f = combination f / trials, trial f % trials): every
// combination has its own list of surviving positions (kept[kept_off .. + n_keep)) and its own output region.
struct RemoveCombo {
    int kept_off, n_keep, words_out, pad;
    long long out_off;   // first output word of the combination
};
__global__ void __launch_bounds__(256) remove_bits_multi_kernel(int words_in, long long trials, const int *kept, const RemoveCombo *info,
                                                                const uint32_t *in, uint32_t *out) {
    const long long f = blockIdx.x;
    const RemoveCombo rc = info[f / trials];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= rc.words_out) return;       // also: combinations without a removal list (words_out == 0)
    const int k = w * 32 + lane;
    uint32_t bit = 0;
    if (k < rc.n_keep) {
        const int p = __ldg(kept + rc.kept_off + k);
        bit = (in[f * words_in + (p >> 5)] >> (p & 31)) & 1u;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, bit != 0);
    if (lane == 0) out[rc.out_off + (f % trials) * rc.words_out + w] = word;
}

}  // namespace qk

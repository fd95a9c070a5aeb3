// Frame preparation (K1), slot scheduler = early stop / retire / refill (K4) with on-device tallies (K5 input),
// and the synthetic key generator.
#pragma once
#include "common.cuh"

namespace qk {

typedef unsigned long long u64;

// Per-batch state (all device pointers).
template <typename T>
struct BatchArgs {
    long long n_frames;
    int words;        // packed words per frame  = ceil(n / 32)
    int swords;       // packed syndrome words   = ceil(m / 32)
    const uint32_t *alice_bits, *bob_bits;
    const double *qber;
    int qber_is_scalar;
    T *frame_llr;              // [n_frames]
    uint32_t *synd_all;        // [n_frames][swords]  Alice's syndrome, packed
    uint32_t *out_bits;
    int32_t *out_iters;
    uint8_t *out_flags;
    u64 *tally;
    u64 *next_frame;           // work queue head
    u64 *n_done;               // frames finished so far
    long long *slot_frame;     // [tiles*FT] frame in the slot, -1 = idle
    int32_t *slot_iter;        // [tiles*FT] completed iterations of that frame; -1 = refilled, VN-side init pending
    int max_iter;
    int adaptive;
};

__device__ __forceinline__ void tally_frame(u64 *tally, bool syn_ok, bool keys_ok, int iters_reported, int iters_run) {
    if (!tally) return;
    atomicAdd(tally + 0, 1ull);
    if (syn_ok) {
        atomicAdd(tally + 1, 1ull);
        if (keys_ok) atomicAdd(tally + 2, 1ull);   // keys counted only when syndromes match (simulation.cpp:596-605)
        atomicAdd(tally + 4 + iters_reported, 1ull);
    }
    atomicAdd(tally + 3, (u64)iters_run);
}

// K1: per frame -- LLR magnitude log((1-q)/q) (qkd_ldpc_algorithm.cpp:1043) and Alice's syndrome
// (calculate_syndrome, array_and_matrix_operations.cpp:936-950), packed 32 checks per word.
template <typename T>
__global__ void __launch_bounds__(256) prep_kernel(int n, int m, const int *row_ptr, const int *col_idx, BatchArgs<T> b) {
    const long long f = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t *al = b.alice_bits + f * b.words;
    if (threadIdx.x == 0) {
        const double q = b.qber_is_scalar ? b.qber[0] : b.qber[f];
        b.frame_llr[f] = (T)log((1. - q) / q);
    }
    for (int jb = warp; jb * 32 < m; jb += nwarps) {
        const int j = jb * 32 + lane;
        uint32_t s = 0;
        if (j < m) {
            const int e1 = row_ptr[j + 1];
            for (int e = row_ptr[j]; e < e1; ++e) {
                const int c = col_idx[e];
                s ^= (al[c >> 5] >> (c & 31)) & 1u;
            }
        }
        const uint32_t sw = __ballot_sync(0xffffffffu, s);
        if (lane == 0) b.synd_all[f * b.swords + jb] = sw;
    }
}

// K4: one CTA per tile, after every CN+VN step.
//  1. all-checks-satisfied test per slot: OR over rows of par (quirk Q9: non-adaptive variants test the decision
//     of iteration t and return t; quirk Q10: adaptive variants test the INITIAL decision too, find success one
//     iteration later (return t+1) and never test the decision of the last allowed iteration);
//  2. retire finished frames: unpack the hard decision into out_bits, compare with Alice's key (arrays_equal,
//     qkd_ldpc_algorithm.cpp:1087), write iterations / flags, update the tallies;
//  3. refill free slots from the frame queue (continuous batching): transpose the new frame's key bits, syndrome
//     and initial check values into the tile's bit masks, and publish the tile's active / new masks.
template <typename T, int V>
__global__ void __launch_bounds__(kSchedThreads) sched_kernel(StepArgs<T> a, BatchArgs<T> b) {
    constexpr int FT = kWarp * V;
    __shared__ uint32_t s_unsat[V], s_act[V], s_newm[V];
    __shared__ int s_ret[FT], s_nret;        // retiring slots: slot | success << 8 | iterations << 9
    __shared__ int s_newslot[FT], s_nnew;
    __shared__ long long s_newframe[FT];
    __shared__ int s_wfree[FT / 32 + 1];
    __shared__ u64 s_base;
    __shared__ int s_run[FT];

    const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    bool any_act = false;
#pragma unroll
    for (int v = 0; v < V; ++v) any_act |= a.tile_active[tile * V + v] != 0;
    if (tid < V) { s_unsat[tid] = 0; s_act[tid] = 0; s_newm[tid] = 0; }
    if (tid == 0) { s_nret = 0; s_nnew = 0; }
    __syncthreads();

    if (any_act) {
        uint32_t acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = 0;
        const uint32_t *par = a.par + (size_t)tile * a.m * V;
        for (int j = tid; j < a.m; j += blockDim.x) {
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] |= par[(size_t)j * V + v];
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t r = __reduce_or_sync(0xffffffffu, acc[v]);
            if (lane == 0 && r) atomicOr(&s_unsat[v], r);
        }
    }
    __syncthreads();

    // per-slot decision; slot position s <-> mask word s % V, bit s / V
    long long frame = -1;
    bool done = false;
    const int sv = tid % V, sl = tid / V;
    if (tid < FT) {
        frame = b.slot_frame[(size_t)tile * FT + tid];
        if (frame >= 0) {
            // iterations completed, this step included; 0 = this step only initialised the slot (VN-side init)
            const int it = b.slot_iter[(size_t)tile * FT + tid] + 1;
            const bool ok = !((s_unsat[sv] >> sl) & 1u);
            int succ = 0, iters = 0;
            if (!b.adaptive) {
                if (it >= 1 && ok) { done = true; succ = 1; iters = it; }
                else if (it >= b.max_iter) { done = true; iters = b.max_iter; }
            } else {
                if (ok && it < b.max_iter) { done = true; succ = 1; iters = it + 1; }
                else if (it >= b.max_iter) { done = true; iters = b.max_iter; }
            }
            if (done) {
                const int idx = atomicAdd(&s_nret, 1);
                s_ret[idx] = tid | (succ << 8) | (iters << 9);
                s_run[idx] = it;
            } else {
                b.slot_iter[(size_t)tile * FT + tid] = it;
            }
        }
    }
    __syncthreads();

    const int nret = s_nret;
    for (int r = 0; r < nret; ++r) {
        const int s = s_ret[r] & 255, succ = (s_ret[r] >> 8) & 1, iters = s_ret[r] >> 9;
        const long long fr = b.slot_frame[(size_t)tile * FT + s];
        const int v = s % V, l = s / V;
        const uint32_t *zm = a.zmask + (size_t)tile * a.n * V + v;
        uint32_t diff = 0;
        for (int w = tid; w < b.words; w += blockDim.x) {
            uint32_t word = 0;
            const int i0 = w * 32;
#pragma unroll 8
            for (int k = 0; k < 32; ++k)
                if (i0 + k < a.n) word |= ((zm[(size_t)(i0 + k) * V] >> l) & 1u) << k;
            if (b.out_bits) b.out_bits[fr * b.words + w] = word;
            diff |= word ^ b.alice_bits[fr * b.words + w];
        }
        const int keys_differ = __syncthreads_or(diff != 0);
        if (tid == 0) {
            if (b.out_iters) b.out_iters[fr] = iters;
            if (b.out_flags) b.out_flags[fr] = (uint8_t)((succ ? 1u : 0u) | (keys_differ ? 0u : 2u));
            tally_frame(b.tally, succ, !keys_differ, iters, s_run[r]);
        }
    }

    // refill: every free slot claims the next frame of the queue
    const bool is_free = tid < FT && (frame < 0 || done);
    const uint32_t fm = __ballot_sync(0xffffffffu, is_free);
    if (tid < FT && lane == 0) s_wfree[tid >> 5] = __popc(fm);
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int w = 0; w < FT / 32; ++w) { const int c = s_wfree[w]; s_wfree[w] = total; total += c; }
        u64 base = (u64)b.n_frames;
        if (total > 0 && *(volatile u64 *)b.next_frame < (u64)b.n_frames) base = atomicAdd(b.next_frame, (u64)total);
        s_base = base;
        if (nret) atomicAdd(b.n_done, (u64)nret);
    }
    __syncthreads();
    if (tid < FT) {
        long long nf = frame;
        if (is_free) {
            const u64 cand = s_base + (u64)(s_wfree[tid >> 5] + __popc(fm & ((1u << lane) - 1u)));
            nf = (cand < (u64)b.n_frames) ? (long long)cand : -1;
            b.slot_frame[(size_t)tile * FT + tid] = nf;
            if (nf >= 0) {
                b.slot_iter[(size_t)tile * FT + tid] = -1;
                a.slot_llr[(size_t)tile * FT + tid] = b.frame_llr[nf];
                const int idx = atomicAdd(&s_nnew, 1);
                s_newslot[idx] = tid;
                s_newframe[idx] = nf;
                atomicOr(&s_newm[sv], 1u << sl);
            }
        }
        if (nf >= 0) atomicOr(&s_act[sv], 1u << sl);
    }
    __syncthreads();
    if (tid < V) {
        a.tile_active[tile * V + tid] = s_act[tid];
        a.tile_new[tile * V + tid] = s_newm[tid];
    }
    const int nnew = s_nnew;
    if (nnew == 0) return;

    uint32_t *bm = a.bobmask + (size_t)tile * a.n * V;
    for (int i = tid; i < a.n; i += blockDim.x) {
        uint32_t mk[V];
#pragma unroll
        for (int v = 0; v < V; ++v) mk[v] = bm[(size_t)i * V + v];
        for (int q = 0; q < nnew; ++q) {
            const int s = s_newslot[q];
            const uint32_t bit = (b.bob_bits[s_newframe[q] * b.words + (i >> 5)] >> (i & 31)) & 1u;
            const int l = s / V;
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v == s % V) mk[v] = (mk[v] & ~(1u << l)) | (bit << l);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) bm[(size_t)i * V + v] = mk[v];
    }
    uint32_t *sy = a.synd + (size_t)tile * a.m * V;
    uint32_t *pa = a.par + (size_t)tile * a.m * V;
    for (int j = tid; j < a.m; j += blockDim.x) {
        uint32_t ms[V], mp[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { ms[v] = sy[(size_t)j * V + v]; mp[v] = pa[(size_t)j * V + v]; }
        for (int q = 0; q < nnew; ++q) {
            const int s = s_newslot[q];
            const size_t off = s_newframe[q] * b.swords + (j >> 5);
            const uint32_t sbit = (b.synd_all[off] >> (j & 31)) & 1u;
            const uint32_t pbit = sbit;   // the VN-side init XORs the parity of the initial decision on top
            const int l = s / V;
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v == s % V) {
                    ms[v] = (ms[v] & ~(1u << l)) | (sbit << l);
                    mp[v] = (mp[v] & ~(1u << l)) | (pbit << l);
                }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) { sy[(size_t)j * V + v] = ms[v]; pa[(size_t)j * V + v] = mp[v]; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tail compaction. Once the frame queue is empty the pool drains: every tile keeps running until its slowest frame is
// done, and at a converging operating point that is a handful of stragglers (FER ~ 1 %: a third of the tiles hold
// one frame that runs to max_iterations while the other 31 or 127 lanes idle). The host calls these kernels between two
// steps when at most half of the slots are still occupied: occupied slots of the HIGH tiles move into free slots of the
// first ceil(active / FT) tiles (sources and destinations are disjoint, so the move is in place and fully parallel)
// and the step kernels are launched on the shrunk tile count from then on. A frame's state is its E messages, one bit
// in each of the 2N + 2M mask words of its tile, and three slot words; results do not depend on where a frame sits.
struct CompactPlan {
    int n_moves;
    int new_tiles;
};

// One thread walks the slot table (a few thousand entries, a few times per batch).
template <int FT>
__global__ void compact_plan_kernel(int tiles, const long long *slot_frame, int2 *moves, CompactPlan *plan) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int slots = tiles * FT;
    int active = 0;
    for (int s = 0; s < slots; ++s) active += slot_frame[s] >= 0;
    const int new_tiles = max(1, (active + FT - 1) / FT);
    int n = 0, dst = 0;
    const int low = new_tiles * FT;
    for (int src = low; src < slots; ++src) {
        if (slot_frame[src] < 0) continue;
        while (dst < low && slot_frame[dst] >= 0) ++dst;   // next free slot of the low tiles (exists: active <= low)
        moves[n++] = make_int2(src, dst++);
    }
    plan->n_moves = n;
    plan->new_tiles = new_tiles;
}

// grid = (moves, chunks): messages of the moved frames. Slot position p of a tile <-> message lane p (common.cuh).
template <typename T, int FT>
__global__ void __launch_bounds__(256) compact_msg_kernel(const int2 *moves, const CompactPlan *plan, T *msg, long long e_stride, int nnz) {
    if ((int)blockIdx.x >= plan->n_moves) return;
    const int2 mv = moves[blockIdx.x];
    const T *src = msg + (long long)(mv.x / FT) * e_stride + (mv.x % FT);
    T *dst = msg + (long long)(mv.y / FT) * e_stride + (mv.y % FT);
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < nnz; e += gridDim.y * blockDim.x) dst[(long long)e * FT] = src[(long long)e * FT];
}

// grid = (moves, chunks): one bit per mask word. Slot p <-> word v = p % V, bit l = p / V (sched_kernel).
template <int V>
__global__ void __launch_bounds__(256) compact_mask_kernel(const int2 *moves, const CompactPlan *plan, int n, int m, uint32_t *bobmask,
                                                           uint32_t *zmask, uint32_t *synd, uint32_t *par) {
    constexpr int FT = 32 * V;
    if ((int)blockIdx.x >= plan->n_moves) return;
    const int2 mv = moves[blockIdx.x];
    const int st = mv.x / FT, sp = mv.x % FT, dt = mv.y / FT, dp = mv.y % FT;
    const int sv = sp % V, sl = sp / V, dv = dp % V, dl = dp / V;
    auto move_bit = [&](uint32_t *arr, int rows, int r) {
        const uint32_t bit = (arr[((size_t)st * rows + r) * V + sv] >> sl) & 1u;
        uint32_t *w = arr + ((size_t)dt * rows + r) * V + dv;
        if (bit) atomicOr(w, 1u << dl);
        else atomicAnd(w, ~(1u << dl));
    };
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n + m; i += gridDim.y * blockDim.x) {
        if (i < n) {
            move_bit(bobmask, n, i);
            move_bit(zmask, n, i);
        } else {
            move_bit(synd, m, i - n);
            move_bit(par, m, i - n);
        }
    }
}

// Slot words of the moved frames, then the active / new masks of every tile of the old pool.
template <typename T, int V>
__global__ void compact_finish_kernel(const int2 *moves, const CompactPlan *plan, int old_tiles, long long *slot_frame, int32_t *slot_iter,
                                      T *slot_llr, uint32_t *tile_active, uint32_t *tile_new) {
    constexpr int FT = 32 * V;
    const int nm = plan->n_moves;
    for (int k = threadIdx.x; k < nm; k += blockDim.x) {   // single CTA: moves are disjoint
        const int2 mv = moves[k];
        slot_frame[mv.y] = slot_frame[mv.x];
        slot_iter[mv.y] = slot_iter[mv.x];
        slot_llr[mv.y] = slot_llr[mv.x];
        slot_frame[mv.x] = -1;
    }
    __syncthreads();
    for (int w = threadIdx.x; w < old_tiles * V; w += blockDim.x) {
        const int tile = w / V, v = w % V;
        uint32_t act = 0, nw = 0;
        for (int l = 0; l < 32; ++l) {
            const int p = tile * FT + l * V + v;
            if (slot_frame[p] >= 0) {
                act |= 1u << l;
                if (slot_iter[p] == -1) nw |= 1u << l;
            }
        }
        tile_active[w] = act;
        tile_new[w] = nw;
    }
}

}  // namespace qk

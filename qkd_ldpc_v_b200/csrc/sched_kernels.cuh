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

// K4, part 1 (`sched_kernel`, one CTA per tile, after every CN+VN step):
//  1. all-checks-satisfied test per slot: OR over rows of par (quirk Q9: non-adaptive variants test the decision
//     of iteration t and return t; quirk Q10: adaptive variants test the INITIAL decision too, find success one
//     iteration later (return t+1) and never test the decision of the last allowed iteration);
//  2. decide which frames retire, hand every free slot the next frame of the queue (continuous batching), update the
//     slot table and the tile's active / new masks, and leave the two work lists of the tile in global memory.
// K4, part 2 (`sched_move_kernel`, several CTAs per tile, each owning a range of bit / check rows): the data movement,
//     which for long codes is most of the scheduler's work and must not be serialised on one CTA per tile:
//     retire -- unpack the hard decision into out_bits, compare with Alice's key (arrays_equal,
//     qkd_ldpc_algorithm.cpp:1087); the last CTA of the tile to finish writes iterations / flags / tallies;
//     refill -- transpose the new frames' key bits, syndromes and initial check values into the tile's bit masks.
template <int FT>
struct TileWork {
    int nret, nnew;
    int ticket;                 // CTAs of sched_move_kernel that have finished this tile (reset by the last one)
    int ret[FT];                // retiring slots: slot | success << 8 | iterations << 9
    int run[FT];                // decoder iterations those frames actually ran
    long long ret_frame[FT];
    unsigned int ret_diff[FT];  // != 0: decoded key differs from Alice's (OR over the CTAs of the tile)
    int newslot[FT];
    long long newframe[FT];
};

template <typename T, int V>
__global__ void __launch_bounds__(kSchedThreads) sched_kernel(StepArgs<T> a, BatchArgs<T> b, TileWork<32 * V> *work) {
    constexpr int FT = kWarp * V;
    __shared__ uint32_t s_unsat[V], s_act[V], s_newm[V];
    __shared__ int s_nret, s_nnew;
    __shared__ int s_wfree[FT / 32 + 1];
    __shared__ u64 s_base;

    const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    TileWork<FT> &tw = work[tile];
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
        // one CTA reads the tile's m x V parity words: on the long codes (m = 52301: 837 KB) this is latency-bound unless
        // many loads are in flight, so each thread issues 8 independent 4V-byte loads per round
        const Vec<uint32_t, V> *par = reinterpret_cast<const Vec<uint32_t, V> *>(a.par + (size_t)tile * a.m * V);
        const int stride = blockDim.x;
        int j = tid;
        for (; j + 7 * stride < a.m; j += 8 * stride) {
            Vec<uint32_t, V> x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = par[j + u * stride];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] |= x[u].v[v];
        }
        for (; j < a.m; j += stride) {
            const Vec<uint32_t, V> x = par[j];
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] |= x.v[v];
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
                tw.ret[idx] = tid | (succ << 8) | (iters << 9);
                tw.run[idx] = it;
                tw.ret_frame[idx] = frame;
                tw.ret_diff[idx] = 0u;
            } else {
                b.slot_iter[(size_t)tile * FT + tid] = it;
            }
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
        if (s_nret) atomicAdd(b.n_done, (u64)s_nret);
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
                tw.newslot[idx] = tid;
                tw.newframe[idx] = nf;
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
    if (tid == 0) {
        tw.nret = s_nret;
        tw.nnew = s_nnew;
    }
}

// grid = (tiles, parts): CTA (tile, p) owns bit rows [p * rows_per_part, ...) and the same range of check rows.
template <typename T, int V>
__global__ void __launch_bounds__(kSchedThreads) sched_move_kernel(StepArgs<T> a, BatchArgs<T> b, TileWork<32 * V> *work, int rows_per_part) {
    constexpr int FT = kWarp * V;
    const int tile = blockIdx.x, tid = threadIdx.x;
    TileWork<FT> &tw = work[tile];
    const int nret = tw.nret, nnew = tw.nnew;
    if (nret == 0 && nnew == 0) return;
    const int r0 = blockIdx.y * rows_per_part;   // multiple of 32
    __shared__ int s_last;

    // retire: a warp owns 32 consecutive bits; every lane reads its bit's V mask words ONCE (coalesced) and the warp
    // peels off one ballot per retiring frame -- n * 4V bytes per tile and step, however many frames retire (reading the
    // masks once per retiring frame cost 32 B of sector traffic per bit and frame: 3.3 MB per frame on the n = 102400 code)
    const int w0 = r0 >> 5, w1 = min(b.words, (r0 + rows_per_part) >> 5);
    __shared__ unsigned int s_diff[FT];
    if (nret > 0) {
        for (int r = tid; r < nret; r += blockDim.x) s_diff[r] = 0u;
        __syncthreads();
        const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
        const Vec<uint32_t, V> *zmv = reinterpret_cast<const Vec<uint32_t, V> *>(a.zmask + (size_t)tile * a.n * V);
        for (int w = w0 + warp; w < w1; w += nwarps) {
            const int i = w * 32 + lane;
            Vec<uint32_t, V> z;
#pragma unroll
            for (int v = 0; v < V; ++v) z.v[v] = 0u;
            if (i < a.n) z = zmv[i];
            for (int r = 0; r < nret; ++r) {
                const int s = tw.ret[r] & 255, sv = s % V, l = s / V;
                uint32_t zw = z.v[0];
#pragma unroll
                for (int v = 1; v < V; ++v) zw = (sv == v) ? z.v[v] : zw;
                const uint32_t word = __ballot_sync(0xffffffffu, (zw >> l) & 1u);
                if (lane == 0) {
                    const long long fr = tw.ret_frame[r];
                    if (b.out_bits) b.out_bits[fr * b.words + w] = word;
                    if (word ^ b.alice_bits[fr * b.words + w]) atomicOr(&s_diff[r], 1u);
                }
            }
        }
        __syncthreads();
        for (int r = tid; r < nret; r += blockDim.x)
            if (s_diff[r]) atomicOr(&tw.ret_diff[r], 1u);
    }

    // refill: transpose the new frames' key bits / syndromes into the tile's mask words. A warp owns 32 consecutive rows,
    // i.e. ONE packed word of every new frame: lane q fetches the word of new frame q (32 frames per round) and the warp
    // hands the words round by shuffle, instead of every lane loading every frame's word itself.
    if (nnew > 0) {
        const int lane = tid & 31;
        uint32_t *bm = a.bobmask + (size_t)tile * a.n * V;
        for (int i0 = r0 + (tid & ~31); i0 < min(a.n, r0 + rows_per_part); i0 += blockDim.x) {
            const int i = i0 + lane;
            const bool live = i < a.n;
            uint32_t mk[V];
#pragma unroll
            for (int v = 0; v < V; ++v) mk[v] = live ? bm[(size_t)i * V + v] : 0u;
            for (int qb = 0; qb < nnew; qb += 32) {
                const int mine = qb + lane;
                const uint32_t myword = mine < nnew ? b.bob_bits[tw.newframe[mine] * b.words + (i0 >> 5)] : 0u;
                const int myslot = mine < nnew ? tw.newslot[mine] : 0;
                const int cnt = min(32, nnew - qb);
                for (int q = 0; q < cnt; ++q) {
                    const uint32_t bit = (__shfl_sync(0xffffffffu, myword, q) >> lane) & 1u;
                    const int s = __shfl_sync(0xffffffffu, myslot, q), l = s / V;
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (v == s % V) mk[v] = (mk[v] & ~(1u << l)) | (bit << l);
                }
            }
            if (live) {
#pragma unroll
                for (int v = 0; v < V; ++v) bm[(size_t)i * V + v] = mk[v];
            }
        }
        uint32_t *sy = a.synd + (size_t)tile * a.m * V;
        uint32_t *pa = a.par + (size_t)tile * a.m * V;
        for (int j0 = r0 + (tid & ~31); j0 < min(a.m, r0 + rows_per_part); j0 += blockDim.x) {
            const int j = j0 + lane;
            const bool live = j < a.m;
            uint32_t ms[V], mp[V];
#pragma unroll
            for (int v = 0; v < V; ++v) { ms[v] = live ? sy[(size_t)j * V + v] : 0u; mp[v] = live ? pa[(size_t)j * V + v] : 0u; }
            for (int qb = 0; qb < nnew; qb += 32) {
                const int mine = qb + lane;
                const uint32_t myword = mine < nnew ? b.synd_all[tw.newframe[mine] * b.swords + (j0 >> 5)] : 0u;
                const int myslot = mine < nnew ? tw.newslot[mine] : 0;
                const int cnt = min(32, nnew - qb);
                for (int q = 0; q < cnt; ++q) {
                    const uint32_t sbit = (__shfl_sync(0xffffffffu, myword, q) >> lane) & 1u;
                    const uint32_t pbit = sbit;   // the VN-side init XORs the parity of the initial decision on top
                    const int s = __shfl_sync(0xffffffffu, myslot, q), l = s / V;
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        if (v == s % V) {
                            ms[v] = (ms[v] & ~(1u << l)) | (sbit << l);
                            mp[v] = (mp[v] & ~(1u << l)) | (pbit << l);
                        }
                }
            }
            if (live) {
#pragma unroll
                for (int v = 0; v < V; ++v) { sy[(size_t)j * V + v] = ms[v]; pa[(size_t)j * V + v] = mp[v]; }
            }
        }
    }

    // the last CTA of the tile to get here publishes the per-frame results of the retired frames
    if (nret == 0) return;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&tw.ticket, 1) == (int)gridDim.y - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int r = tid; r < nret; r += blockDim.x) {
        const int succ = (tw.ret[r] >> 8) & 1, iters = tw.ret[r] >> 9;
        const long long fr = tw.ret_frame[r];
        const bool keys_differ = *(volatile unsigned int *)&tw.ret_diff[r] != 0u;
        if (b.out_iters) b.out_iters[fr] = iters;
        if (b.out_flags) b.out_flags[fr] = (uint8_t)((succ ? 1u : 0u) | (keys_differ ? 0u : 2u));
        tally_frame(b.tally, succ, !keys_differ, iters, tw.run[r]);
    }
    if (tid == 0) tw.ticket = 0;
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

// One CTA plans the compaction: the k-th occupied slot past the tiles that stay is moved into the k-th free slot inside
// them (both in ascending order -- the plan a single thread walking the table would make, which is what this kernel was
// until the polls got frequent enough for its serial loads to show: 13 000 slots took a millisecond).
__device__ __forceinline__ int block_exclusive_scan_1024(int x, int *warp_sums, int &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    __syncthreads();   // warp_sums may still be read by the previous call
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const int v = warp_sums[k];
        base += (k < w) ? v : 0;
        tot += v;
    }
    total = tot;
    return base + incl - x;
}
template <int FT>
__global__ void __launch_bounds__(1024) compact_plan_kernel(int tiles, const long long *slot_frame, int2 *moves, CompactPlan *plan) {
    __shared__ int warp_sums[32];
    const int slots = tiles * FT;
    int mine = 0, active = 0;
    for (int s = threadIdx.x; s < slots; s += 1024) mine += slot_frame[s] >= 0;
    block_exclusive_scan_1024(mine, warp_sums, active);
    const int new_tiles = max(1, (active + FT - 1) / FT);
    const int low = new_tiles * FT;
    int n_src = 0;   // sources: occupied slots of [low, slots)
    for (int base = low; base < slots; base += 1024) {
        const int s = base + threadIdx.x;
        const int occ = (s < slots && slot_frame[s] >= 0) ? 1 : 0;
        int tot;
        const int rank = n_src + block_exclusive_scan_1024(occ, warp_sums, tot);
        if (occ) moves[rank].x = s;
        n_src += tot;
    }
    int n_dst = 0;   // destinations: free slots of [0, low), as many as there are sources (active <= low: they exist)
    for (int base = 0; base < low && n_dst < n_src; base += 1024) {
        const int s = base + threadIdx.x;
        const int fr = (s < low && slot_frame[s] < 0) ? 1 : 0;
        int tot;
        const int rank = n_dst + block_exclusive_scan_1024(fr, warp_sums, tot);
        if (fr && rank < n_src) moves[rank].y = s;
        n_dst += tot;
    }
    if (threadIdx.x == 0) {
        plan->n_moves = n_src;
        plan->new_tiles = new_tiles;
    }
}

// grid = (moves, chunks): messages of the moved frames. Slot position p of a tile <-> message lane p (common.cuh).
template <typename T, int FT>
__global__ void __launch_bounds__(256) compact_msg_kernel(const int2 *moves, const CompactPlan *plan, T *msg, long long e_stride, int nnz) {
    const int nm = plan->n_moves;
    for (int k = blockIdx.x; k < nm; k += gridDim.x) {   // grid-stride: the launch may hold fewer CTAs than there are moves
        const int2 mv = moves[k];
        const T *src = msg + (long long)(mv.x / FT) * e_stride + (mv.x % FT);
        T *dst = msg + (long long)(mv.y / FT) * e_stride + (mv.y % FT);
        for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < nnz; e += gridDim.y * blockDim.x) dst[(long long)e * FT] = src[(long long)e * FT];
    }
}

// grid = (moves, chunks): one bit per mask word. Slot p <-> word v = p % V, bit l = p / V (sched_kernel).
template <int V>
__global__ void __launch_bounds__(256) compact_mask_kernel(const int2 *moves, const CompactPlan *plan, int n, int m, uint32_t *bobmask,
                                                           uint32_t *zmask, uint32_t *synd, uint32_t *par) {
    constexpr int FT = 32 * V;
    const int nm = plan->n_moves;
    for (int k = blockIdx.x; k < nm; k += gridDim.x) {   // grid-stride over the moves, as in compact_msg_kernel
        const int2 mv = moves[k];
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

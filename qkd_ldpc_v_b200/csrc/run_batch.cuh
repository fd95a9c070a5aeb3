// The step loop of one batch: K1 prep -> initial fill -> {CN, VN, scheduler} x steps, replayed as a CUDA graph.
// Instantiated once per (message type, frames per lane) in inst_*.cu so the units compile in parallel; the
// check-node kernels of each algorithm pair live in their own units (launch_cn_alg<T, V, ALG>).
#pragma once
#include "handle.hpp"
#include "common.cuh"
#include "sched_kernels.cuh"
#include "step_kernels.cuh"

namespace qkhost {

using namespace qk;

// FAST check-node / variable-node arithmetic is exact only while no message can become NaN or infinite:
// a finite clamp guarantees it; without the clamp it holds when no factor can scale a message up (offset
// variants subtract; normalized variants need factors <= 1). SPA can produce NaN (0/0, quirk Q3) -> never fast.
inline bool fast_minsum_ok(const qkdldpc_params *P) {
    return P->message_precision == 32 && minsum_factors_ok(P);   // handle.hpp: also rejects negative / non-finite factors
}

inline unsigned ceil_div(int a, int b) { return (unsigned)((a + b - 1) / b); }

// Smallest of 1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48 ... that holds `need` tiles.
inline long long quantised_tiles(long long need) {
    long long q = 1;
    while (q < need) {
        if (q < 4) ++q;
        else if ((q & (q - 1)) == 0) q += q / 2;   // 2^k -> 3 * 2^(k-1)
        else q = q / 3 * 4;                        // 3 * 2^(k-1) -> 2^(k+1)
    }
    return q;
}

// Narrow variable-node buckets: which kernel, how many items per warp, how many resident CTAs per SM (B200, profiles/
// r02_ab_vn_loop.md). float32 with 4 frames per lane: vn_kernel_ell_loop, 4 CTAs per SM for dv <= 4 (64 registers, no
// spills: the 5 / 6-CTA builds spill 180 - 300 bytes inside the loop and lose to vn_kernel_ell), 3 for dv <= 8; a walk as long
// as leaves ~24 waves of CTAs in the grid (n = 102400: VN 0.71 -> 0.86 of the HBM peak at 32 - 64 items; n = 10240: best at
// 8, a longer walk leaves too few CTAs for the tail). float64 (2 frames per lane, the same 16 bytes per lane): 3 CTAs per
// SM for dv <= 4 (77 registers, no spills; uncapped it takes 88 and runs 2 CTAs: A82 SPA float64 0.56 -> 0.54 against
// 0.57 -> 0.60 at 3, n = 102400 SPA float64 0.54 -> 0.67). float32 with 2 frames per lane (the sum-product variants: half
// the bytes per warp): 5 CTAs per SM (48 registers, 16 bytes of spill); no gain before the kernel prefetched the next
// item's messages into L2, with it A82 SPA float32 VN 0.755 -> 0.854 (0.820 at 4 CTAs). 1 frame per lane keeps vn_kernel_ell.
struct VnLoopPlan {
    int items;   // <= 1 with ctas == 0: vn_kernel_ell
    int ctas;    // of the dv <= 4 kernel
};
inline VnLoopPlan vn_loop_plan(const qkdldpc_code *c, size_t elem_bytes, int V, int cnt, int tiles, int warps_per_cta) {
    VnLoopPlan p{c->opt.vn_items_per_warp, c->opt.vn_ctas_per_sm};
    const bool f64 = elem_bytes == 8;
    const int ctas = p.ctas > 0 ? p.ctas : (f64 ? 3 : (V == 2 ? 5 : 4));
    if (p.items == 1) return p;   // vn_kernel_ell, unless a CTA count was asked for too (the walking kernel with one item)
    if (p.items > 1) return VnLoopPlan{std::min(p.items, 64), ctas};
    if (!f64 && V != 4 && V != 2) return VnLoopPlan{1, p.ctas};
    if (c->sm_count <= 0 && cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device) != cudaSuccess) c->sm_count = 148;
    const int sms = c->sm_count;
    const long long per_wave = (long long)sms * ctas * warps_per_cta;
    const long long it = ((long long)cnt * tiles + per_wave * 12) / (per_wave * 24);   // rounded
    return VnLoopPlan{(int)std::max<long long>(1, std::min<long long>(it, 64)), ctas};
}

// One launch per non-empty degree bucket. Returns the number of kernels launched.
template <typename T, int V, int ALG>
int launch_cn_alg(const qkdldpc_code *c, bool fast, int tiles, cudaStream_t s, const StepArgs<T> &a);
template <typename T, int V>
int launch_vn(const qkdldpc_code *c, bool fast, int tiles, cudaStream_t s, const StepArgs<T> &a);

template <typename T, int V>
inline int launch_cn(const qkdldpc_code *c, int alg, bool fast, int tiles, cudaStream_t s, const StepArgs<T> &a) {
    switch (alg) {
        case 0: return launch_cn_alg<T, V, 0>(c, fast, tiles, s, a);
        case 1: return launch_cn_alg<T, V, 1>(c, fast, tiles, s, a);
        case 2: return launch_cn_alg<T, V, 2>(c, fast, tiles, s, a);
        case 3: return launch_cn_alg<T, V, 3>(c, fast, tiles, s, a);
        case 4: return launch_cn_alg<T, V, 4>(c, fast, tiles, s, a);
        default: return launch_cn_alg<T, V, 5>(c, fast, tiles, s, a);
    }
}

#ifdef QK_DEFINE_CN_LAUNCH
template <typename T, int V, int ALG, int B>
inline int launch_cn_bucket(const qkdldpc_code *c, bool fast, int tiles, cudaStream_t s, const StepArgs<T> &a) {
    // float64 sum-product: the two-pass kernel for every bucket -- the row's tanh values go back to the message array
    // between the passes instead of staying in registers (144 registers for 24 edges x 2 frames left 3 warps per scheduler
    // to hide the latency of the double-precision polynomial chains; the two-pass kernel needs under 64). The kernel is
    // bound by the FP64 pipe, not by the extra pass: with the dc <= 8 rows of the n = 102400 code kept in registers (64
    // registers) the check node took 93.0 ms per batch against 92.5 (round 2, call AO).
    constexpr int DCMAX = (sizeof(T) == 8 && ALG == 0) ? 0 : cn_bucket_max(B);
    const int cnt = c->cn_count[B];
    if (cnt == 0) return 0;
    constexpr int threads = cn_threads(sizeof(T), V, DCMAX);
    const dim3 grid(ceil_div(cnt, threads / 32), (unsigned)tiles);
    if constexpr (sizeof(T) == 4 && ALG >= 2 && DCMAX > 0) {
        if (fast) {
            cn_kernel<T, V, ALG, DCMAX, true><<<grid, threads, 0, s>>>(a, c->cn_first[B], cnt);
            return 1;
        }
    }
    cn_kernel<T, V, ALG, DCMAX, false><<<grid, threads, 0, s>>>(a, c->cn_first[B], cnt);
    return 1;
}
template <typename T, int V, int ALG>
int launch_cn_alg(const qkdldpc_code *c, bool fast, int tiles, cudaStream_t s, const StepArgs<T> &a) {
    return launch_cn_bucket<T, V, ALG, 0>(c, fast, tiles, s, a) + launch_cn_bucket<T, V, ALG, 1>(c, fast, tiles, s, a) +
           launch_cn_bucket<T, V, ALG, 2>(c, fast, tiles, s, a) + launch_cn_bucket<T, V, ALG, 3>(c, fast, tiles, s, a) +
           launch_cn_bucket<T, V, ALG, 4>(c, fast, tiles, s, a);
}
#endif

#ifdef QK_DEFINE_RUN_BATCH
template <typename T, int V, int B>
inline int launch_vn_bucket(const qkdldpc_code *c, bool fast, int tiles, cudaStream_t s, const StepArgs<T> &a) {
    constexpr int DVMAX = vn_bucket_max(B);
    const int cnt = c->vn_count[B];
    if (cnt == 0) return 0;
    constexpr int threads = vn_threads(sizeof(T), V, DVMAX);
    const dim3 grid(ceil_div(cnt, threads / 32), (unsigned)tiles);
    if constexpr (DVMAX == 4 || DVMAX == 8) {   // narrow buckets: ELL records, two waves of loads (vn_kernel_ell)
        const int ell_base = (B == 0) ? 0 : c->vn_count[0] * vn_bucket_max(0);
        const VnLoopPlan plan = vn_loop_plan(c, sizeof(T), V, cnt, tiles, threads / 32);
        const int items = plan.items, want = plan.ctas;   // CTAs of the dv <= 4 kernel; the dv <= 8 kernel: 3, or 2 for 3
        if (items > 1 || want > 0) {   // several items per warp, index records one item ahead in L1
            const dim3 lgrid(ceil_div(cnt, (threads / 32) * items), (unsigned)tiles);
#define QK_VN_LOOP(FA, CT) vn_kernel_ell_loop<T, V, DVMAX, FA, CT><<<lgrid, threads, 0, s>>>(a, c->vn_first[B], cnt, ell_base, items)
#define QK_VN_LOOP_CTAS(FA)                                 \
    do {                                                    \
        if constexpr (sizeof(T) != 4) {                     \
            if constexpr (DVMAX == 4) {                     \
                if (want == 1) QK_VN_LOOP(FA, 1);           \
                else QK_VN_LOOP(FA, 3);                     \
            } else QK_VN_LOOP(FA, 1);                       \
        }                                                   \
        else if constexpr (DVMAX == 4) {                    \
            if (want == 6) QK_VN_LOOP(FA, 6);               \
            else if (want == 5) QK_VN_LOOP(FA, 5);          \
            else if (want == 3) QK_VN_LOOP(FA, 3);          \
            else QK_VN_LOOP(FA, 4);                         \
        } else {                                            \
            if (want == 3) QK_VN_LOOP(FA, 2);               \
            else QK_VN_LOOP(FA, 3);                         \
        }                                                   \
    } while (0)
            if constexpr (sizeof(T) == 4) {
                if (fast) {
                    QK_VN_LOOP_CTAS(true);
                    return 1;
                }
            }
            QK_VN_LOOP_CTAS(false);
#undef QK_VN_LOOP_CTAS
#undef QK_VN_LOOP
            return 1;
        }
        if constexpr (sizeof(T) == 4) {
            if (fast) {
                vn_kernel_ell<T, V, DVMAX, true><<<grid, threads, 0, s>>>(a, c->vn_first[B], cnt, ell_base);
                return 1;
            }
        }
        vn_kernel_ell<T, V, DVMAX, false><<<grid, threads, 0, s>>>(a, c->vn_first[B], cnt, ell_base);
        return 1;
    } else {
        if constexpr (DVMAX == 16 || DVMAX == 32) {   // wide buckets: coalesced index loads, the next item's ids and messages ahead
            VnLoopPlan plan = vn_loop_plan(c, sizeof(T), V, cnt, tiles, threads / 32);
            // (float64: 126 / 176 registers leave 2 CTAs of 128 threads per SM, and the I80 workload measured 0.862 -> 0.787
            // Gbit/s with every bucket walking: float64 keeps vn_kernel for the wide buckets)
            if (c->opt.vn_items_per_warp != 1 && V > 1 && sizeof(T) == 4) {
                const dim3 lgrid(ceil_div(cnt, (threads / 32) * plan.items), (unsigned)tiles);
                if constexpr (sizeof(T) == 4) {
                    if (fast) {
                        vn_kernel_wide_loop<T, V, DVMAX, true><<<lgrid, threads, 0, s>>>(a, c->vn_first[B], cnt, plan.items);
                        return 1;
                    }
                }
                vn_kernel_wide_loop<T, V, DVMAX, false><<<lgrid, threads, 0, s>>>(a, c->vn_first[B], cnt, plan.items);
                return 1;
            }
        }
        if constexpr (sizeof(T) == 4) {
            if (fast) {
                vn_kernel<T, V, DVMAX, true><<<grid, threads, 0, s>>>(a, c->vn_first[B], cnt);
                return 1;
            }
        }
        vn_kernel<T, V, DVMAX, false><<<grid, threads, 0, s>>>(a, c->vn_first[B], cnt);
        return 1;
    }
}
// The bucket kernels are latency-bound one by one (48-59 % of DRAM peak each, profiles/r01_c_ncu_summary.md): they touch
// disjoint bits, so they run CONCURRENTLY -- the widest bucket stays on the main stream, the others fork onto side
// streams and join back (also inside a captured graph, where this becomes parallel branches).
template <typename T, int V>
int launch_vn(const qkdldpc_code *c, bool fast, int tiles, cudaStream_t s, const StepArgs<T> &a) {
    int order[kBuckets], nb = 0;
    for (int bk = kBuckets - 1; bk >= 0; --bk)
        if (c->vn_count[bk] > 0) order[nb++] = bk;
    auto launch_bucket = [&](int bk, cudaStream_t st) {
        switch (bk) {
            case 4: return launch_vn_bucket<T, V, 4>(c, fast, tiles, st, a);
            case 3: return launch_vn_bucket<T, V, 3>(c, fast, tiles, st, a);
            case 2: return launch_vn_bucket<T, V, 2>(c, fast, tiles, st, a);
            case 1: return launch_vn_bucket<T, V, 1>(c, fast, tiles, st, a);
            default: return launch_vn_bucket<T, V, 0>(c, fast, tiles, st, a);
        }
    };
    const bool fork = nb > 1 && c->side_streams[0] != nullptr && s != nullptr;
    if (!fork) {
        int nl = 0;
        for (int k = 0; k < nb; ++k) nl += launch_bucket(order[k], s);
        return nl;
    }
    cudaEventRecord(c->ev_fork, s);
    int nl = 0;
    for (int k = 1; k < nb; ++k) {
        cudaStream_t side = c->side_streams[(k - 1) % kSideStreams];
        if (k - 1 < kSideStreams) cudaStreamWaitEvent(side, c->ev_fork, 0);
        nl += launch_bucket(order[k], side);
    }
    nl += launch_bucket(order[0], s);
    for (int k = 0; k < std::min(nb - 1, kSideStreams); ++k) {
        cudaEventRecord(c->ev_join[k], c->side_streams[k]);
        cudaStreamWaitEvent(s, c->ev_join[k], 0);
    }
    return nl;
}

inline EvPair *next_ev(qkdldpc_code *c, int kind) {
    if (c->ev_used == c->ev_pool.size()) {
        EvPair p{};
        cudaEventCreate(&p.a);
        cudaEventCreate(&p.b);
        c->ev_pool.push_back(p);
    }
    EvPair *p = &c->ev_pool[c->ev_used++];
    p->kind = kind;
    return p;
}

template <typename T, int V>
int run_batch(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_frames, const uint32_t *d_alice,
              const uint32_t *d_bob, const double *d_qber, int qber_is_scalar, const int32_t *punct, int n_punct,
              const int32_t *shortd, int n_short, uint32_t *d_out_bits, int32_t *d_out_iters, uint8_t *d_out_flags,
              unsigned long long *d_tally) {
    constexpr int FT = kWarp * V;
    const int n = c->n, m = c->m;
    const int words = (n + 31) / 32, swords = (m + 31) / 32;
    const bool adaptive = P->algorithm >= 4;
    cudaStream_t s = c->stream;

    // ---- pool geometry -------------------------------------------------------------------------------------
    const int64_t tiles_needed = (n_frames + FT - 1) / FT;
    const int64_t per_tile = (int64_t)c->nnz * FT * (int64_t)sizeof(T);
    // default 2 GiB: enough frames in flight to saturate HBM (>= 8k frames of a 10k code), small enough that a batch
    // refills every slot several times -- at converging operating points a pool as large as the batch leaves most lanes
    // idle while each tile waits for its slowest frame. Long codes (n = 100k: 157 MB of messages per 128-frame tile) get
    // at least 32 tiles: the scheduler handles one tile per CTA, and with a dozen tiles its retire / refill work (bit
    // transposes of n-bit frames) costs more than the decoding (measured 75 vs 55 ms per step at 13 tiles).
    int64_t budget = c->opt.pool_bytes > 0 ? c->opt.pool_bytes : (int64_t)2 << 30;
    if (c->opt.pool_bytes <= 0) budget = std::max(budget, std::min<int64_t>(32 * per_tile, (int64_t)8 << 30));
    int64_t tiles = std::max<int64_t>(1, budget / std::max<int64_t>(per_tile, 1));
    if (c->opt.pool_slots > 0) tiles = std::max<int64_t>(1, (c->opt.pool_slots + FT - 1) / FT);
    tiles = std::min<int64_t>(tiles, tiles_needed);
    tiles = std::min<int64_t>(tiles, 65535);
    const size_t slots = (size_t)tiles * FT;

    CK(c->msg.reserve((size_t)tiles * per_tile));
    CK(c->bobmask.reserve((size_t)tiles * n * V));
    CK(c->zmask.reserve((size_t)tiles * n * V));
    CK(c->synd.reserve((size_t)tiles * m * V));
    CK(c->par.reserve((size_t)tiles * m * V));
    CK(c->tile_active.reserve((size_t)tiles * V));
    CK(c->tile_new.reserve((size_t)tiles * V));
    CK(c->slot_llr.reserve(slots * sizeof(T)));
    CK(c->slot_frame.reserve(slots));
    CK(c->slot_iter.reserve(slots));
    CK(c->frame_llr.reserve((size_t)n_frames * sizeof(T)));
    CK(c->synd_all.reserve((size_t)n_frames * swords));
    CK(c->bitclass.reserve(n));
    CK(c->counters.reserve(2));
    c->frames_per_tile = FT;
    c->pool_tiles = (int)tiles;
    c->pool_bytes = tiles * per_tile;

    // ---- per-batch metadata: punctured / shortened classes (H_matrix_params, a&m_ops.hpp:44-48) ---------------
    {
        std::vector<uint8_t> cls(n, 0);
        for (int i = 0; i < n_punct; ++i) {
            if (punct[i] < 0 || punct[i] >= n) return fail(QKDLDPC_ERR_INVALID, "punctured position out of range");
            cls[punct[i]] = 1;
        }
        for (int i = 0; i < n_short; ++i) {
            if (shortd[i] < 0 || shortd[i] >= n) return fail(QKDLDPC_ERR_INVALID, "shortened position out of range");
            if (cls[shortd[i]] == 0) cls[shortd[i]] = 2;   // a position listed twice is punctured (:1150 tested first)
        }
        CK(cudaMemcpyAsync(c->bitclass.p, cls.data(), n, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));   // the host vector dies at scope end
    }

    CK(cudaMemsetAsync(c->tile_active.p, 0, (size_t)tiles * V * sizeof(uint32_t), s));
    CK(cudaMemsetAsync(c->tile_new.p, 0, (size_t)tiles * V * sizeof(uint32_t), s));
    CK(cudaMemsetAsync(c->slot_frame.p, 0xff, slots * sizeof(long long), s));
    CK(cudaMemsetAsync(c->slot_iter.p, 0, slots * sizeof(int32_t), s));
    CK(cudaMemsetAsync(c->counters.p, 0, 2 * sizeof(unsigned long long), s));
    if (d_tally) CK(cudaMemsetAsync(d_tally, 0, (size_t)qkdldpc_tally_len(P->max_iterations) * sizeof(uint64_t), s));

    StepArgs<T> a{};
    a.n = n; a.m = m;
    a.row_ptr = c->row_ptr.p; a.col_idx = c->col_idx.p; a.col_ptr = c->col_ptr.p;
    a.csc_edge = c->csc_edge.p; a.csc_row = c->csc_row.p;
    a.row_order = c->row_order.p; a.col_order = c->col_order.p;
    a.vn_ell_edge = c->vn_ell_edge.p; a.vn_ell_row = c->vn_ell_row.p;
    a.bitclass = c->bitclass.p;
    a.msg = reinterpret_cast<T *>(c->msg.p);
    a.bobmask = c->bobmask.p; a.zmask = c->zmask.p; a.synd = c->synd.p; a.par = c->par.p;
    a.tile_active = c->tile_active.p; a.tile_new = c->tile_new.p;
    a.slot_llr = reinterpret_cast<T *>(c->slot_llr.p);
    a.e_stride = (int64_t)c->nnz * FT;
    a.primary = (T)P->primary; a.secondary = (T)P->secondary; a.enable_thr = P->enable_threshold != 0;
    a.thr = a.enable_thr ? (T)P->threshold : (T)INFINITY;   // the kernels clamp unconditionally; +inf is a no-op

    BatchArgs<T> b{};
    b.n_frames = n_frames; b.words = words; b.swords = swords;
    b.alice_bits = d_alice; b.bob_bits = d_bob; b.qber = d_qber; b.qber_is_scalar = qber_is_scalar;
    b.frame_llr = reinterpret_cast<T *>(c->frame_llr.p);
    b.synd_all = c->synd_all.p;
    b.out_bits = d_out_bits; b.out_iters = d_out_iters; b.out_flags = d_out_flags;
    b.tally = d_tally;
    b.next_frame = c->counters.p; b.n_done = c->counters.p + 1;
    b.slot_frame = c->slot_frame.p; b.slot_iter = c->slot_iter.p;
    b.max_iter = P->max_iterations; b.adaptive = adaptive ? 1 : 0;

    c->ev_used = 0;
    CK(cudaEventRecord(c->ev0, s));
    prep_kernel<T><<<(unsigned)n_frames, 256, 0, s>>>(n, m, c->row_ptr.p, c->col_idx.p, b);
    // scheduler work lists (one per tile) and the split of a tile's rows over the CTAs of sched_move_kernel
    CK(c->sched_work.reserve((size_t)tiles * sizeof(TileWork<FT>)));
    CK(cudaMemsetAsync(c->sched_work.p, 0, (size_t)tiles * sizeof(TileWork<FT>), s));
    auto *work = reinterpret_cast<TileWork<FT> *>(c->sched_work.p);
    const int rows_per_part = std::max(4096, ((std::max(n, m) + 63) / 64 + 31) / 32 * 32);   // multiple of 32, at most 64 parts
    const unsigned move_parts = (unsigned)((std::max(n, m) + rows_per_part - 1) / rows_per_part);
    sched_kernel<T, V><<<(unsigned)tiles, kSchedThreads, 0, s>>>(a, b, work);
    sched_move_kernel<T, V><<<dim3((unsigned)tiles, move_parts), kSchedThreads, 0, s>>>(a, b, work, rows_per_part);
    c->kernel_launches += 3;
    CK(cudaGetLastError());

    const int alg = P->algorithm;
    const bool fast = fast_minsum_ok(P);
    int launches_per_step = 0;
    int cur_tiles = (int)tiles;   // shrinks when the tail of the batch is compacted (sched_kernels.cuh)
    auto one_step = [&](cudaStream_t st, bool prof) {
        EvPair *e = nullptr;
        if (prof) { e = next_ev(c, 0); cudaEventRecord(e->a, st); }
        int nl = launch_cn<T, V>(c, alg, fast, cur_tiles, st, a);
        if (prof) { cudaEventRecord(e->b, st); e = next_ev(c, 1); cudaEventRecord(e->a, st); }
        nl += launch_vn<T, V>(c, fast, cur_tiles, st, a);
        if (prof) { cudaEventRecord(e->b, st); e = next_ev(c, 2); cudaEventRecord(e->a, st); }
        sched_kernel<T, V><<<(unsigned)cur_tiles, kSchedThreads, 0, st>>>(a, b, work);
        sched_move_kernel<T, V><<<dim3((unsigned)cur_tiles, move_parts), kSchedThreads, 0, st>>>(a, b, work, rows_per_part);
        if (prof) cudaEventRecord(e->b, st);
        launches_per_step = nl + 2;
    };

    // steps between two host polls of the done counter: a frame needs at most max_iter steps, and the host
    // only has to look when a whole generation of slots may have drained
    // Auto: about 8 ms of decoding between two polls, 16 steps at most. Every step past the one that retires the last frame
    // still launches the whole grid over empty tiles (n = 102400, 32 tiles: 1.2 ms per step, three of them = 7 % of a batch
    // whose frames all converge within 13 iterations), and the tail compaction can only start at a poll; a poll itself is a
    // stream synchronisation and a graph launch (tens of microseconds), so short steps keep the long interval (A79, 32 tiles,
    // 0.5 ms per step: 7 steps per poll lost 2.7 % against 16; 102 tiles, 1.7 ms: 2..4 steps per poll gain 6 %).
    const double step_seconds = (double)tiles * (double)per_tile * 4.0 / 5e12;   // CN + VN: every message read and written twice
    const int auto_spp = (int)std::max(1.0, std::min(16.0, std::floor(8e-3 / step_seconds + 0.5)));
    int spp = c->opt.steps_per_poll > 0 ? c->opt.steps_per_poll : std::max(1, std::min(P->max_iterations, auto_spp));
    c->last_spp = spp;
    c->last_vn_items = c->vn_count[0] > 0 ? vn_loop_plan(c, sizeof(T), V, c->vn_count[0], (int)tiles, vn_threads(sizeof(T), V, 4) / 32).items : 0;
    // (the legacy default stream cannot be captured: plain launches there)
    const bool use_graph = c->opt.use_graph >= 0 && !c->profiling && s != nullptr;

    // the captured launches embed every pointer and parameter of this batch: the cache is keyed on all of them
    cudaGraphExec_t graph_exec = nullptr;
    auto ensure_graph = [&]() -> int {
        // the key is the launch arguments themselves, byte for byte (the structs are value-initialised, padding included)
        std::string key;
        auto mixin = [&key](const void *p, size_t nbytes) { key.append(static_cast<const char *>(p), nbytes); };
        mixin(&a, sizeof a);
        mixin(&b, sizeof b);
        const int geo[8] = {(int)sizeof(T), V, alg, cur_tiles, spp, fast ? 1 : 0, rows_per_part, (int)move_parts};
        mixin(geo, sizeof geo);
        mixin(&work, sizeof work);
        for (size_t i = 0; i < c->graphs.size(); ++i)
            if (c->graphs[i].key == key) {
                graph_exec = c->graphs[i].exec;
                return QKDLDPC_OK;
            }
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        // from here to EndCapture nothing may return early: a failure would leave the handle's stream (and the forked
        // side streams) in capture mode and every later call on the handle would fail
        for (int k = 0; k < spp; ++k) one_step(s, false);
        cudaError_t ce = cudaGetLastError();   // launch errors of the captured kernels
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(c->h_done, c->counters.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
        const cudaError_t ee = cudaStreamEndCapture(s, &g);   // always: ends the capture even when it is invalidated
        if (ce == cudaSuccess) ce = ee;
        cudaGraphExec_t ex = nullptr;
        if (ce == cudaSuccess) ce = cudaGraphInstantiate(&ex, g, 0);
        if (g) cudaGraphDestroy(g);
        if (ce != cudaSuccess) {
            cudaGetLastError();
            return fail(QKDLDPC_ERR_CUDA, "capture of the step graph failed: %s", cudaGetErrorString(ce));
        }
        if (c->graphs.size() >= qkdldpc_code::kMaxGraphs) {   // oldest out (a sweep over combinations keeps capturing)
            cudaGraphExecDestroy(c->graphs.front().exec);
            c->graphs.erase(c->graphs.begin());
        }
        c->graphs.push_back({key, ex});
        graph_exec = ex;
        return QKDLDPC_OK;
    };
    if (use_graph) {
        const int rc = ensure_graph();
        if (rc) return rc;
    }

    c->h_done[0] = c->h_done[1] = 0;   // [0] frames handed out by the queue, [1] frames finished
    // upper bound on steps: every generation of the pool needs at most max_iter steps (+1 for refills that
    // waited one step); guards against a hang if something is badly wrong
    const int64_t generations = (n_frames + (int64_t)slots - 1) / (int64_t)slots;
    const int64_t step_limit = (generations + 1) * ((int64_t)P->max_iterations + 2) + spp;
    int64_t steps = 0;
    const bool compaction = c->opt.tail_compaction >= 0;
    const long long fill_pct = c->opt.compaction_fill_pct > 0 ? std::min(c->opt.compaction_fill_pct, 99) : 75;
    unsigned long long prev_done = 0;
    while (true) {
        if (use_graph) {
            CK(cudaGraphLaunch(graph_exec, s));
        } else {
            for (int k = 0; k < spp; ++k) one_step(s, c->profiling);
            CK(cudaMemcpyAsync(c->h_done, c->counters.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        }
        steps += spp;
        if (use_graph && launches_per_step == 0) {   // graph came from the cache: count its kernels once
            for (int k = 0; k < kBuckets; ++k) launches_per_step += (c->cn_count[k] > 0) + (c->vn_count[k] > 0);
            launches_per_step += 2;
        }
        c->kernel_launches += (int64_t)launches_per_step * spp;
        c->decoder_steps += spp;
        CK(cudaEventRecord(c->ev_poll, s));
        CK(cudaEventSynchronize(c->ev_poll));
        const unsigned long long handed_out = c->h_done[0], done = c->h_done[1];
        if (done >= (unsigned long long)n_frames) break;
        if (steps > step_limit)
            return fail(QKDLDPC_ERR_STATE, "decoder did not finish: %llu of %lld frames after %lld steps", done, (long long)n_frames,
                        (long long)steps);
        // Tail compaction: the queue is empty and at most fill_pct % (default: 75) of the slots of >= 4 tiles are still occupied
        const long long remaining = (long long)n_frames - (long long)done;
        // (the tile count is rounded up to 1, 2, 3, 4, 6, 8, 12, 16, 24 ...: the step graphs of the next batch's tail are then
        // the ones captured here, and a compaction that would not shrink the grid is skipped)
        const long long q_tiles = quantised_tiles(std::max<long long>(1, (remaining + FT - 1) / FT));
        // A move costs about 64 B of sector traffic per edge (one 4-byte word out of a 512-byte chunk and into another), a
        // tile about 2 KB per edge and step: the moves are paid back after ~4 x occupancy steps on the smaller grid. Frames
        // that are about to converge anyway are not worth moving (n = 102400 NMSA: every frame converges at iteration
        // 10..13, compacting the 28 % left after iteration 11 cost 5 ms and saved 3.5), so the batch must be expected to
        // last twice that long at the rate frames retired since the last poll; stragglers that fail retire at rate 0.
        const double retire_rate = (double)(done - prev_done) / (double)spp;   // frames per step
        const double occupancy = (double)remaining / ((double)cur_tiles * FT);
        const bool lasts = (double)remaining > retire_rate * 8.0 * occupancy;
        prev_done = done;
        if (compaction && handed_out >= (unsigned long long)n_frames && cur_tiles >= 4 && remaining * 100 <= (long long)cur_tiles * FT * fill_pct &&
            q_tiles < cur_tiles && lasts) {
            CK(c->compact_moves.reserve((size_t)cur_tiles * FT));
            CK(c->compact_plan.reserve(2));
            auto *plan = reinterpret_cast<CompactPlan *>(c->compact_plan.p);
            compact_plan_kernel<FT><<<1, 1024, 0, s>>>(cur_tiles, c->slot_frame.p, c->compact_moves.p, plan);
            // both copy kernels stride over plan->n_moves, so a capped grid still moves every frame (a pool of many
            // thousand tiles of a short code can hold more than 65535 stragglers)
            const long long move_cap = c->opt.compaction_max_ctas > 0 ? c->opt.compaction_max_ctas : 65535;
            const unsigned max_moves = (unsigned)std::max<long long>(1, std::min<long long>(remaining, move_cap));
            compact_msg_kernel<T, FT><<<dim3(max_moves, 8), 256, 0, s>>>(c->compact_moves.p, plan, a.msg, a.e_stride, (int)c->nnz);
            compact_mask_kernel<V><<<dim3(max_moves, 4), 256, 0, s>>>(c->compact_moves.p, plan, n, m, a.bobmask, a.zmask, a.synd, a.par);
            compact_finish_kernel<T, V><<<1, 1024, 0, s>>>(c->compact_moves.p, plan, cur_tiles, c->slot_frame.p, c->slot_iter.p, a.slot_llr,
                                                          a.tile_active, a.tile_new);
            c->kernel_launches += 4;
            ++c->tail_compactions;
            CK(cudaGetLastError());
            cur_tiles = (int)q_tiles;   // >= plan->new_tiles (active == remaining); the tiles past it are empty
            if (use_graph) {
                const int rc = ensure_graph();
                if (rc) return rc;
            }
        }
    }
    CK(cudaEventRecord(c->ev1, s));
    CK(cudaEventSynchronize(c->ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_batch_ms = ms;
    c->last_cn_ms = c->last_vn_ms = c->last_sched_ms = 0;
    for (size_t i = 0; i < c->ev_used; ++i) {
        float t = 0;
        cudaEventElapsedTime(&t, c->ev_pool[i].a, c->ev_pool[i].b);
        (c->ev_pool[i].kind == 0 ? c->last_cn_ms : c->ev_pool[i].kind == 1 ? c->last_vn_ms : c->last_sched_ms) += t;
    }
    CK(cudaGetLastError());
    return QKDLDPC_OK;
}

#endif  // QK_DEFINE_RUN_BATCH

}  // namespace qkhost

// The step loop of one batch: K1 prep -> initial fill -> {CN, VN, scheduler} x steps, replayed as a CUDA graph.
// Instantiated once per (message type, frames per lane) in inst_*.cu so the units compile in parallel.
#pragma once
#include "handle.hpp"
#include "common.cuh"
#include "sched_kernels.cuh"
#include "step_kernels.cuh"

namespace qkhost {

using namespace qk;

template <typename T, int V>
inline void launch_cn(int alg, dim3 grid, cudaStream_t s, const StepArgs<T> &a) {
    const dim3 blk(kCnWarps * kWarp);
    switch (alg) {
        case 0: cn_kernel<T, V, 0><<<grid, blk, 0, s>>>(a); break;
        case 1: cn_kernel<T, V, 1><<<grid, blk, 0, s>>>(a); break;
        case 2: cn_kernel<T, V, 2><<<grid, blk, 0, s>>>(a); break;
        case 3: cn_kernel<T, V, 3><<<grid, blk, 0, s>>>(a); break;
        case 4: cn_kernel<T, V, 4><<<grid, blk, 0, s>>>(a); break;
        default: cn_kernel<T, V, 5><<<grid, blk, 0, s>>>(a); break;
    }
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
    int64_t budget = c->opt.pool_bytes > 0 ? c->opt.pool_bytes : (int64_t)8 << 30;
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
    if (adaptive) {
        CK(c->par0_all.reserve((size_t)n_frames * swords));
        CK(c->pre_done.reserve((size_t)n_frames));
    }
    CK(c->payload.reserve(words));
    CK(c->bitclass.reserve(n));
    CK(c->counters.reserve(2));
    c->frames_per_tile = FT;
    c->pool_tiles = (int)tiles;
    c->pool_bytes = tiles * per_tile;

    // ---- per-batch metadata: punctured / shortened classes (H_matrix_params, a&m_ops.hpp:44-48) ---------------
    {
        std::vector<uint8_t> cls(n, 0);
        std::vector<uint32_t> pay(words, 0);
        for (int i = 0; i < n_punct; ++i) {
            if (punct[i] < 0 || punct[i] >= n) return fail(QKDLDPC_ERR_INVALID, "punctured position out of range");
            cls[punct[i]] = 1;
        }
        for (int i = 0; i < n_short; ++i) {
            if (shortd[i] < 0 || shortd[i] >= n) return fail(QKDLDPC_ERR_INVALID, "shortened position out of range");
            if (cls[shortd[i]] == 0) cls[shortd[i]] = 2;   // a position listed twice is punctured (:1150 tested first)
        }
        for (int i = 0; i < n; ++i)
            if (cls[i] == 0) pay[i >> 5] |= 1u << (i & 31);
        CK(cudaMemcpyAsync(c->bitclass.p, cls.data(), n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(c->payload.p, pay.data(), words * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));   // the host vectors die at scope end
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
    a.cn_items = c->cn_items.p; a.vn_items = c->vn_items.p;
    a.bitclass = c->bitclass.p;
    a.msg = reinterpret_cast<T *>(c->msg.p);
    a.bobmask = c->bobmask.p; a.zmask = c->zmask.p; a.synd = c->synd.p; a.par = c->par.p;
    a.tile_active = c->tile_active.p; a.tile_new = c->tile_new.p;
    a.slot_llr = reinterpret_cast<T *>(c->slot_llr.p);
    a.e_stride = (int64_t)c->nnz * FT;
    a.primary = (T)P->primary; a.secondary = (T)P->secondary; a.thr = (T)P->threshold;
    a.enable_thr = P->enable_threshold != 0;

    BatchArgs<T> b{};
    b.n_frames = n_frames; b.words = words; b.swords = swords;
    b.alice_bits = d_alice; b.bob_bits = d_bob; b.qber = d_qber; b.qber_is_scalar = qber_is_scalar;
    b.payload = c->payload.p;
    b.frame_llr = reinterpret_cast<T *>(c->frame_llr.p);
    b.synd_all = c->synd_all.p; b.par0_all = c->par0_all.p; b.pre_done = c->pre_done.p;
    b.out_bits = d_out_bits; b.out_iters = d_out_iters; b.out_flags = d_out_flags;
    b.tally = d_tally;
    b.next_frame = c->counters.p; b.n_done = c->counters.p + 1;
    b.slot_frame = c->slot_frame.p; b.slot_iter = c->slot_iter.p;
    b.max_iter = P->max_iterations; b.adaptive = adaptive ? 1 : 0;

    c->ev_used = 0;
    CK(cudaEventRecord(c->ev0, s));
    prep_kernel<T><<<(unsigned)n_frames, 256, 0, s>>>(n, m, c->row_ptr.p, c->col_idx.p, b);
    sched_kernel<T, V><<<(unsigned)tiles, kSchedThreads, 0, s>>>(a, b);
    c->kernel_launches += 2;
    CK(cudaGetLastError());

    const dim3 cn_grid(c->n_cn_items, (unsigned)tiles), vn_grid(c->n_vn_items, (unsigned)tiles);
    const int alg = P->algorithm;
    auto one_step = [&](cudaStream_t st, bool prof) {
        EvPair *e = nullptr;
        if (prof) { e = next_ev(c, 0); cudaEventRecord(e->a, st); }
        launch_cn<T, V>(alg, cn_grid, st, a);
        if (prof) { cudaEventRecord(e->b, st); e = next_ev(c, 1); cudaEventRecord(e->a, st); }
        vn_kernel<T, V><<<vn_grid, kVnWarps * kWarp, 0, st>>>(a);
        if (prof) { cudaEventRecord(e->b, st); e = next_ev(c, 2); cudaEventRecord(e->a, st); }
        sched_kernel<T, V><<<(unsigned)tiles, kSchedThreads, 0, st>>>(a, b);
        if (prof) cudaEventRecord(e->b, st);
    };

    // steps between two host polls of the done counter: a frame needs at most max_iter steps, and the host
    // only has to look when a whole generation of slots may have drained
    int spp = c->opt.steps_per_poll > 0 ? c->opt.steps_per_poll : std::max(1, std::min(P->max_iterations, 16));
    // (the legacy default stream cannot be captured: plain launches there)
    const bool use_graph = c->opt.use_graph >= 0 && !c->profiling && s != nullptr;

    if (use_graph) {
        // the captured launches embed every pointer and parameter of this batch: key the cache on all of them
        unsigned long long h = 1469598103934665603ull;
        auto mixin = [&h](const void *p, size_t nbytes) {
            const unsigned char *q = static_cast<const unsigned char *>(p);
            for (size_t i = 0; i < nbytes; ++i) h = (h ^ q[i]) * 1099511628211ull;
        };
        mixin(&a, sizeof a);
        mixin(&b, sizeof b);
        const int geo[6] = {(int)sizeof(T), V, alg, (int)tiles, spp, c->n_cn_items};
        mixin(geo, sizeof geo);
        char key[64];
        snprintf(key, sizeof key, "%016llx", h);
        if (c->graph_key != key) {
            if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
            cudaGraph_t g = nullptr;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            for (int k = 0; k < spp; ++k) one_step(s, false);
            CK(cudaMemcpyAsync(c->h_done, c->counters.p + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamEndCapture(s, &g));
            CK(cudaGraphInstantiate(&c->graph_exec, g, 0));
            cudaGraphDestroy(g);
            c->graph_key = key;
        }
    }

    *c->h_done = 0;
    // upper bound on steps: every generation of the pool needs at most max_iter steps (+1 for refills that
    // waited one step); guards against a hang if something is badly wrong
    const int64_t generations = (n_frames + (int64_t)slots - 1) / (int64_t)slots;
    const int64_t step_limit = (generations + 1) * ((int64_t)P->max_iterations + 2) + spp;
    int64_t steps = 0;
    while (true) {
        if (use_graph) {
            CK(cudaGraphLaunch(c->graph_exec, s));
        } else {
            for (int k = 0; k < spp; ++k) one_step(s, c->profiling);
            CK(cudaMemcpyAsync(c->h_done, c->counters.p + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        }
        steps += spp;
        c->kernel_launches += 3 * (int64_t)spp;
        c->decoder_steps += spp;
        CK(cudaEventRecord(c->ev_poll, s));
        CK(cudaEventSynchronize(c->ev_poll));
        if (*c->h_done >= (unsigned long long)n_frames) break;
        if (steps > step_limit)
            return fail(QKDLDPC_ERR_STATE, "decoder did not finish: %llu of %lld frames after %lld steps",
                        *c->h_done, (long long)n_frames, (long long)steps);
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

}  // namespace qkhost

// On-chip min-sum path: host-side launcher (one persistent kernel per batch) and the kernel instantiations (float32 / float64 state).
#include "handle.hpp"
#include "onchip_minsum.cuh"
#include "onchip_minsum64.cuh"

namespace qkhost {

using namespace qk;

cudaError_t onchip_spa_geometry(int alg, int groups_cn, int sms, size_t smem, long long n_frames, int *threads, int *grid);
cudaError_t onchip_spa_launch(int alg, const OnchipArgs &a, int grid, int threads, size_t smem, cudaStream_t s);

bool onchip_usable(const qkdldpc_code *c, const qkdldpc_params *P) {
    if (P->message_precision != 32 && !(P->message_precision == 64 && P->algorithm >= 2)) return false;
    if (P->algorithm < 2) {
        // sum-product kernel (onchip_spa.cuh): one float per edge must fit; NaN / inf messages are handled as in the
        // streaming kernels, so there is no precondition on the parameters
        int dev_smem = 0;
        if (!c->sp_eligible || cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device) != cudaSuccess)
            return false;
        return onchip_spa_smem_bytes(c->n, c->sp_msg_words, c->oc_groups_cn, c->sp_groups_sv) <= (size_t)dev_smem;
    }
    const bool f64 = P->message_precision == 64;
    if (!c->oc2.eligible) return false;
    // same precondition as the FAST streaming kernels (minsum_factors_ok, handle.hpp): no message can become NaN / inf,
    // and the factors are finite and non-negative (the magnitude clamp min(c, thr) covers only the positive side)
    if (!minsum_factors_ok(P)) return false;
    int dev_smem = 0;
    if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device) != cudaSuccess) return false;
    const Oc2Device &d = c->oc2;
    return onchip_staging_fits(c->n, d.l_slots, d.rec_slots) &&
           (f64 ? onchip64_smem_bytes(d.l_slots, d.rec_slots, d.groups_cn) : onchip_smem_bytes(d.l_slots, d.rec_slots, d.groups_cn)) <= (size_t)dev_smem;
}

// float32 launches: 8-byte records when every row fits ONE of them (at most 27 edges: the alist n = 10k codes of
// config 10k NMSA.json) -- three 512-thread CTAs per SM instead of two of 768 and half the record bytes per gather: A79 NMSA @
// 2 % 9.23 -> 9.56 Gbit/s. With two records per row (28..51 edges) the format loses: on the irregular R = 0.8 code every row
// splits, the check phase grows by a quarter and the 32-bit index entries eat half of what the variable phase saves (1.042 ->
// 0.968 Gbit/s, profiles/r02_z_rec8.md), so that case runs only on request (qkdldpc_options.onchip_record_bytes = 8).
static bool use_rec8(const qkdldpc_code *c) {
    int dev_smem = 0;
    if (c->opt.onchip_record_bytes == 16 || !c->oc2r8.eligible || (c->opt.onchip_record_bytes != 8 && c->oc2r8.max_dc > c->oc2r8.prm.rec_cap) ||
        cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device) != cudaSuccess)
        return false;
    const Oc2Device &d = c->oc2r8;
    return onchip_staging_fits(c->n, d.l_slots, d.rec_slots, 12) && onchip_smem_bytes(d.l_slots, d.rec_slots, d.groups_cn, 12) <= (size_t)dev_smem;
}

// threads == 0: pick the CTA size that puts the most warps on an SM (shared memory decides how many CTAs fit; ties go
// to the smaller CTA -- measured: 3 x 512 beats 2 x 768 on n=10240 m=2048, 2 x 768 beats 2 x 512 on m=2201). A CTA
// never has more lanes than the check phase has rows.
typedef void (*OnchipKernel)(const OnchipArgs);

template <int ALG, bool WIDE>
static OnchipKernel kernel_of(bool f64, bool vt16, bool rec8) {
    if (f64) return (OnchipKernel)onchip_minsum64_kernel<ALG, WIDE>;
    if (rec8) return vt16 ? (OnchipKernel)onchip_minsum_kernel<ALG, WIDE, true, true> : (OnchipKernel)onchip_minsum_kernel<ALG, WIDE, false, true>;
    return vt16 ? (OnchipKernel)onchip_minsum_kernel<ALG, WIDE, true, false> : (OnchipKernel)onchip_minsum_kernel<ALG, WIDE, false, false>;
}
static OnchipKernel kernel_of(int alg, bool wide, bool f64, bool vt16, bool rec8) {
    switch (alg) {
        case 2: return wide ? kernel_of<2, true>(f64, vt16, rec8) : kernel_of<2, false>(f64, vt16, rec8);
        case 3: return wide ? kernel_of<3, true>(f64, vt16, rec8) : kernel_of<3, false>(f64, vt16, rec8);
        case 4: return wide ? kernel_of<4, true>(f64, vt16, rec8) : kernel_of<4, false>(f64, vt16, rec8);
        default: return wide ? kernel_of<5, true>(f64, vt16, rec8) : kernel_of<5, false>(f64, vt16, rec8);
    }
}

static cudaError_t pick_geometry(OnchipKernel kern, int max_threads, int m, int sms, size_t smem, long long n_frames, int *threads, int *grid) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (*threads == 0) {
        const int cap = std::max(128, std::min(max_threads, (m + 31) / 32 * 32));
        int best = 0;
        for (int t = 128; t <= cap; t += 128) {
            int k = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, kern, t, smem);
            if (e != cudaSuccess) return e;
            if (k * t > best) {
                best = k * t;
                *threads = t;
            }
        }
        if (*threads == 0) return cudaErrorLaunchOutOfResources;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, *threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid = (int)std::min<long long>(n_frames, (long long)per_sm * sms);   // persistent CTAs pull frames from a queue
    return cudaSuccess;
}

// Variable-phase schedule for `nwarps` warps per CTA: warp w handles entries w, w + nwarps, ... of the returned list.
// Groups differ in cost (degree 2 ... 73), so they are dealt longest-processing-time-first to the least loaded warp
// (round-robin over the degree-sorted list leaves the busiest warp ~10 % above the mean on the irregular codes).
static std::vector<int> vn_schedule(const std::vector<int> &group_degree, int nwarps, const std::vector<int> *cost = nullptr) {
    std::vector<std::vector<int>> per_warp(nwarps);
    std::vector<long long> load(nwarps, 0);
    std::vector<int> order(group_degree.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    auto weight = [&](int g) { return cost ? (*cost)[g] : group_degree[g] + 3; };   // degree + per-group overhead (header, LLR, store)
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return weight(x) > weight(y); });
    for (int g : order) {
        const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        per_warp[w].push_back(g);
        load[w] += weight(g);
    }
    size_t rounds = 0;
    for (auto &v : per_warp) rounds = std::max(rounds, v.size());
    std::vector<int> sched(rounds * nwarps, -1);
    for (int w = 0; w < nwarps; ++w)
        for (size_t i = 0; i < per_warp[w].size(); ++i) sched[i * nwarps + w] = per_warp[w][i];
    return sched;
}

// Packed punctured / shortened masks of one combination into dst[0 .. 2*words): [punctured | shortened without punctured].
int onchip_pack_masks(int n, const int32_t *punct, int n_punct, const int32_t *shortd, int n_short, uint32_t *dst) {
    const int words = (n + 31) / 32;
    for (int i = 0; i < n_punct; ++i) {
        if (punct[i] < 0 || punct[i] >= n) return fail(QKDLDPC_ERR_INVALID, "punctured position out of range");
        dst[punct[i] >> 5] |= 1u << (punct[i] & 31);
    }
    for (int i = 0; i < n_short; ++i) {
        if (shortd[i] < 0 || shortd[i] >= n) return fail(QKDLDPC_ERR_INVALID, "shortened position out of range");
        if (!((dst[shortd[i] >> 5] >> (shortd[i] & 31)) & 1u))   // in both lists: punctured wins (:1150 is tested first)
            dst[(size_t)words + (shortd[i] >> 5)] |= 1u << (shortd[i] & 31);
    }
    return QKDLDPC_OK;
}

// Bookkeeping after the (single) kernel of an on-chip batch has been enqueued on c->stream after ev0.
static int onchip_finish(qkdldpc_code *c, int grid, int threads, size_t smem) {
    c->kernel_launches += 1;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_batch_ms = ms;
    c->last_cn_ms = c->last_vn_ms = c->last_sched_ms = 0;
    if (c->profiling && c->oc2_phase_clk.p) {   // share of the CTAs' clocks spent in the two phases, scaled to the batch time
        unsigned long long clk[4] = {0, 0, 0, 0};
        CK(cudaMemcpy(clk, c->oc2_phase_clk.p, sizeof clk, cudaMemcpyDeviceToHost));
        if (clk[2] > 0) {
            c->last_cn_ms = ms * (double)clk[0] / (double)clk[2];
            c->last_vn_ms = ms * (double)clk[1] / (double)clk[2];
        }
    }
    c->last_path = 2;
    c->frames_per_tile = 1;
    c->oc_threads = threads;
    c->pool_tiles = grid;
    c->pool_bytes = (int64_t)grid * (int64_t)smem;   // bytes of on-chip decoder state in flight
    CK(cudaGetLastError());
    return QKDLDPC_OK;
}

// The launch(es) of one on-chip batch on c->stream. Without a HostPipe: ONE persistent kernel over frames that are already
// in device memory. With one: the batch is cut into pipe->chunks pieces, each its own launch with its own frame queue, on
// two alternating compute streams -- so the CTAs of piece k+1 take over the SM slots that the draining piece k frees (no
// tail between pieces) -- with the piece's packed keys copied in on a copy stream before it and its results copied out on a
// second copy stream after it. Tallies are atomic adds into one vector, so the pieces share it.
template <typename Launch>
static int onchip_launch_all(qkdldpc_code *c, const OnchipArgs &a, int grid, int threads, size_t smem, int64_t n_frames, const HostPipe *pipe,
                             Launch launch) {
    cudaStream_t s = c->stream;
    const int K = pipe ? std::max(1, std::min(pipe->chunks, kMaxPipeChunks)) : 1;
    if (!pipe) {
        CK(cudaEventRecord(c->ev0, s));
        const cudaError_t e = launch(a, grid, s);
        if (e != cudaSuccess) return fail(QKDLDPC_ERR_CUDA, "on-chip kernel launch failed: %s", cudaGetErrorString(e));
        return onchip_finish(c, grid, threads, smem);
    }
    while (c->pipe_ev.size() < (size_t)2 * K) {
        cudaEvent_t ev = nullptr;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->pipe_ev.push_back(ev);
    }
    cudaStream_t s_in = c->side_streams[0], s_out = c->side_streams[1], s_k[2] = {s, c->side_streams[2]};
    const size_t words = (size_t)a.words;
    CK(cudaEventRecord(c->ev0, s));
    CK(cudaEventRecord(c->ev_fork, s));              // table uploads / counter resets on s come first everywhere
    CK(cudaStreamWaitEvent(s_in, c->ev_fork, 0));
    CK(cudaStreamWaitEvent(s_k[1], c->ev_fork, 0));
    for (int k = 0; k < K; ++k) {
        const int64_t f0 = n_frames * k / K, f1 = n_frames * (k + 1) / K, cnt = f1 - f0;
        if (cnt == 0) continue;
        uint32_t *d_al = const_cast<uint32_t *>(a.alice_bits) + f0 * words, *d_bo = const_cast<uint32_t *>(a.bob_bits) + f0 * words;
        CK(cudaMemcpyAsync(d_al, pipe->h_alice + f0 * words, (size_t)cnt * words * 4, cudaMemcpyHostToDevice, s_in));
        CK(cudaMemcpyAsync(d_bo, pipe->h_bob + f0 * words, (size_t)cnt * words * 4, cudaMemcpyHostToDevice, s_in));
        CK(cudaEventRecord(c->pipe_ev[2 * k], s_in));
        cudaStream_t st = s_k[k & 1];
        CK(cudaStreamWaitEvent(st, c->pipe_ev[2 * k], 0));
        OnchipArgs ak = a;
        ak.n_frames = cnt;
        ak.alice_bits = d_al;
        ak.bob_bits = d_bo;
        if (!a.qber_is_scalar && a.qber) ak.qber = a.qber + f0;
        if (a.out_bits) ak.out_bits = a.out_bits + f0 * words;
        if (a.out_iters) ak.out_iters = a.out_iters + f0;
        if (a.out_flags) ak.out_flags = a.out_flags + f0;
        ak.next_frame = c->counters.p + 2 + k;
        ak.frames_per_combo = n_frames;               // a pipelined batch is a single combination: f / frames_per_combo == 0
        const cudaError_t e = launch(ak, (int)std::min<int64_t>(cnt, grid), st);
        if (e != cudaSuccess) {
            cudaDeviceSynchronize();
            return fail(QKDLDPC_ERR_CUDA, "on-chip kernel launch failed: %s", cudaGetErrorString(e));
        }
        c->kernel_launches += 1;
        CK(cudaEventRecord(c->pipe_ev[2 * k + 1], st));
        CK(cudaStreamWaitEvent(s_out, c->pipe_ev[2 * k + 1], 0));
        if (pipe->h_out_bits && a.out_bits)
            CK(cudaMemcpyAsync(pipe->h_out_bits + f0 * words, ak.out_bits, (size_t)cnt * words * 4, cudaMemcpyDeviceToHost, s_out));
        if (pipe->h_out_iters && a.out_iters)
            CK(cudaMemcpyAsync(pipe->h_out_iters + f0, ak.out_iters, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, s_out));
        if (pipe->h_out_flags && a.out_flags) CK(cudaMemcpyAsync(pipe->h_out_flags + f0, ak.out_flags, (size_t)cnt, cudaMemcpyDeviceToHost, s_out));
    }
    // join: the main stream waits for the other compute stream and for the last copy-out
    CK(cudaEventRecord(c->ev_join[0], s_k[1]));
    CK(cudaStreamWaitEvent(s, c->ev_join[0], 0));
    CK(cudaEventRecord(c->ev_join[1], s_out));
    CK(cudaStreamWaitEvent(s, c->ev_join[1], 0));
    c->kernel_launches -= 1;   // onchip_finish counts one
    return onchip_finish(c, grid, threads, smem);
}

// One launch over n_combos x frames_per_combo frames. `combos` / `masks` are HOST tables ([n_combos], [n_combos][2][words]);
// d_tally holds n_combos tally vectors.
int run_onchip_multi(qkdldpc_code *c, const qkdldpc_params *P, int n_combos, int64_t frames_per_combo, const OnchipCombo *combos,
                     const uint32_t *masks, const uint32_t *d_alice, const uint32_t *d_bob, const double *d_qber, int qber_is_scalar,
                     uint32_t *d_out_bits, int32_t *d_out_iters, uint8_t *d_out_flags, unsigned long long *d_tally, const HostPipe *pipe) {
    const int n = c->n, m = c->m, words = (n + 31) / 32;
    const int64_t n_frames = (int64_t)n_combos * frames_per_combo;
    const int tl = (int)qkdldpc_tally_len(P->max_iterations);
    cudaStream_t s = c->stream;
    CK(c->oc_cls.reserve((size_t)n_combos * 2 * words));
    CK(c->oc_combos.reserve((size_t)n_combos * sizeof(OnchipCombo)));
    CK(c->counters.reserve(2 + kMaxPipeChunks));   // [0] frame queue of a single launch, [2 + k] queue of pipeline chunk k
    CK(cudaMemcpyAsync(c->oc_cls.p, masks, (size_t)n_combos * 2 * words * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->oc_combos.p, combos, (size_t)n_combos * sizeof(OnchipCombo), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));   // the caller's host tables may die after this call returns early on an error below
    CK(cudaMemsetAsync(c->counters.p, 0, (2 + kMaxPipeChunks) * sizeof(unsigned long long), s));
    if (d_tally) CK(cudaMemsetAsync(d_tally, 0, (size_t)n_combos * tl * sizeof(uint64_t), s));

    OnchipArgs a{};
    a.n = n; a.m = m; a.words = words; a.rec_slots = c->oc_rec_slots;
    a.n_groups_cn = c->oc_groups_cn;
    a.cn_ginfo = c->oc_cn_ginfo.p; a.cn_row = c->oc_cn_row.p; a.cnT = c->oc_cnT.p;
    a.combos = reinterpret_cast<const OnchipCombo *>(c->oc_combos.p); a.frames_per_combo = frames_per_combo;
    a.cls_masks = c->oc_cls.p; a.tally_len = tl;
    a.n_frames = n_frames; a.alice_bits = d_alice; a.bob_bits = d_bob; a.qber = d_qber; a.qber_is_scalar = qber_is_scalar;
    a.out_bits = d_out_bits; a.out_iters = d_out_iters; a.out_flags = d_out_flags; a.tally = d_tally;
    a.next_frame = c->counters.p;
    a.max_iter = P->max_iterations;
    a.thr = P->enable_threshold ? (float)P->threshold : INFINITY;
    a.thr64 = P->enable_threshold ? P->threshold : (double)INFINITY;

    const bool spa = P->algorithm < 2, f64 = P->message_precision == 64;
    const bool rec8 = !spa && !f64 && use_rec8(c);
    Oc2Device &d = rec8 ? c->oc2r8 : c->oc2;      // layout of this launch (min-sum kernels)
    const int max_threads = (spa || f64) ? 1024 : 768;
    int threads = c->opt.onchip_threads > 0 ? std::max(32, std::min(max_threads, c->opt.onchip_threads / 32 * 32)) : 0;   // 0 = auto
    const size_t smem = spa   ? onchip_spa_smem_bytes(n, c->sp_msg_words, c->oc_groups_cn, c->sp_groups_sv)
                        : f64 ? onchip64_smem_bytes(d.l_slots, d.rec_slots, d.groups_cn)
                              : onchip_smem_bytes(d.l_slots, d.rec_slots, d.groups_cn, rec8 ? 12 : 16);
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));

    int grid = 0;
    cudaError_t e;
    if (spa) {
        e = onchip_spa_geometry(P->algorithm, c->oc_groups_cn, sms, smem, n_frames, &threads, &grid);
        if (e != cudaSuccess) return fail(QKDLDPC_ERR_CUDA, "on-chip sum-product kernel geometry failed: %s", cudaGetErrorString(e));
        if (c->sp_chunk_warps != threads / 32) {
            // variable phase: warp w owns a contiguous run of items; boundaries fall between groups, runs balanced greedily
            const int nw = threads / 32, groups = (int)c->sp_group_item0.size() - 1, total = c->sp_group_item0.back();
            std::vector<int> chunk(nw + 1, total);   // chunk[nw] = total
            chunk[0] = 0;
            int g = 0;   // sp_group_item0[g] <= target < sp_group_item0[g + 1]
            for (int w = 1; w < nw; ++w) {
                const int target = (int)((long long)total * w / nw);
                while (g + 1 < groups && c->sp_group_item0[g + 1] <= target) ++g;
                const int lo = c->sp_group_item0[g], hi = c->sp_group_item0[g + 1];
                chunk[w] = std::max(chunk[w - 1], (target - lo <= hi - target) ? lo : hi);   // the nearer group boundary
            }
            CK(c->sp_sv_chunk.reserve(chunk.size()));
            CK(cudaMemcpyAsync(c->sp_sv_chunk.p, chunk.data(), chunk.size() * sizeof(int), cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));
            c->sp_chunk_warps = nw;
        }
        a.cn_moff = c->sp_cn_moff.p; a.sv_items = c->sp_sv_items.p; a.sv_chunk = c->sp_sv_chunk.p; a.msg_words = c->sp_msg_words;
        a.sv_group_item0 = c->sp_sv_group_item0.p; a.n_groups_sv = c->sp_groups_sv;
        auto launch_spa = [&](const OnchipArgs &args, int g, cudaStream_t st) { return onchip_spa_launch(P->algorithm, args, g, threads, smem, st); };
        return onchip_launch_all(c, a, grid, threads, smem, n_frames, pipe, launch_spa);
    }
    const bool wide = d.max_dc > d.prm.rec_cap;   // rows of two records: separate kernel instantiation
    const bool vt16 = !f64 && d.vT16.p != nullptr;   // 16-bit variable-phase entries (codes with at most 2048 records)
    const OnchipKernel kern = kernel_of(P->algorithm, wide, f64, vt16, rec8);
    e = pick_geometry(kern, max_threads, m, sms, smem, n_frames, &threads, &grid);
    if (e != cudaSuccess) return fail(QKDLDPC_ERR_CUDA, "on-chip kernel geometry failed: %s", cudaGetErrorString(e));
    {
        // the tables of onchip_layout.hpp; the canonical variable-phase groups are dealt to the warps of this CTA size
        // (float32 and float64 launches keep their own deal: 16 and 32 warps per CTA on the n = 10k codes)
        int &sched_warps = d.sched_warps[f64];
        DevBuf<int4> &vn_g = d.vn_g[f64];
        DevBuf<int> &vn_start = d.vn_start[f64];
        if (sched_warps != threads / 32) {
            const int nw = threads / 32;
            std::vector<int> degree;
            for (const Oc2Group &g : d.vn_g_host) degree.push_back(g.deg);
            const std::vector<int> sched = vn_schedule(degree, nw, &d.vn_gcost);   // entry i belongs to warp i % nw; -1 = none
            std::vector<Oc2Group> dealt;
            std::vector<int> start(nw + 1, 0);
            for (int w = 0; w < nw; ++w) {
                for (size_t i = (size_t)w; i < sched.size(); i += (size_t)nw)
                    if (sched[i] >= 0) dealt.push_back(d.vn_g_host[sched[i]]);
                start[w + 1] = (int)dealt.size();
            }
            CK(vn_g.reserve(dealt.size()));
            CK(vn_start.reserve(start.size()));
            CK(cudaMemcpyAsync(vn_g.p, dealt.data(), dealt.size() * sizeof(Oc2Group), cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(vn_start.p, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));
            sched_warps = nw;
        }
        // punctured / shortened masks of every combination into slot order
        const int swords = d.l_slots / 32;
        std::vector<uint32_t> masks2((size_t)n_combos * 2 * swords, 0u);
        for (int cb = 0; cb < n_combos; ++cb) {
            if (!combos[cb].has_cls) continue;
            for (int h = 0; h < 2; ++h) {
                const uint32_t *src = masks + ((size_t)cb * 2 + h) * words;
                uint32_t *dst = masks2.data() + ((size_t)cb * 2 + h) * swords;
                for (int w = 0; w < words; ++w)
                    for (uint32_t bitsw = src[w]; bitsw; bitsw &= bitsw - 1) {
                        const int sl = d.bit_slot_host[(size_t)w * 32 + __builtin_ctz(bitsw)];
                        dst[sl >> 5] |= 1u << (sl & 31);
                    }
            }
        }
        CK(c->oc2_cls.reserve(masks2.size()));
        CK(cudaMemcpyAsync(c->oc2_cls.p, masks2.data(), masks2.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        a.rec_slots = d.rec_slots;
        a.n_groups_cn2 = d.groups_cn; a.l_slots = d.l_slots;
        // float64: the same check-phase table with byte offsets of 8-byte totals
        a.cn_g2 = d.cn_g.p; a.cnT2 = f64 ? d.cnT64.p : d.cnT.p; a.vn_g2 = vn_g.p; a.vn_start = vn_start.p; a.vT2 = d.vT.p;
        a.vT16 = vt16 ? d.vT16.p : nullptr;
        a.slot_bit = d.slot_bit.p; a.bit_slot = d.bit_slot.p; a.cls_masks2 = c->oc2_cls.p;
        if (!f64) c->last_rec_bytes = rec8 ? 8 : 16;
        if (c->profiling) {
            CK(c->oc2_phase_clk.reserve(4));
            CK(cudaMemsetAsync(c->oc2_phase_clk.p, 0, 4 * sizeof(unsigned long long), s));
            a.phase_clk = c->oc2_phase_clk.p;
        }
    }

    auto launch_ms = [&](const OnchipArgs &args, int g, cudaStream_t st) {
        kern<<<(unsigned)g, threads, smem, st>>>(args);
        return cudaGetLastError();
    };
    return onchip_launch_all(c, a, grid, threads, smem, n_frames, pipe, launch_ms);
}

// The single-combination call of qkdldpc_decode_batch_device.
int run_onchip(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_frames, const uint32_t *d_alice, const uint32_t *d_bob,
               const double *d_qber, int qber_is_scalar, const int32_t *punct, int n_punct, const int32_t *shortd, int n_short,
               uint32_t *d_out_bits, int32_t *d_out_iters, uint8_t *d_out_flags, unsigned long long *d_tally, const HostPipe *pipe) {
    const int words = (c->n + 31) / 32;
    std::vector<uint32_t> masks((size_t)2 * words, 0u);
    const int rc = onchip_pack_masks(c->n, punct, n_punct, shortd, n_short, masks.data());
    if (rc) return rc;
    OnchipCombo cb{};
    cb.qber = -1.;   // LLR magnitude from the caller's qber array
    cb.primary = P->primary;
    cb.secondary = P->secondary;
    cb.has_cls = (n_punct > 0 || n_short > 0) ? 1 : 0;
    return run_onchip_multi(c, P, 1, n_frames, &cb, masks.data(), d_alice, d_bob, d_qber, qber_is_scalar, d_out_bits, d_out_iters,
                            d_out_flags, d_tally, pipe);
}

}  // namespace qkhost

// Host-side handle of libqkdldpc_cuda (shared by api.cu and the per-precision instantiation units).
#pragma once
#include "../../include/qkdldpc.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "onchip_layout.hpp"

namespace qkhost {

inline std::string &last_error() {
    static thread_local std::string e;
    return e;
}

inline int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? QKDLDPC_ERR_NOMEM : QKDLDPC_ERR_CUDA,          \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);     \
    } while (0)

template <typename U>
struct DevBuf {
    U *p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(U));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Precondition of the branch-free min-sum arithmetic (FAST streaming kernels, both on-chip min-sum kernels): no message
// can become NaN or infinite -- a finite clamp guarantees it; without the clamp it holds when no factor can scale a
// message up (offset variants subtract; normalized variants need factors <= 1) -- and every factor in use is finite and
// non-negative: the kernels clamp magnitudes with min(c, thr), which would miss the -thr side of threshold_matrix
// (array_and_matrix_operations.cpp:953-972) for a negative factor. Anything else takes the EXACT streaming kernels.
inline bool minsum_factors_ok(const qkdldpc_params *P) {
    if (P->algorithm < 2) return false;
    const bool two = P->algorithm >= 4;
    if (!(P->primary >= 0) || !std::isfinite(P->primary)) return false;
    if (two && (!(P->secondary >= 0) || !std::isfinite(P->secondary))) return false;
    if (P->enable_threshold) return std::isfinite(P->threshold);
    if (P->algorithm == 3 || P->algorithm == 5) return true;
    return P->primary <= 1.0 && (!two || P->secondary <= 1.0);
}

constexpr int kSideStreams = 3;

// Host side of a pipelined on-chip batch (qkdldpc_decode_batch with host buffers): the batch is cut into `chunks` pieces;
// the host-to-device copy of piece k+1 and the device-to-host copy of piece k-1 run on copy streams while piece k decodes.
struct HostPipe {
    const uint32_t *h_alice, *h_bob;
    uint32_t *h_out_bits;
    int32_t *h_out_iters;
    uint8_t *h_out_flags;
    int chunks;
};
constexpr int kMaxPipeChunks = 16;

struct EvPair {
    cudaEvent_t a, b;
    int kind;   // 0 CN, 1 VN, 2 sched
};

}  // namespace qkhost
using namespace qkhost;

namespace qkhost {
// Host tables of the on-chip kernels as built by build_onchip_tables (onchip_tables.cu).
struct OnchipTables {
    bool oc_ok = false, sp_ok = false;   // min-sum / sum-product kernel can address this graph
    int max_dc = 0, rec_slots = 0, sp_msg_words = 0;
    std::vector<int> slot0;              // first record slot of every row (+ total)
    std::vector<int2> cn_ginfo;          // check-phase groups in natural order: the sum-product kernel's (onchip_spa.cuh)
    std::vector<uint16_t> cn_row;
    std::vector<uint2> cnT;
    std::vector<int> sp_cn_moff, sp_group_item0;
    std::vector<uint4> sp_items;
    Oc2Tables oc2;                       // min-sum kernels (float32 / float64 state): storage order = processing order (onchip_layout.hpp)
    Oc2Tables oc2r8;                     // the same for the float32 kernel's 8-byte records (Oc2Params::rec8)
};

// Device copy of one Oc2Tables (+ what the launcher needs of it on the host).
struct Oc2Device {
    bool eligible = false;               // the kernels can address the graph with this layout
    Oc2Params prm;
    int groups_cn = 0, l_slots = 0, rec_slots = 0, max_dc = 0;
    DevBuf<int4> cn_g;
    DevBuf<uint4> cnT, cnT64, vT;        // cnT64: byte offsets of 8-byte totals (float64 kernel; 16-byte records only)
    DevBuf<uint2> vT16;                  // 16-bit variable-phase entries, only for codes with at most 2048 records
    DevBuf<uint16_t> slot_bit, bit_slot;
    // the variable-phase groups dealt to the warps of a CTA: [0] float32 launches, [1] float64 launches (different CTA sizes)
    int sched_warps[2] = {0, 0};
    DevBuf<int4> vn_g[2];
    DevBuf<int> vn_start[2];
    std::vector<Oc2Group> vn_g_host;     // canonical order
    std::vector<int> vn_gcost;           // bank-model cost per group: the weight of that deal
    std::vector<uint16_t> bit_slot_host;
    long long model[4] = {0, 0, 0, 0};   // bank model: check gather, its minimum, variable gather, its minimum (wavefronts / iteration)
    cudaError_t upload(const Oc2Tables &T, bool with_f64) {
        eligible = false;
        prm = T.prm;
        if (!T.ok) return cudaSuccess;
        auto up = [](auto &buf, const auto &vec) -> cudaError_t {
            cudaError_t e = buf.reserve(vec.size());
            if (e != cudaSuccess) return e;
            return cudaMemcpy(buf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice);
        };
        static_assert(sizeof(Oc2Group) == sizeof(int4) && sizeof(Oc2U2) == sizeof(uint2) && sizeof(Oc2U4) == sizeof(uint4), "table layouts");
        bit_slot_host.assign(T.bit_slot.begin(), T.bit_slot.end());
        cudaError_t e;
        if ((e = up(cn_g, T.cn_g)) || (e = up(cnT, T.cnT)) || (e = up(vT, T.vT)) || (T.vt16_ok && (e = up(vT16, T.vT16))) ||
            (e = up(slot_bit, T.slot_bit)) || (e = up(bit_slot, bit_slot_host)))
            return e;
        if (with_f64) {
            std::vector<Oc2U4> t64(T.cnT);   // the same offsets for 8-byte totals
            for (Oc2U4 &w : t64) { w.x *= 2; w.y *= 2; w.z *= 2; w.w *= 2; }
            if ((e = up(cnT64, t64))) return e;
        }
        groups_cn = (int)T.cn_g.size();
        rec_slots = T.rec_slots;
        l_slots = T.l_slots;
        max_dc = T.max_dc;
        vn_g_host = T.vn_g;
        vn_gcost = T.vn_gcost;
        model[0] = T.cn_gather; model[1] = T.cn_gather_min; model[2] = T.vn_gather; model[3] = T.vn_gather_min;
        eligible = true;
        return cudaSuccess;
    }
    void release() {
        cn_g.release(); cnT.release(); cnT64.release(); vT.release(); vT16.release(); slot_bit.release(); bit_slot.release();
        for (int k = 0; k < 2; ++k) { vn_g[k].release(); vn_start[k].release(); }
    }
};
}  // namespace qkhost

struct qkdldpc_code {
    int n = 0, m = 0, device = 0;
    int64_t nnz = 0;
    qkdldpc_options opt{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side_streams[kSideStreams] = {nullptr, nullptr, nullptr};   // concurrent VN bucket kernels (run_batch.cuh)
    cudaEvent_t ev_fork = nullptr, ev_join[kSideStreams] = {nullptr, nullptr, nullptr};
    // graph (device)
    DevBuf<int> row_ptr, col_idx, col_ptr, csc_edge, csc_row, row_order, col_order;
    DevBuf<int> vn_ell_edge, vn_ell_row;   // ELL records of the two narrow VN buckets (step_kernels.cuh: vn_kernel_ell)
    int cn_first[5] = {0}, cn_count[5] = {0}, vn_first[5] = {0}, vn_count[5] = {0};   // degree buckets in row/col_order
    // on-chip paths: check-phase groups in natural order (ELL index arrays per 32-row group) for the sum-product kernel
    int oc_groups_cn = 0, oc_max_dc = 0, oc_rec_slots = 0;
    DevBuf<int2> oc_cn_ginfo;
    DevBuf<uint16_t> oc_cn_row;
    DevBuf<uint2> oc_cnT;
    DevBuf<uint32_t> oc_cls;          // [n_combos][2][words] punctured / shortened bit masks of the current batch, natural order
    DevBuf<unsigned char> oc_combos;  // OnchipCombo table of the current launch
    int oc_threads = 0;               // CTA size of the last on-chip launch
    // min-sum kernels: tables of onchip_layout.hpp -- 16-byte records (float32 and float64 state) and 8-byte records (float32)
    Oc2Device oc2, oc2r8;
    DevBuf<unsigned long long> oc2_phase_clk;   // profiling: clocks per phase of the last on-chip min-sum launch
    DevBuf<uint32_t> oc2_cls;         // [n_combos][2][l_slots/32] punctured / shortened masks of the current batch, slot order
    mutable int sm_count = 0;         // multiprocessors of `device`, read on first use (run_batch.cuh: vn_loop_plan)
    int64_t tail_compactions = 0;     // streaming path
    int last_spp = 0, last_vn_items = 0;
    int last_rec_bytes = 0;           // 16 / 8: record format of the last float32 on-chip min-sum launch
    // on-chip sum-product path (onchip_spa.cuh): one message word per edge; check phase shares oc_cn_* with min-sum
    bool sp_eligible = false;
    int sp_msg_words = 0, sp_chunk_warps = 0;
    DevBuf<int> sp_cn_moff, sp_sv_chunk, sp_sv_group_item0;
    int sp_groups_sv = 0;
    DevBuf<uint4> sp_sv_items;
    std::vector<int> sp_group_item0;  // first item of every variable-phase group (+ total): chunk boundaries lie on these
    void *comm = nullptr;             // ncclComm_t of the tally all-reduce (comm.cu); owned by the handle
    int comm_ranks = 0;
    DevBuf<unsigned long long> comm_buf;
    int last_path = 0;                // 1 streaming, 2 on-chip (of the last batch)
    int last_precision = 0;           // 32 / 64: message precision of the last batch after the policy
    // pool (device, raw bytes reinterpreted per precision)
    DevBuf<unsigned char> msg;
    DevBuf<uint32_t> bobmask, zmask, synd, par, tile_active, tile_new;
    DevBuf<unsigned char> slot_llr;
    DevBuf<long long> slot_frame;
    DevBuf<int32_t> slot_iter;
    // batch (device)
    DevBuf<unsigned char> frame_llr;
    DevBuf<uint32_t> synd_all;
    DevBuf<uint8_t> bitclass;
    DevBuf<unsigned long long> counters;   // [0] next_frame, [1] n_done
    // staging for the host-pointer entry point
    DevBuf<uint32_t> st_alice, st_bob, st_out;
    DevBuf<double> st_qber;
    DevBuf<int32_t> st_iters;
    DevBuf<uint8_t> st_flags;
    DevBuf<unsigned long long> st_tally;
    // reference-compatible trial-input generator (gen_kernels.cuh)
    DevBuf<uint64_t> gen_seeds;
    DevBuf<unsigned char> gen_combos;  // RefKeygenCombo table of the current launch
    DevBuf<uint32_t> gen_masks, gen_scratch;
    DevBuf<int> rb_kept;               // remove_bits: surviving positions
    DevBuf<unsigned char> rb_info;     // RemoveCombo table of the current launch (gen_kernels.cuh)
    DevBuf<uint32_t> st_keys_a, st_keys_b;   // final keys after remove_bits (Alice / Bob)
    DevBuf<unsigned char> sched_work;  // TileWork per tile (sched_kernels.cuh)
    DevBuf<int2> compact_moves;        // tail compaction (sched_kernels.cuh)
    DevBuf<int> compact_plan;
    unsigned long long *h_done = nullptr;   // pinned [2]: frames handed out, frames finished
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_poll = nullptr;
    // captured step graph
    // captured step graphs, newest last: a batch uses one for its full pool and one per tile count the tail compaction
    // leaves (quantised, run_batch.cuh), and the next batch of the same combination finds them again
    struct StepGraph {
        std::string key;
        cudaGraphExec_t exec;
    };
    std::vector<StepGraph> graphs;
    static constexpr size_t kMaxGraphs = 24;
    void drop_graphs() {
        for (auto &g : graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        graphs.clear();
    }
    // stats
    int frames_per_tile = 0, pool_tiles = 0;
    int64_t pool_bytes = 0, kernel_launches = 0, decoder_steps = 0;
    double last_batch_ms = 0, last_cn_ms = 0, last_vn_ms = 0, last_sched_ms = 0;
    bool profiling = false;
    std::vector<EvPair> ev_pool;
    std::vector<cudaEvent_t> pipe_ev;   // pipelined host batches: [2k] copy-in of chunk k done, [2k+1] kernel of chunk k done
    size_t ev_used = 0;
};


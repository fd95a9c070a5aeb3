// libqkdldpc_cuda: C ABI (include/qkdldpc.h) over the sm_100a kernels. Host side: graph re-layout (K0), slot-pool
// management and dispatch to the per-precision step loops (run_batch.cuh).
#include "handle.hpp"
#include "common.cuh"
#include "gen_kernels.cuh"
#include "onchip_minsum.cuh"

namespace qkhost {
int build_onchip_tables(int n, int m, long long nnz, const std::vector<int> &rp, const int *col_idx, const std::vector<int> &col_ptr,
                        const std::vector<int> &csc_edge, const std::vector<int> &csc_row, OnchipTables &T);
template <typename T, int V>
int run_batch(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_frames, const uint32_t *d_alice,
              const uint32_t *d_bob, const double *d_qber, int qber_is_scalar, const int32_t *punct, int n_punct,
              const int32_t *shortd, int n_short, uint32_t *d_out_bits, int32_t *d_out_iters, uint8_t *d_out_flags,
              unsigned long long *d_tally);
extern template int run_batch<float, 4>(qkdldpc_code *, const qkdldpc_params *, int64_t, const uint32_t *, const uint32_t *, const double *, int, const int32_t *, int, const int32_t *, int, uint32_t *, int32_t *, uint8_t *, unsigned long long *);
extern template int run_batch<float, 2>(qkdldpc_code *, const qkdldpc_params *, int64_t, const uint32_t *, const uint32_t *, const double *, int, const int32_t *, int, const int32_t *, int, uint32_t *, int32_t *, uint8_t *, unsigned long long *);
extern template int run_batch<float, 1>(qkdldpc_code *, const qkdldpc_params *, int64_t, const uint32_t *, const uint32_t *, const double *, int, const int32_t *, int, const int32_t *, int, uint32_t *, int32_t *, uint8_t *, unsigned long long *);
extern template int run_batch<double, 2>(qkdldpc_code *, const qkdldpc_params *, int64_t, const uint32_t *, const uint32_t *, const double *, int, const int32_t *, int, const int32_t *, int, uint32_t *, int32_t *, uint8_t *, unsigned long long *);
int run_onchip(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_frames, const uint32_t *d_alice, const uint32_t *d_bob,
               const double *d_qber, int qber_is_scalar, const int32_t *punct, int n_punct, const int32_t *shortd, int n_short,
               uint32_t *d_out_bits, int32_t *d_out_iters, uint8_t *d_out_flags, unsigned long long *d_tally, const HostPipe *pipe);
bool onchip_usable(const qkdldpc_code *c, const qkdldpc_params *P);
void comm_release(qkdldpc_code *c);
int onchip_pack_masks(int n, const int32_t *punct, int n_punct, const int32_t *shortd, int n_short, uint32_t *dst);
int run_onchip_multi(qkdldpc_code *c, const qkdldpc_params *P, int n_combos, int64_t frames_per_combo, const qk::OnchipCombo *combos,
                     const uint32_t *masks, const uint32_t *d_alice, const uint32_t *d_bob, const double *d_qber, int qber_is_scalar,
                     uint32_t *d_out_bits, int32_t *d_out_iters, uint8_t *d_out_flags, unsigned long long *d_tally, const HostPipe *pipe);
}  // namespace qkhost

namespace {

int check_params(const qkdldpc_code *c, const qkdldpc_params *P, int64_t n_frames) {
    if (!c) return fail(QKDLDPC_ERR_INVALID, "null code handle");
    if (!P) return fail(QKDLDPC_ERR_INVALID, "null params");
    if (P->algorithm < 0 || P->algorithm > 5) return fail(QKDLDPC_ERR_INVALID, "algorithm %d not in 0..5", P->algorithm);
    if (P->max_iterations < 1 || P->max_iterations > (1 << 20))
        return fail(QKDLDPC_ERR_INVALID, "max_iterations %d out of range", P->max_iterations);
    if (P->message_precision != 0 && P->message_precision != 32 && P->message_precision != 64)
        return fail(QKDLDPC_ERR_INVALID, "message_precision must be 0 (automatic), 32 or 64");
    if (P->enable_threshold && !(P->threshold > 0)) return fail(QKDLDPC_ERR_INVALID, "threshold must be > 0");
    if (n_frames < 0) return fail(QKDLDPC_ERR_INVALID, "negative frame count");
    return QKDLDPC_OK;
}

}  // namespace

extern "C" {

int qkdldpc_version(void) { return QKDLDPC_VERSION; }
const char *qkdldpc_last_error(void) { return last_error().c_str(); }
int64_t qkdldpc_tally_len(int32_t max_iterations) { return (int64_t)max_iterations + 5; }

int32_t qkdldpc_effective_precision(int32_t algorithm, int32_t n, int32_t message_precision) {
    if (message_precision != 0) return message_precision;
    if (algorithm >= 3) return 64;               // OMSA, ANMSA, AOMSA
    if (algorithm <= 1 && n > 65536) return 64;  // SPA / SPA-lin-approx on long codes
    return 32;
}

int qkdldpc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int qkdldpc_code_create(qkdldpc_code **out, int32_t n, int32_t m, int64_t nnz, const int32_t *row_ptr,
                        const int32_t *col_idx, int32_t device, const qkdldpc_options *options) {
    if (!out) return fail(QKDLDPC_ERR_INVALID, "null out pointer");
    *out = nullptr;
    if (n < 1 || m < 1 || nnz < 1 || !row_ptr || !col_idx) return fail(QKDLDPC_ERR_INVALID, "empty graph");
    if (row_ptr[0] != 0 || row_ptr[m] != nnz) return fail(QKDLDPC_ERR_INVALID, "row_ptr does not span [0, nnz]");
    // K0: validate (ascending, duplicate-free rows -- quirk Q1) and derive the column view
    std::vector<int> col_ptr(n + 1, 0);
    for (int j = 0; j < m; ++j) {
        if (row_ptr[j + 1] < row_ptr[j]) return fail(QKDLDPC_ERR_INVALID, "row_ptr not monotone at row %d", j);
        if (row_ptr[j + 1] == row_ptr[j]) return fail(QKDLDPC_ERR_INVALID, "check node %d has no bits", j);
        for (int e = row_ptr[j]; e < row_ptr[j + 1]; ++e) {
            const int c = col_idx[e];
            if (c < 0 || c >= n) return fail(QKDLDPC_ERR_INVALID, "column index %d out of range in row %d", c, j);
            if (e > row_ptr[j] && col_idx[e - 1] >= c)
                return fail(QKDLDPC_ERR_INVALID, "row %d is not strictly ascending (reference quirk Q1)", j);
            col_ptr[c + 1]++;
        }
    }
    for (int i = 0; i < n; ++i) {
        if (col_ptr[i + 1] == 0) return fail(QKDLDPC_ERR_INVALID, "bit node %d has no checks", i);
        col_ptr[i + 1] += col_ptr[i];
    }
    std::vector<int> csc_edge(nnz), csc_row(nnz), cur(col_ptr.begin(), col_ptr.end() - 1);
    for (int j = 0; j < m; ++j)
        for (int e = row_ptr[j]; e < row_ptr[j + 1]; ++e) {
            const int p = cur[col_idx[e]]++;
            csc_edge[p] = e;
            csc_row[p] = j;   // rows visited ascending => every column lists its checks ascending
        }
    // degree buckets: rows / columns grouped so that every kernel instantiation works inside one register-array size
    auto group = [](int count, const std::vector<int> &ptr, int (*bucket_of)(int), std::vector<int> &order, int *first,
                    int *cnt) {
        order.clear();
        for (int bk = 0; bk < qk::kBuckets; ++bk) {
            first[bk] = (int)order.size();
            for (int i = 0; i < count; ++i)
                if (bucket_of(ptr[i + 1] - ptr[i]) == bk) order.push_back(i);
            cnt[bk] = (int)order.size() - first[bk];
        }
    };
    std::vector<int> row_order, col_order, rp(row_ptr, row_ptr + m + 1);
    int cn_first[5], cn_count[5], vn_first[5], vn_count[5];
    group(m, rp, qk::cn_bucket_of, row_order, cn_first, cn_count);
    group(n, col_ptr, qk::vn_bucket_of, col_order, vn_first, vn_count);
    // ELL records of the two narrow variable-node buckets (dv <= 4, dv <= 8): the edge and check ids of item i of bucket b
    // at [base_b + i * W_b, + W_b), -1 padded, so that vn_kernel_ell needs no col_ptr / csc_edge / csc_row lookups
    std::vector<int> vn_ell_edge, vn_ell_row;
    for (int b = 0; b < 2; ++b) {
        const int W = qk::vn_bucket_max(b);
        for (int i = 0; i < vn_count[b]; ++i) {
            const int bit = col_order[vn_first[b] + i], dv = col_ptr[bit + 1] - col_ptr[bit];
            for (int k = 0; k < W; ++k) {
                vn_ell_edge.push_back(k < dv ? csc_edge[col_ptr[bit] + k] : -1);
                vn_ell_row.push_back(k < dv ? csc_row[col_ptr[bit] + k] : -1);
            }
        }
    }
    if (vn_ell_edge.empty()) { vn_ell_edge.assign(4, -1); vn_ell_row.assign(4, -1); }

    // tables of the two on-chip kernels + their self-check (onchip_tables.cu); host code, runs before the device is touched
    qkhost::OnchipTables T;
    {
        const int rc = qkhost::build_onchip_tables(n, m, nnz, rp, col_idx, col_ptr, csc_edge, csc_row, T);
        if (rc) return rc;
    }
    const bool oc_ok = T.oc_ok, sp_ok = T.sp_ok;

    int ndev = qkdldpc_device_count();
    if (ndev == 0) return fail(QKDLDPC_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(QKDLDPC_ERR_INVALID, "device %d not in [0, %d)", device, ndev);
    CK(cudaSetDevice(device));

    qkdldpc_code *c = new qkdldpc_code();
    c->n = n; c->m = m; c->nnz = nnz; c->device = device;
    if (options) c->opt = *options;
    auto up = [&](auto &buf, const auto &vec) -> cudaError_t {
        cudaError_t e = buf.reserve(vec.size());
        if (e != cudaSuccess) return e;
        return cudaMemcpy(buf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice);
    };
    std::vector<int> ci(col_idx, col_idx + nnz);
    cudaError_t e = cudaSuccess;
    if ((e = up(c->row_ptr, rp)) || (e = up(c->col_idx, ci)) || (e = up(c->col_ptr, col_ptr)) ||
        (e = up(c->csc_edge, csc_edge)) || (e = up(c->csc_row, csc_row)) || (e = up(c->row_order, row_order)) ||
        (e = up(c->col_order, col_order)) || (e = up(c->vn_ell_edge, vn_ell_edge)) || (e = up(c->vn_ell_row, vn_ell_row)) ||
        (oc_ok && ((e = up(c->oc_cn_ginfo, T.cn_ginfo)) || (e = up(c->oc_cnT, T.cnT)) || (e = up(c->oc_cn_row, T.cn_row)))) ||
        (e = c->oc2.upload(T.oc2, true)) || (e = c->oc2r8.upload(T.oc2r8, false)) ||
        (sp_ok && ((e = up(c->sp_cn_moff, T.sp_cn_moff)) || (e = up(c->sp_sv_items, T.sp_items)) || (e = up(c->sp_sv_group_item0, T.sp_group_item0)))) ||
        (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) ||
        (e = cudaMallocHost(&c->h_done, 2 * sizeof(unsigned long long))) || (e = cudaEventCreate(&c->ev0)) ||
        (e = cudaEventCreate(&c->ev1)) || (e = cudaEventCreateWithFlags(&c->ev_poll, cudaEventDisableTiming))) {
        qkdldpc_code_destroy(c);
        return fail(QKDLDPC_ERR_CUDA, "graph upload failed: %s", cudaGetErrorString(e));
    }
    c->own_stream = true;
    for (int k = 0; k < kSideStreams; ++k) {
        if (cudaStreamCreateWithFlags(&c->side_streams[k], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming) != cudaSuccess) {
            qkdldpc_code_destroy(c);
            return fail(QKDLDPC_ERR_CUDA, "side stream creation failed");
        }
    }
    if (cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess) {
        qkdldpc_code_destroy(c);
        return fail(QKDLDPC_ERR_CUDA, "event creation failed");
    }
    c->oc_max_dc = T.max_dc;
    c->oc_rec_slots = T.rec_slots;
    c->oc_groups_cn = (int)T.cn_ginfo.size();
    c->sp_eligible = sp_ok;
    c->sp_group_item0 = T.sp_group_item0;
    c->sp_groups_sv = sp_ok ? (int)T.sp_group_item0.size() - 1 : 0;
    c->sp_msg_words = T.sp_msg_words;
    for (int k = 0; k < 5; ++k) {
        c->cn_first[k] = cn_first[k]; c->cn_count[k] = cn_count[k];
        c->vn_first[k] = vn_first[k]; c->vn_count[k] = vn_count[k];
    }
    *out = c;
    return QKDLDPC_OK;
}

void qkdldpc_code_destroy(qkdldpc_code *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    comm_release(c);
    c->comm_buf.release();
    c->drop_graphs();
    c->row_ptr.release(); c->col_idx.release(); c->col_ptr.release(); c->csc_edge.release(); c->csc_row.release();
    c->row_order.release(); c->col_order.release(); c->vn_ell_edge.release(); c->vn_ell_row.release();
    c->oc_cn_ginfo.release(); c->oc_cnT.release(); c->oc_cn_row.release();
    c->oc_cls.release();
    c->oc2.release(); c->oc2r8.release(); c->oc2_cls.release(); c->oc2_phase_clk.release();
    c->sp_cn_moff.release(); c->sp_sv_items.release(); c->sp_sv_chunk.release(); c->sp_sv_group_item0.release();
    c->msg.release(); c->bobmask.release(); c->zmask.release(); c->synd.release(); c->par.release();
    c->tile_active.release(); c->tile_new.release(); c->slot_llr.release(); c->slot_frame.release();
    c->slot_iter.release(); c->frame_llr.release(); c->synd_all.release();
    c->bitclass.release(); c->counters.release();
    c->st_alice.release(); c->st_bob.release(); c->st_out.release(); c->st_qber.release(); c->st_iters.release();
    c->st_flags.release(); c->st_tally.release();
    c->gen_seeds.release(); c->gen_masks.release(); c->gen_scratch.release(); c->gen_combos.release(); c->oc_combos.release();
    c->compact_moves.release(); c->compact_plan.release(); c->sched_work.release(); c->rb_kept.release();
    c->rb_info.release(); c->st_keys_a.release(); c->st_keys_b.release();
    for (auto &p : c->ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto &ev : c->pipe_ev) cudaEventDestroy(ev);
    if (c->h_done) cudaFreeHost(c->h_done);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_poll) cudaEventDestroy(c->ev_poll);
    for (int k = 0; k < kSideStreams; ++k) {
        if (c->side_streams[k]) { cudaStreamSynchronize(c->side_streams[k]); cudaStreamDestroy(c->side_streams[k]); }
        if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int qkdldpc_code_set_stream(qkdldpc_code *c, void *cuda_stream) {
    if (!c) return fail(QKDLDPC_ERR_INVALID, "null code handle");
    CK(cudaSetDevice(c->device));
    if (c->stream) CK(cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    c->own_stream = false;
    c->drop_graphs();   // they were captured on the old stream
    return QKDLDPC_OK;
}

// pipe != nullptr: the frames are still in HOST memory (qkdldpc_decode_batch) and d_alice_bits / d_bob_bits / d_out_* are
// staging buffers; the on-chip path then copies piece by piece while it decodes (inst_onchip.cu). *used_pipe tells the
// caller whether that happened (otherwise it has to copy in before and out after, as for the streaming path).
static int decode_device_impl(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_frames,
                              const uint32_t *d_alice_bits, const uint32_t *d_bob_bits, const double *d_qber,
                              int32_t qber_is_scalar, const int32_t *punct_pos, int32_t n_punct,
                              const int32_t *short_pos, int32_t n_short, uint32_t *d_out_bits, int32_t *d_out_iters,
                              uint8_t *d_out_flags, uint64_t *d_tally, const HostPipe *pipe, bool *used_pipe);

int qkdldpc_decode_batch_device(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_frames,
                                const uint32_t *d_alice_bits, const uint32_t *d_bob_bits, const double *d_qber,
                                int32_t qber_is_scalar, const int32_t *punct_pos, int32_t n_punct,
                                const int32_t *short_pos, int32_t n_short, uint32_t *d_out_bits, int32_t *d_out_iters,
                                uint8_t *d_out_flags, uint64_t *d_tally) {
    return decode_device_impl(c, P, n_frames, d_alice_bits, d_bob_bits, d_qber, qber_is_scalar, punct_pos, n_punct, short_pos, n_short,
                              d_out_bits, d_out_iters, d_out_flags, d_tally, nullptr, nullptr);
}

static int decode_device_impl(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_frames,
                              const uint32_t *d_alice_bits, const uint32_t *d_bob_bits, const double *d_qber,
                              int32_t qber_is_scalar, const int32_t *punct_pos, int32_t n_punct,
                              const int32_t *short_pos, int32_t n_short, uint32_t *d_out_bits, int32_t *d_out_iters,
                              uint8_t *d_out_flags, uint64_t *d_tally, const HostPipe *pipe, bool *used_pipe) {
    if (used_pipe) *used_pipe = false;
    int rc = check_params(c, P, n_frames);
    if (rc) return rc;
    qkdldpc_params Pe = *P;   // the precision policy resolved (message_precision == 0)
    Pe.message_precision = qkdldpc_effective_precision(P->algorithm, c->n, P->message_precision);
    P = &Pe;
    c->last_precision = Pe.message_precision;
    if (n_frames == 0) {
        CK(cudaSetDevice(c->device));
        if (d_tally)
            CK(cudaMemsetAsync(d_tally, 0, (size_t)qkdldpc_tally_len(P->max_iterations) * sizeof(uint64_t), c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return QKDLDPC_OK;
    }
    if (!d_alice_bits || !d_bob_bits || !d_qber) return fail(QKDLDPC_ERR_INVALID, "null input buffer");
    if ((n_punct > 0 && !punct_pos) || (n_short > 0 && !short_pos) || n_punct < 0 || n_short < 0)
        return fail(QKDLDPC_ERR_INVALID, "bad punctured/shortened position list");
    CK(cudaSetDevice(c->device));
    auto *tl = reinterpret_cast<unsigned long long *>(d_tally);
    // decoder_path: 0 auto (on-chip min-sum when the graph and the parameters allow it), 1 streaming, 2 on-chip or fail
    const bool oc = onchip_usable(c, P);
    if (c->opt.decoder_path == 2 && !oc)
        return fail(QKDLDPC_ERR_INVALID, "on-chip path requested but not usable for this code / these parameters");
    if (oc && c->opt.decoder_path != 1) {
        if (used_pipe) *used_pipe = pipe != nullptr;
        return run_onchip(c, P, n_frames, d_alice_bits, d_bob_bits, d_qber, qber_is_scalar, punct_pos, n_punct, short_pos,
                          n_short, d_out_bits, d_out_iters, d_out_flags, tl, pipe);
    }
    c->last_path = 1;
    if (pipe) {   // streaming path: the step loop needs every frame's syndrome up front -- copy in first, the caller copies out
        const size_t tot = (size_t)n_frames * ((size_t)(c->n + 31) / 32);
        CK(cudaMemcpyAsync(const_cast<uint32_t *>(d_alice_bits), pipe->h_alice, tot * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(const_cast<uint32_t *>(d_bob_bits), pipe->h_bob, tot * 4, cudaMemcpyHostToDevice, c->stream));
    }
#define RUN(T, V)                                                                                                   \
    return run_batch<T, V>(c, P, n_frames, d_alice_bits, d_bob_bits, d_qber, qber_is_scalar, punct_pos, n_punct,    \
                           short_pos, n_short, d_out_bits, d_out_iters, d_out_flags, tl)
    if (P->message_precision == 64) {
        RUN(double, 2);
    } else {
        // tile width: 128-bit accesses (4 frames per lane) for the min-sum kernels. The SPA / SPA-lin check node keeps a
        // whole row of tanh values in registers between its two passes: 2 frames per lane is the best trade between
        // access width and occupancy (n=10240 R=0.82 SPA @ QBER 1.62 %: 2.87 / 3.60 / 3.39 Gbit/s at 1 / 2 / 4)
        int fpl = c->opt.frames_per_lane_f32;
        if (fpl == 0) fpl = P->algorithm <= 1 ? 2 : 4;
        switch (fpl) {
            case 1: RUN(float, 1);
            case 2: RUN(float, 2);
            default: RUN(float, 4);
        }
    }
#undef RUN
}

int qkdldpc_decode_batch(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_frames, const uint32_t *alice_bits,
                         const uint32_t *bob_bits, const double *qber, int32_t qber_is_scalar,
                         const int32_t *punct_pos, int32_t n_punct, const int32_t *short_pos, int32_t n_short,
                         uint32_t *out_bits, int32_t *out_iters, uint8_t *out_flags, uint64_t *tally) {
    int rc = check_params(c, P, n_frames);
    if (rc) return rc;
    const int64_t tl = qkdldpc_tally_len(P->max_iterations);
    if (n_frames == 0) {
        if (tally) memset(tally, 0, tl * sizeof(uint64_t));
        return QKDLDPC_OK;
    }
    if (!alice_bits || !bob_bits || !qber) return fail(QKDLDPC_ERR_INVALID, "null input buffer");
    CK(cudaSetDevice(c->device));
    const size_t words = (size_t)(c->n + 31) / 32, tot = (size_t)n_frames * words;
    const size_t nq = qber_is_scalar ? 1 : (size_t)n_frames;
    CK(c->st_alice.reserve(tot));
    CK(c->st_bob.reserve(tot));
    CK(c->st_qber.reserve(nq));
    if (out_bits) CK(c->st_out.reserve(tot));
    CK(c->st_iters.reserve(n_frames));
    CK(c->st_flags.reserve(n_frames));
    CK(c->st_tally.reserve(tl));
    cudaStream_t s = c->stream;
    CK(cudaMemcpyAsync(c->st_qber.p, qber, nq * sizeof(double), cudaMemcpyHostToDevice, s));
    // copy / compute overlap (on-chip paths): the batch is cut into pieces, see onchip_launch_all (inst_onchip.cu)
    HostPipe pipe{alice_bits, bob_bits, out_bits, out_iters, out_flags, 1};
    pipe.chunks = c->opt.copy_chunks > 0 ? std::min(c->opt.copy_chunks, kMaxPipeChunks) : (n_frames >= 4096 ? 8 : (n_frames >= 1024 ? 2 : 1));
    bool piped = false;
    rc = decode_device_impl(c, P, n_frames, c->st_alice.p, c->st_bob.p, c->st_qber.p, qber_is_scalar, punct_pos, n_punct, short_pos,
                            n_short, out_bits ? c->st_out.p : nullptr, c->st_iters.p, c->st_flags.p,
                            reinterpret_cast<uint64_t *>(c->st_tally.p), &pipe, &piped);
    if (rc) return rc;
    if (!piped) {
        if (out_bits) CK(cudaMemcpyAsync(out_bits, c->st_out.p, tot * 4, cudaMemcpyDeviceToHost, s));
        if (out_iters) CK(cudaMemcpyAsync(out_iters, c->st_iters.p, n_frames * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        if (out_flags) CK(cudaMemcpyAsync(out_flags, c->st_flags.p, n_frames, cudaMemcpyDeviceToHost, s));
    }
    if (tally) CK(cudaMemcpyAsync(tally, c->st_tally.p, tl * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return QKDLDPC_OK;
}

int qkdldpc_generate_keys_device(qkdldpc_code *c, int64_t n_frames, double qber, uint64_t seed,
                                 uint32_t *d_alice_bits, uint32_t *d_bob_bits, double *accurate_qber_out) {
    if (!c) return fail(QKDLDPC_ERR_INVALID, "null code handle");
    if (!(qber > 0.) || !(qber < 1.)) return fail(QKDLDPC_ERR_INVALID, "qber must be in (0, 1)");
    if (n_frames < 0 || (n_frames > 0 && (!d_alice_bits || !d_bob_bits)))
        return fail(QKDLDPC_ERR_INVALID, "bad frame buffers");
    // inject_errors: num_errors = size_t(double(N) * QBER) (array_and_matrix_operations.cpp:913-914)
    const int n_err = (int)(size_t)((double)c->n * qber);
    if (accurate_qber_out) *accurate_qber_out = (double)n_err / (double)c->n;
    if (n_frames == 0) return QKDLDPC_OK;
    CK(cudaSetDevice(c->device));
    const int words = (c->n + 31) / 32;
    qk::gen_keys_kernel<<<(unsigned)n_frames, 128, words * sizeof(uint32_t), c->stream>>>(
        c->n, words, n_err, (unsigned long long)seed, d_alice_bits, d_bob_bits);
    c->kernel_launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    return QKDLDPC_OK;
}

namespace {

// Reference-compatible inputs for n_combos x trials frames (frame c * trials + t = trial t of combination c) into DEVICE
// buffers. accurate[c] receives floor(n * qber_c) / n.
int generate_inputs_multi(qkdldpc_code *c, int n_combos, const qkdldpc_combination *combos, int64_t trials, const uint64_t *trial_seeds,
                          uint32_t *d_alice, uint32_t *d_bob, double *accurate) {
    const int n = c->n, words = (n + 31) / 32;
    std::vector<qk::RefKeygenCombo> table((size_t)n_combos);
    std::vector<uint32_t> masks;
    int max_err = 1;
    bool any_ra = false;
    for (int k = 0; k < n_combos; ++k) {
        const qkdldpc_combination &cb = combos[k];
        if (!(cb.qber >= 0.) || !(cb.qber < 1.)) return fail(QKDLDPC_ERR_INVALID, "qber must be in [0, 1)");
        if ((cb.n_punct > 0 && !cb.punct_pos) || (cb.n_short > 0 && !cb.short_pos) || cb.n_punct < 0 || cb.n_short < 0)
            return fail(QKDLDPC_ERR_INVALID, "bad punctured/shortened position list");
        // inject_errors: num_errors = size_t(double(N) * QBER) (array_and_matrix_operations.cpp:913-914)
        const int n_err = (int)(size_t)((double)n * cb.qber);
        if (accurate) accurate[k] = (double)n_err / (double)n;
        table[k].seed_offset = cb.seed_offset;
        table[k].n_err = n_err;
        table[k].rate_adapt = (cb.n_punct > 0 || cb.n_short > 0) ? 1 : 0;
        max_err = std::max(max_err, n_err);
        any_ra |= table[k].rate_adapt != 0;
    }
    const int64_t n_frames = (int64_t)n_combos * trials;
    if (n_frames == 0) return QKDLDPC_OK;
    if (any_ra) {
        masks.assign((size_t)n_combos * 2 * words, 0u);
        for (int k = 0; k < n_combos; ++k) {
            const int rc = onchip_pack_masks(n, combos[k].punct_pos, combos[k].n_punct, combos[k].short_pos, combos[k].n_short,
                                             masks.data() + (size_t)k * 2 * words);
            if (rc) return rc;
        }
    }
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    CK(c->gen_seeds.reserve((size_t)trials));
    CK(c->gen_combos.reserve(table.size() * sizeof(qk::RefKeygenCombo)));
    CK(c->gen_masks.reserve(std::max<size_t>(masks.size(), 1)));
    CK(cudaMemcpyAsync(c->gen_seeds.p, trial_seeds, (size_t)trials * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->gen_combos.p, table.data(), table.size() * sizeof(qk::RefKeygenCombo), cudaMemcpyHostToDevice, s));
    if (any_ra) CK(cudaMemcpyAsync(c->gen_masks.p, masks.data(), masks.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    // one thread per frame; sub-batches bound the per-thread scratch (raw keys for rate adaptation + shuffle prefix)
    const long long per_thread = (any_ra ? 2ll * words : 0ll) + max_err;
    const long long sub = std::max<long long>(128, std::min<long long>(n_frames, ((long long)256 << 20) / (per_thread * 4)) / 128 * 128);
    CK(c->gen_scratch.reserve((size_t)(per_thread * sub)));
    for (long long f0 = 0; f0 < n_frames; f0 += sub) {
        qk::RefKeygenArgs a{};
        a.n = n; a.words = words;
        a.n_frames = std::min<long long>(sub, n_frames - f0);
        a.first_frame = f0;
        a.trials = trials;
        a.seeds = reinterpret_cast<const qk::u64 *>(c->gen_seeds.p);
        a.combos = reinterpret_cast<const qk::RefKeygenCombo *>(c->gen_combos.p);
        a.alice = d_alice + f0 * words; a.bob = d_bob + f0 * words;
        a.masks = c->gen_masks.p;
        a.any_rate_adapt = any_ra ? 1 : 0;
        a.max_err = max_err;
        a.scratch = c->gen_scratch.p; a.scratch_stride = sub;
        qk::ref_keygen_kernel<<<(unsigned)((a.n_frames + 127) / 128), 128, 0, s>>>(a);
        c->kernel_launches += 1;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));   // the host tables die at scope end
    return QKDLDPC_OK;
}

}  // namespace

namespace {

// remove_bits for the frames of a (multi-combination) launch, on the device: plan = every combination's surviving
// positions back to back + where its final keys go.
struct RemovePlan {
    std::vector<int> kept;
    std::vector<qk::RemoveCombo> info;
    long long total_words = 0;   // output words per party
    int max_words_out = 0;
    bool any = false;
};

int plan_removal(int n, int64_t trials, int n_combos, const qkdldpc_combination *combos, RemovePlan &pl) {
    pl.info.assign((size_t)n_combos, qk::RemoveCombo{0, 0, 0, 0, 0});
    for (int k = 0; k < n_combos; ++k) {
        const qkdldpc_combination &cb = combos[k];
        if (cb.n_remove < 0 || cb.n_remove > n || (cb.n_remove > 0 && !cb.remove_pos)) return fail(QKDLDPC_ERR_INVALID, "bad removal list");
        if (cb.n_remove == 0) continue;
        qk::RemoveCombo &rc = pl.info[k];
        rc.kept_off = (int)pl.kept.size();
        for (int i = 0, r = 0; i < n; ++i) {
            if (r < cb.n_remove && cb.remove_pos[r] == i) {
                ++r;
                if (r < cb.n_remove && cb.remove_pos[r] <= i) return fail(QKDLDPC_ERR_INVALID, "bits_to_remove must be strictly ascending");
            } else {
                pl.kept.push_back(i);
            }
        }
        rc.n_keep = (int)pl.kept.size() - rc.kept_off;
        if (rc.n_keep != n - cb.n_remove) return fail(QKDLDPC_ERR_INVALID, "bits_to_remove must be strictly ascending positions in [0, n)");
        rc.words_out = (rc.n_keep + 31) / 32;
        rc.out_off = pl.total_words;
        pl.total_words += (long long)rc.words_out * trials;
        pl.max_words_out = std::max(pl.max_words_out, rc.words_out);
        pl.any = true;
    }
    return QKDLDPC_OK;
}

// Final keys of Alice (from d_alice) and Bob (from d_solution) into c->st_keys_a / st_keys_b; copies them to the host
// pointers of combinations [0, n_combos) where given.
int run_removal(qkdldpc_code *c, const RemovePlan &pl, int64_t trials, int n_combos, const uint32_t *d_alice, const uint32_t *d_solution,
                uint32_t *const *out_a, uint32_t *const *out_b) {
    if (!pl.any || trials == 0 || pl.max_words_out == 0) return QKDLDPC_OK;
    cudaStream_t s = c->stream;
    const int words_in = (c->n + 31) / 32;
    CK(c->rb_kept.reserve(pl.kept.size()));
    CK(c->rb_info.reserve(pl.info.size() * sizeof(qk::RemoveCombo)));
    CK(c->st_keys_a.reserve((size_t)pl.total_words));
    CK(c->st_keys_b.reserve((size_t)pl.total_words));
    CK(cudaMemcpyAsync(c->rb_kept.p, pl.kept.data(), pl.kept.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->rb_info.p, pl.info.data(), pl.info.size() * sizeof(qk::RemoveCombo), cudaMemcpyHostToDevice, s));
    const dim3 grid((unsigned)((int64_t)n_combos * trials), (unsigned)((pl.max_words_out + 7) / 8));
    const auto *info = reinterpret_cast<const qk::RemoveCombo *>(c->rb_info.p);
    qk::remove_bits_multi_kernel<<<grid, 256, 0, s>>>(words_in, (long long)trials, c->rb_kept.p, info, d_alice, c->st_keys_a.p);
    qk::remove_bits_multi_kernel<<<grid, 256, 0, s>>>(words_in, (long long)trials, c->rb_kept.p, info, d_solution, c->st_keys_b.p);
    c->kernel_launches += 2;
    CK(cudaGetLastError());
    for (int k = 0; k < n_combos; ++k) {
        const qk::RemoveCombo &rc = pl.info[k];
        const size_t bytes = (size_t)rc.words_out * trials * 4;
        if (bytes == 0) continue;
        if (out_a && out_a[k]) CK(cudaMemcpyAsync(out_a[k], c->st_keys_a.p + rc.out_off, bytes, cudaMemcpyDeviceToHost, s));
        if (out_b && out_b[k]) CK(cudaMemcpyAsync(out_b[k], c->st_keys_b.p + rc.out_off, bytes, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));   // the plan's host tables may die after this
    return QKDLDPC_OK;
}

}  // namespace

int qkdldpc_generate_trial_inputs_device(qkdldpc_code *c, int64_t n_frames, const uint64_t *trial_seeds, uint64_t seed_offset,
                                         double qber, const int32_t *punct_pos, int32_t n_punct, const int32_t *short_pos,
                                         int32_t n_short, uint32_t *d_alice_bits, uint32_t *d_bob_bits, double *accurate_qber_out) {
    if (!c) return fail(QKDLDPC_ERR_INVALID, "null code handle");
    if (n_frames < 0 || (n_frames > 0 && (!trial_seeds || !d_alice_bits || !d_bob_bits))) return fail(QKDLDPC_ERR_INVALID, "bad frame buffers");
    qkdldpc_combination cb{};
    cb.qber = qber;
    cb.punct_pos = punct_pos; cb.n_punct = n_punct;
    cb.short_pos = short_pos; cb.n_short = n_short;
    cb.seed_offset = seed_offset;
    return generate_inputs_multi(c, 1, &cb, n_frames, trial_seeds, d_alice_bits, d_bob_bits, accurate_qber_out);
}

static int run_trials_impl(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_trials, const uint64_t *trial_seeds, uint64_t seed_offset,
                           double qber, const int32_t *punct_pos, int32_t n_punct, const int32_t *short_pos, int32_t n_short,
                           bool keep_bits_on_device, uint32_t *out_bits, int32_t *out_iters, uint8_t *out_flags, uint64_t *tally,
                           double *accurate_qber_out);

int qkdldpc_run_trials_multi(qkdldpc_code *c, const qkdldpc_params *P, int32_t n_combinations, const qkdldpc_combination *combos,
                             int64_t n_trials, const uint64_t *trial_seeds, int32_t *out_iters, uint8_t *out_flags, uint64_t *tallies,
                             double *accurate_qber_out) {
    return qkdldpc_run_trials_multi_keys(c, P, n_combinations, combos, n_trials, trial_seeds, out_iters, out_flags, tallies, accurate_qber_out,
                                         nullptr, nullptr);
}

int qkdldpc_run_trials_multi_keys(qkdldpc_code *c, const qkdldpc_params *P, int32_t n_combinations, const qkdldpc_combination *combos,
                                  int64_t n_trials, const uint64_t *trial_seeds, int32_t *out_iters, uint8_t *out_flags, uint64_t *tallies,
                                  double *accurate_qber_out, uint32_t *const *out_alice_keys, uint32_t *const *out_bob_keys) {
    int rc = check_params(c, P, n_trials);
    if (rc) return rc;
    if (n_combinations < 0 || (n_combinations > 0 && !combos)) return fail(QKDLDPC_ERR_INVALID, "bad combination table");
    const int64_t tl = qkdldpc_tally_len(P->max_iterations);
    if (n_combinations == 0) return QKDLDPC_OK;
    if (n_trials > 0 && !trial_seeds) return fail(QKDLDPC_ERR_INVALID, "null seed array");
    qkdldpc_params Pe = *P;   // the precision policy resolved (message_precision == 0)
    Pe.message_precision = qkdldpc_effective_precision(P->algorithm, c->n, P->message_precision);
    P = &Pe;
    qkdldpc_params Pk = *P;
    Pk.primary = combos[0].primary;
    Pk.secondary = combos[0].secondary;
    bool onchip = c->opt.decoder_path != 1;
    for (int k = 0; k < n_combinations && onchip; ++k) {   // every combination must satisfy the on-chip preconditions
        Pk.primary = combos[k].primary;
        Pk.secondary = combos[k].secondary;
        onchip = onchip_usable(c, &Pk);
    }
    if (!onchip || n_trials == 0) {
        // streaming path (SPA, float64, long codes): one combination after the other
        for (int k = 0; k < n_combinations; ++k) {
            Pk.primary = combos[k].primary;
            Pk.secondary = combos[k].secondary;
            RemovePlan pl;
            rc = plan_removal(c->n, n_trials, 1, combos + k, pl);
            if (rc) return rc;
            rc = run_trials_impl(c, &Pk, n_trials, trial_seeds, combos[k].seed_offset, combos[k].qber, combos[k].punct_pos, combos[k].n_punct,
                                 combos[k].short_pos, combos[k].n_short, pl.any, nullptr, out_iters ? out_iters + (int64_t)k * n_trials : nullptr,
                                 out_flags ? out_flags + (int64_t)k * n_trials : nullptr, tallies ? tallies + (int64_t)k * tl : nullptr,
                                 accurate_qber_out ? accurate_qber_out + k : nullptr);
            if (rc) return rc;
            // the combination's frames are still in the staging buffers: Alice's (extended) keys and bob_solution
            rc = run_removal(c, pl, n_trials, 1, c->st_alice.p, c->st_out.p, out_alice_keys ? out_alice_keys + k : nullptr,
                             out_bob_keys ? out_bob_keys + k : nullptr);
            if (rc) return rc;
        }
        return QKDLDPC_OK;
    }
    RemovePlan pl;
    rc = plan_removal(c->n, n_trials, n_combinations, combos, pl);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    const int words = (c->n + 31) / 32;
    const int64_t n_frames = (int64_t)n_combinations * n_trials;
    const size_t tot = (size_t)n_frames * words;
    CK(c->st_alice.reserve(tot));
    CK(c->st_bob.reserve(tot));
    std::vector<double> acc((size_t)n_combinations, 0.);
    rc = generate_inputs_multi(c, n_combinations, combos, n_trials, trial_seeds, c->st_alice.p, c->st_bob.p, acc.data());
    if (rc) return rc;
    std::vector<qk::OnchipCombo> table((size_t)n_combinations);
    std::vector<uint32_t> masks((size_t)n_combinations * 2 * words, 0u);
    for (int k = 0; k < n_combinations; ++k) {
        if (acc[k] == 0.) return fail(QKDLDPC_ERR_INVALID, "Key size '%d' is too small for QBER.", c->n);   // simulation.cpp:556-557
        if (accurate_qber_out) accurate_qber_out[k] = acc[k];
        table[k].qber = acc[k];
        table[k].primary = combos[k].primary;
        table[k].secondary = combos[k].secondary;
        table[k].has_cls = (combos[k].n_punct > 0 || combos[k].n_short > 0) ? 1 : 0;
        rc = onchip_pack_masks(c->n, combos[k].punct_pos, combos[k].n_punct, combos[k].short_pos, combos[k].n_short,
                               masks.data() + (size_t)k * 2 * words);
        if (rc) return rc;
    }
    CK(c->st_iters.reserve(n_frames));
    CK(c->st_flags.reserve(n_frames));
    CK(c->st_tally.reserve((size_t)n_combinations * tl));
    c->last_precision = Pe.message_precision;
    if (pl.any) CK(c->st_out.reserve(tot));   // bob_solution stays on the device for remove_bits
    rc = run_onchip_multi(c, P, n_combinations, n_trials, table.data(), masks.data(), c->st_alice.p, c->st_bob.p, nullptr, 1,
                          pl.any ? c->st_out.p : nullptr, c->st_iters.p, c->st_flags.p, c->st_tally.p, nullptr);
    if (rc) return rc;
    rc = run_removal(c, pl, n_trials, n_combinations, c->st_alice.p, c->st_out.p, out_alice_keys, out_bob_keys);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    if (out_iters) CK(cudaMemcpyAsync(out_iters, c->st_iters.p, n_frames * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (out_flags) CK(cudaMemcpyAsync(out_flags, c->st_flags.p, n_frames, cudaMemcpyDeviceToHost, s));
    if (tallies) CK(cudaMemcpyAsync(tallies, c->st_tally.p, (size_t)n_combinations * tl * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return QKDLDPC_OK;
}

int qkdldpc_run_trials(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_trials, const uint64_t *trial_seeds, uint64_t seed_offset,
                       double qber, const int32_t *punct_pos, int32_t n_punct, const int32_t *short_pos, int32_t n_short,
                       uint32_t *out_bits, int32_t *out_iters, uint8_t *out_flags, uint64_t *tally, double *accurate_qber_out) {
    return run_trials_impl(c, P, n_trials, trial_seeds, seed_offset, qber, punct_pos, n_punct, short_pos, n_short, false, out_bits, out_iters,
                           out_flags, tally, accurate_qber_out);
}

// keep_bits_on_device: decode with bob_solution written to the staging buffer even when the caller wants no host copy
// (remove_bits follows on the device).
static int run_trials_impl(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_trials, const uint64_t *trial_seeds, uint64_t seed_offset,
                           double qber, const int32_t *punct_pos, int32_t n_punct, const int32_t *short_pos, int32_t n_short,
                           bool keep_bits_on_device, uint32_t *out_bits, int32_t *out_iters, uint8_t *out_flags, uint64_t *tally,
                           double *accurate_qber_out) {
    int rc = check_params(c, P, n_trials);
    if (rc) return rc;
    const int64_t tl = qkdldpc_tally_len(P->max_iterations);
    double acc = 0.;
    const size_t words = (size_t)(c->n + 31) / 32, tot = (size_t)n_trials * words;
    CK(cudaSetDevice(c->device));
    CK(c->st_alice.reserve(std::max<size_t>(tot, 1)));
    CK(c->st_bob.reserve(std::max<size_t>(tot, 1)));
    rc = qkdldpc_generate_trial_inputs_device(c, n_trials, trial_seeds, seed_offset, qber, punct_pos, n_punct, short_pos, n_short,
                                              c->st_alice.p, c->st_bob.p, &acc);
    if (rc) return rc;
    if (accurate_qber_out) *accurate_qber_out = acc;
    if (n_trials == 0) {
        if (tally) memset(tally, 0, tl * sizeof(uint64_t));
        return QKDLDPC_OK;
    }
    if (acc == 0.)   // run_trial throws here (simulation.cpp:556-557)
        return fail(QKDLDPC_ERR_INVALID, "Key size '%d' is too small for QBER.", c->n);
    CK(c->st_qber.reserve(1));
    if (out_bits || keep_bits_on_device) CK(c->st_out.reserve(tot));
    CK(c->st_iters.reserve(n_trials));
    CK(c->st_flags.reserve(n_trials));
    CK(c->st_tally.reserve(tl));
    cudaStream_t s = c->stream;
    CK(cudaMemcpyAsync(c->st_qber.p, &acc, sizeof(double), cudaMemcpyHostToDevice, s));
    rc = qkdldpc_decode_batch_device(c, P, n_trials, c->st_alice.p, c->st_bob.p, c->st_qber.p, 1, punct_pos, n_punct, short_pos, n_short,
                                     (out_bits || keep_bits_on_device) ? c->st_out.p : nullptr, c->st_iters.p, c->st_flags.p,
                                     reinterpret_cast<uint64_t *>(c->st_tally.p));
    if (rc) return rc;
    if (out_bits) CK(cudaMemcpyAsync(out_bits, c->st_out.p, tot * 4, cudaMemcpyDeviceToHost, s));
    if (out_iters) CK(cudaMemcpyAsync(out_iters, c->st_iters.p, n_trials * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (out_flags) CK(cudaMemcpyAsync(out_flags, c->st_flags.p, n_trials, cudaMemcpyDeviceToHost, s));
    if (tally) CK(cudaMemcpyAsync(tally, c->st_tally.p, tl * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return QKDLDPC_OK;
}

// synthetic keys generated on the device, decoded, timed -- the throughput entry point.
int qkdldpc_bench_synthetic(qkdldpc_code *c, const qkdldpc_params *P, int64_t n_frames, double qber, uint64_t seed,
                            uint64_t *tally, double *seconds_out) {
    int rc = check_params(c, P, n_frames);
    if (rc) return rc;
    const int64_t tl = qkdldpc_tally_len(P->max_iterations);
    if (seconds_out) *seconds_out = 0.;
    if (n_frames == 0) {
        if (tally) memset(tally, 0, tl * sizeof(uint64_t));
        return QKDLDPC_OK;
    }
    const size_t words = (size_t)(c->n + 31) / 32, tot = (size_t)n_frames * words;
    CK(cudaSetDevice(c->device));
    CK(c->st_alice.reserve(tot));
    CK(c->st_bob.reserve(tot));
    double acc = 0.;
    rc = qkdldpc_generate_keys_device(c, n_frames, qber, seed, c->st_alice.p, c->st_bob.p, &acc);
    if (rc) return rc;
    if (acc == 0.) return fail(QKDLDPC_ERR_INVALID, "Key size '%d' is too small for QBER.", c->n);
    CK(c->st_qber.reserve(1));
    CK(c->st_iters.reserve(n_frames));
    CK(c->st_flags.reserve(n_frames));
    CK(c->st_tally.reserve(tl));
    cudaStream_t s = c->stream;
    CK(cudaMemcpyAsync(c->st_qber.p, &acc, sizeof(double), cudaMemcpyHostToDevice, s));
    cudaEvent_t t0 = nullptr, t1 = nullptr;   // the handle's own events are used inside the decode call
    CK(cudaEventCreate(&t0));
    CK(cudaEventCreate(&t1));
    cudaEventRecord(t0, s);
    rc = qkdldpc_decode_batch_device(c, P, n_frames, c->st_alice.p, c->st_bob.p, c->st_qber.p, 1, nullptr, 0, nullptr, 0, nullptr,
                                     c->st_iters.p, c->st_flags.p, reinterpret_cast<uint64_t *>(c->st_tally.p));
    cudaEventRecord(t1, s);
    cudaEventSynchronize(t1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    if (rc) return rc;
    if (seconds_out) *seconds_out = (double)ms * 1e-3;   // decode only: keys resident in device memory, CUDA events
    if (tally) CK(cudaMemcpy(tally, c->st_tally.p, tl * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return QKDLDPC_OK;
}

int qkdldpc_remove_bits(qkdldpc_code *c, int64_t n_frames, const uint32_t *keys, const int32_t *bits_to_remove, int32_t n_remove,
                        uint32_t *out_keys) {
    if (!c) return fail(QKDLDPC_ERR_INVALID, "null code handle");
    if (n_frames < 0 || n_remove < 0 || n_remove > c->n || (n_remove > 0 && !bits_to_remove))
        return fail(QKDLDPC_ERR_INVALID, "bad remove_bits arguments");
    const int n = c->n, words_in = (n + 31) / 32, n_keep = n - n_remove, words_out = (n_keep + 31) / 32;
    if (n_frames == 0 || words_out == 0) return QKDLDPC_OK;
    if (!keys || !out_keys) return fail(QKDLDPC_ERR_INVALID, "null key buffer");
    std::vector<int> kept;
    kept.reserve((size_t)n_keep);
    for (int i = 0, r = 0; i < n; ++i) {
        if (r < n_remove && bits_to_remove[r] == i) {
            ++r;
            if (r < n_remove && bits_to_remove[r] <= i) return fail(QKDLDPC_ERR_INVALID, "bits_to_remove must be strictly ascending");
        } else {
            kept.push_back(i);
        }
    }
    if ((int)kept.size() != n_keep) return fail(QKDLDPC_ERR_INVALID, "bits_to_remove must be strictly ascending positions in [0, n)");
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t tot_in = (size_t)n_frames * words_in, tot_out = (size_t)n_frames * words_out;
    CK(c->st_alice.reserve(tot_in));
    CK(c->st_out.reserve(tot_out));
    CK(c->rb_kept.reserve(kept.size()));
    CK(cudaMemcpyAsync(c->rb_kept.p, kept.data(), kept.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->st_alice.p, keys, tot_in * 4, cudaMemcpyHostToDevice, s));
    for (int64_t f0 = 0; f0 < n_frames; f0 += 65535 * 1024ll) {   // gridDim.x limit is far away; keep the launch simple
        const unsigned nf = (unsigned)std::min<int64_t>(n_frames - f0, 65535 * 1024ll);
        qk::remove_bits_kernel<<<dim3(nf, (unsigned)((words_out + 7) / 8)), 256, 0, s>>>(words_in, n_keep, words_out, c->rb_kept.p,
                                                                                      c->st_alice.p + f0 * words_in, c->st_out.p + f0 * words_out);
        c->kernel_launches += 1;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_keys, c->st_out.p, tot_out * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));   // also keeps `kept` alive until the upload is done
    return QKDLDPC_OK;
}

int qkdldpc_code_info(const qkdldpc_code *c, qkdldpc_info *info) {
    if (!c || !info) return fail(QKDLDPC_ERR_INVALID, "null argument");
    info->n = c->n; info->m = c->m; info->nnz = c->nnz; info->device = c->device;
    info->frames_per_tile = c->frames_per_tile; info->pool_tiles = c->pool_tiles; info->pool_bytes = c->pool_bytes;
    info->kernel_launches = c->kernel_launches; info->decoder_steps = c->decoder_steps;
    info->last_batch_ms = c->last_batch_ms;
    info->last_path = c->last_path;
    info->onchip_threads = c->oc_threads;
    info->last_precision = c->last_precision;
    info->onchip_record_bytes = c->last_rec_bytes;
    info->tail_compactions = c->tail_compactions;
    info->last_steps_per_poll = c->last_spp;
    info->last_vn_items_per_warp = c->last_vn_items;
    info->last_cn_ms = c->last_cn_ms; info->last_vn_ms = c->last_vn_ms; info->last_sched_ms = c->last_sched_ms;
    return QKDLDPC_OK;
}

int qkdldpc_code_set_profiling(qkdldpc_code *c, int32_t enabled) {
    if (!c) return fail(QKDLDPC_ERR_INVALID, "null code handle");
    c->profiling = enabled != 0;
    return QKDLDPC_OK;
}

}  // extern "C"

// K5, the only collective of the path: the all-reduce (sum) of the tally vectors -- success counts and the iteration
// histogram that process_trials_results consumes (simulation.cpp:580-624) -- over the GPUs that share a batch of frames.
// The handle owns its NCCL communicator. NCCL is bound at run time (dlopen of libnccl.so.2: inside a
// torch process that is torch's own copy, in qkdldpc_sim the system library), so the decoder itself has no link-time
// dependency on it and single-GPU users never load it. The message is < 1 KB: latency only, NVLink bandwidth is irrelevant.
#include "handle.hpp"

#include <dlfcn.h>
#include <mutex>
#include <nccl.h>

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    std::string error;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
            return;
        }
        auto sym = [&](const char *n) {
            void *p = dlsym(api.lib, n);
            if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + n;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    return api;
}

int nccl_fail(const char *what, ncclResult_t r) {
    return fail(QKDLDPC_ERR_CUDA, "%s failed: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
}

int need_nccl() {
    if (!nccl().error.empty() || !nccl().lib) return fail(QKDLDPC_ERR_STATE, "NCCL unavailable: %s", nccl().error.c_str());
    return QKDLDPC_OK;
}

}  // namespace

namespace qkhost {
void comm_release(qkdldpc_code *c) {
    if (c->comm && nccl().CommDestroy) nccl().CommDestroy(static_cast<ncclComm_t>(c->comm));
    c->comm = nullptr;
    c->comm_ranks = 0;
}
}  // namespace qkhost

extern "C" {

int qkdldpc_comm_nccl_version(void) {
    int v = 0;
    if (need_nccl() != QKDLDPC_OK || nccl().GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

int qkdldpc_comm_get_unique_id(uint8_t *id_out) {
    static_assert(sizeof(ncclUniqueId) == QKDLDPC_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    if (!id_out) return fail(QKDLDPC_ERR_INVALID, "null id buffer");
    int rc = need_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    const ncclResult_t r = nccl().GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id_out, &id, sizeof id);
    return QKDLDPC_OK;
}

int qkdldpc_comm_init_rank(qkdldpc_code *c, const uint8_t *id_bytes, int32_t n_ranks, int32_t rank) {
    if (!c || !id_bytes) return fail(QKDLDPC_ERR_INVALID, "null argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(QKDLDPC_ERR_INVALID, "rank %d not in [0, %d)", rank, n_ranks);
    int rc = need_nccl();
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    comm_release(c);
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ncclComm_t comm = nullptr;
    const ncclResult_t r = nccl().CommInitRank(&comm, n_ranks, id, rank);
    if (r != ncclSuccess) return nccl_fail("ncclCommInitRank", r);
    c->comm = comm;
    c->comm_ranks = n_ranks;
    return QKDLDPC_OK;
}

int qkdldpc_comm_init_all(qkdldpc_code *const *codes, int32_t n_codes) {
    if (!codes || n_codes < 1) return fail(QKDLDPC_ERR_INVALID, "no handles");
    int rc = need_nccl();
    if (rc) return rc;
    std::vector<int> devs((size_t)n_codes);
    for (int i = 0; i < n_codes; ++i) {
        if (!codes[i]) return fail(QKDLDPC_ERR_INVALID, "null code handle");
        devs[i] = codes[i]->device;
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return fail(QKDLDPC_ERR_INVALID, "qkdldpc_comm_init_all needs one handle per DISTINCT device (device %d twice)", devs[i]);
    }
    std::vector<ncclComm_t> comms((size_t)n_codes, nullptr);
    const ncclResult_t r = nccl().CommInitAll(comms.data(), n_codes, devs.data());
    if (r != ncclSuccess) return nccl_fail("ncclCommInitAll", r);
    for (int i = 0; i < n_codes; ++i) {
        comm_release(codes[i]);
        codes[i]->comm = comms[i];
        codes[i]->comm_ranks = n_codes;
    }
    return QKDLDPC_OK;
}

int qkdldpc_comm_size(const qkdldpc_code *c) { return c ? c->comm_ranks : 0; }

int qkdldpc_tally_allreduce_device(qkdldpc_code *c, uint64_t *d_tally, int64_t count) {
    if (!c || !d_tally || count < 0) return fail(QKDLDPC_ERR_INVALID, "bad tally buffer");
    if (!c->comm) return fail(QKDLDPC_ERR_STATE, "the handle has no communicator (qkdldpc_comm_init_rank / _init_all)");
    CK(cudaSetDevice(c->device));
    const ncclResult_t r = nccl().AllReduce(d_tally, d_tally, (size_t)count, ncclUint64, ncclSum, static_cast<ncclComm_t>(c->comm), c->stream);
    if (r != ncclSuccess) return nccl_fail("ncclAllReduce", r);
    return QKDLDPC_OK;
}

int qkdldpc_tally_allreduce(qkdldpc_code *c, uint64_t *tally, int64_t count) {
    if (!c || !tally || count < 0) return fail(QKDLDPC_ERR_INVALID, "bad tally buffer");
    if (!c->comm) return fail(QKDLDPC_ERR_STATE, "the handle has no communicator (qkdldpc_comm_init_rank / _init_all)");
    CK(cudaSetDevice(c->device));
    CK(c->comm_buf.reserve((size_t)count));
    CK(cudaMemcpyAsync(c->comm_buf.p, tally, (size_t)count * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    const int rc = qkdldpc_tally_allreduce_device(c, reinterpret_cast<uint64_t *>(c->comm_buf.p), count);
    if (rc) return rc;
    CK(cudaMemcpyAsync(tally, c->comm_buf.p, (size_t)count * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return QKDLDPC_OK;
}

}  // extern "C"

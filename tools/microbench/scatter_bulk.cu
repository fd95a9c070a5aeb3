// Micro-benchmark: the scattered 512-byte read-modify-write of the streaming variable-node kernel (one chunk per edge, DV
// chunks per item) with the chunks moved by the bulk-copy engine instead of by the lanes: cp.async.bulk global -> shared
// (completion on an mbarrier), the lanes work on shared memory, cp.async.bulk shared -> global. A warp walks ITEMS items
// with a STAGES-deep ring of 2 KB buffers, so that the messages of the next STAGES - 1 items are in flight while one item
// is processed -- bytes in flight without registers. Question: does this beat the 5.7 - 5.9 TB/s of plain 16-byte loads?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scatter_bulk scatter_bulk.cu && ./scatter_bulk
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kChunk = 512, kMaxDv = 4;

template <int STAGES>
__global__ void __launch_bounds__(256) rmw_bulk(float4 *pool, const uint32_t *slot, int n_items, int dv, int items) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned char *buf = smem + (size_t)w * STAGES * kMaxDv * kChunk;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nw * STAGES * kMaxDv * kChunk) + w * STAGES;
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int first = (blockIdx.x * nw + w) * items;
    const int mine = min(items, n_items - first);
    if (mine <= 0) return;
    const uint32_t bytes = (uint32_t)dv * kChunk;
    auto issue = [&](int i) {   // lane 0: request the chunks of item first + i into stage i % STAGES
        const int st = i % STAGES;
        mbar_expect_tx(bars + st, bytes);
        for (int k = 0; k < dv; ++k) {
            const size_t base = (size_t)__ldg(slot + (size_t)(first + i) * dv + k) * (kChunk / 16);
            bulk_g2s(buf + (st * kMaxDv + k) * kChunk, pool + base, kChunk, bars + st);
        }
    };
    if (lane == 0)
        for (int i = 0; i < STAGES - 1 && i < mine; ++i) issue(i);
    for (int i = 0; i < mine; ++i) {
        const int st = i % STAGES;
        if (lane == 0 && i + STAGES - 1 < mine) {
            bulk_wait_read<0>();   // the stores of item i - 1 have read their stage: it is free for item i + STAGES - 1
            issue(i + STAGES - 1);
        }
        mbar_wait(bars + st, (uint32_t)(i / STAGES) & 1u);
        float4 *sb = reinterpret_cast<float4 *>(buf + st * kMaxDv * kChunk) + lane;
        float4 acc = make_float4(0, 0, 0, 0);
        for (int k = 0; k < dv; ++k) {
            const float4 v = sb[k * 32];
            acc.x += v.x; acc.y += v.y;
        }
        for (int k = 0; k < dv; ++k) {
            float4 o = sb[k * 32];
            o.x = acc.x - o.x; o.y = acc.y - o.y;
            sb[k * 32] = o;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            for (int k = 0; k < dv; ++k) {
                const size_t base = (size_t)__ldg(slot + (size_t)(first + i) * dv + k) * (kChunk / 16);
                bulk_s2g(pool + base, buf + (st * kMaxDv + k) * kChunk, kChunk);
            }
            bulk_commit();
        }
    }
    if (lane == 0) bulk_wait_read<0>();
}

// the plain version (16-byte loads and stores by the lanes), walking `items` items per warp
__global__ void __launch_bounds__(256) rmw_plain(float4 *pool, const uint32_t *slot, int n_items, int dv, int items) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int first = (blockIdx.x * nw + w) * items;
    const int mine = min(items, n_items - first);
    for (int i = 0; i < mine; ++i) {
        float4 v[kMaxDv], acc = make_float4(0, 0, 0, 0);
        size_t base[kMaxDv];
#pragma unroll
        for (int k = 0; k < kMaxDv; ++k)
            if (k < dv) {
                base[k] = (size_t)__ldg(slot + (size_t)(first + i) * dv + k) * (kChunk / 16);
                v[k] = __ldcg(pool + base[k] + lane);
            }
#pragma unroll
        for (int k = 0; k < kMaxDv; ++k)
            if (k < dv) { acc.x += v[k].x; acc.y += v[k].y; }
#pragma unroll
        for (int k = 0; k < kMaxDv; ++k)
            if (k < dv) {
                float4 o = v[k];
                o.x = acc.x - o.x; o.y = acc.y - o.y;
                __stcg(pool + base[k] + lane, o);
            }
    }
}

static float time_launch(void (*launch)(void *), void *ctx) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        launch(ctx);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    return ms;
}

struct Ctx {
    float4 *pool; uint32_t *slot; int n_items, dv, items, stages, warps;
};
static void launch_bulk(void *p) {
    Ctx *c = (Ctx *)p;
    const int per_cta = c->warps * c->items;
    const int grid = (c->n_items + per_cta - 1) / per_cta;
    const size_t sm = (size_t)c->warps * c->stages * kMaxDv * kChunk + (size_t)c->warps * c->stages * 8;
    if (c->stages == 2) {
        cudaFuncSetAttribute(rmw_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        rmw_bulk<2><<<grid, c->warps * 32, sm>>>(c->pool, c->slot, c->n_items, c->dv, c->items);
    } else if (c->stages == 3) {
        cudaFuncSetAttribute(rmw_bulk<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        rmw_bulk<3><<<grid, c->warps * 32, sm>>>(c->pool, c->slot, c->n_items, c->dv, c->items);
    } else {
        cudaFuncSetAttribute(rmw_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        rmw_bulk<4><<<grid, c->warps * 32, sm>>>(c->pool, c->slot, c->n_items, c->dv, c->items);
    }
}
static void launch_plain(void *p) {
    Ctx *c = (Ctx *)p;
    const int per_cta = c->warps * c->items;
    rmw_plain<<<(c->n_items + per_cta - 1) / per_cta, c->warps * 32>>>(c->pool, c->slot, c->n_items, c->dv, c->items);
}

int main() {
    const size_t pool_bytes = 4ull << 30;
    const size_t n_chunks = pool_bytes / kChunk;
    float4 *pool;
    uint32_t *d_slot, *h_slot = (uint32_t *)malloc(n_chunks * 4);
    cudaMalloc(&pool, pool_bytes);
    cudaMemset(pool, 0, pool_bytes);
    cudaMalloc(&d_slot, n_chunks * 4);
    for (size_t i = 0; i < n_chunks; ++i) h_slot[i] = (uint32_t)i;
    uint64_t s = 88172645463325252ull;
    for (size_t i = n_chunks - 1; i > 0; --i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const size_t j = s % (i + 1);
        const uint32_t t = h_slot[i]; h_slot[i] = h_slot[j]; h_slot[j] = t;
    }
    cudaMemcpy(d_slot, h_slot, n_chunks * 4, cudaMemcpyHostToDevice);
    for (int dv = 3; dv <= 4; ++dv) {
        Ctx c{pool, d_slot, (int)(n_chunks / dv), dv, 1, 2, 8};
        const double bytes = 2.0 * c.n_items * dv * kChunk;
        for (int items : {1, 8, 32}) {
            c.items = items;
            const float ms = time_launch(launch_plain, &c);
            printf("plain  dv %d items %2d            : %7.1f GB/s  %s\n", dv, items, bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
        for (int stages : {2, 3, 4})
            for (int warps : {4, 8})
                for (int items : {8, 32}) {
                    c.items = items; c.stages = stages; c.warps = warps;
                    const float ms = time_launch(launch_bulk, &c);
                    printf("bulk   dv %d items %2d stages %d warps %d: %7.1f GB/s  %s\n", dv, items, stages, warps, bytes / (ms * 1e-3) / 1e9,
                           cudaGetErrorString(cudaGetLastError()));
                }
    }
    return 0;
}

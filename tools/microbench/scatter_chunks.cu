// Micro-benchmark behind a design question of the streaming variable-node kernel: how does HBM3e bandwidth depend on the
// size of the contiguous chunk when a warp reads and rewrites `DV` chunks at scattered places of a 2 GiB pool?
// (The VN kernel touches one FT x 4-byte chunk per edge: 512 B at 4 frames per lane.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scatter_chunks scatter_chunks.cu && ./scatter_chunks
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CHUNK>   // bytes per chunk, multiple of 512 handled as CHUNK/512 float4 per lane; smaller: partial warp
__global__ void rmw(float4 *pool, const uint32_t *slot, int n_items, int dv) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= n_items) return;
    constexpr int PER_LANE = (CHUNK >= 512) ? CHUNK / 512 : 1;
    constexpr int LANES = (CHUNK >= 512) ? 32 : CHUNK / 16;
    float4 acc = make_float4(0, 0, 0, 0);
    float4 v[8][PER_LANE];
    for (int k = 0; k < dv; ++k) {
        const size_t base = (size_t)slot[item * dv + k] * (CHUNK / 16);
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j)
            if (lane < LANES) {
                v[k][j] = __ldcg(pool + base + j * 32 + lane);
                acc.x += v[k][j].x; acc.y += v[k][j].y;
            }
    }
    for (int k = 0; k < dv; ++k) {
        const size_t base = (size_t)slot[item * dv + k] * (CHUNK / 16);
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j)
            if (lane < LANES) {
                float4 o = v[k][j];
                o.x = acc.x - o.x; o.y = acc.y - o.y;
                __stcg(pool + base + j * 32 + lane, o);
            }
    }
}

template <int CHUNK>
static void run(float4 *pool, size_t pool_bytes, uint32_t *d_slot, uint32_t *h_slot, int dv) {
    const size_t n_chunks = pool_bytes / CHUNK;
    const int n_items = (int)(n_chunks / dv);
    // a random permutation of the chunks: every chunk is touched exactly once per launch (like the edge slots of a code)
    for (size_t i = 0; i < n_chunks; ++i) h_slot[i] = (uint32_t)i;
    uint64_t s = 88172645463325252ull;
    for (size_t i = n_chunks - 1; i > 0; --i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const size_t j = s % (i + 1);
        const uint32_t t = h_slot[i]; h_slot[i] = h_slot[j]; h_slot[j] = t;
    }
    cudaMemcpy(d_slot, h_slot, n_chunks * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int warps = 8;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        rmw<CHUNK><<<(n_items + warps - 1) / warps, warps * 32>>>(pool, d_slot, n_items, dv);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("chunk %5d B  dv %d : %8.1f GB/s (read + write)  %s\n", CHUNK, dv, 2.0 * n_items * dv * CHUNK / (ms * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t pool_bytes = 2ull << 30;
    float4 *pool;
    uint32_t *d_slot, *h_slot;
    cudaMalloc(&pool, pool_bytes);
    cudaMemset(pool, 0, pool_bytes);
    cudaMalloc(&d_slot, pool_bytes / 128 * 4);
    h_slot = (uint32_t *)malloc(pool_bytes / 128 * 4);
    for (int dv = 3; dv <= 6; dv += 3) {
        run<128>(pool, pool_bytes, d_slot, h_slot, dv);
        run<256>(pool, pool_bytes, d_slot, h_slot, dv);
        run<512>(pool, pool_bytes, d_slot, h_slot, dv);
        run<1024>(pool, pool_bytes, d_slot, h_slot, dv);
        run<2048>(pool, pool_bytes, d_slot, h_slot, dv);
    }
    return 0;
}

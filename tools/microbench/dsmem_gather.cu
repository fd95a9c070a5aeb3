// Micro-benchmark behind the question "can codes that do not fit one SM be decoded on chip by a thread-block cluster?"
// (VERDICT round 1, item 5): how fast are the on-chip kernels' gathers -- a 4-byte word per lane (L[bit], check phase) and a
// 16-byte record per lane (variable phase) at scattered addresses -- when the array is spread over the shared memory of the
// 2, 4 or 8 CTAs of a cluster, against the same gathers from the CTA's own shared memory?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dsmem_gather dsmem_gather.cu && ./dsmem_gather
// Every CTA owns 96 KB; a thread gathers from position hash(i) of the cluster-wide array (rank = pos / per-CTA size), so a
// fraction (C-1)/C of the accesses is remote. 8 independent gathers in flight per thread (the on-chip kernels' unrolling).
// Output: gathers per clock and SM, for local-only (cluster 1) and clusters of 2 / 4 / 8.
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

constexpr int kBytesPerCta = 96 * 1024;

template <typename V>
__device__ __forceinline__ V ld_cluster(unsigned addr);
template <>
__device__ __forceinline__ float ld_cluster<float>(unsigned addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
template <>
__device__ __forceinline__ float4 ld_cluster<float4>(unsigned addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <typename V>
__global__ void gather(unsigned long long *out_clocks, float *sink, int iters, int cluster_size) {
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    V *mine = reinterpret_cast<V *>(smem);
    constexpr int kPer = kBytesPerCta / sizeof(V);
    for (int i = threadIdx.x; i < kPer; i += blockDim.x) {
        V v;
        float *f = reinterpret_cast<float *>(&v);
        for (unsigned j = 0; j < sizeof(V) / 4; ++j) f[j] = (float)(i + j);
        mine[i] = v;
    }
    // shared::cluster window addresses of the 8 (or fewer) ranks' arrays: mapa + ld.shared::cluster, no generic loads
    unsigned base[8];
    const unsigned local = (unsigned)__cvta_generic_to_shared(mine);
    for (int r = 0; r < 8; ++r) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(base[r]) : "r"(local), "r"(r % cluster_size));
    cluster.sync();
    const unsigned total = (unsigned)kPer * (unsigned)cluster_size;
    unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        unsigned pos[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x = x * 1664525u + 1013904223u;
            pos[u] = (unsigned)(((unsigned long long)x * total) >> 32);
        }
        V v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ld_cluster<V>(base[pos[u] / kPer] + (pos[u] % kPer) * (unsigned)sizeof(V));
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += *reinterpret_cast<float *>(&v[u]);
    }
    const long long t1 = clock64();
    cluster.sync();   // no CTA may leave while its shared memory is still being read
    if (threadIdx.x == 0) out_clocks[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 123.456f) *sink = acc;
}

template <typename V>
static void run(const char *what, int cluster_size, int threads, int iters) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms / cluster_size * cluster_size;   // one CTA per SM
    unsigned long long *d_clk;
    float *d_sink;
    cudaMalloc(&d_clk, grid * sizeof(unsigned long long));
    cudaMalloc(&d_sink, 4);
    cudaFuncSetAttribute(gather<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytesPerCta);
    cudaFuncSetAttribute(gather<V>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = kBytesPerCta;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchKernelEx(&cfg, gather<V>, d_clk, d_sink, iters, cluster_size);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            printf("%-22s cluster %d: launch failed (%s)\n", what, cluster_size, cudaGetErrorString(e));
            return;
        }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    unsigned long long *h = new unsigned long long[grid];
    cudaMemcpy(h, d_clk, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    const double gathers = (double)threads / 32.0 * iters * 8;   // warp-wide gathers per CTA
    printf("%-22s cluster %d  %4d threads: %7.3f warp-gathers / clk / SM  = %6.1f B / clk / SM  (%.3f ms, remote share %.2f)\n", what,
           cluster_size, threads, gathers / mean, gathers * 32 * sizeof(V) / mean, ms, (cluster_size - 1.0) / cluster_size);
    delete[] h;
    cudaFree(d_clk);
    cudaFree(d_sink);
}

int main() {
    for (int threads : {512, 1024})
        for (int c : {1, 2, 4, 8}) {
            run<float>("4-byte gather", c, threads, 2000);
            run<float4>("16-byte gather", c, threads, 2000);
        }
    return 0;
}

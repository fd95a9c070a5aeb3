// On-chip sum-product path: the two kernel instantiations (SPA, SPA-lin-approx) and their launch geometry.
#include "handle.hpp"
#include "onchip_spa.cuh"

namespace qkhost {

using namespace qk;

template <int ALG>
static cudaError_t spa_geometry(int groups_cn, int sms, size_t smem, long long n_frames, int *threads, int *grid) {
    cudaError_t e = cudaFuncSetAttribute(onchip_spa_kernel<ALG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (*threads == 0) {
        // the CTA size that puts the most warps on an SM (shared memory decides how many CTAs fit), never more warps
        // than the check phase has groups; ties go to the smaller CTA
        const int cap = std::max(128, std::min(1024, (groups_cn * 32 + 127) / 128 * 128));
        int best = 0;
        for (int t = 128; t <= cap; t += 128) {
            int k = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, onchip_spa_kernel<ALG>, t, smem);
            if (e != cudaSuccess) return e;
            if (k * t > best) {
                best = k * t;
                *threads = t;
            }
        }
        if (*threads == 0) return cudaErrorLaunchOutOfResources;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, onchip_spa_kernel<ALG>, *threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid = (int)std::min<long long>(n_frames, (long long)per_sm * sms);   // persistent CTAs pull frames from a queue
    return cudaSuccess;
}

cudaError_t onchip_spa_geometry(int alg, int groups_cn, int sms, size_t smem, long long n_frames, int *threads, int *grid) {
    return alg == 0 ? spa_geometry<0>(groups_cn, sms, smem, n_frames, threads, grid)
                    : spa_geometry<1>(groups_cn, sms, smem, n_frames, threads, grid);
}

cudaError_t onchip_spa_launch(int alg, const OnchipArgs &a, int grid, int threads, size_t smem, cudaStream_t s) {
    if (alg == 0) onchip_spa_kernel<0><<<(unsigned)grid, threads, smem, s>>>(a);
    else onchip_spa_kernel<1><<<(unsigned)grid, threads, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace qkhost

// Shared device-side definitions of libqkdldpc_cuda (sm_100a).
//
// HBM layout of the frame-slot pool (DESIGN.md "Data layout"): frames are decoded in TILES of FT = 32*V slots; a
// warp's 32 lanes are 32*V frames of the same tile (lane l owns tile positions l*V .. l*V+V-1), so every access
// to a message of edge e is one fully coalesced 128-bit (V*sizeof(T) = 16 B) load/store per lane regardless of
// how irregular the Tanner graph is, and graph indices are read once per tile, not once per frame.
//   msg     [tile][E][FT]   T      messages, IN PLACE: bit->check after the VN kernel, check->bit after the CN
//                                  kernel, addressed by the CSR edge index e in both directions
//   bobmask [tile][N][V]    u32    Bob's key bits      (word v, bit l  <->  tile position l*V+v)
//   zmask   [tile][N][V]    u32    current hard decision
//   synd    [tile][M][V]    u32    Alice's syndrome
//   par     [tile][M][V]    u32    syndrome XOR parity of the current hard decision (0 <=> check satisfied)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace qk {

constexpr int kWarp = 32;
constexpr int kBuckets = 5;   // degree buckets per node type
constexpr int kSchedThreads = 256;

// Degree buckets: pick the register-array size of the CN (8/16/24/32/re-read) and VN (4/8/16/32/re-read) kernels.
__host__ __device__ constexpr int cn_bucket_max(int b) { return b == 0 ? 8 : b == 1 ? 16 : b == 2 ? 24 : b == 3 ? 32 : 0; }
__host__ __device__ constexpr int vn_bucket_max(int b) { return b == 0 ? 4 : b == 1 ? 8 : b == 2 ? 16 : b == 3 ? 32 : 0; }
__host__ __device__ inline int cn_bucket_of(int dc) { return dc <= 8 ? 0 : dc <= 16 ? 1 : dc <= 24 ? 2 : dc <= 32 ? 3 : 4; }
__host__ __device__ inline int vn_bucket_of(int dv) { return dv <= 4 ? 0 : dv <= 8 ? 1 : dv <= 16 ? 2 : dv <= 32 ? 3 : 4; }

template <typename T, int V>
struct alignas(sizeof(T) * V) Vec {
    T v[V];
};

template <typename T> struct Lim;
template <> struct Lim<float>  { __host__ __device__ static constexpr float  max() { return FLT_MAX; } };
template <> struct Lim<double> { __host__ __device__ static constexpr double max() { return DBL_MAX; } };

// L2-only ("cache global") accesses: the message stream has no L1 reuse.
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ld_msg(const T *p) {
    Vec<T, V> r;
    if constexpr (sizeof(T) * V == 16) {
        float4 t = __ldcg(reinterpret_cast<const float4 *>(p));
        r = *reinterpret_cast<Vec<T, V> *>(&t);
    } else if constexpr (sizeof(T) * V == 8) {
        float2 t = __ldcg(reinterpret_cast<const float2 *>(p));
        r = *reinterpret_cast<Vec<T, V> *>(&t);
    } else {
        static_assert(sizeof(T) * V == 4, "unsupported vector width");
        float t = __ldcg(reinterpret_cast<const float *>(p));
        r = *reinterpret_cast<Vec<T, V> *>(&t);
    }
    return r;
}
template <typename T, int V>
__device__ __forceinline__ void st_msg(T *p, const Vec<T, V> &x) {
    if constexpr (sizeof(T) * V == 16) {
        __stcg(reinterpret_cast<float4 *>(p), *reinterpret_cast<const float4 *>(&x));
    } else if constexpr (sizeof(T) * V == 8) {
        __stcg(reinterpret_cast<float2 *>(p), *reinterpret_cast<const float2 *>(&x));
    } else {
        __stcg(reinterpret_cast<float *>(p), *reinterpret_cast<const float *>(&x));
    }
}

// threshold_matrix (array_and_matrix_operations.cpp:953-972): NaN passes through.
template <typename T>
__device__ __forceinline__ T clamp_msg(T x, T thr) {
    if constexpr (sizeof(T) == 4) {
        // NaN-propagating min / max (FMNMX.NAN): the same result as the compare chain below for every input,
        // +-inf, -0 and NaN included, without the two branches
        float r;
        asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(-thr));
        asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(thr));
        return r;
    } else {
        if (x > thr) return thr;
        if (x < -thr) return -thr;
        return x;
    }
}

// Per-batch launch arguments shared by the step kernels.
template <typename T>
struct StepArgs {
    // graph (device)
    int n, m;
    const int *row_ptr;    // [m+1]
    const int *col_idx;    // [nnz]  CSR
    const int *col_ptr;    // [n+1]
    const int *csc_edge;   // [nnz]  k-th check of bit i (ascending) -> CSR edge id
    const int *csc_row;    // [nnz]  ... -> check id
    const int *row_order;  // [m]    rows grouped by degree bucket
    const int *col_order;  // [n]    bits grouped by degree bucket
    const int *vn_ell_edge, *vn_ell_row;   // narrow VN buckets (dv <= 4, dv <= 8): DVMAX edge / check ids per item, -1 padded
    const uint8_t *bitclass;  // [n] 0 payload, 1 punctured, 2 shortened
    // pool (device)
    T *msg;
    uint32_t *bobmask, *zmask, *synd, *par;
    uint32_t *tile_active;  // [tiles][V]
    uint32_t *tile_new;     // [tiles][V]
    T *slot_llr;            // [tiles][FT]  log((1-q)/q) of the frame in the slot
    int64_t e_stride;       // elements of msg per tile = nnz * FT
    // decoder parameters
    T primary, secondary, thr;
    int enable_thr;
};

}  // namespace qk

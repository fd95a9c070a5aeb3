// Check-node (K2a-f) and fused variable-node / hard-decision / syndrome-accumulate (K3) kernels.
//
// One warp = one Tanner-graph node x one tile of FT = 32*V frame slots; lane l handles slots l*V..l*V+V-1 with one
// 128-bit access per edge. Because all 32 lanes work on the same node, node degree is warp-uniform: irregular
// graphs cause no divergence, and rows / columns are only bucketed by degree to pick the register-array size.
//
// Arithmetic follows the reference operation by operation (same order, same comparisons, no FMA contraction:
// the library is compiled with -fmad=false), so float64 messages reproduce the reference bit for bit for the
// min-sum family and the linear-approximation SPA, and float32 messages reproduce "the reference with every
// double replaced by float" (oracle/ldpc_oracle_body.inc, f32 flavour) bit for bit.
#pragma once
#include "common.cuh"

namespace qk {

// ---------------------------------------------------------------------------------------------------------------
// A-priori LLR of bit `col` for this lane's V slots (QKD_LDPC, qkd_ldpc_algorithm.cpp:1043-1049; rate adaptation
// :1148-1174: punctured -> ALMOST_ZERO = 1e-4, shortened -> largest finite value).
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> lane_llr(const StepArgs<T> &a, int tile, int col, int lane, const Vec<T, V> &lp) {
    Vec<T, V> r;
    const uint8_t cls = __ldg(a.bitclass + col);
    if (cls == 0) {
        const uint32_t *bm = a.bobmask + ((size_t)tile * a.n + col) * V;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t w = __ldg(bm + v);
            r.v[v] = ((w >> lane) & 1u) ? -lp.v[v] : lp.v[v];
        }
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) r.v[v] = (cls == 1) ? (T)1e-4 : Lim<T>::max();
    }
    return r;
}

// tanh / atanh flavours. ALG 0: libm-accurate evaluation in double on the stored value (for float messages this is
// exactly what the f32 oracle does); ALG 1: the reference's piecewise-linear tables (qkd_ldpc_algorithm.cpp:146-172).
template <typename T, int ALG>
__device__ __forceinline__ T cn_tanh_half(T x) {
    const T h = x / (T)2;
    if constexpr (ALG == 0) {
        return (T)tanh((double)h);
    } else {
        const T ax = fabs(h);
        T r;
        if (ax < (T)0.5) r = (T)0.9242 * ax;
        else if (ax < (T)0.9) r = (T)0.6355 * ax + (T)0.1444;
        else if (ax < (T)1.2) r = (T)0.3912 * ax + (T)0.3642;
        else if (ax < (T)1.75) r = (T)0.1958 * ax + (T)0.5986;
        else if (ax < (T)2.5) r = (T)0.0603 * ax + (T)0.8358;
        else if (ax < (T)3.5) r = (T)0.0115 * ax + (T)0.9577;
        else if (ax < (T)8) r = (T)0.0004 * ax + (T)0.9967;
        else r = (T)1;
        return (h < (T)0) ? -r : r;
    }
}
template <typename T, int ALG>
__device__ __forceinline__ T cn_two_atanh(T y) {
    if constexpr (ALG == 0) {
        return (T)2 * (T)atanh((double)y);
    } else {
        const T ay = fabs(y);
        T r;
        if (ay < (T)0.7) r = (T)1.196 * ay - (T)0.0323;
        else if (ay < (T)0.9) r = (T)2.9187 * ay - (T)1.214;
        else if (ay < (T)0.999) r = (T)10.8717 * ay - (T)8.3717;
        else r = (T)2510.9 * ay - (T)2505.9;
        return (T)2 * ((y < (T)0) ? -r : r);
    }
}

// Running state of one check node for one slot.
template <typename T, int ALG>
struct RowState {
    T a, b;   // min-sum: min1, min2 ; SPA: row product, unused
    int neg;
    __device__ __forceinline__ void init(bool syn) {
        if constexpr (ALG <= 1) { a = syn ? (T)-1 : (T)1; b = (T)0; }
        else { a = Lim<T>::max(); b = Lim<T>::max(); }
        neg = 0;
    }
    // First pass over the row (qkd_ldpc_algorithm.cpp:55-63 / :381-397). For SPA returns tanh(m/2) (stored in place).
    __device__ __forceinline__ T absorb(T msg) {
        if constexpr (ALG <= 1) {
            const T t = cn_tanh_half<T, ALG>(msg);
            a *= t;
            return t;
        } else {
            if (msg < (T)0) ++neg;
            const T ab = fabs(msg);
            if (ab < a) { b = a; a = ab; }
            else if (ab < b) { b = ab; }
            return msg;
        }
    }
    // Second pass (:64-70 / :400-408, :573-574, :759-767, :949-956). `kept` is what absorb() returned for this edge.
    __device__ __forceinline__ T emit(T kept, bool syn, T factor) const {
        if constexpr (ALG <= 1) {
            return cn_two_atanh<T, ALG>(a / kept);
        } else {
            T sign_prod = syn ? (T)-1 : (T)1;
            sign_prod *= (neg % 2 == 0) ? (T)1 : (T)-1;
            const T prod = sign_prod * ((kept > (T)0) ? (T)1 : (T)-1);
            const T sel = (fabs(kept) == a) ? b : a;
            if constexpr (ALG == 2 || ALG == 4) {
                return factor * prod * sel;
            } else {
                const T diff = sel - factor;
                return prod * ((diff < (T)0) ? (T)0 : diff);
            }
        }
    }
};

// One check node (row) for this lane's V slots. DCMAX > 0: the row lives in registers between the two passes
// (one read + one write of every message); DCMAX == 0: rows wider than 32 edges are re-read in the second pass.
template <typename T, int V, int ALG, int DCMAX>
__device__ __forceinline__ void cn_row(const StepArgs<T> &a, int tile, int row, int lane, bool lane_act, bool any_new,
                                       const uint32_t (&newm)[V], const Vec<T, V> &lp) {
    constexpr int FT = kWarp * V;
    constexpr bool kAdaptive = (ALG >= 4);
    const int e0 = __ldg(a.row_ptr + row);
    const int dc = __ldg(a.row_ptr + row + 1) - e0;
    T *base = a.msg + (size_t)tile * a.e_stride + (size_t)e0 * FT + lane * V;

    // syndrome bit, row-satisfied bit of the previous hard decision, and reset of the parity accumulator
    const size_t rb = ((size_t)tile * a.m + row) * V;
    bool syn[V];
    T factor[V];
    {
        uint32_t sw[V], pw[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            sw[v] = a.synd[rb + v];
            pw[v] = kAdaptive ? a.par[rb + v] : 0u;
        }
        __syncwarp();
        if (lane < V) a.par[rb + lane] = sw[lane];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            syn[v] = (sw[v] >> lane) & 1u;
            // ANMSA/AOMSA: nu / sigma when the check is violated by the previous decision (:749-757, :939-947)
            factor[v] = (kAdaptive && ((pw[v] >> lane) & 1u)) ? a.secondary : a.primary;
        }
    }
    if (!lane_act) return;

    auto fetch = [&](int k) -> Vec<T, V> {
        Vec<T, V> x = ld_msg<T, V>(base + (size_t)k * FT);
        if (any_new) {   // slots refilled by the scheduler: first iteration reads the a-priori LLR (:21-29)
            const int col = __ldg(a.col_idx + e0 + k);
            const Vec<T, V> l = lane_llr<T, V>(a, tile, col, lane, lp);
#pragma unroll
            for (int v = 0; v < V; ++v)
                if ((newm[v] >> lane) & 1u) x.v[v] = l.v[v];
        }
        return x;
    };

    RowState<T, ALG> st[V];
#pragma unroll
    for (int v = 0; v < V; ++v) st[v].init(syn[v]);

    if constexpr (DCMAX > 0) {
        Vec<T, V> x[DCMAX];
#pragma unroll
        for (int k = 0; k < DCMAX; ++k)
            if (k < dc) x[k] = fetch(k);
#pragma unroll
        for (int k = 0; k < DCMAX; ++k)
            if (k < dc) {
#pragma unroll
                for (int v = 0; v < V; ++v) x[k].v[v] = st[v].absorb(x[k].v[v]);
            }
#pragma unroll
        for (int k = 0; k < DCMAX; ++k)
            if (k < dc) {
                Vec<T, V> o;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    T c = st[v].emit(x[k].v[v], syn[v], factor[v]);
                    o.v[v] = a.enable_thr ? clamp_msg(c, a.thr) : c;   // threshold_matrix(check_to_bit_msg), :73-74
                }
                st_msg<T, V>(base + (size_t)k * FT, o);
            }
    } else {
        for (int k = 0; k < dc; ++k) {
            Vec<T, V> x = fetch(k);
#pragma unroll
            for (int v = 0; v < V; ++v) st[v].absorb(x.v[v]);
        }
        for (int k = 0; k < dc; ++k) {
            Vec<T, V> x = fetch(k);
            Vec<T, V> o;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                T kept = x.v[v];
                if constexpr (ALG <= 1) kept = cn_tanh_half<T, ALG>(kept);
                T c = st[v].emit(kept, syn[v], factor[v]);
                o.v[v] = a.enable_thr ? clamp_msg(c, a.thr) : c;
            }
            st_msg<T, V>(base + (size_t)k * FT, o);
        }
    }
}

// K2: flooding check-node update, one kernel instantiation per algorithm variant.
// grid = (CTAs over degree-bucketed rows, tiles); block = kCnWarps warps, one row per warp.
template <typename T, int V, int ALG>
__global__ void __launch_bounds__(kCnWarps * kWarp) cn_kernel(const StepArgs<T> a) {
    constexpr int FT = kWarp * V;
    const int tile = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t newm[V];
    bool any_act = false, any_new = false, lane_act = false;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const uint32_t act = a.tile_active[tile * V + v];
        newm[v] = a.tile_new[tile * V + v];
        any_act |= act != 0;
        any_new |= newm[v] != 0;
        lane_act |= (act >> lane) & 1u;
    }
    if (!any_act) return;
    const int2 item = a.cn_items[blockIdx.x];
    if (warp >= (item.y >> 8)) return;
    const int row = __ldg(a.row_order + item.x + warp);
    const Vec<T, V> lp = *reinterpret_cast<const Vec<T, V> *>(a.slot_llr + (size_t)tile * FT + lane * V);
    switch (item.y & 255) {
        case 0: cn_row<T, V, ALG, 8>(a, tile, row, lane, lane_act, any_new, newm, lp); break;
        case 1: cn_row<T, V, ALG, 16>(a, tile, row, lane, lane_act, any_new, newm, lp); break;
        case 2: cn_row<T, V, ALG, 24>(a, tile, row, lane, lane_act, any_new, newm, lp); break;
        case 3: cn_row<T, V, ALG, 32>(a, tile, row, lane, lane_act, any_new, newm, lp); break;
        default: cn_row<T, V, ALG, 0>(a, tile, row, lane, lane_act, any_new, newm, lp); break;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: variable-node update + hard decision + syndrome accumulation for one bit and this lane's V slots.
//   L = llr + sum_k c2b_k in ascending check order (std::accumulate from the LLR, :78), z = (L <= 0) (:80-83),
//   b2c_k = clamp(L - c2b_k) (:109-123). The parity of z is XOR-ed into par[] of every check of the bit, so that
//   par == 0 over all rows <=> calculate_syndrome(z) == syndrome (:86,101).
template <typename T, int V, int DVMAX>
__device__ __forceinline__ void vn_col(const StepArgs<T> &a, int tile, int bit, int lane, bool lane_act,
                                       const uint32_t (&act)[V], const Vec<T, V> &lp) {
    constexpr int FT = kWarp * V;
    const int c0 = __ldg(a.col_ptr + bit);
    const int dv = __ldg(a.col_ptr + bit + 1) - c0;
    T *tbase = a.msg + (size_t)tile * a.e_stride + lane * V;
    Vec<T, V> L = lane_llr<T, V>(a, tile, bit, lane, lp);
    bool z[V];

    if constexpr (DVMAX > 0) {
        Vec<T, V> c[DVMAX];
        int e[DVMAX];
#pragma unroll
        for (int k = 0; k < DVMAX; ++k)
            if (k < dv) e[k] = __ldg(a.csc_edge + c0 + k);
        if (lane_act) {
#pragma unroll
            for (int k = 0; k < DVMAX; ++k)
                if (k < dv) c[k] = ld_msg<T, V>(tbase + (size_t)e[k] * FT);
#pragma unroll
            for (int k = 0; k < DVMAX; ++k)
                if (k < dv) {
#pragma unroll
                    for (int v = 0; v < V; ++v) L.v[v] = L.v[v] + c[k].v[v];
                }
#pragma unroll
            for (int k = 0; k < DVMAX; ++k)
                if (k < dv) {
                    Vec<T, V> o;
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        T s = L.v[v] - c[k].v[v];
                        o.v[v] = a.enable_thr ? clamp_msg(s, a.thr) : s;
                    }
                    st_msg<T, V>(tbase + (size_t)e[k] * FT, o);
                }
        }
    } else {
        if (lane_act) {
            for (int k = 0; k < dv; ++k) {
                const Vec<T, V> c = ld_msg<T, V>(tbase + (size_t)__ldg(a.csc_edge + c0 + k) * FT);
#pragma unroll
                for (int v = 0; v < V; ++v) L.v[v] = L.v[v] + c.v[v];
            }
            for (int k = 0; k < dv; ++k) {
                T *p = tbase + (size_t)__ldg(a.csc_edge + c0 + k) * FT;
                const Vec<T, V> c = ld_msg<T, V>(p);
                Vec<T, V> o;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    T s = L.v[v] - c.v[v];
                    o.v[v] = a.enable_thr ? clamp_msg(s, a.thr) : s;
                }
                st_msg<T, V>(p, o);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) z[v] = lane_act && (L.v[v] <= (T)0);   // NaN decides 0 (quirk Q2)

    // pack the decisions of the tile: word v, bit lane; only active slots contribute
    uint32_t zw = 0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const uint32_t w = __ballot_sync(0xffffffffu, z[v]) & act[v];
        if (lane == v) zw = w;
    }
    if (lane < V) {
        a.zmask[((size_t)tile * a.n + bit) * V + lane] = zw;
        if (zw != 0) {
            for (int k = 0; k < dv; ++k) {
                const int row = __ldg(a.csc_row + c0 + k);
                atomicXor(a.par + ((size_t)tile * a.m + row) * V + lane, zw);
            }
        }
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(kVnWarps * kWarp) vn_kernel(const StepArgs<T> a) {
    constexpr int FT = kWarp * V;
    const int tile = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t act[V];
    bool any_act = false, lane_act = false;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        act[v] = a.tile_active[tile * V + v];
        any_act |= act[v] != 0;
        lane_act |= (act[v] >> lane) & 1u;
    }
    if (!any_act) return;
    const int2 item = a.vn_items[blockIdx.x];
    if (warp >= (item.y >> 8)) return;
    const int bit = __ldg(a.col_order + item.x + warp);
    const Vec<T, V> lp = *reinterpret_cast<const Vec<T, V> *>(a.slot_llr + (size_t)tile * FT + lane * V);
    switch (item.y & 255) {
        case 0: vn_col<T, V, 4>(a, tile, bit, lane, lane_act, act, lp); break;
        case 1: vn_col<T, V, 8>(a, tile, bit, lane, lane_act, act, lp); break;
        case 2: vn_col<T, V, 16>(a, tile, bit, lane, lane_act, act, lp); break;
        case 3: vn_col<T, V, 32>(a, tile, bit, lane, lane_act, act, lp); break;
        default: vn_col<T, V, 0>(a, tile, bit, lane, lane_act, act, lp); break;
    }
}

}  // namespace qk

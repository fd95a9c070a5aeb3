// Check-node (K2a-f) and fused variable-node / hard-decision / syndrome-accumulate (K3) kernels.
//
// One warp = one Tanner-graph node x one tile of FT = 32*V frame slots; lane l handles slots l*V..l*V+V-1 with one
// 128-bit access per edge. All 32 lanes work on the same node, so node degree is warp-uniform: irregular graphs
// cause no divergence. Rows / columns are bucketed by degree and every bucket is its OWN kernel instantiation
// (register array sized for the bucket, own launch bounds), so a few wide nodes do not dictate the occupancy of
// the many narrow ones.
//
// Two arithmetic flavours, both operation-for-operation equal to the reference on their domain:
//  * EXACT  (any message type, any algorithm): the reference's comparison chains literally (NaN / inf / -0 behave
//    as in the C++ code). Used for float64 parity mode, for SPA / SPA-lin-approx, and whenever the fast flavour's
//    precondition does not hold.
//  * FAST   (float32 min-sum family): branch-free bit arithmetic -- sign parity by XOR of raw bits, min1/min2 by
//    FMNMX, the two possible output magnitudes computed once per row. Identical results to EXACT whenever no
//    message is NaN, which the host guarantees before selecting it (clamp enabled, or all factors <= 1; see
//    fast_minsum_ok in run_batch.cuh). Exact zeros (quirk Q4) are handled: a zero message is "positive" for the
//    row parity and "negative" for its own exclusion.
// The library is compiled with -fmad=false: no FMA contraction anywhere.
#pragma once
#include "common.cuh"
#include "spa_f64_math.hpp"

namespace qk {

// ---------------------------------------------------------------------------------------------------------------
// A-priori LLR of bit `col` for this lane's V slots (QKD_LDPC, qkd_ldpc_algorithm.cpp:1043-1049; rate adaptation
// :1148-1174: punctured -> ALMOST_ZERO = 1e-4, shortened -> largest finite value).
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> lane_llr(const StepArgs<T> &a, int tile, int col, int lane, const Vec<T, V> &lp) {
    Vec<T, V> r;
    const uint8_t cls = __ldg(a.bitclass + col);
    if (cls == 0) {
        const Vec<uint32_t, V> bm = *reinterpret_cast<const Vec<uint32_t, V> *>(a.bobmask + ((size_t)tile * a.n + col) * V);
#pragma unroll
        for (int v = 0; v < V; ++v) r.v[v] = ((bm.v[v] >> lane) & 1u) ? -lp.v[v] : lp.v[v];
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) r.v[v] = (cls == 1) ? (T)1e-4 : Lim<T>::max();
    }
    return r;
}

// tanh / atanh flavours.
//  ALG 0, double messages: branch-free double evaluation, <= 5 ulp from glibc (spa_f64_math.hpp; parity mode).
//  ALG 0, float messages:  float tanh (spa_tanh_half_f32, <= 3 ulp) and a fused divide + logarithm for 2 atanh (RowState::emit).
//         An earlier version evaluated
//         them in double on the float value (bit-equal to the f32 oracle) but spent 92 % of the SPA step in FP64
//         transcendentals (2.1 s of a 2.3 s step); the float versions keep >= 99 % equal iteration counts against the
//         float64 reference (tests/test_gpu_large.py) at a tenth of the time.
//  ALG 1: the reference's piecewise-linear tables (qkd_ldpc_algorithm.cpp:146-172), written branch-free: slope and
//         intercept are selected with the same `<` comparisons (a NaN fails them all and lands in the last segment,
//         like the reference's else-chain), then one multiply and one add -- the same two roundings as the C++ code.
// tanh(m / 2) in float32, <= 3 ulp: 1 - 2 / (exp(2|h|) + 1) with the sign of h for |h| >= 0.6 (exp2 and reciprocal on the
// MUFU unit; exp = +inf gives exactly 1), an odd minimax polynomial h + h u p(u), u = h^2, below (coefficients fitted
// for this file: tools/fit_tanh_poly.py). Both branches are evaluated and selected, NaN falls through to the polynomial.
__device__ __forceinline__ float spa_tanh_half_f32(float m) {
    const float h = 0.5f * m, ah = fabsf(h);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ah * 2.885390043f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    const float big = copysignf(fmaf(-2.f, r, 1.f), h);
    const float u = h * h;
    float p = fmaf(u, 0.015640417113900185f, -0.052231062203645706f);
    p = fmaf(u, p, 0.13313519954681396f);
    p = fmaf(u, p, -0.33332622051239014f);
    const float small = fmaf(h, p * u, h);
    return (ah >= 0.6f) ? big : small;
}

template <typename T, int ALG>
__device__ __forceinline__ T cn_tanh_half(T x) {
    const T h = x / (T)2;
    if constexpr (ALG == 0) {
        if constexpr (sizeof(T) == 4) return spa_tanh_half_f32(x);
        else return (T)spa_tanh_half_f64((double)x);
    } else {
        const T ax = fabs(h);
        T a = (T)0.9242, b = (T)0;
        if (!(ax < (T)0.5)) { a = (T)0.6355; b = (T)0.1444; }
        if (!(ax < (T)0.9)) { a = (T)0.3912; b = (T)0.3642; }
        if (!(ax < (T)1.2)) { a = (T)0.1958; b = (T)0.5986; }
        if (!(ax < (T)1.75)) { a = (T)0.0603; b = (T)0.8358; }
        if (!(ax < (T)2.5)) { a = (T)0.0115; b = (T)0.9577; }
        if (!(ax < (T)3.5)) { a = (T)0.0004; b = (T)0.9967; }
        T r = a * ax + b;                       // first segment: 0.9242 * ax + 0 == 0.9242 * ax (ax >= 0)
        if (!(ax < (T)8)) r = (T)1;
        return (h < (T)0) ? -r : r;
    }
}
template <typename T, int ALG>
__device__ __forceinline__ T cn_two_atanh(T y) {
    if constexpr (ALG == 0) {
        if constexpr (sizeof(T) == 4) return 2.f * atanhf(y);
        else return (T)spa_two_atanh_f64((double)y);
    } else {
        const T ay = fabs(y);
        T a = (T)1.196, b = (T)0.0323;
        if (!(ay < (T)0.7)) { a = (T)2.9187; b = (T)1.214; }
        if (!(ay < (T)0.9)) { a = (T)10.8717; b = (T)8.3717; }
        if (!(ay < (T)0.999)) { a = (T)2510.9; b = (T)2505.9; }
        const T r = a * ay - b;
        return (T)2 * ((y < (T)0) ? -r : r);
    }
}

// EXACT flavour: running state of one check node for one slot.
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
            const bool lt1 = ab < a, lt2 = ab < b;
            b = lt1 ? a : (lt2 ? ab : b);
            a = lt1 ? ab : a;
            return msg;
        }
    }
    // Second pass (:64-70 / :400-408, :573-574, :759-767, :949-956). `kept` is what absorb() returned for this edge.
    __device__ __forceinline__ T emit(T kept, bool syn, T factor) const {
        if constexpr (ALG == 0 && sizeof(T) == 4) {
            // float SPA: 2 atanh(P / t) = ln((|t| + P') / (|t| - P')) with P' = P * sgn(t): one approximate reciprocal
            // and one MUFU logarithm instead of a divide, a divide inside atanhf and a log1pf (the accurate logf alone
            // was 28 of the 77 instructions per edge). Special values follow the reference's formula (quirk Q3):
            //   P' == |t| != 0 (all other tanh saturated to 1)  ->  x / +0 = +inf -> ln = +inf  (atanh(1), clamped)
            //   P' == -|t|                                     ->  0 / 2|t| = 0  -> ln = -inf  (atanh(-1), clamped)
            //   t == 0                                         ->  P / -P = -1, or 0 * inf -> NaN  (0/0 there)
            // |P| <= |t| always holds (a running product of factors <= 1 never grows), so the quotient is never
            // negative for another reason. Folding sgn(t) into P keeps the +0 denominator's sign from flipping the
            // infinity when t < 0.
            const float ps = __uint_as_float(__float_as_uint(a) ^ (__float_as_uint(kept) & 0x80000000u));
            const float at = fabsf(kept);
            float r, l;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(at - ps));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"((at + ps) * r));
            return l * 0.693147182464599609375f;
        } else if constexpr (ALG <= 1) {
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

__host__ __device__ constexpr int cn_threads(int elem_bytes, int V, int DCMAX) {
    return (DCMAX == 0 || DCMAX * V * elem_bytes <= 256) ? 256 : 128;
}
__host__ __device__ constexpr int vn_threads(int elem_bytes, int V, int DVMAX) {
    return (DVMAX == 0 || DVMAX * V * elem_bytes <= 128) ? 256 : 128;
}

// K2: flooding check-node update. One kernel instantiation per (algorithm variant, degree bucket, flavour).
// grid = (ceil(count / warps per CTA), tiles); one row per warp; rows [first, first+count) of row_order.
// DCMAX > 0: the row lives in registers between the two passes (one read + one write of every message);
// DCMAX == 0: rows wider than 32 edges are re-read in the second pass.
template <typename T, int V, int ALG, int DCMAX, bool FAST>
__global__ void __launch_bounds__(cn_threads(sizeof(T), V, DCMAX))
cn_kernel(const StepArgs<T> a, const int first, const int count) {
    constexpr int FT = kWarp * V;
    constexpr bool kAdaptive = (ALG >= 4);
    static_assert(!FAST || (sizeof(T) == 4 && ALG >= 2), "fast flavour: float32 min-sum family only");
    const int tile = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // slots that take part in this check-node pass: active and not waiting for their VN-side initialisation
    bool any_act = false, lane_act = false;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const uint32_t act = a.tile_active[tile * V + v] & ~a.tile_new[tile * V + v];
        any_act |= act != 0;
        lane_act |= (act >> lane) & 1u;
    }
    if (!any_act || idx >= count) return;
    const int row = __ldg(a.row_order + first + idx);
    const int e0 = __ldg(a.row_ptr + row);
    const int dc = __ldg(a.row_ptr + row + 1) - e0;
    T *base = a.msg + (size_t)tile * a.e_stride + (size_t)e0 * FT + lane * V;

    // syndrome bit, check value of the previous hard decision, and reset of the parity accumulator
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

    if constexpr (FAST) {
        // ---- float32 min-sum, branch-free ---------------------------------------------------------------------
        static_assert(DCMAX > 0, "fast flavour keeps the row in registers");
        Vec<float, V> x[DCMAX];
#pragma unroll
        for (int k = 0; k < DCMAX; ++k)
            if (k < dc) x[k] = ld_msg<float, V>(base + (size_t)k * FT);
        float m1[V], m2[V];
        uint32_t sx[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { m1[v] = FLT_MAX; m2[v] = FLT_MAX; sx[v] = syn[v] ? 0x80000000u : 0u; }
#pragma unroll
        for (int k = 0; k < DCMAX; ++k)
            if (k < dc) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float xv = x[k].v[v];
                    sx[v] ^= __float_as_uint(xv);                 // sign parity lives in bit 31 (:383,398)
                    const float ax = fabsf(xv);
                    m2[v] = fminf(m2[v], fmaxf(ax, m1[v]));       // == the reference's if / else-if chain (:386-396)
                    m1[v] = fminf(m1[v], ax);
                }
            }
        uint32_t o1[V], o2[V];   // signed outputs for "own |msg| != min1" / "== min1", row sign already applied
#pragma unroll
        for (int v = 0; v < V; ++v) {
            float c1, c2;
            if constexpr (ALG == 2 || ALG == 4) {
                c1 = factor[v] * m1[v];
                c2 = factor[v] * m2[v];
            } else {
                const float d1 = m1[v] - factor[v], d2 = m2[v] - factor[v];
                c1 = (d1 < 0.f) ? 0.f : d1;
                c2 = (d2 < 0.f) ? 0.f : d2;
            }
            c1 = fminf(c1, a.thr); c2 = fminf(c2, a.thr);   // magnitudes: clamp is symmetric (thr = +inf when disabled)
            const uint32_t rs = sx[v] & 0x80000000u;
            o1[v] = __float_as_uint(c1) ^ rs;
            // a message that is exactly 0 has min1 == 0, takes the "== min1" output and counts as negative for its
            // own exclusion ((m > 0) ? 1 : -1, :402) although its sign bit is clear (quirk Q4)
            o2[v] = __float_as_uint(c2) ^ rs ^ ((m1[v] == 0.f) ? 0x80000000u : 0u);
        }
#pragma unroll
        for (int k = 0; k < DCMAX; ++k)
            if (k < dc) {
                Vec<float, V> o;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float xv = x[k].v[v];
                    const uint32_t sel = (fabsf(xv) == m1[v]) ? o2[v] : o1[v];
                    o.v[v] = __uint_as_float(sel ^ (__float_as_uint(xv) & 0x80000000u));
                }
                st_msg<float, V>(base + (size_t)k * FT, o);
            }
    } else {
        // ---- exact flavour ------------------------------------------------------------------------------------
        RowState<T, ALG> st[V];
#pragma unroll
        for (int v = 0; v < V; ++v) st[v].init(syn[v]);
        if constexpr (DCMAX > 0) {
            Vec<T, V> x[DCMAX];
#pragma unroll
            for (int k = 0; k < DCMAX; ++k)
                if (k < dc) x[k] = ld_msg<T, V>(base + (size_t)k * FT);
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
                        o.v[v] = clamp_msg(c, a.thr);   // threshold_matrix(check_to_bit_msg), :73-74; +inf = off
                    }
                    st_msg<T, V>(base + (size_t)k * FT, o);
                }
        } else {
            // float64 sum-product: tanh is the expensive half of the row, so it is stored in place by the first pass and
            // read back by the second (one more write of the row; the kernel is far from the HBM roofline). Everything
            // else recomputes it (float32 tanh and the piecewise-linear table cost less than the extra traffic).
            constexpr bool kKeepT = (ALG == 0 && sizeof(T) == 8);
#pragma unroll 2
            for (int k = 0; k < dc; ++k) {
                Vec<T, V> x = ld_msg<T, V>(base + (size_t)k * FT);
#pragma unroll
                for (int v = 0; v < V; ++v) x.v[v] = st[v].absorb(x.v[v]);
                if constexpr (kKeepT) st_msg<T, V>(base + (size_t)k * FT, x);
            }
#pragma unroll 2
            for (int k = 0; k < dc; ++k) {
                Vec<T, V> x = ld_msg<T, V>(base + (size_t)k * FT);
                Vec<T, V> o;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    T kept = x.v[v];
                    if constexpr (ALG <= 1 && !kKeepT) kept = cn_tanh_half<T, ALG>(kept);
                    T c = st[v].emit(kept, syn[v], factor[v]);
                    o.v[v] = clamp_msg(c, a.thr);
                }
                st_msg<T, V>(base + (size_t)k * FT, o);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: variable-node update + hard decision + syndrome accumulation; one bit per warp, this lane's V slots.
//   L = llr + sum_k c2b_k in ascending check order (std::accumulate from the LLR, :78), z = (L <= 0) (:80-83),
//   b2c_k = clamp(L - c2b_k) (:109-123). The parity of z is XOR-ed into par[] of every check of the bit, so that
//   par == 0 over all rows <=> calculate_syndrome(z) == syndrome (:86,101).
// Slots the scheduler has just refilled ("new") are initialised here instead: b2c_k = llr (not clamped, :21-29),
// z = (llr <= 0) (the adaptive variants' initial decision, :683-691).
// FAST: clamp by FMNMX (valid when no message can be NaN); otherwise the reference's compare chain.
template <typename T, int V, bool FAST, bool HASNEW>
__device__ __forceinline__ Vec<T, V> vn_out(const StepArgs<T> &a, const Vec<T, V> &L, const Vec<T, V> &c,
                                            const Vec<T, V> &llr, const bool (&isnew)[V]) {
    Vec<T, V> o;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        T s = L.v[v] - c.v[v];
        if constexpr (FAST) s = fminf(fmaxf(s, -a.thr), a.thr);   // thr = +inf when the clamp is disabled
        else s = clamp_msg(s, a.thr);
        if constexpr (HASNEW) s = isnew[v] ? llr.v[v] : s;
        o.v[v] = s;
    }
    return o;
}

template <typename T, int V, int DVMAX, bool FAST, bool HASNEW>
__device__ __forceinline__ void vn_body(const StepArgs<T> &a, int tile, int bit, int lane, bool lane_act,
                                        const uint32_t (&act)[V], const uint32_t (&newm)[V]) {
    constexpr int FT = kWarp * V;
    const int c0 = __ldg(a.col_ptr + bit);
    const int dv = __ldg(a.col_ptr + bit + 1) - c0;
    T *tbase = a.msg + (size_t)tile * a.e_stride + lane * V;
    const Vec<T, V> lp = *reinterpret_cast<const Vec<T, V> *>(a.slot_llr + (size_t)tile * FT + lane * V);
    const Vec<T, V> llr = lane_llr<T, V>(a, tile, bit, lane, lp);
    Vec<T, V> L = llr;
    bool isnew[V];
#pragma unroll
    for (int v = 0; v < V; ++v) isnew[v] = HASNEW && ((newm[v] >> lane) & 1u);

    if (lane_act) {
        if constexpr (DVMAX > 0) {
            Vec<T, V> c[DVMAX];
#pragma unroll
            for (int k = 0; k < DVMAX; ++k)
                if (k < dv) c[k] = ld_msg<T, V>(tbase + (size_t)__ldg(a.csc_edge + c0 + k) * FT);
#pragma unroll
            for (int k = 0; k < DVMAX; ++k)
                if (k < dv) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if constexpr (HASNEW) c[k].v[v] = isnew[v] ? (T)0 : c[k].v[v];
                        L.v[v] = L.v[v] + c[k].v[v];
                    }
                }
#pragma unroll
            for (int k = 0; k < DVMAX; ++k)
                if (k < dv)
                    st_msg<T, V>(tbase + (size_t)__ldg(a.csc_edge + c0 + k) * FT,
                                 vn_out<T, V, FAST, HASNEW>(a, L, c[k], llr, isnew));
        } else {
            for (int k = 0; k < dv; ++k) {
                Vec<T, V> c = ld_msg<T, V>(tbase + (size_t)__ldg(a.csc_edge + c0 + k) * FT);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    if constexpr (HASNEW) c.v[v] = isnew[v] ? (T)0 : c.v[v];
                    L.v[v] = L.v[v] + c.v[v];
                }
            }
            for (int k = 0; k < dv; ++k) {
                T *p = tbase + (size_t)__ldg(a.csc_edge + c0 + k) * FT;
                Vec<T, V> c = ld_msg<T, V>(p);
                if constexpr (HASNEW) {
#pragma unroll
                    for (int v = 0; v < V; ++v) c.v[v] = isnew[v] ? (T)0 : c.v[v];
                }
                st_msg<T, V>(p, vn_out<T, V, FAST, HASNEW>(a, L, c, llr, isnew));
            }
        }
    }
    // pack the decisions of the tile: word v, bit lane; only active slots contribute. NaN decides 0 (quirk Q2).
    uint32_t zw = 0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const uint32_t w = __ballot_sync(0xffffffffu, lane_act && (L.v[v] <= (T)0)) & act[v];
        if (lane == v) zw = w;
    }
    if (lane < V) {
        a.zmask[((size_t)tile * a.n + bit) * V + lane] = zw;
        if (zw != 0) {
            for (int k = 0; k < dv; ++k)
                atomicXor(a.par + ((size_t)tile * a.m + __ldg(a.csc_row + c0 + k)) * V + lane, zw);
        }
    }
}

template <typename T, int V, int DVMAX, bool FAST>
__global__ void __launch_bounds__(vn_threads(sizeof(T), V, DVMAX))
vn_kernel(const StepArgs<T> a, const int first, const int count) {
    const int tile = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t act[V], newm[V];
    bool any_act = false, any_new = false, lane_act = false;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        act[v] = a.tile_active[tile * V + v];
        newm[v] = a.tile_new[tile * V + v];
        any_act |= act[v] != 0;
        any_new |= newm[v] != 0;
        lane_act |= (act[v] >> lane) & 1u;
    }
    if (!any_act || idx >= count) return;
    const int bit = __ldg(a.col_order + first + idx);
    if (any_new) vn_body<T, V, DVMAX, FAST, true>(a, tile, bit, lane, lane_act, act, newm);
    else vn_body<T, V, DVMAX, FAST, false>(a, tile, bit, lane, lane_act, act, newm);
}

// ---------------------------------------------------------------------------------------------------------------
// K3 for the narrow buckets (DVMAX = 4, 8): the same arithmetic as vn_body, with the loads arranged in TWO dependent
// round trips instead of six. A bit of 3-4 edges moves only 3-4 KB, so the kernel lives or dies by how many loads are in
// flight: vn_body chases col_order -> col_ptr -> (bitclass -> bobmask) -> csc_edge -> messages -> csc_row -> RED one after
// the other (ncu: 75 % of all stall samples sit on those six waits, 4.3 TB/s on the n = 102400 code, while a bare
// scattered 512-byte read-modify-write kernel reaches 5.7 TB/s, tools/microbench/scatter_chunks.cu). Here everything
// that depends only on (tile, item) is fetched at once -- tile masks, slot LLR magnitudes, the bit id and the item's
// ELL record of edge and check ids (vn_ell_edge / vn_ell_row, -1 = padding) -- and everything that needs the bit id or
// the edge ids in a second wave (bit class, Bob's mask word unconditionally, the messages).
template <int V>
__device__ __forceinline__ void vn_lane_words(const uint32_t (&act)[V], const uint32_t (&newm)[V], int lane,
                                              uint32_t &actl, uint32_t &newl, bool &stale) {
    actl = 0;
    newl = 0;
    uint32_t old = 0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        actl = (lane == v) ? act[v] : actl;
        newl |= ((newm[v] >> lane) & 1u) << v;
        old |= (act[v] & ~newm[v]) >> lane;
    }
    // stale: some active slot of this lane carries messages of earlier iterations. A lane whose active slots were all just
    // refilled (every lane in the first step of a batch) writes b2c = llr without reading the pool's leftovers.
    stale = (old & 1u) != 0;
}
template <typename T, int V, int DVMAX, bool FAST, bool HASNEW>
__device__ __forceinline__ void vn_body_ell(const StepArgs<T> &a, int tile, int bit, int lane, bool lane_act,
                                            const uint32_t actl, const uint32_t newl, const bool stale,
                                            const int (&e)[DVMAX], const int (&r)[DVMAX], const Vec<T, V> &lp) {
    // actl: the tile's active mask of word `lane` (lanes < V, else 0); newl: bit v set <=> this lane's slot v was just refilled
    constexpr int FT = kWarp * V;
    T *tbase = a.msg + (size_t)tile * a.e_stride + lane * V;
    // second wave of loads
    const uint8_t cls = __ldg(a.bitclass + bit);
    const Vec<uint32_t, V> bm = *reinterpret_cast<const Vec<uint32_t, V> *>(a.bobmask + ((size_t)tile * a.n + bit) * V);
    Vec<T, V> c[DVMAX];
    const bool ld = lane_act && (!HASNEW || stale);
#pragma unroll
    for (int k = 0; k < DVMAX; ++k) {
        if constexpr (HASNEW) c[k] = Vec<T, V>{};
        if (ld && e[k] >= 0) c[k] = ld_msg<T, V>(tbase + (size_t)e[k] * FT);
    }
    // a-priori LLR (lane_llr)
    Vec<T, V> llr;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const T pay = ((bm.v[v] >> lane) & 1u) ? -lp.v[v] : lp.v[v];
        llr.v[v] = (cls == 0) ? pay : ((cls == 1) ? (T)1e-4 : Lim<T>::max());
    }
    Vec<T, V> L = llr;
    bool isnew[V];
#pragma unroll
    for (int v = 0; v < V; ++v) isnew[v] = HASNEW && ((newl >> v) & 1u);
    if (lane_act) {
#pragma unroll
        for (int k = 0; k < DVMAX; ++k)
            if (e[k] >= 0) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    if constexpr (HASNEW) c[k].v[v] = isnew[v] ? (T)0 : c[k].v[v];
                    L.v[v] = L.v[v] + c[k].v[v];
                }
            }
#pragma unroll
        for (int k = 0; k < DVMAX; ++k)
            if (e[k] >= 0) st_msg<T, V>(tbase + (size_t)e[k] * FT, vn_out<T, V, FAST, HASNEW>(a, L, c[k], llr, isnew));
    }
    uint32_t zw = 0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const uint32_t w = __ballot_sync(0xffffffffu, lane_act && (L.v[v] <= (T)0));
        if (lane == v) zw = w;
    }
    zw &= actl;
    if (lane < V) {
        a.zmask[((size_t)tile * a.n + bit) * V + lane] = zw;
        if (zw != 0) {
#pragma unroll
            for (int k = 0; k < DVMAX; ++k)
                if (r[k] >= 0) atomicXor(a.par + ((size_t)tile * a.m + r[k]) * V + lane, zw);
        }
    }
}

// dv <= 4, float32: 6 CTAs of 256 threads per SM (<= 42 registers); dv <= 8: 3 CTAs (<= 85) -- occupancy is bytes in
// flight here (measured on n = 10240 irregular: VN 0.74 -> 0.78 of the HBM roofline at 3 CTAs, 0.77 at 4)
#ifndef QK_VN4_CTAS
#define QK_VN4_CTAS 6
#endif
#ifndef QK_VN8_CTAS
#define QK_VN8_CTAS 3
#endif
__host__ __device__ constexpr int vn_ell_min_ctas(int elem_bytes, int V, int DVMAX) {
    return (elem_bytes == 4 && DVMAX == 4) ? QK_VN4_CTAS : (elem_bytes == 4 && DVMAX == 8) ? QK_VN8_CTAS : 1;
}
template <typename T, int V, int DVMAX, bool FAST>
__global__ void __launch_bounds__(vn_threads(sizeof(T), V, DVMAX), vn_ell_min_ctas(sizeof(T), V, DVMAX))
vn_kernel_ell(const StepArgs<T> a, const int first, const int count, const int ell_base) {
    static_assert(DVMAX == 4 || DVMAX == 8, "ELL records exist for the two narrow buckets");
    constexpr int FT = kWarp * V;
    const int tile = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= count) return;
    // first wave of loads: nothing here depends on another load
    uint32_t act[V], newm[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        act[v] = a.tile_active[tile * V + v];
        newm[v] = a.tile_new[tile * V + v];
    }
    const int bit = __ldg(a.col_order + first + idx);
    int e[DVMAX], r[DVMAX];
    {
        const int4 *pe = reinterpret_cast<const int4 *>(a.vn_ell_edge + ell_base + (size_t)idx * DVMAX);
        const int4 *pr = reinterpret_cast<const int4 *>(a.vn_ell_row + ell_base + (size_t)idx * DVMAX);
#pragma unroll
        for (int j = 0; j < DVMAX / 4; ++j) {
            const int4 x = __ldg(pe + j), y = __ldg(pr + j);
            e[4 * j] = x.x; e[4 * j + 1] = x.y; e[4 * j + 2] = x.z; e[4 * j + 3] = x.w;
            r[4 * j] = y.x; r[4 * j + 1] = y.y; r[4 * j + 2] = y.z; r[4 * j + 3] = y.w;
        }
    }
    const Vec<T, V> lp = *reinterpret_cast<const Vec<T, V> *>(a.slot_llr + (size_t)tile * FT + lane * V);
    bool any_act = false, any_new = false, lane_act = false;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        any_act |= act[v] != 0;
        any_new |= newm[v] != 0;
        lane_act |= (act[v] >> lane) & 1u;
    }
    if (!any_act) return;
    uint32_t actl, newl;
    bool stale;
    vn_lane_words<V>(act, newm, lane, actl, newl, stale);
    if (any_new) vn_body_ell<T, V, DVMAX, FAST, true>(a, tile, bit, lane, lane_act, actl, newl, stale, e, r, lp);
    else vn_body_ell<T, V, DVMAX, FAST, false>(a, tile, bit, lane, lane_act, actl, newl, stale, e, r, lp);
}


// The same kernel with SEVERAL items per warp (qkdldpc_options.vn_items_per_warp): a warp of vn_kernel_ell spends about a
// third of its life in the first wave (index loads, L2 hits) with no message bytes in flight, and every CTA pays its launch
// for eight items only. Here a warp walks items idx, idx + W, idx + 2 W ... (W warps per CTA, so the CTA's warps stay on
// adjacent ELL records: one 128-byte line of edge ids per eight items). Right behind an item's message loads the warp
// PREFETCHES INTO L1 the lines it will need next -- the next item's bit id and edge record, this item's check ids -- so that
// from the second item on the index loads are L1 hits and only the second wave (the messages) is exposed. A prefetch costs
// no registers (holding the next record in registers did: 48 registers still spilled 170 bytes); the check ids are loaded
// last, just before the parity RED, and the slots' LLR magnitudes come from L1 with every item instead of staying live.
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <typename T, int V, int DVMAX>
__device__ __forceinline__ void vn_ell_record(const int *tab, int (&x)[DVMAX]) {
    const int4 *p = reinterpret_cast<const int4 *>(tab);
#pragma unroll
    for (int j = 0; j < DVMAX / 4; ++j) {
        const int4 t = __ldg(p + j);
        x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
    }
}
template <typename T, int V, int DVMAX, bool FAST, bool HASNEW, int L2PF>
__device__ __forceinline__ void vn_items_ell(const StepArgs<T> &a, int tile, int lane, bool lane_act,
                                             const uint32_t actl, const uint32_t newl, const bool stale, const int first,
                                             const int count, const int ell_base, int idx, const int stride, int items) {
    constexpr int FT = kWarp * V;
    T *tbase = a.msg + (size_t)tile * a.e_stride + lane * V;
    const Vec<T, V> *lpp = reinterpret_cast<const Vec<T, V> *>(a.slot_llr + (size_t)tile * FT + lane * V);
    bool isnew[V];
#pragma unroll
    for (int v = 0; v < V; ++v) isnew[v] = HASNEW && ((newl >> v) & 1u);
    const bool ld = lane_act && (!HASNEW || stale);
    for (;;) {
        // first wave: L1 hits from the warp's second item on
        const int bit = __ldg(a.col_order + first + idx);
        int e[DVMAX];
        vn_ell_record<T, V, DVMAX>(a.vn_ell_edge + ell_base + (size_t)idx * DVMAX, e);
        const int *rows = a.vn_ell_row + ell_base + (size_t)idx * DVMAX;
        // second wave
        const uint8_t cls = __ldg(a.bitclass + bit);
        const Vec<uint32_t, V> bm = *reinterpret_cast<const Vec<uint32_t, V> *>(a.bobmask + ((size_t)tile * a.n + bit) * V);
        const Vec<T, V> lp = *lpp;
        Vec<T, V> c[DVMAX];
#pragma unroll
        for (int k = 0; k < DVMAX; ++k) {
            if constexpr (HASNEW) c[k] = Vec<T, V>{};
            if (ld && e[k] >= 0) c[k] = ld_msg<T, V>(tbase + (size_t)e[k] * FT);
        }
        idx += stride;
        const bool more = --items > 0 && idx < count;   // warp-uniform
        prefetch_l1(rows);
        if constexpr (DVMAX == 8) prefetch_l1(rows + 4);
        if constexpr (L2PF == 0) {
            if (more) {
                prefetch_l1(a.col_order + first + idx);
                prefetch_l1(a.vn_ell_edge + ell_base + (size_t)idx * DVMAX);
                if constexpr (DVMAX == 8) prefetch_l1(a.vn_ell_edge + ell_base + (size_t)idx * DVMAX + 4);
            }
        } else {
            // L2PF = D (default 1): the edge record of the item D ahead is read now (an L1 hit: its line was prefetched one
            // item earlier) and its message chunks are requested into L2 -- one 128-byte line per lane --, so that its loads
            // find them a short latency away: bytes in flight without registers. The index lines are prefetched D + 1 items
            // ahead. Measured (B200, profiles/r02_ab_vn_loop.md 5): float64 VN 0.66 -> 0.75 of the HBM peak on the n = 102400
            // code (3 CTAs per SM leave few loads in flight), float32 0.873 -> 0.885; D = 2, 3 are 1-2 % behind D = 1.
            const int ip = idx + (L2PF - 1) * stride;   // idx is the next item already
            if (items >= L2PF && ip < count) {
                int en[DVMAX];
                vn_ell_record<T, V, DVMAX>(a.vn_ell_edge + ell_base + (size_t)ip * DVMAX, en);
                constexpr int LPC = FT * (int)sizeof(T) / 128;   // lines per chunk
                const int kk = lane / LPC;
                int ek = -1;
#pragma unroll
                for (int k = 0; k < DVMAX; ++k) ek = (kk == k) ? en[k] : ek;
                if (ek >= 0)
                    prefetch_l2(reinterpret_cast<const char *>(a.msg + (size_t)tile * a.e_stride + (size_t)ek * FT) + (lane % LPC) * 128);
                const int iq = ip + stride;
                if (items > L2PF && iq < count) {
                    prefetch_l1(a.col_order + first + iq);
                    prefetch_l1(a.vn_ell_edge + ell_base + (size_t)iq * DVMAX);
                    if constexpr (DVMAX == 8) prefetch_l1(a.vn_ell_edge + ell_base + (size_t)iq * DVMAX + 4);
                }
            }
        }
        // from here on: vn_body_ell's arithmetic, operation by operation
        Vec<T, V> llr;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const T pay = ((bm.v[v] >> lane) & 1u) ? -lp.v[v] : lp.v[v];
            llr.v[v] = (cls == 0) ? pay : ((cls == 1) ? (T)1e-4 : Lim<T>::max());
        }
        Vec<T, V> L = llr;
        if (lane_act) {
#pragma unroll
            for (int k = 0; k < DVMAX; ++k)
                if (e[k] >= 0) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if constexpr (HASNEW) c[k].v[v] = isnew[v] ? (T)0 : c[k].v[v];
                        L.v[v] = L.v[v] + c[k].v[v];
                    }
                }
#pragma unroll
            for (int k = 0; k < DVMAX; ++k)
                if (e[k] >= 0) st_msg<T, V>(tbase + (size_t)e[k] * FT, vn_out<T, V, FAST, HASNEW>(a, L, c[k], llr, isnew));
        }
        uint32_t zw = 0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t w = __ballot_sync(0xffffffffu, lane_act && (L.v[v] <= (T)0));
            if (lane == v) zw = w;
        }
        zw &= actl;
        if (lane < V) {
            a.zmask[((size_t)tile * a.n + bit) * V + lane] = zw;
            if (zw != 0) {
                int r[DVMAX];
                vn_ell_record<T, V, DVMAX>(rows, r);
#pragma unroll
                for (int k = 0; k < DVMAX; ++k)
                    if (r[k] >= 0) atomicXor(a.par + ((size_t)tile * a.m + r[k]) * V + lane, zw);
            }
        }
        if (!more) break;
    }
}
template <typename T, int V, int DVMAX, bool FAST, int CTAS, int L2PF = 1>
__global__ void __launch_bounds__(vn_threads(sizeof(T), V, DVMAX), CTAS)
vn_kernel_ell_loop(const StepArgs<T> a, const int first, const int count, const int ell_base, const int items) {
    static_assert(DVMAX == 4 || DVMAX == 8, "ELL records exist for the two narrow buckets");
    const int tile = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int idx = blockIdx.x * wpc * items + (threadIdx.x >> 5);
    if (idx >= count) return;
    uint32_t act[V], newm[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        act[v] = a.tile_active[tile * V + v];
        newm[v] = a.tile_new[tile * V + v];
    }
    bool any_act = false, any_new = false, lane_act = false;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        any_act |= act[v] != 0;
        any_new |= newm[v] != 0;
        lane_act |= (act[v] >> lane) & 1u;
    }
    if (!any_act) return;
    uint32_t actl, newl;
    bool stale;
    vn_lane_words<V>(act, newm, lane, actl, newl, stale);
    if (any_new) vn_items_ell<T, V, DVMAX, FAST, true, L2PF>(a, tile, lane, lane_act, actl, newl, stale, first, count, ell_base, idx, wpc, items);
    else vn_items_ell<T, V, DVMAX, FAST, false, L2PF>(a, tile, lane, lane_act, actl, newl, stale, first, count, ell_base, idx, wpc, items);
}


// ---------------------------------------------------------------------------------------------------------------
// K3 for the wide buckets (8 < dv <= 32), walking: vn_kernel chases col_order -> col_ptr -> csc_edge[k] -> message for every
// edge (one uniform index load per edge, again before the stores and before the parity RED). Here the edge and check ids of an
// item are ONE coalesced load each (lane k holds the k-th id; the edges read them by shuffle), and while an item's messages
// are in flight the warp fetches the NEXT item's ids the same way and asks for its message chunks in L2 (one chunk per lane),
// as vn_kernel_ell_loop does for the narrow buckets. The arithmetic is vn_body's, operation by operation.
template <typename T, int V, int DVMAX, bool FAST, bool HASNEW>
__device__ __forceinline__ void vn_items_wide(const StepArgs<T> &a, int tile, int lane, bool lane_act, const uint32_t actl,
                                              const uint32_t newl, const bool stale, const int first, const int count, int idx,
                                              const int stride, int items) {
    constexpr int FT = kWarp * V;
    constexpr int LPC = FT * (int)sizeof(T) / 128;   // 128-byte lines per message chunk
    T *tbase = a.msg + (size_t)tile * a.e_stride + lane * V;
    const char *pbase = reinterpret_cast<const char *>(a.msg + (size_t)tile * a.e_stride);
    const Vec<T, V> *lpp = reinterpret_cast<const Vec<T, V> *>(a.slot_llr + (size_t)tile * FT + lane * V);
    bool isnew[V];
#pragma unroll
    for (int v = 0; v < V; ++v) isnew[v] = HASNEW && ((newl >> v) & 1u);
    const bool ld = lane_act && (!HASNEW || stale);
    int bit = __ldg(a.col_order + first + idx);
    int c0 = __ldg(a.col_ptr + bit);
    int dv = __ldg(a.col_ptr + bit + 1) - c0;
    int e_l = lane < dv ? __ldg(a.csc_edge + c0 + lane) : -1;
    int r_l = lane < dv ? __ldg(a.csc_row + c0 + lane) : -1;
    for (;;) {
        const uint8_t cls = __ldg(a.bitclass + bit);
        const Vec<uint32_t, V> bm = *reinterpret_cast<const Vec<uint32_t, V> *>(a.bobmask + ((size_t)tile * a.n + bit) * V);
        const Vec<T, V> lp = *lpp;
        Vec<T, V> c[DVMAX];
#pragma unroll
        for (int k = 0; k < DVMAX; ++k) {
            if constexpr (HASNEW) c[k] = Vec<T, V>{};
            if (k < dv) {   // warp-uniform
                const int ek = __shfl_sync(0xffffffffu, e_l, k);
                if (ld) c[k] = ld_msg<T, V>(tbase + (size_t)ek * FT);
            }
        }
        idx += stride;
        const bool more = --items > 0 && idx < count;   // warp-uniform
        int bit_n = 0, dv_n = 0, e_n = -1, r_n = -1;
        if (more) {
            bit_n = __ldg(a.col_order + first + idx);
            const int c0n = __ldg(a.col_ptr + bit_n);
            dv_n = __ldg(a.col_ptr + bit_n + 1) - c0n;
            if (lane < dv_n) {
                e_n = __ldg(a.csc_edge + c0n + lane);
                r_n = __ldg(a.csc_row + c0n + lane);
#pragma unroll
                for (int j = 0; j < LPC; ++j) prefetch_l2(pbase + (size_t)e_n * (FT * sizeof(T)) + j * 128);
            }
        }
        Vec<T, V> llr;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const T pay = ((bm.v[v] >> lane) & 1u) ? -lp.v[v] : lp.v[v];
            llr.v[v] = (cls == 0) ? pay : ((cls == 1) ? (T)1e-4 : Lim<T>::max());
        }
        Vec<T, V> L = llr;
#pragma unroll
        for (int k = 0; k < DVMAX; ++k)
            if (k < dv && lane_act) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    if constexpr (HASNEW) c[k].v[v] = isnew[v] ? (T)0 : c[k].v[v];
                    L.v[v] = L.v[v] + c[k].v[v];
                }
            }
#pragma unroll
        for (int k = 0; k < DVMAX; ++k)
            if (k < dv) {
                const int ek = __shfl_sync(0xffffffffu, e_l, k);
                if (lane_act) st_msg<T, V>(tbase + (size_t)ek * FT, vn_out<T, V, FAST, HASNEW>(a, L, c[k], llr, isnew));
            }
        uint32_t zw = 0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t w = __ballot_sync(0xffffffffu, lane_act && (L.v[v] <= (T)0));
            if (lane == v) zw = w;
        }
        zw &= actl;
        if (lane < V) a.zmask[((size_t)tile * a.n + bit) * V + lane] = zw;
        for (int k = 0; k < dv; ++k) {
            const int rk = __shfl_sync(0xffffffffu, r_l, k);
            if (lane < V && zw != 0) atomicXor(a.par + ((size_t)tile * a.m + rk) * V + lane, zw);
        }
        if (!more) break;
        bit = bit_n; dv = dv_n; e_l = e_n; r_l = r_n;
    }
}
template <typename T, int V, int DVMAX, bool FAST>
__global__ void __launch_bounds__(vn_threads(sizeof(T), V, DVMAX))
vn_kernel_wide_loop(const StepArgs<T> a, const int first, const int count, const int items) {
    static_assert(DVMAX == 16 || DVMAX == 32, "wide buckets whose edges fit a register array and a warp's lanes");
    const int tile = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int idx = blockIdx.x * wpc * items + (threadIdx.x >> 5);
    if (idx >= count) return;
    uint32_t act[V], newm[V];
    bool any_act = false, any_new = false, lane_act = false;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        act[v] = a.tile_active[tile * V + v];
        newm[v] = a.tile_new[tile * V + v];
        any_act |= act[v] != 0;
        any_new |= newm[v] != 0;
        lane_act |= (act[v] >> lane) & 1u;
    }
    if (!any_act) return;
    uint32_t actl, newl;
    bool stale;
    vn_lane_words<V>(act, newm, lane, actl, newl, stale);
    if (any_new) vn_items_wide<T, V, DVMAX, FAST, true>(a, tile, lane, lane_act, actl, newl, stale, first, count, idx, wpc, items);
    else vn_items_wide<T, V, DVMAX, FAST, false>(a, tile, lane, lane_act, actl, newl, stale, first, count, idx, wpc, items);
}

}  // namespace qk

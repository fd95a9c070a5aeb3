// On-chip min-sum decoder with FLOAT64 state: the kernel of onchip_minsum.cuh with `double` totals and `double` row
// records, so that NMSA / OMSA / ANMSA / AOMSA reproduce the reference's double arithmetic bit for bit at on-chip speed
// (the float32 kernel agrees with the reference on 95.7 .. 99.2 % of the iteration counts for the offset / adaptive
// variants: exact-tie sums are resolved by rounding noise -- SURVEY.md 7, hard part 1). Same phases, same tables, same
// iteration accounting; only the state is wider:
//     L[n]   double  total LLR after the last variable-node phase (llr before the first iteration)
//     rec[m] 24 B    {bits(c1), bits(c2), sign bit per edge of the row, position of the first minimum}
// = 8n + 24m bytes (131 KB for n = 10240, m = 2048; 205 KB for the R = 0.5 code of ADAPTIVE R.json): one CTA of up to
// 1024 threads per SM. Every double operation is the reference's, on the same operands in the same order
// (qkd_ldpc_algorithm.cpp:374-461, 542-577, 719-768, 909-958), under the same precondition as the float32 kernel (no
// message can become NaN / inf; onchip_usable, inst_onchip.cu), so the results are bit-identical to the reference and to
// the streaming float64 kernels (tests/test_gpu_onchip.py).
#pragma once
#include "onchip_minsum.cuh"

namespace qk {

constexpr int kRec64Bytes = 24;

// Shared-memory layout: L[n+1] double (padded to 16 B) | rec[rec_slots+2] 24 B | bob[words] | alice[words] | syn[groups_cn] | misc
__host__ __device__ inline size_t onchip64_l_slots(int n) { return ((size_t)n + 1 + 1) / 2 * 2; }
__host__ __device__ inline size_t onchip64_smem_bytes(int n, int rec_slots, int groups_cn) {
    const size_t words = (size_t)(n + 31) / 32;
    return onchip64_l_slots(n) * 8 + ((size_t)rec_slots + 2) * kRec64Bytes + (2 * words + (size_t)groups_cn) * 4 + 128;
}

// Per-frame parameters of the float64 kernel (shared memory, written by thread 0 when the CTA takes a frame).
struct FrameCtx64 {
    double lp, primary, secondary;
    int has_cls, pad;
    const uint32_t *cls_punct, *cls_short;
    u64 *tally;
};

__device__ __forceinline__ uint32_t hi32(double x) { return (uint32_t)__double2hiint(x); }
// sign bit of (bits(x) - 1): x <= 0 for every x that is neither NaN nor -0
__device__ __forceinline__ uint32_t le0_word(double x) { return (uint32_t)(((u64)__double_as_longlong(x) - 1ull) >> 32); }

// One edge of a check node (see QK_CN_EDGE of the float32 kernel): the old record's magnitudes are c1o / c2o (raw
// bits, non-negative), the old signs shift through bit 31 of zs.
#define QK_CN_EDGE64(J, COL)                                                                                            \
    {                                                                                                                   \
        const double Lv = L[(COL)];                                                                                     \
        const u64 mag = (rel == (J)) ? c2o : c1o;                                                                       \
        /* clamp(L - c2b) (:447-461); first iteration: zero record and thr_b = +inf leave the unclamped LLR (:336-350) */ \
        const double braw = Lv - __hiloint2double((int)((uint32_t)(mag >> 32) ^ (zs & 0x80000000u)), (int)(uint32_t)mag); \
        zs <<= 1;                                                                                                       \
        zacc ^= le0_word(Lv);                             /* parity of the hard decision L <= 0 (:414-422) */            \
        pacc ^= hi32(braw);                               /* parity of m < 0 (:383); the clamp keeps the sign */         \
        own = __funnelshift_l(le0_word(braw), own, 1);    /* (m > 0) ? +1 : -1 (:402): zero is negative (Q4) */          \
        const double ab = fmin(fabs(braw), thr_b);        /* |clamp(x)| == min(|x|, thr) */                              \
        arg = (ab < m1) ? (kb + (J)) : arg;               /* first minimum */                                            \
        m2 = fmin(m2, fmax(ab, m1));                      /* == the if / else-if chain (:386-396) */                     \
        m1 = fmin(m1, ab);                                                                                              \
    }

template <int ALG, bool WIDE>
__device__ __forceinline__ bool onchip64_cn_phase(const OnchipArgs &a, const FrameCtx64 *ctx, const double *L, unsigned char *recb,
                                                  const uint32_t *synw, double thr_b, int warp, int lane, int nwarps) {
    bool unsat = false;
    for (int g = warp; g < a.n_groups_cn; g += nwarps) {
        const int2 gi = __ldg(a.cn_ginfo + g);
        const int dc_row = gi.y;                                  // degree of the group's rows (warp-uniform)
        const bool two = WIDE && dc_row > 32;                     // two records per row
        const int dc = two ? 32 : dc_row;                         // edges covered by the first record
        const uint32_t row = __ldg(a.cn_row + g * 32 + lane);     // first record slot of the lane's row
        u64 *rp = reinterpret_cast<u64 *>(recb + (size_t)row * kRec64Bytes);
        u64 c1o = rp[0], c2o = rp[1];
        const uint2 zw = *reinterpret_cast<const uint2 *>(rp + 2);
        const uint2 *cp = a.cnT + gi.x + lane;
        double m1 = DBL_MAX, m2 = DBL_MAX;                        // (:378-379)
        uint32_t zs = zw.x << (32 - dc);          // sign of the old message on the current edge in bit 31
        const int arg_old = (int)zw.y - (32 - dc);
        uint32_t own = 0, pacc = 0, zacc = 0;
        int arg = 0, kb = 0;
#pragma unroll 2
        for (; kb + 4 <= dc; kb += 4) {
            const uint2 cw = __ldg(cp + (kb >> 2) * 32);
            const int rel = arg_old - kb;
            QK_CN_EDGE64(0, cw.x & 0xFFFFu)
            QK_CN_EDGE64(1, cw.x >> 16)
            QK_CN_EDGE64(2, cw.y & 0xFFFFu)
            QK_CN_EDGE64(3, cw.y >> 16)
        }
        if (kb < dc) {                            // warp-uniform tail of 1..3 edges
            const uint2 cw = __ldg(cp + (kb >> 2) * 32);
            const int rel = arg_old - kb, left = dc - kb;
            QK_CN_EDGE64(0, cw.x & 0xFFFFu)
            if (left > 1) QK_CN_EDGE64(1, cw.x >> 16)
            if (left > 2) QK_CN_EDGE64(2, cw.y & 0xFFFFu)
        }
        uint32_t own_first = own;
        if constexpr (WIDE) {
            if (two) {                            // edges 32..dc_row-1: the row's second record (warp-uniform branch)
                const int dc2 = dc_row - 32;
                c1o = rp[3];
                c2o = rp[4];
                const uint2 zw2 = *reinterpret_cast<const uint2 *>(rp + 5);
                zs = zw2.x << (32 - dc2);
                const int arg_old2 = (int)zw2.y - (32 - dc2);   // kNoArg gives a value no edge index reaches
                own = 0;
                for (; kb + 4 <= dc_row; kb += 4) {
                    const uint2 cw = __ldg(cp + (kb >> 2) * 32);
                    const int rel = arg_old2 - (kb - 32);
                    QK_CN_EDGE64(0, cw.x & 0xFFFFu)
                    QK_CN_EDGE64(1, cw.x >> 16)
                    QK_CN_EDGE64(2, cw.y & 0xFFFFu)
                    QK_CN_EDGE64(3, cw.y >> 16)
                }
                if (kb < dc_row) {
                    const uint2 cw = __ldg(cp + (kb >> 2) * 32);
                    const int rel = arg_old2 - (kb - 32), left = dc_row - kb;
                    QK_CN_EDGE64(0, cw.x & 0xFFFFu)
                    if (left > 1) QK_CN_EDGE64(1, cw.x >> 16)
                    if (left > 2) QK_CN_EDGE64(2, cw.y & 0xFFFFu)
                }
            }
        }
        const uint32_t syn = (synw[g] >> lane) & 1u;
        const bool viol = (((zacc >> 31) ^ syn) & 1u) != 0;    // check not satisfied by the current hard decision
        unsat |= viol && row < (uint32_t)a.rec_slots;
        const double factor = (ALG >= 4 && viol) ? ctx->secondary : ctx->primary;   // (:749-757, :939-947)
        double c1, c2;
        if constexpr (ALG == 2 || ALG == 4) {
            c1 = factor * m1;
            c2 = factor * m2;
        } else {
            const double d1 = m1 - factor, d2 = m2 - factor;   // max(min - beta, 0) (:573-574)
            c1 = (d1 < 0.) ? 0. : d1;
            c2 = (d2 < 0.) ? 0. : d2;
        }
        c1 = fmin(c1, a.thr64);                                // threshold_matrix(check_to_bit), magnitudes (:411-412)
        c2 = fmin(c2, a.thr64);
        const uint32_t rowneg = ((pacc >> 31) ^ syn) & 1u;     // (syndrome ? -1 : 1) * (-1)^negatives (:398-399)
        rp[0] = (u64)__double_as_longlong(c1);
        rp[1] = (u64)__double_as_longlong(c2);
        *reinterpret_cast<uint2 *>(rp + 2) =
            make_uint2(own_first ^ (0u - rowneg), (!two || arg < 32) ? (uint32_t)(arg + 32 - dc) : kNoArg);
        if constexpr (WIDE) {
            if (two) {
                rp[3] = (u64)__double_as_longlong(c1);
                rp[4] = (u64)__double_as_longlong(c2);
                *reinterpret_cast<uint2 *>(rp + 5) =
                    make_uint2(own ^ (0u - rowneg), (arg >= 32) ? (uint32_t)(arg - 32 + 32 - (dc_row - 32)) : kNoArg);
            }
        }
    }
    return unsat;
}
#undef QK_CN_EDGE64

__device__ __forceinline__ double onchip64_llr(const FrameCtx64 *ctx, const uint32_t *bobw, uint32_t bit, double lp) {
    const uint32_t w = bit >> 5, s = bit & 31u;
    double v = ((bobw[w] >> s) & 1u) ? 0. - lp : lp;           // qkd_ldpc_algorithm.cpp:1043-1049 (never -0)
    if (ctx->has_cls) {
        if ((__ldg(ctx->cls_punct + w) >> s) & 1u) v = 1e-4;   // punctured: ALMOST_ZERO (:1155)
        else if ((__ldg(ctx->cls_short + w) >> s) & 1u) v = DBL_MAX;   // shortened (:1164)
    }
    return v;
}

// The check-to-bit message addressed by table entry `ent` = row << 9 | sh, added to the running sum: the sign / argmin
// word first, then the one magnitude that applies.
#define QK_VN_EDGE64(ENT)                                                                                               \
    {                                                                                                                   \
        const unsigned char *rq = recb + ((ENT) >> 9) * kRec64Bytes;                                                    \
        const uint2 zw = *reinterpret_cast<const uint2 *>(rq + 16);                                                     \
        const uint2 mg = *reinterpret_cast<const uint2 *>(rq + ((((ENT) ^ zw.y) & 0x1FFu) ? 0 : 8));   /* kNoArg never matches */ \
        acc = acc + __hiloint2double((int)(mg.y ^ (__funnelshift_l(0u, zw.x, (ENT)) & 0x80000000u)), (int)mg.x);        \
    }

__device__ __forceinline__ void onchip64_vn_phase(const OnchipArgs &a, const FrameCtx64 *ctx, double *L, const unsigned char *recb,
                                                  const uint32_t *bobw, double lp, int warp, int lane, int nwarps) {
    for (int g = warp; g < a.n_groups_vn; g += nwarps) {
        const int2 gi = __ldg(a.vn_ginfo + g);
        const int dv = gi.y;
        const uint32_t bit = __ldg(a.vn_bit + g * 32 + lane);
        double acc = onchip64_llr(ctx, bobw, bit < (uint32_t)a.n ? bit : 0u, lp);
        const uint4 *ep = a.vT + gi.x + lane;
        int kb = 0;
        // ascending check order, starting from the LLR (std::accumulate, :414-417)
#pragma unroll 2
        for (; kb + 4 <= dv; kb += 4) {
            const uint4 ew = __ldg(ep + (kb >> 2) * 32);
            QK_VN_EDGE64(ew.x)
            QK_VN_EDGE64(ew.y)
            QK_VN_EDGE64(ew.z)
            QK_VN_EDGE64(ew.w)
        }
        if (kb < dv) {                            // warp-uniform tail of 1..3 checks
            const uint4 ew = __ldg(ep + (kb >> 2) * 32);
            const int left = dv - kb;
            QK_VN_EDGE64(ew.x)
            if (left > 1) QK_VN_EDGE64(ew.y)
            if (left > 2) QK_VN_EDGE64(ew.z)
        }
        L[bit] = acc;                             // padding lanes write the scratch slot L[n]
    }
}
#undef QK_VN_EDGE64

template <int ALG, bool WIDE>
__global__ void __launch_bounds__(1024, 1) onchip_minsum64_kernel(const OnchipArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *L = reinterpret_cast<double *>(smem_raw);
    unsigned char *recb = reinterpret_cast<unsigned char *>(L + onchip64_l_slots(a.n));
    uint32_t *bobw = reinterpret_cast<uint32_t *>(recb + ((size_t)a.rec_slots + 2) * kRec64Bytes);
    uint32_t *alw = bobw + a.words;
    uint32_t *synw = alw + a.words;
    uint32_t *tail = synw + a.n_groups_cn;
    long long *s_frame = reinterpret_cast<long long *>(tail + ((2 * a.words + a.n_groups_cn) & 1));
    FrameCtx64 *ctx = reinterpret_cast<FrameCtx64 *>(s_frame + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    constexpr bool kAdaptive = (ALG >= 4);
    const double inf = __longlong_as_double(0x7ff0000000000000ll);

    for (;;) {
        __syncthreads();   // previous frame fully written out before the state is reused
        if (tid == 0) {
            const long long f = (long long)atomicAdd(a.next_frame, 1ull);
            *s_frame = f;
            if (f < a.n_frames) {
                const long long combo = f / a.frames_per_combo;
                const OnchipCombo cb = a.combos[combo];
                const double q = cb.qber >= 0. ? cb.qber : (a.qber_is_scalar ? a.qber[0] : a.qber[f]);
                ctx->lp = log((1. - q) / q);
                ctx->primary = cb.primary;
                ctx->secondary = cb.secondary;
                ctx->has_cls = cb.has_cls;
                ctx->cls_punct = a.cls_masks + combo * 2 * a.words;
                ctx->cls_short = ctx->cls_punct + a.words;
                ctx->tally = a.tally ? a.tally + combo * a.tally_len : nullptr;
            }
        }
        __syncthreads();
        const long long f = *s_frame;
        if (f >= a.n_frames) break;
        const double lp = ctx->lp;
        for (int w = tid; w < a.words; w += blockDim.x) {
            bobw[w] = a.bob_bits[f * a.words + w];
            alw[w] = a.alice_bits[f * a.words + w];
        }
        __syncthreads();
        // L = a-priori LLR; Alice's syndrome (calculate_syndrome, array_and_matrix_operations.cpp:936-950); records = 0
        for (int i = tid; i <= a.n; i += blockDim.x) L[i] = (i < a.n) ? onchip64_llr(ctx, bobw, (uint32_t)i, lp) : 1.;
        for (int g = warp; g < a.n_groups_cn; g += nwarps) {
            const int2 gi = __ldg(a.cn_ginfo + g);
            const uint32_t row = __ldg(a.cn_row + g * 32 + lane);
            const uint2 *cp = a.cnT + gi.x + lane;
            uint32_t s = 0;
            for (int kb = 0; kb < gi.y; kb += 4) {
                const uint2 cw = __ldg(cp + (kb >> 2) * 32);
                const int left = gi.y - kb;
                const uint32_t c0 = cw.x & 0xFFFFu, c1 = cw.x >> 16, c2 = cw.y & 0xFFFFu, c3 = cw.y >> 16;
                s ^= alw[c0 >> 5] >> (c0 & 31u);
                if (left > 1) s ^= alw[c1 >> 5] >> (c1 & 31u);
                if (left > 2) s ^= alw[c2 >> 5] >> (c2 & 31u);
                if (left > 3) s ^= alw[c3 >> 5] >> (c3 & 31u);
            }
            const uint32_t sw = __ballot_sync(0xffffffffu, (s & 1u) != 0 && row < (uint32_t)a.rec_slots);
            if (lane == 0) synw[g] = sw;
            u64 *rp = reinterpret_cast<u64 *>(recb + (size_t)row * kRec64Bytes);
            rp[0] = rp[1] = rp[2] = 0ull;
            if (WIDE && gi.y > 32) rp[3] = rp[4] = rp[5] = 0ull;
        }
        __syncthreads();

        int iters = a.max_iter, run = a.max_iter;
        bool success = false;
        for (int it = 1;; ++it) {
            // the check-node pass of iteration `it`; at it = max_iter + 1 it only serves as the syndrome test of the
            // last hard decision (non-adaptive variants, :424-445)
            const bool unsat = onchip64_cn_phase<ALG, WIDE>(a, ctx, L, recb, synw, it == 1 ? inf : a.thr64, warp, lane, nwarps);
            const bool any_unsat = __syncthreads_or(unsat) != 0;
            if (!kAdaptive) {
                if (it > 1 && !any_unsat) { success = true; iters = run = it - 1; break; }   // z of iteration it-1 (:439-445)
                if (it > a.max_iter) break;
            } else {
                if (!any_unsat) { success = true; iters = it; run = it - 1; break; }         // exit test before the VN step (:770-776)
            }
            onchip64_vn_phase(a, ctx, L, recb, bobw, lp, warp, lane, nwarps);
            __syncthreads();
            if (kAdaptive && it == a.max_iter) break;          // the decision of the last iteration is never tested (Q10)
        }

        // bob_solution = last hard decision (L <= 0), packed; keys compare (arrays_equal, :1087)
        uint32_t diff = 0;
        for (int w = warp; w < a.words; w += nwarps) {
            const int i = w * 32 + lane;
            const uint32_t word = __ballot_sync(0xffffffffu, i < a.n && L[i < a.n ? i : 0] <= 0.);
            if (lane == 0) {
                if (a.out_bits) a.out_bits[f * a.words + w] = word;
                diff |= word ^ alw[w];
            }
        }
        const bool keys_differ = __syncthreads_or(diff != 0) != 0;
        if (tid == 0) {
            if (a.out_iters) a.out_iters[f] = iters;
            if (a.out_flags) a.out_flags[f] = (uint8_t)((success ? 1u : 0u) | (keys_differ ? 0u : 2u));
            u64 *tally = ctx->tally;
            if (tally) {
                atomicAdd(tally + 0, 1ull);
                if (success) {
                    atomicAdd(tally + 1, 1ull);
                    if (!keys_differ) atomicAdd(tally + 2, 1ull);
                    atomicAdd(tally + 4 + iters, 1ull);
                }
                atomicAdd(tally + 3, (u64)run);
            }
        }
    }
}

}  // namespace qk

// On-chip sum-product decoder (SPA and SPA-lin-approx, float32 messages): one frame per CTA, every message of the
// frame lives in shared memory for all iterations; HBM is touched only for the packed key bits going in and the packed
// decision coming out.
//
// Unlike the min-sum family (onchip_minsum.cuh) a sum-product check node sends dc different magnitudes, so the state
// cannot be compressed into a row record: the kernel keeps ONE float per edge, the check-to-bit message c2b, plus the
// bit totals L[n]. The bit-to-check message is not stored: b2c = clamp(L - c2b) (qkd_ldpc_algorithm.cpp:109-123) is
// rebuilt from the bit's total when the check node needs it, with the operands and the order of the reference, so the
// results are bit-identical to the streaming float32 kernels (step_kernels.cuh), whose arithmetic (RowState) is used
// as is. State per frame: 4 E' + 4 n bytes (E' = edges padded to whole 32-row groups), 205 KB for n = 10240, E = 40960:
// one CTA of up to 1024 threads per SM. Codes whose state does not fit (E' > 65535 message words, or more than the
// 227 KB of shared memory) take the streaming path.
//
// Phases per iteration (two __syncthreads):
//   CN  thread per row, two passes over the row's message words (layout [group][k][lane]: conflict-free):
//       pass 1  b2c = clamp(L[bit] - c2b_old), t = tanh(b2c / 2) stored in place, P *= t (:55-63); the parity of the
//               hard decision z = (L <= 0) falls out of the same gather = the syndrome test of the previous iteration
//               (:86,101-107)
//       pass 2  c2b = clamp(2 atanh(P / t)) stored in place (:64-74)
//   VN  thread per bit: L = llr + sum of the bit's c2b in ascending check order (:76-84)
// The check-phase tables (groups of 32 rows of one degree, bit indices in blocks of 4) are the ones of the min-sum
// kernel; the variable phase has its own groups, packed so that the 32 lanes' k-th messages lie in different banks.
#pragma once
#include "onchip_minsum.cuh"
#include "step_kernels.cuh"

namespace qk {

// Shared-memory layout (onchip_spa_smem_bytes, onchip_minsum.cuh):
//   msg[msg_words] float | L[n+1] float (padded to 16 B) | bob[words] | alice[words] | syn[groups_cn] | misc

#define QK_SPA_CN_EDGE(J, COL)                                                                                          \
    {                                                                                                                   \
        const float Lv = L[(COL)];                                                                                      \
        zpar ^= (Lv <= 0.f) ? 1u : 0u;                    /* hard decision (:80-83); NaN decides 0 (quirk Q2) */         \
        float *w = mp + (kb + (J)) * 32;                                                                                \
        /* clamp(L - c2b) (:109-123); first iteration: zero message and thr_b = +inf leave the unclamped LLR (:21-29) */ \
        *w = st.absorb(clamp_msg(Lv - *w, thr_b));        /* tanh(m / 2), in place (:58-62) */                           \
    }
#define QK_SPA_CN_EMIT(J)                                                                                               \
    {                                                                                                                   \
        float *w = mp + (kb + (J)) * 32;                                                                                \
        *w = clamp_msg(st.emit(*w, syn != 0, 0.f), a.thr);   /* 2 atanh(P / t), threshold_matrix (:64-74) */             \
    }

template <int ALG>
__device__ __forceinline__ bool onchip_spa_cn_phase(const OnchipArgs &a, const float *L, float *msg, const uint32_t *synw,
                                                    float thr_b, int warp, int lane, int nwarps) {
    bool unsat = false;
    for (int g = warp; g < a.n_groups_cn; g += nwarps) {
        const int2 gi = __ldg(a.cn_ginfo + g);
        const int dc = gi.y;                                      // degree of the group's rows (warp-uniform)
        const uint32_t row = __ldg(a.cn_row + g * 32 + lane);     // record slot of the min-sum kernel: only its validity is used
        const uint2 *cp = a.cnT + gi.x + lane;
        float *mp = msg + __ldg(a.cn_moff + g) + lane;
        const uint32_t syn = (synw[g] >> lane) & 1u;
        RowState<float, ALG> st;
        st.init(syn != 0);                                        // P = syndrome ? -1 : 1 (:56-57)
        uint32_t zpar = 0;
        int kb = 0;
#pragma unroll 1
        for (; kb + 4 <= dc; kb += 4) {
            const uint2 cw = __ldg(cp + (kb >> 2) * 32);
            QK_SPA_CN_EDGE(0, cw.x & 0xFFFFu)
            QK_SPA_CN_EDGE(1, cw.x >> 16)
            QK_SPA_CN_EDGE(2, cw.y & 0xFFFFu)
            QK_SPA_CN_EDGE(3, cw.y >> 16)
        }
        if (kb < dc) {                            // warp-uniform tail of 1..3 edges
            const uint2 cw = __ldg(cp + (kb >> 2) * 32);
            const int left = dc - kb;
            QK_SPA_CN_EDGE(0, cw.x & 0xFFFFu)
            if (left > 1) QK_SPA_CN_EDGE(1, cw.x >> 16)
            if (left > 2) QK_SPA_CN_EDGE(2, cw.y & 0xFFFFu)
        }
        unsat |= ((zpar ^ syn) & 1u) != 0 && row < (uint32_t)a.rec_slots;
#pragma unroll 1
        for (kb = 0; kb + 4 <= dc; kb += 4) {
            QK_SPA_CN_EMIT(0)
            QK_SPA_CN_EMIT(1)
            QK_SPA_CN_EMIT(2)
            QK_SPA_CN_EMIT(3)
        }
        if (kb < dc) {
            const int left = dc - kb;
            QK_SPA_CN_EMIT(0)
            if (left > 1) QK_SPA_CN_EMIT(1)
            if (left > 2) QK_SPA_CN_EMIT(2)
        }
    }
    return unsat;
}
#undef QK_SPA_CN_EDGE
#undef QK_SPA_CN_EMIT

__device__ __forceinline__ void onchip_spa_vn_phase(const OnchipArgs &a, const FrameCtx *ctx, float *L, const float *msg,
                                                    const uint32_t *bobw, float lp, int warp, int lane, int nwarps) {
    for (int g = warp; g < a.n_groups_sv; g += nwarps) {
        const int2 gi = __ldg(a.sv_ginfo + g);
        const int dv = gi.y;
        const uint32_t bit = __ldg(a.sv_bit + g * 32 + lane);
        float acc = onchip_llr(ctx, bobw, bit < (uint32_t)a.n ? bit : 0u, lp);
        const uint2 *ep = a.svT + gi.x + lane;
        int kb = 0;
        // ascending check order, starting from the LLR (std::accumulate, :78)
#pragma unroll 2
        for (; kb + 4 <= dv; kb += 4) {
            const uint2 ew = __ldg(ep + (kb >> 2) * 32);
            acc = acc + msg[ew.x & 0xFFFFu];
            acc = acc + msg[ew.x >> 16];
            acc = acc + msg[ew.y & 0xFFFFu];
            acc = acc + msg[ew.y >> 16];
        }
        if (kb < dv) {                            // warp-uniform tail of 1..3 checks
            const uint2 ew = __ldg(ep + (kb >> 2) * 32);
            const int left = dv - kb;
            acc = acc + msg[ew.x & 0xFFFFu];
            if (left > 1) acc = acc + msg[ew.x >> 16];
            if (left > 2) acc = acc + msg[ew.y & 0xFFFFu];
        }
        L[bit] = acc;                             // padding lanes write the scratch slot L[n]
    }
}

template <int ALG>
__global__ void __launch_bounds__(1024, 1) onchip_spa_kernel(const OnchipArgs a) {
    static_assert(ALG == 0 || ALG == 1, "sum-product variants only");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *msg = reinterpret_cast<float *>(smem_raw);
    float *L = msg + (a.msg_words + 3) / 4 * 4;
    uint32_t *bobw = reinterpret_cast<uint32_t *>(L + onchip_l_slots(a.n));
    uint32_t *alw = bobw + a.words;
    uint32_t *synw = alw + a.words;
    uint32_t *tail = synw + a.n_groups_cn;
    long long *s_frame = reinterpret_cast<long long *>(tail + ((2 * a.words + a.n_groups_cn) & 1));
    FrameCtx *ctx = reinterpret_cast<FrameCtx *>(s_frame + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float inf = __int_as_float(0x7f800000);

    for (;;) {
        __syncthreads();   // previous frame fully written out before the state is reused
        if (tid == 0) {
            const long long f = (long long)atomicAdd(a.next_frame, 1ull);
            *s_frame = f;
            if (f < a.n_frames) {
                const long long combo = f / a.frames_per_combo;
                const OnchipCombo cb = a.combos[combo];
                const double q = cb.qber >= 0. ? cb.qber : (a.qber_is_scalar ? a.qber[0] : a.qber[f]);
                ctx->lp = (float)log((1. - q) / q);
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
        const float lp = ctx->lp;
        for (int w = tid; w < a.words; w += blockDim.x) {
            bobw[w] = a.bob_bits[f * a.words + w];
            alw[w] = a.alice_bits[f * a.words + w];
        }
        for (int i = tid; i < a.msg_words; i += blockDim.x) msg[i] = 0.f;
        __syncthreads();
        // L = a-priori LLR; Alice's syndrome (calculate_syndrome, array_and_matrix_operations.cpp:936-950)
        for (int i = tid; i <= a.n; i += blockDim.x) L[i] = (i < a.n) ? onchip_llr(ctx, bobw, (uint32_t)i, lp) : 1.f;
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
        }
        __syncthreads();

        int iters = a.max_iter, run = a.max_iter;
        bool success = false;
        for (int it = 1;; ++it) {
            // the check-node pass of iteration `it`; at it = max_iter + 1 it only serves as the syndrome test of the
            // last hard decision (:101-107)
            const bool unsat = onchip_spa_cn_phase<ALG>(a, L, msg, synw, it == 1 ? inf : a.thr, warp, lane, nwarps);
            const bool any_unsat = __syncthreads_or(unsat) != 0;
            if (it > 1 && !any_unsat) { success = true; iters = run = it - 1; break; }   // z of iteration it-1
            if (it > a.max_iter) break;
            onchip_spa_vn_phase(a, ctx, L, msg, bobw, lp, warp, lane, nwarps);
            __syncthreads();
        }

        // bob_solution = last hard decision (L <= 0), packed; keys compare (arrays_equal, :1087)
        uint32_t diff = 0;
        for (int w = warp; w < a.words; w += nwarps) {
            const int i = w * 32 + lane;
            const uint32_t word = __ballot_sync(0xffffffffu, i < a.n && L[i < a.n ? i : 0] <= 0.f);
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

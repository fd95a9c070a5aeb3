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
// kernel; the variable phase has its own groups, packed so that the 32 lanes' k-th messages lie in different banks,
// stored as a flat list of 4-check items.
#pragma once
#include "onchip_minsum.cuh"
#include "step_kernels.cuh"

namespace qk {

// Shared-memory layout (onchip_spa_smem_bytes, onchip_minsum.cuh):
//   L[n+1] float (padded to 16 B) | msg[msg_words + 1] float (padded; the last word stays 0) | bob[words] | alice[words] |
//   syn[groups_cn] | bobg[groups_sv] | frame id, FrameCtx (64 B) | LinLut (SPA-lin-approx only).  Every word of it has an index < 65536 (227 KB / 4), which is what the
//   variable-phase table stores.

// SPA-lin-approx: the reference's piecewise-linear tanh (qkd_ldpc_algorithm.cpp:146-160) by table instead of a chain of
// 7 compares and 14 selects. |h| is bucketed by its exponent and two mantissa bits (quarter octaves from 0.5 to 8, one
// bucket below, one above that also takes NaN): all breakpoints but 0.9 and 1.2 fall on bucket borders, so a bucket holds
// at most one breakpoint `thr` and the pair of (slope, intercept) on either side of it. The selected slope and intercept
// are the reference's constants and the arithmetic is its one multiply and one add, so the result is bit-identical to
// cn_tanh_half<float, 1> (the streaming kernels keep the select chain; tests compare the two paths frame by frame).
constexpr int kLinBuckets = 18;
struct LinLut {
    float thr[kLinBuckets];        // breakpoint inside the bucket, +inf when there is none
    float2 ab[2 * kLinBuckets];    // [2 * bucket + (|h| >= thr)] = {slope, intercept}
};
__device__ __forceinline__ void spa_lin_lut_fill(LinLut *lut, int tid) {
    if (tid >= kLinBuckets) return;
    const float T[7] = {0.5f, 0.9f, 1.2f, 1.75f, 2.5f, 3.5f, 8.f};
    const float A[8] = {0.9242f, 0.6355f, 0.3912f, 0.1958f, 0.0603f, 0.0115f, 0.0004f, 0.f};
    const float B[8] = {0.f, 0.1444f, 0.3642f, 0.5986f, 0.8358f, 0.9577f, 0.9967f, 1.f};
    float lo, hi;                  // the bucket covers [lo, hi)
    if (tid == 0) { lo = 0.f; hi = 0.5f; }
    else if (tid == kLinBuckets - 1) { lo = 8.f; hi = __int_as_float(0x7f800000); }
    else {
        const int e = (tid - 1) / 4 - 1, q = (tid - 1) % 4;
        const float scale = (e < 0) ? 0.5f : (float)(1 << e);
        lo = scale * (1.f + 0.25f * (float)q);
        hi = scale * (1.f + 0.25f * (float)(q + 1));
    }
    int seg = 0;                   // segment of `lo`: number of breakpoints <= lo
    for (int k = 0; k < 7; ++k) seg += (T[k] <= lo) ? 1 : 0;
    const bool inside = seg < 7 && T[seg] > lo && T[seg] < hi;
    lut->thr[tid] = inside ? T[seg] : __int_as_float(0x7f800000);
    lut->ab[2 * tid] = make_float2(A[seg], B[seg]);
    lut->ab[2 * tid + 1] = make_float2(A[inside ? seg + 1 : seg], B[inside ? seg + 1 : seg]);
}
__device__ __forceinline__ float spa_tanh_lin_lut(const LinLut *lut, float x) {
    const float h = x / 2.f, ax = fabsf(h);
    const int idx = min(max((int)(__float_as_uint(ax) >> 21), 503), 520) - 503;   // 0.5 = 504 << 21, 8.0 = 520 << 21
    const float thr = lut->thr[idx];
    const float2 ab = lut->ab[2 * idx + ((ax >= thr) ? 1 : 0)];
    const float r = ab.x * fminf(ax, 8.f) + ab.y;      // |h| >= 8 and NaN: 0 * 8 + 1, as `if (!(ax < 8)) r = 1`
    return (h < 0.f) ? -r : r;
}

// Pass 1 of a check node for a block of edges: all gathers first (independent loads in flight together), then the
// arithmetic, then the stores -- the in-place update would otherwise serialise the edges of a block.
#define QK_SPA_LOAD(J, COL) const float Lv##J = L[(COL)]; const float c##J = mp[(kb + (J)) * 32];
/* hard decision (:80-83), NaN decides 0 (quirk Q2); clamp(L - c2b) (:109-123) -- first iteration: zero message and     */
/* thr_b = +inf leave the unclamped LLR (:21-29); tanh(m / 2) kept in place (:58-62), P *= t in edge order               */
#define QK_SPA_ABSORB(J)                                                                                                \
    zpar ^= (Lv##J <= 0.f) ? 1u : 0u;                                                                                   \
    const float t##J = spa_absorb<ALG>(st, lut, clamp_msg(Lv##J - c##J, thr_b));
#define QK_SPA_STORE(J) mp[(kb + (J)) * 32] = t##J;
/* pass 2: 2 atanh(P / t), threshold_matrix (:64-74) */
#define QK_SPA_EMIT(J) const float e##J = clamp_msg(st.emit(mp[(kb + (J)) * 32], syn != 0, 0.f), a.thr);
#define QK_SPA_ESTORE(J) mp[(kb + (J)) * 32] = e##J;

template <int ALG>
__device__ __forceinline__ float spa_absorb(RowState<float, ALG> &st, const LinLut *lut, float b2c) {
    if constexpr (ALG == 1) {
        const float t = spa_tanh_lin_lut(lut, b2c);
        st.a *= t;                                                // RowState::absorb with the tabulated tanh
        return t;
    } else {
        return st.absorb(b2c);
    }
}

template <int ALG>
__device__ __forceinline__ bool onchip_spa_cn_phase(const OnchipArgs &a, const float *__restrict__ L, float *__restrict__ msg,
                                                    const uint32_t *synw, const LinLut *lut, float thr_b, int warp, int lane,
                                                    int nwarps) {
    bool unsat = false;
    for (int g = warp; g < a.n_groups_cn; g += nwarps) {
        const int2 gi = __ldg(a.cn_ginfo + g);
        const int dc = gi.y;                                      // degree of the group's rows (warp-uniform)
        const uint32_t row = __ldg(a.cn_row + g * 32 + lane);     // record slot of the min-sum kernel: only its validity is used
        const uint2 *cp = a.cnT + gi.x + lane;
        float *__restrict__ mp = msg + __ldg(a.cn_moff + g) + lane;
        const uint32_t syn = (synw[g] >> lane) & 1u;
        RowState<float, ALG> st;
        st.init(syn != 0);                                        // P = syndrome ? -1 : 1 (:56-57)
        uint32_t zpar = 0;
        int kb = 0;
        uint2 cw = __ldg(cp);                                     // indices of the first block; the next block is
#pragma unroll 1                                                  // fetched while the current one is processed
        for (; kb + 4 <= dc; kb += 4) {
            const uint2 nx = __ldg(cp + (((kb + 4 < dc) ? kb + 4 : kb) >> 2) * 32);
            QK_SPA_LOAD(0, cw.x & 0xFFFFu) QK_SPA_LOAD(1, cw.x >> 16) QK_SPA_LOAD(2, cw.y & 0xFFFFu) QK_SPA_LOAD(3, cw.y >> 16)
            QK_SPA_ABSORB(0) QK_SPA_ABSORB(1) QK_SPA_ABSORB(2) QK_SPA_ABSORB(3)
            QK_SPA_STORE(0) QK_SPA_STORE(1) QK_SPA_STORE(2) QK_SPA_STORE(3)
            cw = nx;
        }
        if (kb < dc) {                            // warp-uniform tail of 1..3 edges
            const int left = dc - kb;
            QK_SPA_LOAD(0, cw.x & 0xFFFFu)
            QK_SPA_ABSORB(0)
            QK_SPA_STORE(0)
            if (left > 1) {
                QK_SPA_LOAD(1, cw.x >> 16)
                QK_SPA_ABSORB(1)
                QK_SPA_STORE(1)
            }
            if (left > 2) {
                QK_SPA_LOAD(2, cw.y & 0xFFFFu)
                QK_SPA_ABSORB(2)
                QK_SPA_STORE(2)
            }
        }
        unsat |= ((zpar ^ syn) & 1u) != 0 && row < (uint32_t)a.rec_slots;
#pragma unroll 1
        for (kb = 0; kb + 4 <= dc; kb += 4) {
            QK_SPA_EMIT(0) QK_SPA_EMIT(1) QK_SPA_EMIT(2) QK_SPA_EMIT(3)
            QK_SPA_ESTORE(0) QK_SPA_ESTORE(1) QK_SPA_ESTORE(2) QK_SPA_ESTORE(3)
        }
        if (kb < dc) {
            const int left = dc - kb;
            { QK_SPA_EMIT(0) QK_SPA_ESTORE(0) }
            if (left > 1) { QK_SPA_EMIT(1) QK_SPA_ESTORE(1) }
            if (left > 2) { QK_SPA_EMIT(2) QK_SPA_ESTORE(2) }
        }
    }
    return unsat;
}
#undef QK_SPA_LOAD
#undef QK_SPA_ABSORB
#undef QK_SPA_STORE
#undef QK_SPA_EMIT
#undef QK_SPA_ESTORE

// Variable phase. The work is a flat list of ITEMS, one per (group of 32 bits, block of 4 checks): a lane's 16-byte
// entry {4 x uint16 shared-memory word of the message, bit id | first << 16 | last << 17, group} is everything it needs,
// so there is no group header to chase, and the entries are fetched two items ahead. Warp w owns the contiguous items
// [sv_chunk[w], sv_chunk[w+1]) (whole groups, balanced on the host). Blocks of fewer than 4 checks are padded with the
// message word that is always 0.0f (x + 0.0f == x for every x that can occur here), so a block is 4 unconditional adds
// in ascending check order, starting from the LLR (std::accumulate, :78). Bob's bits are kept in group order
// (bobg[group], bit = lane), so the LLR sign is one broadcast load away.
__device__ __forceinline__ void onchip_spa_vn_phase(const OnchipArgs &a, const FrameCtx *ctx, float *__restrict__ smf,
                                                    const uint32_t *__restrict__ bobg, float lp, int warp, int lane) {
    int i = __ldg(a.sv_chunk + warp);
    const int end = __ldg(a.sv_chunk + warp + 1);
    if (i >= end) return;
    const int has_cls = ctx->has_cls;
    const uint32_t *cls_punct = ctx->cls_punct, *cls_short = ctx->cls_short;
    const uint4 *items = a.sv_items + lane;
    uint4 it = __ldg(items + (size_t)i * 32);
    uint4 nx = __ldg(items + (size_t)min(i + 1, end - 1) * 32);
    float acc = 0.f;
    for (; i < end; ++i) {
        const uint4 nx2 = __ldg(items + (size_t)min(i + 2, end - 1) * 32);
        const uint32_t bit = it.z & 0xFFFFu;                       // padding lanes: n (the scratch slot L[n])
        float llr = ((bobg[it.w] >> lane) & 1u) ? -lp : lp;        // qkd_ldpc_algorithm.cpp:1043-1049
        if (has_cls) {                                             // rate adaptation (warp-uniform)
            const uint32_t bi = bit < (uint32_t)a.n ? bit : 0u, w = bi >> 5, sh = bi & 31u;
            if ((__ldg(cls_punct + w) >> sh) & 1u) llr = 1e-4f;                // punctured: ALMOST_ZERO (:1155)
            else if ((__ldg(cls_short + w) >> sh) & 1u) llr = FLT_MAX;         // shortened: largest finite value (:1164)
        }
        acc = (it.z & 0x10000u) ? llr : acc;                       // first block of the bit: start from the LLR
        const float m0 = smf[it.x & 0xFFFFu], m1 = smf[it.x >> 16], m2 = smf[it.y & 0xFFFFu], m3 = smf[it.y >> 16];
        acc = acc + m0;
        acc = acc + m1;
        acc = acc + m2;
        acc = acc + m3;
        if (it.z & 0x20000u) smf[bit] = acc;                       // last block of the bit: L[bit]
        it = nx;
        nx = nx2;
    }
}

template <int ALG>
__global__ void __launch_bounds__(1024, 1) onchip_spa_kernel(const OnchipArgs a) {
    static_assert(ALG == 0 || ALG == 1, "sum-product variants only");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *L = reinterpret_cast<float *>(smem_raw);          // offset 0: the gather address is one shift-add
    float *msg = L + onchip_l_slots(a.n);
    uint32_t *bobw = reinterpret_cast<uint32_t *>(msg + (a.msg_words + 4) / 4 * 4);
    uint32_t *alw = bobw + a.words;
    uint32_t *synw = alw + a.words;
    uint32_t *bobg = synw + a.n_groups_cn;
    uint32_t *tail = bobg + a.n_groups_sv;
    long long *s_frame = reinterpret_cast<long long *>(tail + ((2 * a.words + a.n_groups_cn + a.n_groups_sv) & 1));
    FrameCtx *ctx = reinterpret_cast<FrameCtx *>(s_frame + 1);
    LinLut *lut = reinterpret_cast<LinLut *>(reinterpret_cast<unsigned char *>(s_frame) + 64);   // 8-byte aligned, inside the tail

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float inf = __int_as_float(0x7f800000);
    if constexpr (ALG == 1) spa_lin_lut_fill(lut, tid);   // visible to all warps after the first barrier below

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
                ctx->primary = (float)cb.primary;
                ctx->secondary = (float)cb.secondary;
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
        for (int i = tid; i <= a.msg_words; i += blockDim.x) msg[i] = 0.f;   // + the always-zero word msg[msg_words]
        __syncthreads();
        // L = a-priori LLR; Bob's bits in variable-group order; Alice's syndrome (calculate_syndrome,
        // array_and_matrix_operations.cpp:936-950)
        for (int i = tid; i <= a.n; i += blockDim.x) L[i] = (i < a.n) ? onchip_llr(ctx, bobw, (uint32_t)i, lp) : 1.f;
        for (int g = warp; g < a.n_groups_sv; g += nwarps) {
            const uint32_t bit = __ldg(a.sv_items + (size_t)__ldg(a.sv_group_item0 + g) * 32 + lane).z & 0xFFFFu;
            const uint32_t bw = __ballot_sync(0xffffffffu, bit < (uint32_t)a.n && ((bobw[(bit < (uint32_t)a.n ? bit : 0u) >> 5] >> (bit & 31u)) & 1u));
            if (lane == 0) bobg[g] = bw;
        }
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
            const bool unsat = onchip_spa_cn_phase<ALG>(a, L, msg, synw, lut, it == 1 ? inf : a.thr, warp, lane, nwarps);
            const bool any_unsat = __syncthreads_or(unsat) != 0;
            if (it > 1 && !any_unsat) { success = true; iters = run = it - 1; break; }   // z of iteration it-1
            if (it > a.max_iter) break;
            onchip_spa_vn_phase(a, ctx, L, bobg, lp, warp, lane);
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

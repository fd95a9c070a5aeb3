// On-chip min-sum decoder: one frame per CTA, the whole belief-propagation state of the frame lives in shared memory
// for all iterations; HBM is touched only for the packed key bits going in and the packed decision coming out.
//
// Why this is possible. For the min-sum family a check node's dc outgoing messages take only two magnitudes
// (qkd_ldpc_algorithm.cpp:400-408): factor*min1 for every edge except the one holding the minimum, which gets
// factor*min2 (ties: min2 == min1, so "first minimum" and the reference's by-value test `|m| == min1` select the
// same number). A row is therefore fully described by a 16-byte RECORD {c1, c2, sign bits, argmin}, and the
// bit-to-check message need not be stored at all: b2c = clamp(L - c2b) (:447-461) is recomputed from the bit's
// total L and the row's previous record when the check node needs it. State per frame:
//     L[n]   float   total LLR after the last variable-node phase (llr before the first iteration)
//     rec[m] uint4   {bits(c1), bits(c2), sign bit per edge of the row, position of the first minimum}
// = 4n + 16m bytes (72 KB for n=10240, m=2048 instead of 242 KB of float messages), so two frames fit one SM.
// Every float operation of the reference is performed on the same operands in the same order (the sum over a
// bit's checks runs in ascending check order starting from the LLR, :414-422), so the results are bit-identical
// to the streaming float32 kernels (step_kernels.cuh) and to the f32 oracle.
//
// Phases per iteration (two __syncthreads):
//   CN  thread per row: gather L of the row's bits, rebuild b2c with the OLD record, min1/min2/signs -> NEW record.
//       The parity of the current hard decision z = (L <= 0) falls out of the same gather, which gives the syndrome
//       test of the previous iteration (:424-445) and the adaptive variants' per-row factor (:745-757) for free.
//   VN  thread per bit: L = llr + sum over the bit's checks of the message rebuilt from the row record.
// Graph indices are stored per 32-node group in ELL form ([k][lane], coalesced) and shared by all CTAs through L1/L2;
// rows and bits are sorted by degree and a group never mixes degrees, so the inner loops carry no validity tests.
//
// Eligibility (host): float32 messages, min-sum family with the FAST precondition of run_batch.cuh (no NaN possible),
// every check degree <= 64 (rows of 33..64 edges own two records), fewer than 65533 records, state fits the 227 KB of
// shared memory. SPA / SPA-lin-approx have their own on-chip kernel (onchip_spa.cuh); float64 and n = 100k codes stream.
#pragma once
#include "common.cuh"

namespace qk {

typedef unsigned long long u64;

struct OnchipCombo {
    double qber;               // accurate QBER of the combination; < 0: take the per-frame / scalar `qber` array instead
    double primary, secondary; // decoding_scaling_factors (the float32 kernels round them to float)
    int has_cls;               // rate adaptation: punctured / shortened masks present
    int pad;
};

struct OnchipArgs {
    int n, m, words;
    int rec_slots;              // record slots in use (m + number of rows wider than 32 edges); + 2 scratch slots
    int n_groups_cn, n_groups_vn;
    // Index tables: one entry per (group, block of 4 edges, lane), so a lane fetches the indices of 4 edges with ONE
    // 8- or 16-byte load. A group holds 32 rows (bits) of ONE degree; which nodes share a group is chosen on the host
    // so that the lanes' shared-memory gathers fall into different banks (conflict-aware grouping, onchip_tables.cu).
    const int2 *cn_ginfo;       // [groups] {offset into cnT (in uint2), degree of the group's rows}
    const uint16_t *cn_row;     // [groups*32] row handled by (group, lane); padding lanes hold m (a scratch record slot)
    const uint2 *cnT;           // [off + kb*32 + lane] 4 x uint16: bit index of edges 4kb..4kb+3 of that row (padding: 0)
    const int2 *vn_ginfo;       // [groups] {offset into vT (in uint4), degree of the group's bits}, in schedule order
                                //          (degree 0 = an empty slot of the schedule)
    const uint16_t *vn_bit;     // [groups*32] bit handled by (group, lane); padding lanes hold n (a scratch L slot)
    const uint4 *vT;            // [off + kb*32 + lane] 4 x uint32: row << 9 | sh of checks 4kb..4kb+3 of that bit,
                                //                      sh = 32 - dc(row) + position in the row (padding: scratch row m)
    // Sum-product kernel only (onchip_spa.cuh): one float per edge in shared memory, word cn_moff[group] + k * 32 + lane
    // for edge k of the row handled by (group, lane); the variable phase has its own groups and addresses the words directly.
    const int *cn_moff;         // [groups_cn] first message word of the group
    const uint4 *sv_items;      // [items*32] variable phase: {4 x uint16 shared-memory WORD of the message (L starts at word
                                //            0, the messages follow), bit | first << 16 | last << 17, variable group}
    const int *sv_chunk;        // [warps per CTA + 1] items [sv_chunk[w], sv_chunk[w+1]) belong to warp w
    const int *sv_group_item0;  // [groups_sv] first item of the group
    int n_groups_sv;
    int msg_words;              // word msg_words is kept at 0.0f (padding of short blocks)
    // One launch may span several parameter COMBINATIONS of a sweep (frames [c * frames_per_combo, (c+1) * ...) belong to
    // combination c): scaling factors, QBER, punctured / shortened masks and the tally vector are per combination.
    const OnchipCombo *combos;  // [n_combos]
    long long frames_per_combo; // frames of one combination (n_frames when there is a single one)
    const uint32_t *cls_masks;  // [n_combos][2][words] packed punctured / shortened positions (shortened excludes punctured)
    int tally_len;
    long long n_frames;
    const uint32_t *alice_bits, *bob_bits;
    const double *qber;
    int qber_is_scalar;
    uint32_t *out_bits;
    int32_t *out_iters;
    uint8_t *out_flags;
    u64 *tally;
    u64 *next_frame;
    int max_iter;
    float thr;                  // +inf when the clamp is disabled
    double thr64;               // the same for the float64 kernel (onchip_minsum64.cuh)
};

// Shared-memory layout: rec[rec_slots+2] uint4 | L[n+1] float (padded to 16 B) | bob[words] | alice[words] | syn[groups_cn] | misc
__host__ __device__ inline size_t onchip_l_slots(int n) { return ((size_t)n + 1 + 3) / 4 * 4; }
__host__ __device__ inline size_t onchip_smem_bytes(int n, int rec_slots, int groups_cn) {
    const size_t words = (size_t)(n + 31) / 32;
    return ((size_t)rec_slots + 2) * 16 + onchip_l_slots(n) * 4 + (2 * words + (size_t)groups_cn) * 4 + 96;   // + frame id, FrameCtx
}

// The sum-product kernel (onchip_spa.cuh): msg[msg_words] float instead of the records, the rest alike.
__host__ __device__ inline size_t onchip_spa_smem_bytes(int n, int msg_words, int groups_cn, int groups_sv) {
    const size_t words = (size_t)(n + 31) / 32;
    return ((size_t)msg_words + 4) / 4 * 16 + onchip_l_slots(n) * 4 + (2 * words + (size_t)groups_cn + (size_t)groups_sv) * 4 + 96 +
           384;   // + the piecewise-linear tanh table of SPA-lin-approx (LinLut, 360 B)
}
constexpr size_t kOnchipSmemMax = 227 * 1024;   // opt-in shared memory per CTA on sm_100

// Record of a row: x = bits(c1), y = bits(c2) (non-negative magnitudes), z = final sign of the message on edge k in bit
// (dc-1-k), w = 32 - dc + position of the first minimum. A reader that knows sh = 32 - dc + k gets the sign with
// (z << sh) & 0x80000000 and the magnitude with (sh == w) ? c2 : c1.

// Per-frame parameters (shared memory, written by thread 0 when the CTA takes a frame).
struct FrameCtx {
    float lp, primary, secondary;
    int has_cls;
    const uint32_t *cls_punct, *cls_short;
    u64 *tally;
};

// One edge of a check node: gather L of the bit, rebuild the bit-to-check message with the OLD record, update the
// running min1 / min2 / argmin / sign state.  `rel` = old argmin - index of the first edge of the current block.
#define QK_CN_EDGE(J, COL)                                                                                              \
    {                                                                                                                   \
        const float Lv = L[(COL)];                                                                                      \
        const uint32_t mag = (rel == (J)) ? ro.y : ro.x;                                                                \
        /* clamp(L - c2b) (:447-461); first iteration: zero record and thr_b = +inf leave the unclamped LLR (:336-350) */ \
        const float braw = Lv - __uint_as_float(mag ^ (zs & 0x80000000u));                                              \
        zs <<= 1;                                                                                                       \
        /* for x neither NaN nor -0: (x <= 0) == sign bit of (bits(x) - 1); L and b are never -0 */                     \
        zacc ^= __float_as_uint(Lv) - 1u;                 /* parity of the hard decision L <= 0 (:414-422) */            \
        pacc ^= __float_as_uint(braw);                    /* parity of m < 0 (:383); the clamp keeps the sign */         \
        own = __funnelshift_l(__float_as_uint(braw) - 1u, own, 1);   /* (m > 0) ? +1 : -1 (:402): zero is negative (Q4) */ \
        const float ab = fminf(fabsf(braw), thr_b);       /* |clamp(x)| == min(|x|, thr) */                              \
        arg = (ab < m1) ? (kb + (J)) : arg;               /* first minimum */                                            \
        m2 = fminf(m2, fmaxf(ab, m1));                    /* == the if / else-if chain (:386-396) */                     \
        m1 = fminf(m1, ab);                                                                                             \
    }

// A row of 33..64 edges owns two consecutive records (edges 0..31 and 32..dc-1) with the same c1 / c2; the record that
// does not hold the first minimum carries kNoArg in w, which no table entry matches.
constexpr uint32_t kNoArg = 0x100u;

template <int ALG, bool WIDE>
__device__ __forceinline__ bool onchip_cn_phase(const OnchipArgs &a, const FrameCtx *ctx, const float *L, uint4 *rec,
                                                const uint32_t *synw, float thr_b, int warp, int lane, int nwarps) {
    bool unsat = false;
    for (int g = warp; g < a.n_groups_cn; g += nwarps) {
        const int2 gi = __ldg(a.cn_ginfo + g);
        const int dc_row = gi.y;                                  // degree of the group's rows (warp-uniform)
        const bool two = WIDE && dc_row > 32;                     // two records per row
        const int dc = two ? 32 : dc_row;                         // edges covered by the first record
        const uint32_t row = __ldg(a.cn_row + g * 32 + lane);     // first record slot of the lane's row
        const uint4 ro = rec[row];
        const uint2 *cp = a.cnT + gi.x + lane;
        float m1 = FLT_MAX, m2 = FLT_MAX;
        uint32_t zs = ro.z << (32 - dc);          // sign of the old message on the current edge in bit 31
        const int arg_old = (int)ro.w - (32 - dc);
        uint32_t own = 0, pacc = 0, zacc = 0;
        int arg = 0, kb = 0;
#pragma unroll 2
        for (; kb + 4 <= dc; kb += 4) {
            const uint2 cw = __ldg(cp + (kb >> 2) * 32);
            const int rel = arg_old - kb;
            QK_CN_EDGE(0, cw.x & 0xFFFFu)
            QK_CN_EDGE(1, cw.x >> 16)
            QK_CN_EDGE(2, cw.y & 0xFFFFu)
            QK_CN_EDGE(3, cw.y >> 16)
        }
        if (kb < dc) {                            // warp-uniform tail of 1..3 edges
            const uint2 cw = __ldg(cp + (kb >> 2) * 32);
            const int rel = arg_old - kb, left = dc - kb;
            QK_CN_EDGE(0, cw.x & 0xFFFFu)
            if (left > 1) QK_CN_EDGE(1, cw.x >> 16)
            if (left > 2) QK_CN_EDGE(2, cw.y & 0xFFFFu)
        }
        uint32_t own_first = own;
        if constexpr (WIDE) {
            if (two) {                            // edges 32..dc_row-1: the row's second record (warp-uniform branch)
                const int dc2 = dc_row - 32;
                const uint4 ro2 = rec[row + 1];
                zs = ro2.z << (32 - dc2);
                const int arg_old2 = (int)ro2.w - (32 - dc2);   // kNoArg gives a value no edge index reaches
                own = 0;
                for (; kb + 4 <= dc_row; kb += 4) {
                    const uint2 cw = __ldg(cp + (kb >> 2) * 32);
                    const int rel = arg_old2 - (kb - 32);
                    const uint4 &ro = ro2;
                    QK_CN_EDGE(0, cw.x & 0xFFFFu)
                    QK_CN_EDGE(1, cw.x >> 16)
                    QK_CN_EDGE(2, cw.y & 0xFFFFu)
                    QK_CN_EDGE(3, cw.y >> 16)
                }
                if (kb < dc_row) {
                    const uint2 cw = __ldg(cp + (kb >> 2) * 32);
                    const int rel = arg_old2 - (kb - 32), left = dc_row - kb;
                    const uint4 &ro = ro2;
                    QK_CN_EDGE(0, cw.x & 0xFFFFu)
                    if (left > 1) QK_CN_EDGE(1, cw.x >> 16)
                    if (left > 2) QK_CN_EDGE(2, cw.y & 0xFFFFu)
                }
            }
        }
        const uint32_t syn = (synw[g] >> lane) & 1u;
        const bool viol = (((zacc >> 31) ^ syn) & 1u) != 0;    // check not satisfied by the current hard decision
        unsat |= viol && row < (uint32_t)a.rec_slots;
        const float factor = (ALG >= 4 && viol) ? ctx->secondary : ctx->primary;   // (:749-757, :939-947)
        float c1, c2;
        if constexpr (ALG == 2 || ALG == 4) {
            c1 = factor * m1;
            c2 = factor * m2;
        } else {
            const float d1 = m1 - factor, d2 = m2 - factor;    // max(min - beta, 0) (:573-574)
            c1 = (d1 < 0.f) ? 0.f : d1;
            c2 = (d2 < 0.f) ? 0.f : d2;
        }
        c1 = fminf(c1, a.thr);                                 // threshold_matrix(check_to_bit), magnitudes (:411-412)
        c2 = fminf(c2, a.thr);
        const uint32_t rowneg = ((pacc >> 31) ^ syn) & 1u;     // (syndrome ? -1 : 1) * (-1)^negatives (:398-399)
        uint4 rn;
        rn.x = __float_as_uint(c1);
        rn.y = __float_as_uint(c2);
        rn.z = own_first ^ (0u - rowneg);
        rn.w = (!two || arg < 32) ? (uint32_t)(arg + 32 - dc) : kNoArg;
        rec[row] = rn;
        if constexpr (WIDE) {
            if (two) {
                rn.z = own ^ (0u - rowneg);
                rn.w = (arg >= 32) ? (uint32_t)(arg - 32 + 32 - (dc_row - 32)) : kNoArg;
                rec[row + 1] = rn;
            }
        }
    }
    return unsat;
}
#undef QK_CN_EDGE

__device__ __forceinline__ float onchip_llr(const FrameCtx *ctx, const uint32_t *bobw, uint32_t bit, float lp) {
    const uint32_t w = bit >> 5, s = bit & 31u;
    // qkd_ldpc_algorithm.cpp:1043-1049; `0 - lp` instead of `-lp`: never -0 (QBER 0.5 gives lp = 0), which the
    // bit tricks of the check phase rely on -- the decision (L <= 0) is the reference's either way
    float v = ((bobw[w] >> s) & 1u) ? 0.f - lp : lp;
    if (ctx->has_cls) {
        if ((__ldg(ctx->cls_punct + w) >> s) & 1u) v = 1e-4f;  // punctured: ALMOST_ZERO (:1155)
        else if ((__ldg(ctx->cls_short + w) >> s) & 1u) v = FLT_MAX;   // shortened: largest finite value (:1164)
    }
    return v;
}

// The check-to-bit message addressed by table entry `ent` = row << 9 | sh, added to the running sum.
#define QK_VN_EDGE(ENT)                                                                                                 \
    {                                                                                                                   \
        const uint4 r = *reinterpret_cast<const uint4 *>(recb + ((ENT) >> 5));                                          \
        const uint32_t mag = (((ENT) ^ r.w) & 0x1FFu) ? r.x : r.y;   /* r.w == kNoArg never matches */                                                         \
        acc = acc + __uint_as_float(mag ^ (__funnelshift_l(0u, r.z, (ENT)) & 0x80000000u));   /* r.z << sh */           \
    }

__device__ __forceinline__ void onchip_vn_phase(const OnchipArgs &a, const FrameCtx *ctx, float *L, const uint4 *rec, const uint32_t *bobw,
                                                float lp, int warp, int lane, int nwarps) {
    const unsigned char *recb = reinterpret_cast<const unsigned char *>(rec);
    // groups come in SCHEDULE order: entry g is handled by warp g % nwarps, and the host dealt the groups to the warps
    // longest-first so that all warps of the CTA finish the phase together (inst_onchip.cu)
    for (int g = warp; g < a.n_groups_vn; g += nwarps) {
        const int2 gi = __ldg(a.vn_ginfo + g);
        const int dv = gi.y;
        const uint32_t bit = __ldg(a.vn_bit + g * 32 + lane);
        float acc = onchip_llr(ctx, bobw, bit < (uint32_t)a.n ? bit : 0u, lp);
        const uint4 *ep = a.vT + gi.x + lane;
        int kb = 0;
        // ascending check order, starting from the LLR (std::accumulate, :414-417)
#pragma unroll 2
        for (; kb + 4 <= dv; kb += 4) {
            const uint4 ew = __ldg(ep + (kb >> 2) * 32);
            QK_VN_EDGE(ew.x)
            QK_VN_EDGE(ew.y)
            QK_VN_EDGE(ew.z)
            QK_VN_EDGE(ew.w)
        }
        if (kb < dv) {                            // warp-uniform tail of 1..3 checks
            const uint4 ew = __ldg(ep + (kb >> 2) * 32);
            const int left = dv - kb;
            QK_VN_EDGE(ew.x)
            if (left > 1) QK_VN_EDGE(ew.y)
            if (left > 2) QK_VN_EDGE(ew.z)
        }
        L[bit] = acc;                             // padding lanes write the scratch slot L[n]
    }
}
#undef QK_VN_EDGE

template <int ALG, bool WIDE>
__global__ void __launch_bounds__(768, 2) onchip_minsum_kernel(const OnchipArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *rec = reinterpret_cast<uint4 *>(smem_raw);
    float *L = reinterpret_cast<float *>(rec + a.rec_slots + 2);
    uint32_t *bobw = reinterpret_cast<uint32_t *>(L + onchip_l_slots(a.n));
    uint32_t *alw = bobw + a.words;
    uint32_t *synw = alw + a.words;
    uint32_t *tail = synw + a.n_groups_cn;
    long long *s_frame = reinterpret_cast<long long *>(tail + ((2 * a.words + a.n_groups_cn) & 1));
    FrameCtx *ctx = reinterpret_cast<FrameCtx *>(s_frame + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    constexpr bool kAdaptive = (ALG >= 4);
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
        __syncthreads();
        // L = a-priori LLR; Alice's syndrome (calculate_syndrome, array_and_matrix_operations.cpp:936-950); records = 0
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
            rec[row] = make_uint4(0u, 0u, 0u, 0u);
            if (WIDE && gi.y > 32) rec[row + 1] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();

        int iters = a.max_iter, run = a.max_iter;
        bool success = false;
        for (int it = 1;; ++it) {
            // the check-node pass of iteration `it`; at it = max_iter + 1 it only serves as the syndrome test of the
            // last hard decision (non-adaptive variants, :424-445)
            const bool unsat = onchip_cn_phase<ALG, WIDE>(a, ctx, L, rec, synw, it == 1 ? inf : a.thr, warp, lane, nwarps);
            const bool any_unsat = __syncthreads_or(unsat) != 0;
            if (!kAdaptive) {
                if (it > 1 && !any_unsat) { success = true; iters = run = it - 1; break; }   // z of iteration it-1 (:439-445)
                if (it > a.max_iter) break;
            } else {
                if (!any_unsat) { success = true; iters = it; run = it - 1; break; }         // exit test before the VN step (:770-776)
            }
            onchip_vn_phase(a, ctx, L, rec, bobw, lp, warp, lane, nwarps);
            __syncthreads();
            if (kAdaptive && it == a.max_iter) break;          // the decision of the last iteration is never tested (Q10)
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

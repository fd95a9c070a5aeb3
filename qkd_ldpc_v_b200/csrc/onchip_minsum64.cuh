// On-chip min-sum decoder with FLOAT64 state: the kernel of onchip_minsum.cuh with `double` totals and `double` row
// magnitudes, so that NMSA / OMSA / ANMSA / AOMSA reproduce the reference's double arithmetic bit for bit at on-chip speed
// (the float32 kernel agrees with the reference on 95.7 .. 99.2 % of the iteration counts for the offset / adaptive
// variants: exact-tie sums are resolved by rounding noise -- SURVEY.md 7, hard part 1). Same phases, same iteration
// accounting, same layout tables (onchip_layout.hpp: storage order = processing order, conflict-aware lanes and check-node
// edge order); only the state is wider:
//     L[l_slots]  double  total LLR after the last variable-node phase (llr before the first iteration), 8-byte gathers
//     rec[slots]  16 B    {bits(c1) lo, hi, sign bit per edge of the row, position of the first minimum}: what EVERY edge of
//                         the variable phase needs, gathered exactly like the float32 kernel's 16-byte record
//     c2[slots]   double  the second magnitude: read by the check phase (coalesced) and by the one edge per row that holds
//                         the first minimum (a predicated 8-byte load, ~1 edge in dc)
// = 8 l_slots + 24 rec_slots bytes (126 KB for n = 10240, m = 1801; 206 KB for the R = 0.5 code of ADAPTIVE R.json): one CTA
// of up to 1024 threads per SM. Every double operation is the reference's, on the same operands in the same order
// (qkd_ldpc_algorithm.cpp:374-461, 542-577, 719-768, 909-958), under the same precondition as the float32 kernel (no
// message can become NaN / inf; onchip_usable, inst_onchip.cu), so the results are bit-identical to the reference and to
// the streaming float64 kernels (tests/test_gpu_onchip.py).
//
// Instruction budget of the check phase (it is bound by the ALU pipe; B200 has no double min / max instruction): sign tests
// come from double subtractions on the FP64 pipe -- for x neither NaN nor -0 the sign bit of (0 - x) is (x > 0) --, the
// magnitude clamp is applied to min1 / min2 after the loop (min over clamped values == clamped min), the first minimum's
// position is the last "|m| < min1" event of a bit word, and min1 / min2 follow the reference's if / else-if chain
// (:386-396) as two compares and three selects.
#pragma once
#include "onchip_minsum.cuh"

namespace qk {

// Shared-memory layout: L[l_slots + 2] double (slot l_slots holds +inf: the total that padding edges of mixed-degree check
// groups gather) | rec[rec_slots + 2] uint4 | c2[rec_slots + 2] double | Bob's bits in slot order [l_slots / 32] |
// syn[groups_cn] | frame id, FrameCtx64, phase clocks. While a frame is set up the record array doubles as staging space for
// the key words (onchip_staging_fits, as in the float32 kernel).
__host__ __device__ inline size_t onchip64_l_bytes(int l_slots) { return ((size_t)l_slots + 2) * 8; }
__host__ __device__ inline size_t onchip64_misc_offset(int l_slots, int rec_slots, int groups_cn) {
    return (onchip64_l_bytes(l_slots) + ((size_t)rec_slots + 2) * 24 + ((size_t)l_slots / 32 + (size_t)groups_cn) * 4 + 7) / 8 * 8;
}
__host__ __device__ inline size_t onchip64_smem_bytes(int l_slots, int rec_slots, int groups_cn) {
    return onchip64_misc_offset(l_slots, rec_slots, groups_cn) + 8 + 64 + 24;   // + frame id, FrameCtx64, phase clocks
}

// Per-frame parameters of the float64 kernel (shared memory, written by thread 0 when the CTA takes a frame).
struct FrameCtx64 {
    double lp, primary, secondary;
    int has_cls, pad;
    const uint32_t *cls_punct, *cls_short;
    u64 *tally;
};

__device__ __forceinline__ uint32_t hi32(double x) { return (uint32_t)__double2hiint(x); }

// One edge of a check node (see QK_CN_EDGE of the float32 kernel): OFF = byte offset of the bit's total, the old record's
// magnitudes are c1h:c1l / c2h:c2l (raw bits, non-negative), the old signs shift through bit 31 of zs.
//   * m = L - c2b is never -0 (L is never -0: the a-priori LLR is built as 0 - lp, and a sum that starts from it cannot
//     come out as -0), so the sign bit of (0 - m) is exactly (m > 0) -- also when c2b is a -0 of the offset variants.
#define QK_CN_EDGE64(J, OFF)                                                                                            \
    {                                                                                                                   \
        const double Lv = *reinterpret_cast<const double *>(smem + (OFF));                                              \
        const bool am = (rel == (J));                                                                                   \
        const uint32_t mh = am ? c2h : c1h, ml = am ? c2l : c1l;                                                        \
        const double c2b = __hiloint2double((int)(mh ^ (zs & 0x80000000u)), (int)ml);                                   \
        zs <<= 1;                                                                                                       \
        const double braw = Lv - c2b;       /* L - c2b (:447-461); first iteration: zero record leaves the LLR (:336-350) */ \
        const double nb = 0. - braw;                                                                                    \
        zpos ^= hi32(0. - Lv);                            /* parity of L > 0: the hard decision is z = !(L > 0) (:414-422) */ \
        pacc ^= hi32(braw);                               /* parity of m < 0 (:383): zero is positive here (Q4) */       \
        own = __funnelshift_l(hi32(nb), own, 1);          /* (m > 0) ? +1 : -1 (:402): zero is negative (Q4) */          \
        const double ab = fabs(nb);                                                                                     \
        lt = __funnelshift_l(hi32(ab - m1), lt, 1);       /* |m| < min1: a new first minimum */                          \
        const bool p1 = ab < m1, p2 = ab < m2;            /* the if / else-if chain (:386-396) */                        \
        m2 = p1 ? m1 : (p2 ? ab : m2);                                                                                  \
        m1 = p1 ? ab : m1;                                                                                              \
    }

template <int ALG, bool WIDE>
__device__ __forceinline__ bool onchip64_cn_phase(const OnchipArgs &a, const FrameCtx64 *ctx, const unsigned char *smem, uint4 *rec,
                                                  double *c2a, const uint32_t *synw, double thr_b, int warp, int lane, int nwarps) {
    bool unsat = false;
    for (int g = warp; g < a.n_groups_cn2; g += nwarps) {
        const int4 gi = __ldg(a.cn_g2 + g);
        const int dc_row = gi.y;                                  // degree of the group's rows (warp-uniform)
        const bool two = WIDE && dc_row > 32;                     // two records per row
        const int dc = two ? 32 : dc_row;                         // edges covered by the first record
        const bool valid = lane < gi.w;
        const int slot = valid ? gi.z + lane : a.rec_slots;       // padding lanes work on the scratch record
        const uint4 ro = rec[slot];
        const double c2o = c2a[slot];
        uint32_t c1l = ro.x, c1h = ro.y, c2l = (uint32_t)__double2loint(c2o), c2h = hi32(c2o);
        const uint4 *cp = a.cnT2 + gi.x + lane;
        double m1 = DBL_MAX, m2 = DBL_MAX;                        // (:378-379)
        uint32_t zs = ro.z << (32 - dc);          // sign of the old message on the current edge in bit 31
        const int arg_old = (int)ro.w - (32 - dc);
        uint32_t own = 0, pacc = 0, zpos = 0, lt = 0;
        int kb = 0;
#pragma unroll 2
        for (; kb + 4 <= dc; kb += 4) {
            const uint4 cw = __ldg(cp + (kb >> 2) * 32);
            const int rel = arg_old - kb;
            QK_CN_EDGE64(0, cw.x)
            QK_CN_EDGE64(1, cw.y)
            QK_CN_EDGE64(2, cw.z)
            QK_CN_EDGE64(3, cw.w)
        }
        if (kb < dc) {                            // warp-uniform tail of 1..3 edges
            const uint4 cw = __ldg(cp + (kb >> 2) * 32);
            const int rel = arg_old - kb, left = dc - kb;
            QK_CN_EDGE64(0, cw.x)
            if (left > 1) QK_CN_EDGE64(1, cw.y)
            if (left > 2) QK_CN_EDGE64(2, cw.z)
        }
        // the last "new minimum" event is the first minimum: edge k sits in bit dc-1-k of lt (none: every |m| is DBL_MAX)
        int arg = lt ? dc - __ffs((int)lt) : 0;
        uint32_t own_first = own;
        int slot2 = 0;
        if constexpr (WIDE) {
            if (two) {                            // edges 32..dc_row-1: the row's second record (warp-uniform branch)
                const int dc2 = dc_row - 32;
                slot2 = valid ? slot + gi.w : a.rec_slots;
                const uint4 ro2 = rec[slot2];
                const double c2o2 = c2a[slot2];
                c1l = ro2.x; c1h = ro2.y; c2l = (uint32_t)__double2loint(c2o2); c2h = hi32(c2o2);
                zs = ro2.z << (32 - dc2);
                const int arg_old2 = (int)ro2.w - (32 - dc2);   // kNoArg gives a value no edge index reaches
                own = 0;
                lt = 0;
                for (; kb + 4 <= dc_row; kb += 4) {
                    const uint4 cw = __ldg(cp + (kb >> 2) * 32);
                    const int rel = arg_old2 - (kb - 32);
                    QK_CN_EDGE64(0, cw.x)
                    QK_CN_EDGE64(1, cw.y)
                    QK_CN_EDGE64(2, cw.z)
                    QK_CN_EDGE64(3, cw.w)
                }
                if (kb < dc_row) {
                    const uint4 cw = __ldg(cp + (kb >> 2) * 32);
                    const int rel = arg_old2 - (kb - 32), left = dc_row - kb;
                    QK_CN_EDGE64(0, cw.x)
                    if (left > 1) QK_CN_EDGE64(1, cw.y)
                    if (left > 2) QK_CN_EDGE64(2, cw.z)
                }
                if (lt) arg = 32 + dc2 - __ffs((int)lt);
            }
        }
        m1 = fmin(m1, thr_b);                                     // threshold_matrix(bit_to_check), magnitudes (:447-461)
        m2 = fmin(m2, thr_b);
        const uint32_t syn = (synw[g] >> lane) & 1u;
        const bool viol = (((zpos >> 31) ^ (uint32_t)dc_row ^ syn) & 1u) != 0;   // parity of (L <= 0) = parity of dc + parity of (L > 0)
        unsat |= viol && valid;
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
        uint4 rn;
        rn.x = (uint32_t)__double2loint(c1);
        rn.y = hi32(c1);
        rn.z = own_first ^ (rowneg - 1u);                      // bit = message negative: !(m > 0) xor row sign
        rn.w = (!two || arg < 32) ? (uint32_t)(arg + 32 - dc) : kNoArg;
        rec[slot] = rn;
        c2a[slot] = c2;
        if constexpr (WIDE) {
            if (two) {
                rn.z = own ^ (rowneg - 1u);
                rn.w = (arg >= 32) ? (uint32_t)(arg - 32 + 32 - (dc_row - 32)) : kNoArg;
                rec[slot2] = rn;
                c2a[slot2] = c2;
            }
        }
    }
    return unsat;
}
#undef QK_CN_EDGE64

// a-priori LLR of the bit with key bit `bob` (qkd_ldpc_algorithm.cpp:1043-1049); never -0. `pos` indexes the punctured /
// shortened masks (slot order).
__device__ __forceinline__ double onchip64_llr_of(const FrameCtx64 *ctx, uint32_t bob, uint32_t pos, double lp) {
    double v = bob ? 0. - lp : lp;
    if (ctx->has_cls) {
        const uint32_t w = pos >> 5, s = pos & 31u;
        if ((__ldg(ctx->cls_punct + w) >> s) & 1u) v = 1e-4;   // punctured: ALMOST_ZERO (:1155)
        else if ((__ldg(ctx->cls_short + w) >> s) & 1u) v = DBL_MAX;   // shortened: largest finite value (:1164)
    }
    return v;
}

// The check-to-bit message addressed by table entry `ent` = (16 * record slot) << 5 | sh, added to the running sum: the
// 16-byte record first, then -- only on the edge that holds the row's first minimum -- the second magnitude.
#define QK_VN_EDGE64(ENT)                                                                                               \
    {                                                                                                                   \
        const uint4 r = *reinterpret_cast<const uint4 *>(recb + ((ENT) >> 5));                                          \
        uint2 mg = make_uint2(r.x, r.y);                                                                                \
        if ((((ENT) ^ r.w) & 0x1FFu) == 0) mg = *reinterpret_cast<const uint2 *>(c2b + ((ENT) >> 6));   /* kNoArg never matches */ \
        acc = acc + __hiloint2double((int)(mg.y ^ (__funnelshift_l(0u, r.z, (ENT)) & 0x80000000u)), (int)mg.x);         \
    }

// (Measured and not kept, B200, profiles/r02_u_onchip64.md: straight-line code per degree <= 8 with the 16-bit table of the
// float32 kernel, -1 .. -3 %; index blocks requested one trip ahead and the next group's header / first block during the
// current group, +1.7 % on the irregular code, -6 % on the dv = 4 codes.)
__device__ __forceinline__ void onchip64_vn_phase(const OnchipArgs &a, const FrameCtx64 *ctx, double *L, const uint4 *rec, const double *c2a,
                                                  const uint32_t *bobs, double lp, int warp, int lane) {
    const unsigned char *recb = reinterpret_cast<const unsigned char *>(rec);
    const unsigned char *c2b = reinterpret_cast<const unsigned char *>(c2a);
    const double nlp = 0. - lp;
    const int has_cls = ctx->has_cls;
    // the host dealt the groups to the warps longest-first (inst_onchip.cu); warp w owns a contiguous run of the dealt list
    const int g_end = __ldg(a.vn_start + warp + 1);
    for (int g = __ldg(a.vn_start + warp); g < g_end; ++g) {
        const int4 gi = __ldg(a.vn_g2 + g);
        const int dv = gi.y;
        const uint32_t s = (uint32_t)gi.z + (uint32_t)lane;       // the group's 32 slots start on a multiple of 32
        double acc = ((bobs[gi.z >> 5] >> lane) & 1u) ? nlp : lp; // a-priori LLR; padding lanes see Bob bit 0
        if (has_cls) acc = onchip64_llr_of(ctx, (bobs[gi.z >> 5] >> lane) & 1u, s, lp);
        const uint4 *ep = a.vT2 + gi.x + lane;
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
        L[s] = acc;                               // consecutive slots: coalesced; padding lanes own padding slots
    }
}
#undef QK_VN_EDGE64

template <int ALG, bool WIDE>
__global__ void __launch_bounds__(1024, 1) onchip_minsum64_kernel(const OnchipArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *L = reinterpret_cast<double *>(smem_raw);
    uint4 *rec = reinterpret_cast<uint4 *>(smem_raw + onchip64_l_bytes(a.l_slots));
    double *c2a = reinterpret_cast<double *>(rec + a.rec_slots + 2);
    uint32_t *bobs = reinterpret_cast<uint32_t *>(c2a + a.rec_slots + 2);
    uint32_t *synw = bobs + a.l_slots / 32;
    long long *s_frame = reinterpret_cast<long long *>(smem_raw + onchip64_misc_offset(a.l_slots, a.rec_slots, a.n_groups_cn2));
    FrameCtx64 *ctx = reinterpret_cast<FrameCtx64 *>(s_frame + 1);
    long long *s_clk = reinterpret_cast<long long *>(reinterpret_cast<unsigned char *>(ctx) + 64);   // profiling: [0] check, [1] variable, [2] start
    if (a.phase_clk && threadIdx.x == 0) {
        s_clk[0] = s_clk[1] = 0;
        s_clk[2] = clock64();
    }
    // frame set-up only: the key words as they come from HBM and Alice's bits in slot order, inside the record array
    uint32_t *st_bob = reinterpret_cast<uint32_t *>(rec), *st_alice = st_bob + a.words, *alice_s = st_alice + a.words;   // [l_slots/32 + 1]

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
                ctx->cls_punct = a.cls_masks2 + combo * 2 * (a.l_slots / 32);   // slot order
                ctx->cls_short = ctx->cls_punct + a.l_slots / 32;
                ctx->tally = a.tally ? a.tally + combo * a.tally_len : nullptr;
            }
        }
        __syncthreads();
        const long long f = *s_frame;
        if (f >= a.n_frames) break;
        const double lp = ctx->lp;
        for (int w = tid; w < a.words; w += blockDim.x) {
            st_bob[w] = a.bob_bits[f * a.words + w];
            st_alice[w] = a.alice_bits[f * a.words + w];
        }
        __syncthreads();
        // key bits into slot order; L = a-priori LLR (qkd_ldpc_algorithm.cpp:1043-1049). Four words per trip, their slot -> bit
        // entries (L2) requested together: frame set-up is a chain of latencies, not of work
        for (int s0 = warp * 32; s0 < a.l_slots; s0 += nwarps * 128) {
            uint32_t sbs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int s = s0 + u * nwarps * 32 + lane;
                sbs[u] = s < a.l_slots ? (uint32_t)__ldg(a.slot_bit + s) : 0xFFFFu;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int sw = s0 + u * nwarps * 32, s = sw + lane;
                if (sw >= a.l_slots) break;           // warp-uniform
                const uint32_t sb = sbs[u];
                const bool v = sb != 0xFFFFu;             // else a padding slot: never gathered, holds a harmless finite value
                const uint32_t bit = v ? sb : 0u;
                const uint32_t bb = (st_bob[bit >> 5] >> (bit & 31u)) & 1u, ab = (st_alice[bit >> 5] >> (bit & 31u)) & 1u;
                const uint32_t wb = __ballot_sync(0xffffffffu, v && bb), wa = __ballot_sync(0xffffffffu, v && ab);
                if (lane == 0) {
                    bobs[sw >> 5] = wb;
                    alice_s[sw >> 5] = wa;
                }
                L[s] = v ? onchip64_llr_of(ctx, bb, (uint32_t)s, lp) : 1.;
            }
        }
        if (tid == 0) {
            L[a.l_slots] = inf;                   // gathered by the padding edges of mixed-degree check groups, never written
            alice_s[a.l_slots >> 5] = 0u;         // ... and no bit of Alice's key for the syndrome
        }
        __syncthreads();
        // Alice's syndrome (calculate_syndrome, array_and_matrix_operations.cpp:936-950) over the check-phase table
        for (int g = warp; g < a.n_groups_cn2; g += nwarps) {
            const int4 gi = __ldg(a.cn_g2 + g);
            const uint4 *cp = a.cnT2 + gi.x + lane;
            uint32_t sy = 0;
            const int last_block = (gi.y - 1) >> 2;
            for (int kb = 0; kb < gi.y; kb += 16) {       // four index blocks (L2) in flight
                uint4 cws[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) cws[u] = __ldg(cp + min((kb >> 2) + u, last_block) * 32);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int left = gi.y - kb - 4 * u;
                    if (left <= 0) break;                 // warp-uniform
                    const uint32_t c0 = cws[u].x >> 3, c1 = cws[u].y >> 3, c2 = cws[u].z >> 3, c3 = cws[u].w >> 3;   // slots
                    sy ^= alice_s[c0 >> 5] >> (c0 & 31u);
                    if (left > 1) sy ^= alice_s[c1 >> 5] >> (c1 & 31u);
                    if (left > 2) sy ^= alice_s[c2 >> 5] >> (c2 & 31u);
                    if (left > 3) sy ^= alice_s[c3 >> 5] >> (c3 & 31u);
                }
            }
            const uint32_t sw = __ballot_sync(0xffffffffu, (sy & 1u) != 0 && lane < gi.w);
            if (lane == 0) synw[g] = sw;
        }
        __syncthreads();
        for (int i = tid; i < a.rec_slots + 2; i += blockDim.x) {   // over the staging words
            rec[i] = make_uint4(0u, 0u, 0u, 0u);
            c2a[i] = 0.;
        }
        __syncthreads();

        int iters = a.max_iter, run = a.max_iter;
        bool success = false;
        for (int it = 1;; ++it) {
            // the check-node pass of iteration `it`; at it = max_iter + 1 it only serves as the syndrome test of the
            // last hard decision (non-adaptive variants, :424-445)
            if (a.phase_clk && tid == 0) s_clk[0] -= clock64();
            const bool unsat = onchip64_cn_phase<ALG, WIDE>(a, ctx, smem_raw, rec, c2a, synw, it == 1 ? inf : a.thr64, warp, lane, nwarps);
            const bool any_unsat = __syncthreads_or(unsat) != 0;
            if (a.phase_clk && tid == 0) s_clk[0] += clock64();
            if (!kAdaptive) {
                if (it > 1 && !any_unsat) { success = true; iters = run = it - 1; break; }   // z of iteration it-1 (:439-445)
                if (it > a.max_iter) break;
            } else {
                if (!any_unsat) { success = true; iters = it; run = it - 1; break; }         // exit test before the VN step (:770-776)
            }
            if (a.phase_clk && tid == 0) s_clk[1] -= clock64();
            onchip64_vn_phase(a, ctx, L, rec, c2a, bobs, lp, warp, lane);
            __syncthreads();
            if (a.phase_clk && tid == 0) s_clk[1] += clock64();
            if (kAdaptive && it == a.max_iter) break;          // the decision of the last iteration is never tested (Q10)
        }

        // bob_solution = last hard decision (L <= 0), packed in natural bit order; keys compare (arrays_equal, :1087)
        // (lane r of a warp fetches Alice's word of the warp's round r up front; bit -> slot entries four rounds at a time)
        uint32_t diff = 0;
        for (int w0 = warp; w0 < a.words; w0 += nwarps * 32) {
            const int wl = w0 + lane * nwarps;
            const uint32_t aw = wl < a.words ? a.alice_bits[f * a.words + wl] : 0u;
            for (int r0 = 0; r0 < 32 && w0 + r0 * nwarps < a.words; r0 += 4) {
                uint32_t sl[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = (w0 + (r0 + u) * nwarps) * 32 + lane;
                    sl[u] = i < a.n ? (uint32_t)__ldg(a.bit_slot + i) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int w = w0 + (r0 + u) * nwarps, i = w * 32 + lane;
                    if (w >= a.words) break;          // warp-uniform
                    const uint32_t word = __ballot_sync(0xffffffffu, i < a.n && L[sl[u]] <= 0.);
                    const uint32_t al = __shfl_sync(0xffffffffu, aw, r0 + u);
                    if (lane == 0) {
                        if (a.out_bits) a.out_bits[f * a.words + w] = word;
                        diff |= word ^ al;
                    }
                }
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
    if (a.phase_clk && tid == 0) {
        atomicAdd(a.phase_clk + 0, (u64)s_clk[0]);
        atomicAdd(a.phase_clk + 1, (u64)s_clk[1]);
        atomicAdd(a.phase_clk + 2, (u64)(clock64() - s_clk[2]));
    }
}

}  // namespace qk

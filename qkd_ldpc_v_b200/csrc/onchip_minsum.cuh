// On-chip min-sum decoder (float32 state): one frame per CTA, the whole belief-propagation state of the frame lives in
// shared memory for all iterations; HBM is touched only for the packed key bits going in and the packed decision coming out.
//
// Why this is possible. For the min-sum family a check node's dc outgoing messages take only two magnitudes
// (qkd_ldpc_algorithm.cpp:400-408): factor*min1 for every edge except the one holding the minimum, which gets
// factor*min2 (ties: min2 == min1, so "first minimum" and the reference's by-value test `|m| == min1` select the
// same number). A row is therefore fully described by a 16-byte RECORD {c1, c2, sign bits, argmin}, and the
// bit-to-check message need not be stored at all: b2c = clamp(L - c2b) (:447-461) is recomputed from the bit's
// total L and the row's previous record when the check node needs it. State per frame:
//     L[n]   float   total LLR after the last variable-node phase (llr before the first iteration)
//     rec[m] uint4   {bits(c1), bits(c2), sign bit per edge of the row, position of the first minimum}
// = 4n + 16m bytes (72 KB for n=10240, m=2048 instead of 242 KB of float messages), so three frames fit one SM.
// Every float operation of the reference is performed on the same operands in the same order (the sum over a
// bit's checks runs in ascending check order starting from the LLR, :414-422), so the results are bit-identical
// to the streaming float32 kernels (step_kernels.cuh) and to the f32 oracle.
//
// Phases per iteration (two __syncthreads):
//   CN  thread per row: gather L of the row's bits, rebuild b2c with the OLD record, min1/min2/signs -> NEW record.
//       The parity of the current hard decision z = (L <= 0) falls out of the same gather, which gives the syndrome
//       test of the previous iteration (:424-445) and the adaptive variants' per-row factor (:745-757) for free.
//   VN  thread per bit: L = llr + sum over the bit's checks of the message rebuilt from the row record.
// Storage order = processing order (onchip_layout.hpp): the totals L are stored group by group as the variable phase
// produces them and the records group by group as the check phase produces them, so both phases write coalesced and
// conflict-free without an index; Bob's bits (the LLR signs) and the punctured / shortened masks are permuted into the
// same slot order once per frame / per combination. Which nodes share a warp, the lane of every node and the order in
// which a check node walks its edges (free: min1 / min2 / sign parity do not depend on it) are chosen on the host so that
// the 4-byte gathers of the check phase hit 32 different banks (1.2 wavefronts per 32 edges in the bank model) and the
// 16-byte record gathers of the variable phase collide as little as the graph allows. Graph indices come per 32-node
// group in ELL form ([block of 4 edges][lane], one 8- / 16-byte load per lane and block) through L1/L2 and carry the
// shared-memory byte offset itself, so a gather costs no address arithmetic.
//
// Eligibility (host): float32 messages, min-sum family with the FAST precondition of run_batch.cuh (no NaN possible),
// every check degree <= 64 (rows of 33..64 edges own two records), n < 65535, state fits the 227 KB
// of shared memory. SPA / SPA-lin-approx have their own on-chip kernel (onchip_spa.cuh, natural-order tables of
// onchip_tables.cu); float64 state: onchip_minsum64.cuh on the same tables as this kernel; n = 100k codes stream.
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
    int n_groups_cn;
    // Check-phase index table in natural order (sum-product kernel): one entry per (group, block of 4 edges, lane), so a lane
    // fetches the indices of 4 edges with ONE 8-byte load. A group holds 32 rows of ONE degree.
    const int2 *cn_ginfo;       // [groups] {offset into cnT (in uint2), degree of the group's rows}
    const uint16_t *cn_row;     // [groups*32] row handled by (group, lane); padding lanes hold m
    const uint2 *cnT;           // [off + kb*32 + lane] 4 x uint16: bit index of edges 4kb..4kb+3 of that row (padding: 0)
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
    // min-sum kernels (float32 and float64 state): tables of onchip_layout.hpp (storage order = processing order)
    int n_groups_cn2, l_slots;  // l_slots: 32 per variable-phase group
    const int4 *cn_g2;          // [groups] {offset into cnT2, degree, first record slot, rows in the group}
    const uint4 *cnT2;          // [off + kb*32 + lane] 4 x uint32: shared-memory BYTE offset of the totals of edges 4kb..4kb+3
                                //                      (4 bytes per total for the float32 kernel, 8 for the float64 one)
    const int4 *vn_g2;          // [groups, dealt to the warps] {offset into vT2, degree, first total slot (multiple of 32), bits in the group}
    const int *vn_start;        // [warps per CTA + 1] warp w handles vn_g2[vn_start[w] .. vn_start[w+1])
    const uint4 *vT2;           // [off + kb*32 + lane] 4 x uint32: (16 * record slot) << 5 | sh; entries past the degree: the all-zero record
    const uint2 *vT16;          // the same in 4 x uint16 sh << 11 | record slot (codes with at most 2048 records; else null)
    const uint16_t *slot_bit;   // [l_slots] bit whose total lives in slot s; 0xFFFF = padding slot
    const uint16_t *bit_slot;   // [n] inverse
    const uint32_t *cls_masks2; // [n_combos][2][l_slots/32] punctured / shortened masks in SLOT order
    u64 *phase_clk;             // profiling: [0] check phases, [1] variable phases, [2] whole kernel, SM clocks summed over the CTAs
};

// Shared-memory layout: L[l_slots + 4] float (slot l_slots holds +inf: the total that padding edges of mixed-degree check groups
// gather) | rec[rec_slots+2] uint4 (REC8: uint2 records, then c2[rec_slots+2] float) | bob bits in slot order [l_slots/32] |
// syn[groups_cn] | frame id, FrameCtx. While a frame is set
// up the record array doubles as staging space for the key words (2 * words as they come from HBM + l_slots/32 + 1 of
// Alice's bits in slot order, the last word zero for the +inf slot).
__host__ __device__ inline size_t onchip_l_slots(int n) { return ((size_t)n + 1 + 3) / 4 * 4; }   // sum-product kernel
__host__ __device__ inline size_t onchip_l_bytes(int l_slots) { return ((size_t)l_slots + 4) * 4; }
// rec_bytes: 16 (one uint4 record per slot) or 12 (REC8: an 8-byte record {c1, signs << 5 | position code} + a float c2 per slot)
__host__ __device__ inline size_t onchip_misc_offset(int l_slots, int rec_slots, int groups_cn, int rec_bytes = 16) {
    return (onchip_l_bytes(l_slots) + ((size_t)rec_slots + 2) * (size_t)rec_bytes + ((size_t)l_slots / 32 + (size_t)groups_cn) * 4 + 7) / 8 * 8;
}
__host__ __device__ inline size_t onchip_smem_bytes(int l_slots, int rec_slots, int groups_cn, int rec_bytes = 16) {
    return onchip_misc_offset(l_slots, rec_slots, groups_cn, rec_bytes) + 8 + 48 + 24;   // + frame id, FrameCtx, phase clocks
}
__host__ __device__ inline bool onchip_staging_fits(int n, int l_slots, int rec_slots, int rec_bytes = 16) {
    return ((size_t)rec_slots + 2) * (size_t)rec_bytes >= (2 * ((size_t)(n + 31) / 32) + (size_t)l_slots / 32 + 1) * 4;
}

// The sum-product kernel (onchip_spa.cuh): msg[msg_words] float instead of the records, the rest alike.
__host__ __device__ inline size_t onchip_spa_smem_bytes(int n, int msg_words, int groups_cn, int groups_sv) {
    const size_t words = (size_t)(n + 31) / 32;
    return ((size_t)msg_words + 4) / 4 * 16 + onchip_l_slots(n) * 4 + (2 * words + (size_t)groups_cn + (size_t)groups_sv) * 4 + 96 +
           384;   // + the piecewise-linear tanh table of SPA-lin-approx (LinLut, 360 B)
}
constexpr size_t kOnchipSmemMax = 227 * 1024;   // opt-in shared memory per CTA on sm_100

// Record of a row (16-byte format): x = bits(c1), y = bits(c2) (non-negative magnitudes), z = final sign of the message on
// edge k in bit (dc-1-k), w = 32 - dc + position of the first minimum. A reader that knows sh = 32 - dc + k gets the sign with
// (z << sh) & 0x80000000 and the magnitude with (sh == w) ? c2 : c1.
// REC8 format (template parameter; codes whose rows have at most 51 edges): the variable phase is bound by the LSU pipe, and
// 16 of the 32 lanes' 16 bytes are c2 / padding that only the first-minimum edge wants. An 8-byte record {bits(c1), signs <<
// 5 | code} is gathered instead (a half-warp per wavefront), code = 27 - dc + position of the first minimum (31: not in this
// record), sign of edge k = (word << (27 - dc + k)) & 0x80000000; c2 lives in a float array beside it, read by the check
// phase (coalesced) and by the one variable-phase edge per row whose code matches. 27 sign bits per record: rows of 28..51
// edges own two records (24 edges + the rest), like the rows of 33..64 edges of the 16-byte format.

// Per-frame parameters (shared memory, written by thread 0 when the CTA takes a frame).
struct FrameCtx {
    float lp, primary, secondary;
    int has_cls;
    const uint32_t *cls_punct, *cls_short;
    u64 *tally;
};

// One edge of a check node: gather L of the bit (OFF = byte offset of the total in shared memory), rebuild the bit-to-check
// message with the OLD record, update the running min1 / min2 / sign state. `rel` = old argmin - index of the first edge of
// the current block. Everything that only needs a sign bit is taken from a float subtraction (FMA pipe) instead of an
// integer one (ALU pipe, the busier of the two): for x neither NaN nor -0, the sign bit of (0 - x) is (x > 0).
//   * The magnitude clamp |clamp(x)| == min(|x|, thr) (:447-461) is applied to min1 / min2 after the loop instead of to
//     every edge: min over clamped values == clamped min, and the first minimum only changes among edges that all sit at
//     the clamp, where min2 == min1 and the position is irrelevant.
#define QK_CN_EDGE(J, OFF)                                                                                              \
    {                                                                                                                   \
        const float Lv = *reinterpret_cast<const float *>(smem + (OFF));                                                \
        const uint32_t mag = (rel == (J)) ? c2o : c1o;                                                                  \
        const float c2b = __uint_as_float(mag ^ (zs & 0x80000000u));                                                    \
        zs <<= 1;                                                                                                       \
        const float braw = Lv - c2b;        /* L - c2b (:447-461); first iteration: zero record leaves the LLR (:336-350) */ \
        const float nb = 0.f - braw;        /* m is never -0 (L never is), so the sign bit of 0 - m is exactly (m > 0) */ \
        zpos ^= __float_as_uint(0.f - Lv);                /* parity of L > 0: the hard decision is z = !(L > 0) (:414-422) */ \
        pacc ^= __float_as_uint(braw);                    /* parity of m < 0 (:383): zero is positive here (Q4) */       \
        own = __funnelshift_l(__float_as_uint(nb), own, 1);              /* (m > 0) ? +1 : -1 (:402): zero is negative (Q4) */ \
        lt = __funnelshift_l(__float_as_uint(fabsf(nb) - m1), lt, 1);    /* |m| < min1: a new first minimum */            \
        m2 = fminf(m2, fmaxf(fabsf(nb), m1));             /* == the if / else-if chain (:386-396) */                     \
        m1 = fminf(m1, fabsf(nb));                                                                                      \
    }

// A row of 33..64 edges owns two records (edges 0..31 and 32..dc-1) with the same c1 / c2; the record that does not hold
// the first minimum carries kNoArg in w, which no table entry matches (REC8: code 31).
constexpr uint32_t kNoArg = 0x100u;

template <int ALG, bool WIDE, bool REC8>
__device__ __forceinline__ bool onchip_cn_phase(const OnchipArgs &a, const FrameCtx *ctx, const unsigned char *smem, void *recv, float *c2a,
                                                const uint32_t *synw, float thr_b, int warp, int lane, int nwarps) {
    constexpr int kCap = REC8 ? 27 : 32, kSplit = REC8 ? 24 : 32;   // edges per record / in the first of two records
    uint4 *rec = reinterpret_cast<uint4 *>(recv);
    uint2 *rec8 = reinterpret_cast<uint2 *>(recv);
    bool unsat = false;
    for (int g = warp; g < a.n_groups_cn2; g += nwarps) {
        const int4 gi = __ldg(a.cn_g2 + g);
        const int dc_row = gi.y;                                  // degree of the group's rows (warp-uniform)
        const bool two = WIDE && dc_row > kCap;                   // two records per row
        const int dc = two ? kSplit : dc_row;                     // edges covered by the first record
        const bool valid = lane < gi.w;
        const int slot = valid ? gi.z + lane : a.rec_slots;       // padding lanes work on the scratch record
        uint32_t c1o, c2o, zs;                    // old magnitudes; sign of the old message on the current edge in bit 31 of zs
        int arg_old;
        if constexpr (REC8) {
            const uint2 ro = rec8[slot];
            c1o = ro.x;
            c2o = __float_as_uint(c2a[slot]);
            zs = (ro.y & ~31u) << (kCap - dc);
            arg_old = (int)(ro.y & 31u) - (kCap - dc);
        } else {
            const uint4 ro = rec[slot];
            c1o = ro.x;
            c2o = ro.y;
            zs = ro.z << (32 - dc);
            arg_old = (int)ro.w - (32 - dc);
        }
        const uint4 *cp = a.cnT2 + gi.x + lane;
        float m1 = FLT_MAX, m2 = FLT_MAX;
        uint32_t own = 0, pacc = 0, zpos = 0, lt = 0;
        int kb = 0;
#pragma unroll 2
        for (; kb + 4 <= dc; kb += 4) {
            const uint4 cw = __ldg(cp + (kb >> 2) * 32);
            const int rel = arg_old - kb;
            QK_CN_EDGE(0, cw.x)
            QK_CN_EDGE(1, cw.y)
            QK_CN_EDGE(2, cw.z)
            QK_CN_EDGE(3, cw.w)
        }
        if (kb < dc) {                            // warp-uniform tail of 1..3 edges
            const uint4 cw = __ldg(cp + (kb >> 2) * 32);
            const int rel = arg_old - kb, left = dc - kb;
            QK_CN_EDGE(0, cw.x)
            if (left > 1) QK_CN_EDGE(1, cw.y)
            if (left > 2) QK_CN_EDGE(2, cw.z)
        }
        // the last "new minimum" event is the first minimum: edge k sits in bit dc-1-k of lt (none: every |m| is FLT_MAX)
        int arg = lt ? dc - __ffs((int)lt) : 0;
        uint32_t own_first = own;
        int slot2 = 0;
        if constexpr (WIDE) {
            if (two) {                            // edges kSplit..dc_row-1: the row's second record (warp-uniform branch)
                const int dc2 = dc_row - kSplit;
                slot2 = valid ? slot + gi.w : a.rec_slots;
                int arg_old2;                     // the "not here" code gives a value no edge index reaches
                if constexpr (REC8) {
                    const uint32_t w2 = rec8[slot2].y;
                    zs = (w2 & ~31u) << (kCap - dc2);
                    arg_old2 = (int)(w2 & 31u) - (kCap - dc2);
                } else {
                    const uint4 ro2 = rec[slot2];
                    zs = ro2.z << (32 - dc2);
                    arg_old2 = (int)ro2.w - (32 - dc2);
                }
                own = 0;
                lt = 0;
                for (; kb + 4 <= dc_row; kb += 4) {
                    const uint4 cw = __ldg(cp + (kb >> 2) * 32);
                    const int rel = arg_old2 - (kb - kSplit);
                    QK_CN_EDGE(0, cw.x)
                    QK_CN_EDGE(1, cw.y)
                    QK_CN_EDGE(2, cw.z)
                    QK_CN_EDGE(3, cw.w)
                }
                if (kb < dc_row) {
                    const uint4 cw = __ldg(cp + (kb >> 2) * 32);
                    const int rel = arg_old2 - (kb - kSplit), left = dc_row - kb;
                    QK_CN_EDGE(0, cw.x)
                    if (left > 1) QK_CN_EDGE(1, cw.y)
                    if (left > 2) QK_CN_EDGE(2, cw.z)
                }
                if (lt) arg = kSplit + dc2 - __ffs((int)lt);
            }
        }
        m1 = fminf(m1, thr_b);                                    // threshold_matrix(bit_to_check), magnitudes (:447-461)
        m2 = fminf(m2, thr_b);
        const uint32_t syn = (synw[g] >> lane) & 1u;
        const bool viol = (((zpos >> 31) ^ (uint32_t)dc_row ^ syn) & 1u) != 0;   // parity of (L <= 0) = parity of dc + parity of (L > 0)
        unsat |= viol && valid;
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
        const uint32_t flip = rowneg - 1u;                     // bit = message negative: !(m > 0) xor row sign
        if constexpr (REC8) {
            const bool first = !two || arg < kSplit;
            rec8[slot] = make_uint2(__float_as_uint(c1), ((own_first ^ flip) << 5) | (first ? (uint32_t)(arg + kCap - dc) : 31u));
            c2a[slot] = c2;
            if constexpr (WIDE) {
                if (two) {
                    rec8[slot2] = make_uint2(__float_as_uint(c1), ((own ^ flip) << 5) | (first ? 31u : (uint32_t)(arg - kSplit + kCap - (dc_row - kSplit))));
                    c2a[slot2] = c2;
                }
            }
        } else {
            uint4 rn;
            rn.x = __float_as_uint(c1);
            rn.y = __float_as_uint(c2);
            rn.z = own_first ^ flip;
            rn.w = (!two || arg < 32) ? (uint32_t)(arg + 32 - dc) : kNoArg;
            rec[slot] = rn;
            if constexpr (WIDE) {
                if (two) {
                    rn.z = own ^ flip;
                    rn.w = (arg >= 32) ? (uint32_t)(arg - 32 + 32 - (dc_row - 32)) : kNoArg;
                    rec[slot2] = rn;
                }
            }
        }
    }
    return unsat;
}
#undef QK_CN_EDGE

// a-priori LLR of the bit with key bit `bob` (qkd_ldpc_algorithm.cpp:1043-1049); `0 - lp` instead of `-lp`: never -0 (QBER 0.5
// gives lp = 0), which the sign tricks of the check phase rely on -- the decision (L <= 0) is the reference's either way.
// `pos` indexes the punctured / shortened masks (natural order for the sum-product kernel, slot order for min-sum).
__device__ __forceinline__ float onchip_llr_of(const FrameCtx *ctx, uint32_t bob, uint32_t pos, float lp) {
    float v = bob ? 0.f - lp : lp;
    if (ctx->has_cls) {
        const uint32_t w = pos >> 5, s = pos & 31u;
        if ((__ldg(ctx->cls_punct + w) >> s) & 1u) v = 1e-4f;  // punctured: ALMOST_ZERO (:1155)
        else if ((__ldg(ctx->cls_short + w) >> s) & 1u) v = FLT_MAX;   // shortened: largest finite value (:1164)
    }
    return v;
}
__device__ __forceinline__ float onchip_llr(const FrameCtx *ctx, const uint32_t *bobw, uint32_t bit, float lp) {
    return onchip_llr_of(ctx, (bobw[bit >> 5] >> (bit & 31u)) & 1u, bit, lp);
}

// The check-to-bit message addressed by table entry `ent` = (16 * record slot) << 5 | sh, added to the running sum.
#define QK_VN_EDGE(ENT)                                                                                                 \
    {                                                                                                                   \
        const uint4 r = *reinterpret_cast<const uint4 *>(recb + ((ENT) >> 5));                                          \
        const uint32_t mag = (((ENT) ^ r.w) & 0x1FFu) ? r.x : r.y;   /* r.w == kNoArg never matches */                   \
        acc = acc + __uint_as_float(mag ^ (__funnelshift_l(0u, r.z, (ENT)) & 0x80000000u));   /* r.z << sh */           \
    }

// The same for the 16-bit table entry E = sh << 11 | record slot (upper half of the register clear).
#define QK_VN_EDGE16(E)                                                                                                 \
    {                                                                                                                   \
        const uint4 r = *reinterpret_cast<const uint4 *>(recb + (((E) & 0x7FFu) << 4));                                 \
        const uint32_t sh = (E) >> 11;                                                                                  \
        const uint32_t mag = (sh == r.w) ? r.y : r.x;                /* r.w == kNoArg never matches */                   \
        acc = acc + __uint_as_float(mag ^ (__funnelshift_l(0u, r.z, sh) & 0x80000000u));                                \
    }

// REC8: `ent` = (8 * record slot) << 5 | sh, record {c1, signs << 5 | code}; the edge whose sh equals the code (the row's
// first minimum; code 31 matches none) fetches c2 from the float array beside the records (byte offset 4 * slot = ent >> 6).
#define QK_VN_EDGE_R8(ENT)                                                                                              \
    {                                                                                                                   \
        const uint2 r = *reinterpret_cast<const uint2 *>(recb + ((ENT) >> 5));                                          \
        uint32_t mag = r.x;                                                                                             \
        if ((((ENT) ^ r.y) & 31u) == 0) mag = *reinterpret_cast<const uint32_t *>(c2b + ((ENT) >> 6));                  \
        acc = acc + __uint_as_float(mag ^ (__funnelshift_l(0u, r.y, (ENT)) & 0x80000000u));   /* r.y << sh */           \
    }
#define QK_VN_EDGE16_R8(E)                                                                                              \
    {                                                                                                                   \
        const uint2 r = *reinterpret_cast<const uint2 *>(recb + (((E) & 0x7FFu) << 3));                                 \
        const uint32_t sh = (E) >> 11;                                                                                  \
        uint32_t mag = r.x;                                                                                             \
        if (sh == (r.y & 31u)) mag = *reinterpret_cast<const uint32_t *>(c2b + (((E) & 0x7FFu) << 2));                  \
        acc = acc + __uint_as_float(mag ^ (__funnelshift_l(0u, r.y, sh) & 0x80000000u));                                \
    }

// Index blocks of 4 entries: 16 bytes (VT16 = false) or 8 bytes per lane.
template <bool VT16, bool REC8>
struct VnBlock;
template <bool REC8>
struct VnBlock<false, REC8> {
    uint4 w;
    __device__ __forceinline__ void load(const void *tab, int idx) { w = __ldg(reinterpret_cast<const uint4 *>(tab) + idx); }
    template <int J>
    __device__ __forceinline__ float add(const unsigned char *recb, const unsigned char *c2b, float acc) const {
        const uint32_t ent = (J == 0) ? w.x : (J == 1) ? w.y : (J == 2) ? w.z : w.w;
        if constexpr (REC8) QK_VN_EDGE_R8(ent) else QK_VN_EDGE(ent)
        return acc;
    }
};
template <bool REC8>
struct VnBlock<true, REC8> {
    uint2 w;
    __device__ __forceinline__ void load(const void *tab, int idx) { w = __ldg(reinterpret_cast<const uint2 *>(tab) + idx); }
    template <int J>
    __device__ __forceinline__ float add(const unsigned char *recb, const unsigned char *c2b, float acc) const {
        const uint32_t ent = (J == 0) ? (w.x & 0xFFFFu) : (J == 1) ? (w.x >> 16) : (J == 2) ? (w.y & 0xFFFFu) : (w.y >> 16);
        if constexpr (REC8) QK_VN_EDGE16_R8(ent) else QK_VN_EDGE16(ent)
        return acc;
    }
};

// A group of 32 bits of degree DV <= 8, straight line: all index blocks first, then exactly DV messages.
template <int DV, bool VT16, bool REC8>
__device__ __forceinline__ float onchip_vn_fixed(const void *tab, int idx, const unsigned char *recb, const unsigned char *c2b, float acc) {
    VnBlock<VT16, REC8> b0, b1;
    b0.load(tab, idx);
    if constexpr (DV > 4) b1.load(tab, idx + 32);
    acc = b0.template add<0>(recb, c2b, acc);
    if constexpr (DV > 1) acc = b0.template add<1>(recb, c2b, acc);
    if constexpr (DV > 2) acc = b0.template add<2>(recb, c2b, acc);
    if constexpr (DV > 3) acc = b0.template add<3>(recb, c2b, acc);
    if constexpr (DV > 4) acc = b1.template add<0>(recb, c2b, acc);
    if constexpr (DV > 5) acc = b1.template add<1>(recb, c2b, acc);
    if constexpr (DV > 6) acc = b1.template add<2>(recb, c2b, acc);
    if constexpr (DV > 7) acc = b1.template add<3>(recb, c2b, acc);
    return acc;
}

template <bool VT16, bool REC8>
__device__ __forceinline__ void onchip_vn_phase(const OnchipArgs &a, const FrameCtx *ctx, float *L, const void *rec, const float *c2a,
                                                const uint32_t *bobs, float lp, int warp, int lane) {
    const unsigned char *recb = reinterpret_cast<const unsigned char *>(rec);
    const unsigned char *c2b = reinterpret_cast<const unsigned char *>(c2a);
    const void *tab = VT16 ? static_cast<const void *>(a.vT16) : static_cast<const void *>(a.vT2);
    const float nlp = 0.f - lp;
    const int has_cls = ctx->has_cls;
    // the host dealt the groups to the warps longest-first so that all warps of the CTA finish the phase together
    // (inst_onchip.cu); warp w owns a contiguous run of the dealt list
    const int g_end = __ldg(a.vn_start + warp + 1);
    for (int g = __ldg(a.vn_start + warp); g < g_end; ++g) {
        const int4 gi = __ldg(a.vn_g2 + g);
        const int dv = gi.y;
        const uint32_t s = (uint32_t)gi.z + (uint32_t)lane;       // the group's 32 slots start on a multiple of 32
        // a-priori LLR (qkd_ldpc_algorithm.cpp:1043-1049, onchip_llr_of); padding lanes see Bob bit 0
        float acc = ((bobs[gi.z >> 5] >> lane) & 1u) ? nlp : lp;
        if (has_cls) acc = onchip_llr_of(ctx, (bobs[gi.z >> 5] >> lane) & 1u, s, lp);
        const int idx = gi.x + lane;
        // ascending check order, starting from the LLR (std::accumulate, :414-417)
        switch (dv) {
            case 1: acc = onchip_vn_fixed<1, VT16, REC8>(tab, idx, recb, c2b, acc); break;
            case 2: acc = onchip_vn_fixed<2, VT16, REC8>(tab, idx, recb, c2b, acc); break;
            case 3: acc = onchip_vn_fixed<3, VT16, REC8>(tab, idx, recb, c2b, acc); break;
            case 4: acc = onchip_vn_fixed<4, VT16, REC8>(tab, idx, recb, c2b, acc); break;
            case 5: acc = onchip_vn_fixed<5, VT16, REC8>(tab, idx, recb, c2b, acc); break;
            case 6: acc = onchip_vn_fixed<6, VT16, REC8>(tab, idx, recb, c2b, acc); break;
            case 7: acc = onchip_vn_fixed<7, VT16, REC8>(tab, idx, recb, c2b, acc); break;
            case 8: acc = onchip_vn_fixed<8, VT16, REC8>(tab, idx, recb, c2b, acc); break;
            default: {
                int kb = 0;
#pragma unroll 2
                for (; kb + 4 <= dv; kb += 4) {
                    VnBlock<VT16, REC8> b;
                    b.load(tab, idx + (kb >> 2) * 32);
                    acc = b.template add<0>(recb, c2b, acc);
                    acc = b.template add<1>(recb, c2b, acc);
                    acc = b.template add<2>(recb, c2b, acc);
                    acc = b.template add<3>(recb, c2b, acc);
                }
                if (kb < dv) {                    // warp-uniform tail of 1..3 checks
                    VnBlock<VT16, REC8> b;
                    b.load(tab, idx + (kb >> 2) * 32);
                    const int left = dv - kb;
                    acc = b.template add<0>(recb, c2b, acc);
                    if (left > 1) acc = b.template add<1>(recb, c2b, acc);
                    if (left > 2) acc = b.template add<2>(recb, c2b, acc);
                }
            }
        }
        L[s] = acc;                               // consecutive slots: coalesced; padding lanes own padding slots
    }
}
#undef QK_VN_EDGE
#undef QK_VN_EDGE16
#undef QK_VN_EDGE_R8
#undef QK_VN_EDGE16_R8

template <int ALG, bool WIDE, bool VT16, bool REC8>
__global__ void __launch_bounds__(768, 2) onchip_minsum_kernel(const OnchipArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kRecBytes = REC8 ? 12 : 16;     // per record slot: uint4, or uint2 + a float of the c2 array behind the records
    float *L = reinterpret_cast<float *>(smem_raw);
    unsigned char *rec = smem_raw + onchip_l_bytes(a.l_slots);
    float *c2a = reinterpret_cast<float *>(rec + ((size_t)a.rec_slots + 2) * 8);   // REC8 only
    uint32_t *bobs = reinterpret_cast<uint32_t *>(rec + ((size_t)a.rec_slots + 2) * kRecBytes);
    uint32_t *synw = bobs + a.l_slots / 32;
    long long *s_frame = reinterpret_cast<long long *>(smem_raw + onchip_misc_offset(a.l_slots, a.rec_slots, a.n_groups_cn2, kRecBytes));
    FrameCtx *ctx = reinterpret_cast<FrameCtx *>(s_frame + 1);
    long long *s_clk = reinterpret_cast<long long *>(reinterpret_cast<unsigned char *>(ctx) + 48);   // profiling: [0] check, [1] variable, [2] start
    if (a.phase_clk && threadIdx.x == 0) {
        s_clk[0] = s_clk[1] = 0;
        s_clk[2] = clock64();
    }
    // frame set-up only: the key words as they come from HBM and Alice's bits in slot order, inside the record array
    uint32_t *st_bob = reinterpret_cast<uint32_t *>(rec), *st_alice = st_bob + a.words, *alice_s = st_alice + a.words;   // [l_slots/32]

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
                ctx->cls_punct = a.cls_masks2 + combo * 2 * (a.l_slots / 32);   // slot order
                ctx->cls_short = ctx->cls_punct + a.l_slots / 32;
                ctx->tally = a.tally ? a.tally + combo * a.tally_len : nullptr;
            }
        }
        __syncthreads();
        const long long f = *s_frame;
        if (f >= a.n_frames) break;
        const float lp = ctx->lp;
        for (int w = tid; w < a.words; w += blockDim.x) {
            st_bob[w] = a.bob_bits[f * a.words + w];
            st_alice[w] = a.alice_bits[f * a.words + w];
        }
        __syncthreads();
        // key bits into slot order; L = a-priori LLR (qkd_ldpc_algorithm.cpp:1043-1049)
        for (int s0 = warp * 32; s0 < a.l_slots; s0 += nwarps * 32) {
            const int s = s0 + lane;
            const uint32_t sb = (uint32_t)__ldg(a.slot_bit + s);
            const bool v = sb != 0xFFFFu;             // else a padding slot: never gathered, holds a harmless finite value
            const uint32_t bit = v ? sb : 0u;
            const uint32_t bb = (st_bob[bit >> 5] >> (bit & 31u)) & 1u, ab = (st_alice[bit >> 5] >> (bit & 31u)) & 1u;
            const uint32_t wb = __ballot_sync(0xffffffffu, v && bb), wa = __ballot_sync(0xffffffffu, v && ab);
            if (lane == 0) {
                bobs[s0 >> 5] = wb;
                alice_s[s0 >> 5] = wa;
            }
            L[s] = v ? onchip_llr_of(ctx, bb, (uint32_t)s, lp) : 1.f;
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
            for (int kb = 0; kb < gi.y; kb += 4) {
                const uint4 cw = __ldg(cp + (kb >> 2) * 32);
                const int left = gi.y - kb;
                const uint32_t c0 = cw.x >> 2, c1 = cw.y >> 2, c2 = cw.z >> 2, c3 = cw.w >> 2;   // slots
                sy ^= alice_s[c0 >> 5] >> (c0 & 31u);
                if (left > 1) sy ^= alice_s[c1 >> 5] >> (c1 & 31u);
                if (left > 2) sy ^= alice_s[c2 >> 5] >> (c2 & 31u);
                if (left > 3) sy ^= alice_s[c3 >> 5] >> (c3 & 31u);
            }
            const uint32_t sw = __ballot_sync(0xffffffffu, (sy & 1u) != 0 && lane < gi.w);
            if (lane == 0) synw[g] = sw;
        }
        __syncthreads();
        if constexpr (REC8) {                                 // records + c2 array: zero over the staging words
            for (int i = tid; i < (a.rec_slots + 2) * 3; i += blockDim.x) reinterpret_cast<uint32_t *>(rec)[i] = 0u;
        } else {
            for (int i = tid; i < a.rec_slots + 2; i += blockDim.x) reinterpret_cast<uint4 *>(rec)[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();

        int iters = a.max_iter, run = a.max_iter;
        bool success = false;
        for (int it = 1;; ++it) {
            // the check-node pass of iteration `it`; at it = max_iter + 1 it only serves as the syndrome test of the
            // last hard decision (non-adaptive variants, :424-445)
            if (a.phase_clk && tid == 0) s_clk[0] -= clock64();
            const bool unsat = onchip_cn_phase<ALG, WIDE, REC8>(a, ctx, smem_raw, rec, c2a, synw, it == 1 ? inf : a.thr, warp, lane, nwarps);
            const bool any_unsat = __syncthreads_or(unsat) != 0;
            if (a.phase_clk && tid == 0) s_clk[0] += clock64();
            if (!kAdaptive) {
                if (it > 1 && !any_unsat) { success = true; iters = run = it - 1; break; }   // z of iteration it-1 (:439-445)
                if (it > a.max_iter) break;
            } else {
                if (!any_unsat) { success = true; iters = it; run = it - 1; break; }         // exit test before the VN step (:770-776)
            }
            if (a.phase_clk && tid == 0) s_clk[1] -= clock64();
            onchip_vn_phase<VT16, REC8>(a, ctx, L, rec, c2a, bobs, lp, warp, lane);
            __syncthreads();
            if (a.phase_clk && tid == 0) s_clk[1] += clock64();
            if (kAdaptive && it == a.max_iter) break;          // the decision of the last iteration is never tested (Q10)
        }

        // bob_solution = last hard decision (L <= 0), packed in natural bit order; keys compare (arrays_equal, :1087)
        uint32_t diff = 0;
        for (int w = warp; w < a.words; w += nwarps) {
            const int i = w * 32 + lane;
            const uint32_t s = i < a.n ? (uint32_t)__ldg(a.bit_slot + i) : 0u;
            const uint32_t word = __ballot_sync(0xffffffffu, i < a.n && L[s] <= 0.f);
            if (lane == 0) {
                if (a.out_bits) a.out_bits[f * a.words + w] = word;
                diff |= word ^ a.alice_bits[f * a.words + w];
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

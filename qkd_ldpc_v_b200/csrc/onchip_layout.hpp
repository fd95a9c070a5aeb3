// Storage layout and index tables of the float32 on-chip min-sum kernel (onchip_minsum.cuh), chosen on the host so that the
// kernel's shared-memory gathers are (nearly) free of bank conflicts. Pure host C++ (no CUDA types): onchip_tables.cu
// includes it for qkdldpc_code_create, tools/oc_layout_model.cpp and the CPU test-suite run it without a device.
//
// What the kernel is free to choose -- none of it changes a single floating-point operation of the reference:
//   * WHERE a bit's total L and a row's record live in shared memory. Totals are stored in the order the variable phase
//     produces them (slot = position of the bit in its 32-lane group: the store is coalesced and needs no index), records
//     in the order the check phase produces them (slot = position of the row in its group). The bank of a total is then
//     the bit's lane, the 16-byte bank group of a record is the row's lane mod 8 -- both are ours to pick.
//   * WHICH nodes of one degree share a warp, and in which lanes.
//   * The ORDER in which a check node walks its edges: min1 / min2 / sign parity do not depend on it, and among equal
//     minima "the first" receives min2 == min1 (qkd_ldpc_algorithm.cpp:386-408), so any order yields the same record
//     values. (The variable node has no such freedom: its sum runs in ascending check order, :414-417.)
// Bank model (B300_MICROARCH "LDS/STS"): a 4-byte gather costs max-over-banks(distinct words in the bank) wavefronts per
// warp; a 16-byte gather is served per quarter-warp, each costing max-over-bank-groups(distinct records in the group).
// Stages: (1) rows and bits sorted into degree classes; (2) variable phase: greedy packing of bits into octets (quarter-
// warps); (3) check phase: lanes inside an octet and octets inside a class are permuted (free for the variable phase) to
// level every check group's load per bank; (4) every check group's edges are scheduled step by step with a bipartite
// matching rows -> banks. The model cost of the result is kept in the tables (tests/test_onchip_layout.py holds it
// against bounds).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

namespace qkhost {

struct Oc2U2 { uint32_t x, y; };
struct Oc2U4 { uint32_t x, y, z, w; };
struct Oc2Group { int off, deg, base, cnt; };   // int4 on the device: table offset, degree, first slot, nodes in the group

// Record format the tables are built for. REC16 (float32 and float64 kernels): one 16-byte record per row of up to 32 edges
// (gathered per quarter-warp: 8 lanes on 8 bank groups), rows of 33..64 edges own two (32 + the rest). REC8 (float32 kernel
// only): 8-byte records {c1, signs << 5 | position code} gathered per half-warp (16 lanes on 16 bank pairs) -- half the
// variable phase's shared-memory traffic --, which leaves 27 sign bits: rows of 28..51 edges own two records (24 + the rest).
struct Oc2Params {
    int rec_cap = 32;     // edges one record can describe
    int split = 32;       // edges in the first record of a two-record row (a multiple of 4: index blocks do not straddle)
    int vn_w = 8;         // lanes served together by one record gather = bank classes of the record slots
    int rec_bytes = 16;
    static Oc2Params rec8() { return Oc2Params{27, 24, 16, 8}; }
};

struct Oc2Tables {
    bool ok = false;
    Oc2Params prm;
    int n = 0, m = 0, max_dc = 0;
    int l_slots = 0;       // slots of the totals: 32 per variable-phase group (the last group of a degree class is padded);
                           // slot l_slots itself holds +inf (padding edges of mixed-degree check groups)
    int rec_slots = 0;     // records in use (m + rows wider than 32 edges); slot rec_slots is scratch for the check phase's
                           // padding lanes, slot rec_slots + 1 stays all-zero (the +0.0f message of padding table entries)
    std::vector<int> bit_slot;        // [n] slot of a bit's total
    std::vector<uint16_t> slot_bit;   // [l_slots] inverse; 0xFFFF = padding slot
    std::vector<int> row_slot;        // [m] record of edges at positions 0..31
    std::vector<int> row_slot2;       // [m] record of positions 32..dc-1, or -1
    std::vector<int> row_gdeg;        // [m] degree of the check group the row sits in (>= its own: the leftovers of several
                                      // degree classes share groups; the missing edges gather the +inf total in slot l_slots)
    std::vector<int> edge_pos;        // [nnz] position of CSR edge e in its row's processing order
    std::vector<Oc2Group> cn_g, vn_g; // vn_g in canonical (slot) order; the launcher deals it to the warps
    std::vector<int> vn_gcost;        // [vn_g] shared-memory + index wavefronts of the group in the bank model: the weight the
                                      // launcher balances the warps of a CTA with (the variable phase is bound by the LSU pipe)
    std::vector<Oc2U4> cnT;           // [off + kb*32 + lane] 4 x u32: BYTE offset (slot*4) of the totals of edges 4kb..4kb+3 -- 32-bit
                                      // entries although 16 would do: the check phase is bound by the ALU pipe, not by loads, and
                                      // an entry that IS the address costs no unpacking
    std::vector<Oc2U4> vT;            // [off + kb*32 + lane] 4 x u32: (rec_bytes * record slot) << 5 | sh, sh = rec_cap - edges in the record + position
    bool vt16_ok = false;             // at most 2048 records: the same table in 16-bit entries sh << 11 | record slot -- the
    std::vector<Oc2U2> vT16;          // variable phase is bound by the LSU pipe, so it unpacks rather than loads twice the bytes
    // bank model, wavefronts per decoder iteration
    long long cn_gather = 0, cn_gather_min = 0, vn_gather = 0, vn_gather_min = 0;
};

namespace oc2 {

struct Rng {   // xorshift64*: layout search only, no relation to the trial generator
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint32_t next() {
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        return (uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32);
    }
    uint32_t below(uint32_t k) { return (uint32_t)(((uint64_t)next() * k) >> 32); }
};

// wavefronts of one record gather by `cnt` = vn_w lanes (a quarter-warp of 16-byte records, a half-warp of 8-byte ones): per
// bank class (slot mod w) the number of DISTINCT slots, maximum over the classes
inline int octet_cost(const int *slots, int cnt, int w = 8) {
    int best = 0;
    for (int i = 0; i < cnt; ++i) {
        if (slots[i] < 0) continue;
        int c = 0;
        for (int j = 0; j < cnt; ++j) {
            if (slots[j] < 0 || (slots[j] & (w - 1)) != (slots[i] & (w - 1))) continue;
            bool dup = false;
            for (int t = 0; t < j; ++t) dup |= slots[t] == slots[j];
            c += !dup;
        }
        best = std::max(best, c);
    }
    return best;
}

// wavefronts of one warp-wide 4-byte gather
inline int warp_cost(const int *slots, int cnt) {
    int per_bank[32] = {0}, best = 0;
    for (int i = 0; i < cnt; ++i) {
        if (slots[i] < 0) continue;
        bool dup = false;
        for (int t = 0; t < i; ++t) dup |= slots[t] == slots[i];
        if (!dup) best = std::max(best, ++per_bank[slots[i] & 31]);
    }
    return best;
}

// Greedy packing of `members` (nodes of ONE degree d) into sets of `width`: start a set with the first free node, add the
// node whose k-th neighbours collide least, step by step, with the bank classes (`cls_of[neighbour] % ncol`) already in the set.
inline void pack(const std::vector<int> &members, const int *ptr, const int *nbr, const int *cls_of, int ncol, int width,
                 std::vector<int> &order) {
    const int d = ptr[members[0] + 1] - ptr[members[0]];
    const size_t cnt = members.size();
    std::vector<unsigned char> col(cnt * d);
    for (size_t i = 0; i < cnt; ++i)
        for (int k = 0; k < d; ++k) col[i * d + k] = (unsigned char)(cls_of[nbr[ptr[members[i]] + k]] % ncol);
    std::vector<char> used(cnt, 0);
    std::vector<int> load((size_t)d * ncol);
    size_t next_free = 0, left = cnt;
    order.clear();
    while (left > 0) {
        while (used[next_free]) ++next_free;
        size_t in_set = 1;
        used[next_free] = 1;
        --left;
        order.push_back(members[next_free]);
        std::fill(load.begin(), load.end(), 0);
        for (int k = 0; k < d; ++k) load[(size_t)k * ncol + col[next_free * d + k]]++;
        while (in_set < (size_t)width && left > 0) {
            size_t best = cnt;
            int best_cost = 1 << 30;
            for (size_t i = next_free + 1; i < cnt; ++i) {
                if (used[i]) continue;
                int cost = 0;
                for (int k = 0; k < d; ++k) cost += load[(size_t)k * ncol + col[i * d + k]];
                if (cost < best_cost) {
                    best_cost = cost;
                    best = i;
                    if (cost == 0) break;
                }
            }
            used[best] = 1;
            --left;
            ++in_set;
            order.push_back(members[best]);
            for (int k = 0; k < d; ++k) load[(size_t)k * ncol + col[best * d + k]]++;
        }
    }
}

}  // namespace oc2

// effort: 0 = greedy stages only (no bank levelling, no matching), 1 = default, larger = more levelling sweeps
inline void build_oc2_layout(int n, int m, long long nnz, const int *rp, const int *col_idx, const int *col_ptr, const int *csc_edge,
                             const int *csc_row, int effort, Oc2Tables &T, const Oc2Params prm = Oc2Params()) {
    using namespace oc2;
    T = Oc2Tables();
    T.prm = prm;
    T.n = n;
    T.m = m;
    const int W = prm.vn_w, CAP = prm.rec_cap, SPLIT = prm.split;
    for (int j = 0; j < m; ++j) T.max_dc = std::max(T.max_dc, rp[j + 1] - rp[j]);
    int wide_rows = 0;
    for (int j = 0; j < m; ++j) wide_rows += (rp[j + 1] - rp[j]) > CAP;
    // byte offsets of the totals are 16-bit table entries; record addresses (<< 5) must fit 32 bits with room to spare
    T.ok = T.max_dc <= SPLIT + CAP && (long long)m + wide_rows + 2 < (1 << 20);
    if (!T.ok) return;
    T.rec_slots = m + wide_rows;

    // ---- stage 1: degree classes (widest first), natural order inside
    auto classes_of = [](int count, const int *ptr) {
        std::vector<int> order(count);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return ptr[x + 1] - ptr[x] > ptr[y + 1] - ptr[y]; });
        std::vector<std::vector<int>> cls;
        for (int i = 0; i < count; ++i) {
            if (i == 0 || ptr[order[i] + 1] - ptr[order[i]] != ptr[order[i - 1] + 1] - ptr[order[i - 1]]) cls.emplace_back();
            cls.back().push_back(order[i]);
        }
        return cls;
    };
    std::vector<std::vector<int>> rcls = classes_of(m, rp), bcls = classes_of(n, col_ptr);

    // Check groups: 32 rows of ONE degree, in class order. The leftover rows of the classes (fewer than 32 each) are merged,
    // widest first, into groups of mixed degree -- a row then walks `group degree` edges, the missing ones gathering a total
    // that is +inf (min1 / min2, the sign parity and the first-minimum position ignore it; it counts as a positive total on
    // both sides of the syndrome parity) -- as long as no row is padded by more than 3 edges. I80: 67 -> 64 groups, which
    // 16 warps finish in 4 rounds instead of 5. Rows of two records keep pure groups.
    struct CnGroup { int deg; std::vector<int> rows; };
    std::vector<CnGroup> cgroups;
    {
        std::vector<int> left;
        for (const auto &cls : rcls) {
            const int dc = rp[cls[0] + 1] - rp[cls[0]];
            size_t g0 = 0;
            for (; g0 + 32 <= cls.size(); g0 += 32) cgroups.push_back(CnGroup{dc, std::vector<int>(cls.begin() + g0, cls.begin() + g0 + 32)});
            if (g0 < cls.size()) {
                if (dc > CAP) cgroups.push_back(CnGroup{dc, std::vector<int>(cls.begin() + g0, cls.end())});
                else left.insert(left.end(), cls.begin() + g0, cls.end());   // classes come widest first
            }
        }
        for (size_t i = 0; i < left.size();) {
            CnGroup g{rp[left[i] + 1] - rp[left[i]], {}};
            while (i < left.size() && g.rows.size() < 32 && rp[left[i] + 1] - rp[left[i]] >= g.deg - 3) g.rows.push_back(left[i++]);   // g.deg <= CAP
            cgroups.push_back(std::move(g));
        }
    }
    // record slots are handed out group by group: slot = first slot of the group + lane
    T.row_slot.assign(m, 0);
    T.row_slot2.assign(m, -1);
    T.row_gdeg.assign(m, 0);
    {
        int next = 0;
        for (const CnGroup &g : cgroups) {
            const int cnt = (int)g.rows.size();
            for (int l = 0; l < cnt; ++l) {
                T.row_slot[g.rows[l]] = next + l;
                if (g.deg > CAP) T.row_slot2[g.rows[l]] = next + cnt + l;
                T.row_gdeg[g.rows[l]] = g.deg;
            }
            next += g.deg > CAP ? 2 * cnt : cnt;
        }
    }

    // ---- stage 2: variable phase. Octets = W consecutive positions of a bit class (8 for 16-byte records, 16 for 8-byte ones);
    // cell (octet, k) gathers W records. (The second record of a two-record row sits a whole group -- 32 slots for all but
    // the last group of a class -- behind the first, i.e. in the same bank class.)
    // greedy start
    for (auto &cls : bcls) {
        if ((int)cls.size() <= W) continue;
        std::vector<int> order;
        pack(cls, col_ptr, csc_row, T.row_slot.data(), W, W, order);
        cls = order;
    }
    // (A local search over octet membership and record slots -- swaps of two bits / two rows, wavefront count or pair
    // collisions as objective, greedy or annealed -- was tried on top of the greedy packing and gained under 2 %: a bit of
    // degree 17 or 26 meets a different octet of rows at every step, and 8 lanes on 8 bank groups collide like a birthday
    // problem whatever the colouring. The variable phase keeps the greedy result.)

    // ---- stage 3: slots of the totals = class order; lanes inside an octet / octets inside a class are permuted to level,
    // for every check group, the number of its edges per bank. Check groups: 32 consecutive positions of a row class.
    std::vector<int> row_group(m);
    const int n_cn_groups = (int)cgroups.size();
    for (int g = 0; g < n_cn_groups; ++g)
        for (int r : cgroups[g].rows) row_group[r] = g;
    std::vector<int> class_base(bcls.size() + 1, 0);   // every class starts on a multiple of 32: slot & 31 == lane == bank
    for (size_t c = 0; c < bcls.size(); ++c) class_base[c + 1] = class_base[c] + (int)((bcls[c].size() + 31) / 32 * 32);
    T.l_slots = class_base.back();
    if (T.l_slots > 65535) {   // slot <-> bit maps are 16-bit
        T.ok = false;
        return;
    }
    T.bit_slot.assign(n, 0);
    auto refresh_bit_slots = [&]() {
        for (size_t c = 0; c < bcls.size(); ++c)
            for (size_t i = 0; i < bcls[c].size(); ++i) T.bit_slot[bcls[c][i]] = class_base[c] + (int)i;
    };
    refresh_bit_slots();
    if (effort > 0) {
        std::vector<int> load((size_t)n_cn_groups * 32, 0);
        for (int b = 0; b < n; ++b)
            for (int p = col_ptr[b]; p < col_ptr[b + 1]; ++p) load[(size_t)row_group[csc_row[p]] * 32 + (T.bit_slot[b] & 31)]++;
        // change of sum(load^2) when bit b moves from bank `from` to bank `to`
        auto delta_move = [&](int b, int from, int to) {
            long long d = 0;
            for (int p = col_ptr[b]; p < col_ptr[b + 1]; ++p) {
                const size_t g = (size_t)row_group[csc_row[p]] * 32;
                d += 2 * (load[g + to] - load[g + from]) + 2;
            }
            return d;
        };
        auto apply_move = [&](int b, int from, int to) {
            for (int p = col_ptr[b]; p < col_ptr[b + 1]; ++p) {
                const size_t g = (size_t)row_group[csc_row[p]] * 32;
                load[g + from]--;
                load[g + to]++;
            }
        };
        for (int sweep = 0; sweep < 4 * effort; ++sweep) {
            long long gained = 0;
            for (size_t c = 0; c < bcls.size(); ++c) {
                auto &v = bcls[c];
                const int base = class_base[c];
                // lanes inside an octet
                for (size_t o = 0; o < v.size(); o += W) {
                    const size_t end = std::min(v.size(), o + W);
                    for (size_t i = o; i < end; ++i)
                        for (size_t j = i + 1; j < end; ++j) {
                            const int bi = (base + (int)i) & 31, bj = (base + (int)j) & 31;
                            // the two moves touch different banks, but may share check groups: evaluate sequentially
                            const long long d1 = delta_move(v[i], bi, bj);
                            apply_move(v[i], bi, bj);
                            const long long d2 = delta_move(v[j], bj, bi);
                            if (d1 + d2 < 0) {
                                apply_move(v[j], bj, bi);
                                std::swap(v[i], v[j]);
                                gained -= d1 + d2;
                            } else {
                                apply_move(v[i], bj, bi);
                            }
                        }
                }
                // whole octets at positions with a different quarter (bank offset differs by a multiple of W)
                const size_t full = v.size() / W;
                for (size_t a = 0; a + 1 < full; ++a)
                    for (size_t t = 0; t < 6; ++t) {
                        const size_t b2 = a + 1 + (size_t)((a * 7 + t * 13 + (size_t)sweep * 5) % (full - a - 1));
                        if (((a ^ b2) & (size_t)(32 / W - 1)) == 0) continue;
                        long long d = 0;
                        for (int l = 0; l < W; ++l) {
                            const int ba = (base + (int)(a * W) + l) & 31, bb = (base + (int)(b2 * W) + l) & 31;
                            d += delta_move(v[a * W + l], ba, bb);
                            apply_move(v[a * W + l], ba, bb);
                            d += delta_move(v[b2 * W + l], bb, ba);
                            apply_move(v[b2 * W + l], bb, ba);
                        }
                        if (d < 0) {
                            for (int l = 0; l < W; ++l) std::swap(v[a * W + l], v[b2 * W + l]);
                            gained -= d;
                        } else {
                            for (int l = 0; l < W; ++l) {
                                const int ba = (base + (int)(a * W) + l) & 31, bb = (base + (int)(b2 * W) + l) & 31;
                                apply_move(v[a * W + l], bb, ba);
                                apply_move(v[b2 * W + l], ba, bb);
                            }
                        }
                    }
            }
            if (gained == 0) break;
        }
        refresh_bit_slots();
    }
    T.slot_bit.assign(T.l_slots, (uint16_t)0xFFFF);
    for (int b = 0; b < n; ++b) T.slot_bit[T.bit_slot[b]] = (uint16_t)b;

    // ---- stage 4: edge order of every check group. Step k: maximum bipartite matching rows -> banks over the edges not
    // yet placed, heaviest remaining banks first (a bank holding more edges than steps are left must be hit every step);
    // a row left unmatched takes the edge whose bank is least used in this step.
    T.edge_pos.assign(nnz, -1);
    T.cn_gather = T.cn_gather_min = 0;
    {
        for (const CnGroup &cg : cgroups) {
            const int dc = cg.deg, blocks = (dc + 3) / 4;
            const int cnt = (int)cg.rows.size();
            const int base = T.row_slot[cg.rows[0]];
            Oc2Group gi{(int)T.cnT.size(), dc, base, cnt};
            T.cn_g.push_back(gi);
            // remaining edges per row, bank of each
            std::vector<std::vector<int>> rem(cnt);
            int bank_left[32] = {0};
            for (int l = 0; l < cnt; ++l)
                for (int e = rp[cg.rows[l]]; e < rp[cg.rows[l] + 1]; ++e) {
                    rem[l].push_back(e);
                    bank_left[T.bit_slot[col_idx[e]] & 31]++;
                }
            std::vector<int> sched((size_t)cnt * dc, -1);   // [l*dc + k] = CSR edge, -1 = past the row's own degree
            for (int k = 0; k < dc; ++k) {
                int row_edge[32], bank_row[32];
                std::fill(row_edge, row_edge + 32, -1);
                std::fill(bank_row, bank_row + 32, -1);
                if (effort > 0) {
                    int banks[32];
                    std::iota(banks, banks + 32, 0);
                    std::stable_sort(banks, banks + 32, [&](int x, int y) { return bank_left[x] > bank_left[y]; });
                    // Kuhn's augmenting paths from the bank side
                    for (int bi = 0; bi < 32; ++bi) {
                        const int b0 = banks[bi];
                        if (bank_left[b0] == 0) break;
                        char seen[32] = {0};
                        // iterative DFS with explicit stack would be overkill for 32 x 32: recursion via lambda
                        struct Aug {
                            const std::vector<std::vector<int>> &rem;
                            const std::vector<int> &slot;
                            const int *col;
                            int *row_edge, *bank_row;
                            char *seen;
                            int cnt;
                            bool run(int bank) {
                                for (int l = 0; l < cnt; ++l) {
                                    if (seen[l]) continue;
                                    int edge = -1;
                                    for (int e : rem[l])
                                        if ((slot[col[e]] & 31) == bank) { edge = e; break; }
                                    if (edge < 0) continue;
                                    seen[l] = 1;
                                    const int old = row_edge[l];
                                    if (old < 0 || run_from_row(l, old)) {
                                        row_edge[l] = edge;
                                        bank_row[bank] = l;
                                        return true;
                                    }
                                }
                                return false;
                            }
                            // row l currently holds edge `old` (bank ob): move that bank to another row
                            bool run_from_row(int l, int old) {
                                const int ob = slot[col[old]] & 31;
                                (void)l;
                                return run(ob);
                            }
                        } aug{rem, T.bit_slot, col_idx, row_edge, bank_row, seen, cnt};
                        aug.run(b0);
                    }
                }
                int used[32] = {0};
                for (int l = 0; l < cnt; ++l)
                    if (row_edge[l] >= 0) used[T.bit_slot[col_idx[row_edge[l]]] & 31]++;
                for (int l = 0; l < cnt; ++l) {
                    if (row_edge[l] >= 0 || rem[l].empty()) continue;   // a row of a lower degree has run out of edges
                    int best = -1, best_use = 1 << 30;
                    for (int e : rem[l]) {
                        const int u = used[T.bit_slot[col_idx[e]] & 31];
                        if (u < best_use) { best_use = u; best = e; }
                    }
                    row_edge[l] = best;
                    used[T.bit_slot[col_idx[best]] & 31]++;
                }
                int slots[32];
                for (int l = 0; l < cnt; ++l) {
                    const int e = row_edge[l];
                    sched[(size_t)l * dc + k] = e;
                    slots[l] = e >= 0 ? T.bit_slot[col_idx[e]] : T.l_slots;   // the +inf total: one address, a broadcast
                    if (e < 0) continue;
                    T.edge_pos[e] = k;
                    rem[l].erase(std::find(rem[l].begin(), rem[l].end(), e));
                    bank_left[T.bit_slot[col_idx[e]] & 31]--;
                }
                T.cn_gather += warp_cost(slots, cnt);
                T.cn_gather_min += 1;
            }
            for (int kb = 0; kb < blocks; ++kb)
                for (int l = 0; l < 32; ++l) {
                    uint32_t c[4] = {0, 0, 0, 0};   // padding: the total in slot 0 (a broadcast)
                    if (l < cnt)
                        for (int j = 0; j < 4 && kb * 4 + j < dc; ++j) {
                            const int e = sched[(size_t)l * dc + kb * 4 + j];
                            c[j] = (uint32_t)(e >= 0 ? T.bit_slot[col_idx[e]] : T.l_slots) * 4u;
                        }
                    T.cnT.push_back(Oc2U4{c[0], c[1], c[2], c[3]});
                }
        }
    }

    // ---- variable-phase tables (canonical order: class by class, 32 consecutive slots per group) and their model cost
    T.vn_gather = T.vn_gather_min = 0;
    T.vt16_ok = T.rec_slots <= 2048;
    for (size_t c = 0; c < bcls.size(); ++c) {
        const auto &v = bcls[c];
        const int dv = col_ptr[v[0] + 1] - col_ptr[v[0]], blocks = (dv + 3) / 4;
        for (size_t g0 = 0; g0 < v.size(); g0 += 32) {
            const int cnt = (int)std::min<size_t>(32, v.size() - g0);
            T.vn_g.push_back(Oc2Group{(int)T.vT.size(), dv, class_base[c] + (int)g0, cnt});
            for (int kb = 0; kb < blocks; ++kb)
                for (int l = 0; l < 32; ++l) {
                    uint32_t e[4];
                    for (int j = 0; j < 4; ++j) {
                        e[j] = ((uint32_t)(T.rec_slots + 1) * (uint32_t)prm.rec_bytes) << 5;   // padding: the all-zero record (adds +0.0f)
                        const int k = kb * 4 + j;
                        if (l < cnt && k < dv) {
                            const int p = col_ptr[v[g0 + l]] + k, r = csc_row[p], pos = T.edge_pos[csc_edge[p]], dcr = T.row_gdeg[r];
                            const bool two = dcr > CAP;
                            const int half = two && pos >= SPLIT ? 1 : 0, in_rec = !two ? dcr : (half == 0 ? SPLIT : dcr - SPLIT);
                            const int slot = half == 0 ? T.row_slot[r] : T.row_slot2[r];
                            e[j] = (((uint32_t)slot * (uint32_t)prm.rec_bytes) << 5) | (uint32_t)(CAP - in_rec + pos - half * SPLIT);
                        }
                    }
                    T.vT.push_back(Oc2U4{e[0], e[1], e[2], e[3]});
                    if (T.vt16_ok) {   // entries past the degree and padding lanes: record 0 (the kernel reads exactly `deg` entries)
                        uint32_t h[4];
                        for (int j = 0; j < 4; ++j) h[j] = (l < cnt && kb * 4 + j < dv) ? ((e[j] & 31u) << 11) | ((e[j] >> 5) / (uint32_t)prm.rec_bytes) : 0u;
                        T.vT16.push_back(Oc2U2{h[0] | (h[1] << 16), h[2] | (h[3] << 16)});
                    }
                }
            int gcost = 4 + 2 * blocks;   // header, Bob's bits, the store of the totals; index blocks
            for (int k = 0; k < dv; ++k)
                for (int q = 0; q < 32 / W; ++q) {
                    int slots[16], c8 = 0;
                    for (int l = q * W; l < q * W + W; ++l) {
                        const uint32_t ent = (&T.vT[(size_t)T.vn_g.back().off + (size_t)(k / 4) * 32 + l].x)[k % 4];
                        slots[c8++] = (int)((ent >> 5) / (uint32_t)prm.rec_bytes);
                    }
                    const int oc = octet_cost(slots, c8, W);
                    T.vn_gather += oc;
                    T.vn_gather_min += 1;
                    gcost += oc;
                }
            T.vn_gcost.push_back(gcost);
        }
    }
}

// Everything the kernel indexes shared memory with, checked against the graph (the kernel does no bounds tests). Returns
// nullptr when the tables are sound, else a description of the first fault.
inline const char *check_oc2_layout(int n, int m, long long nnz, const int *rp, const int *col_idx, const int *col_ptr, const int *csc_edge,
                                    const int *csc_row, const Oc2Tables &T) {
    if (!T.ok) return nullptr;
    const int CAP = T.prm.rec_cap, SPLIT = T.prm.split;
    const uint32_t RB = (uint32_t)T.prm.rec_bytes;
    if ((int)T.bit_slot.size() != n || (int)T.slot_bit.size() != T.l_slots || (int)T.row_slot.size() != m || n >= 0xFFFF) return "table sizes";
    std::vector<char> seen_slot(T.l_slots, 0), seen_rec(T.rec_slots, 0);
    for (int b = 0; b < n; ++b) {
        const int s = T.bit_slot[b];
        if (s < 0 || s >= T.l_slots || seen_slot[s] || T.slot_bit[s] != b) return "bit -> slot is not injective";
        seen_slot[s] = 1;
    }
    for (int s = 0; s < T.l_slots; ++s)
        if (!seen_slot[s] && T.slot_bit[s] != 0xFFFF) return "padding slot not marked";
    for (int j = 0; j < m; ++j) {
        const int dc = rp[j + 1] - rp[j];
        if ((dc > CAP) != (T.row_slot2[j] >= 0)) return "second record of a row";
        for (int h = 0; h < (dc > CAP ? 2 : 1); ++h) {
            const int s = h ? T.row_slot2[j] : T.row_slot[j];
            if (s < 0 || s >= T.rec_slots || seen_rec[s]) return "row -> record slot is not injective";
            seen_rec[s] = 1;
        }
        if ((int)T.row_gdeg.size() != m || T.row_gdeg[j] < dc) return "group degree of a row";
        std::vector<char> pos_seen(dc, 0);
        for (int e = rp[j]; e < rp[j + 1]; ++e) {
            const int p = T.edge_pos[e];
            if (p < 0 || p >= dc || pos_seen[p]) return "edge order of a row is not a permutation";   // own edges come first
            pos_seen[p] = 1;
        }
    }
    // check-phase table: group g, lane l = the row whose first record is base + l; entry k = total of the edge at position k
    std::vector<int> row_of_slot(T.rec_slots, -1);
    for (int j = 0; j < m; ++j) row_of_slot[T.row_slot[j]] = j;
    long long edges_cn = 0, edges_vn = 0;
    for (const Oc2Group &g : T.cn_g) {
        const int blocks = (g.deg + 3) / 4;
        if (g.deg < 1 || g.deg > SPLIT + CAP || g.cnt < 1 || g.cnt > 32 || g.off < 0 || (size_t)g.off + (size_t)blocks * 32 > T.cnT.size()) return "check group header";
        for (int l = 0; l < g.cnt; ++l) {
            const int j = g.base + l < T.rec_slots ? row_of_slot[g.base + l] : -1;
            if (j < 0 || T.row_gdeg[j] != g.deg || rp[j + 1] - rp[j] > g.deg || rp[j + 1] - rp[j] < g.deg - 3 ||
                (g.deg > CAP && rp[j + 1] - rp[j] != g.deg))
                return "check group: record slot without a row of (nearly) the group's degree";
            if (g.deg > CAP && T.row_slot2[j] != g.base + g.cnt + l) return "check group: second record not at base + cnt + lane";
            std::vector<int> by_pos(g.deg, -1);
            for (int e = rp[j]; e < rp[j + 1]; ++e) by_pos[T.edge_pos[e]] = e;
            for (int k = 0; k < g.deg; ++k) {
                const uint32_t off = (&T.cnT[(size_t)g.off + (size_t)(k / 4) * 32 + l].x)[k % 4];
                if (by_pos[k] < 0) {   // past the row's own degree: the +inf total
                    if (k < rp[j + 1] - rp[j] || off != (uint32_t)T.l_slots * 4u) return "check table: padding edge of a mixed group";
                    continue;
                }
                if (off != (uint32_t)T.bit_slot[col_idx[by_pos[k]]] * 4u) return "check table: wrong total for an edge";
                ++edges_cn;
            }
        }
        for (int kb = 0; kb < blocks; ++kb)
            for (int l = 0; l < 32; ++l) {
                const Oc2U4 w = T.cnT[(size_t)g.off + (size_t)kb * 32 + l];
                for (uint32_t off : {w.x, w.y, w.z, w.w})
                    if ((off & 3u) || off / 4 > (uint32_t)T.l_slots || (off / 4 < (uint32_t)T.l_slots && T.slot_bit[off / 4] == 0xFFFF))
                        return "check table: offset out of range";
            }
    }
    int next_slot = 0;
    for (const Oc2Group &g : T.vn_g) {
        const int blocks = (g.deg + 3) / 4;
        if (g.deg < 1 || g.cnt < 1 || g.cnt > 32 || g.base != next_slot || g.off < 0 || (size_t)g.off + (size_t)blocks * 32 > T.vT.size()) return "variable group header";
        next_slot += 32;
        for (int l = 0; l < 32; ++l)
            for (int k = 0; k < blocks * 4; ++k) {
                const uint32_t ent = (&T.vT[(size_t)g.off + (size_t)(k / 4) * 32 + l].x)[k % 4];
                const int slot = (int)((ent >> 5) / RB), sh = (int)(ent & 31u);
                if ((ent >> 5) % RB || slot > T.rec_slots + 1 || sh >= CAP) return "variable table: bad entry";
                if (l < g.cnt && k < g.deg) {
                    const int b = T.slot_bit[g.base + l];
                    if (col_ptr[b + 1] - col_ptr[b] != g.deg) return "variable group: bit of another degree";
                    const int p = col_ptr[b] + k, r = csc_row[p], pos = T.edge_pos[csc_edge[p]], dcr = T.row_gdeg[r];
                    const bool two = dcr > CAP, second = two && pos >= SPLIT;
                    const int in_rec = !two ? dcr : (second ? dcr - SPLIT : SPLIT);
                    if (slot != (second ? T.row_slot2[r] : T.row_slot[r]) || sh != CAP - in_rec + pos - (second ? SPLIT : 0)) return "variable table: wrong record or shift";
                    if (T.vt16_ok) {
                        const Oc2U2 w16 = T.vT16[(size_t)g.off + (size_t)(k / 4) * 32 + l];
                        const uint32_t h = (k % 4 == 0) ? (w16.x & 0xFFFFu) : (k % 4 == 1) ? (w16.x >> 16) : (k % 4 == 2) ? (w16.y & 0xFFFFu) : (w16.y >> 16);
                        if ((int)(h & 0x7FFu) != slot || (int)(h >> 11) != sh) return "16-bit variable table disagrees with the 32-bit one";
                    }
                    ++edges_vn;
                } else if (slot != T.rec_slots + 1) {
                    return "variable table: padding entry does not point at the all-zero record";
                }
            }
    }
    if (next_slot != T.l_slots) return "variable groups do not cover the totals";
    if (T.vt16_ok && (T.vT16.size() != T.vT.size() || T.rec_slots > 2048)) return "16-bit variable table size";
    if (edges_cn != nnz || edges_vn != nnz) return "tables do not cover every edge once";
    return nullptr;
}

}  // namespace qkhost

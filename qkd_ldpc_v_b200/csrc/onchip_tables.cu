// Host-side tables of the two on-chip kernels (onchip_minsum.cuh, onchip_spa.cuh): degree classes, conflict-aware packing of
// rows / bits into 32-lane groups, 4-edge index blocks, the item list of the sum-product variable phase -- and the
// self-check of everything the kernels index shared memory with (the kernels themselves do no bounds tests). Plain host
// code: qkdldpc_code_create runs it before it touches the device, so the CPU test-suite exercises it on every golden code.
#include "handle.hpp"
#include "onchip_minsum.cuh"

namespace qkhost {

// n, m, nnz, row pointers, column indices (CSR, ascending) and the column view (col_ptr, csc_edge: k-th check of a bit ->
// CSR edge id, csc_row: ... -> check id). Fills `T`; returns QKDLDPC_OK or QKDLDPC_ERR_STATE (with qkdldpc_last_error set)
// when a table fails its self-check.
int build_onchip_tables(int n, int m, long long nnz, const std::vector<int> &rp, const int *col_idx, const std::vector<int> &col_ptr,
                        const std::vector<int> &csc_edge, const std::vector<int> &csc_row, OnchipTables &T) {
    const int *row_ptr = rp.data();
    int &max_dc = T.max_dc;
    std::vector<int> &slot0 = T.slot0;
    std::vector<int2> &oc_cn_ginfo = T.cn_ginfo;
    std::vector<uint16_t> &oc_cn_row = T.cn_row;
    std::vector<uint2> &oc_cnT = T.cnT;
    std::vector<int> &sp_cn_moff = T.sp_cn_moff, &sp_group_item0 = T.sp_group_item0;
    std::vector<uint4> &sp_items = T.sp_items;
    int &sp_msg_words = T.sp_msg_words;
    bool &sp_ok = T.sp_ok;
    // On-chip path layout (onchip_minsum.cuh): rows and bits are split into degree classes; inside a class the nodes are
    // packed into groups of 32 lanes by a greedy conflict-aware heuristic (below); index tables are stored per group in
    // blocks of 4 edges per lane.
    max_dc = 0;
    for (int j = 0; j < m; ++j) max_dc = std::max(max_dc, row_ptr[j + 1] - row_ptr[j]);
    // Record slots: a row of up to 32 edges owns one 16-byte record; a row of 33..64 edges owns two consecutive ones
    // (edges 0..31 and 32..dc-1), so that every record still carries 32 sign bits and the variable phase is unchanged.
    slot0.assign(m + 1, 0);
    for (int j = 0; j < m; ++j) slot0[j + 1] = slot0[j] + ((row_ptr[j + 1] - row_ptr[j]) > 32 ? 2 : 1);
    const int rec_slots = T.rec_slots = slot0[m];   // slots rec_slots and rec_slots + 1 are scratch (padding lanes)
    const bool oc_ok = T.oc_ok = n < 65535 && rec_slots + 2 < 65535 && max_dc <= 64;
    // sum-product on-chip path (onchip_spa.cuh): message word of edge k of the row at (group g, lane l) = moff[g] + 32 k + l
    std::vector<int> row_word0(m, 0), row_lane(m, 0);
    if (oc_ok) {
        // Shared-memory bank model: a warp-wide 4-byte gather (L[bit], check phase) is conflict-free when the 32 lanes hit
        // 32 different banks, i.e. bit index mod 32 all different; a 16-byte gather (row record, variable phase) is
        // served per quarter-warp, conflict-free when its 8 lanes hit 8 different 16-byte bank groups, i.e. row index
        // mod 8 all different. `pack` builds sets of `width` nodes of one degree so that, step by step (k-th neighbour of
        // every member), as few members as possible share a bank: greedy -- start from the first free node, add the
        // node whose neighbours collide least with the banks already used at each step. Measured effect on n=10240
        // codes: 9.9 -> 6.9 wavefronts per 32 record gathers (irregular R=0.8), 9.8 -> 4.5 (alist R=0.79).
        auto pack = [](const std::vector<int> &members, const std::vector<int> &ptr, const int *nbr, int ncol, int width,
                       std::vector<std::vector<int>> &out, const int *remap /* neighbour id -> storage slot, or null */) {
            const int d = ptr[members[0] + 1] - ptr[members[0]];
            std::vector<unsigned char> col((size_t)members.size() * d);
            for (size_t i = 0; i < members.size(); ++i)
                for (int k = 0; k < d; ++k) {
                    const int id = nbr[ptr[members[i]] + k];
                    col[i * d + k] = (unsigned char)((remap ? remap[id] : id) % ncol);
                }
            std::vector<char> used(members.size(), 0);
            std::vector<int> cnt((size_t)d * ncol);
            size_t next_free = 0, left = members.size();
            while (left > 0) {
                while (used[next_free]) ++next_free;
                std::vector<int> cur{members[next_free]};
                used[next_free] = 1;
                --left;
                std::fill(cnt.begin(), cnt.end(), 0);
                for (int k = 0; k < d; ++k) cnt[(size_t)k * ncol + col[next_free * d + k]]++;
                while ((int)cur.size() < width && left > 0) {
                    size_t best = members.size();
                    int best_cost = 1 << 30;
                    for (size_t i = next_free + 1; i < members.size(); ++i) {
                        if (used[i]) continue;
                        int cost = 0;
                        for (int k = 0; k < d; ++k) cost += cnt[(size_t)k * ncol + col[i * d + k]];
                        if (cost < best_cost) {
                            best_cost = cost;
                            best = i;
                            if (cost == 0) break;
                        }
                    }
                    used[best] = 1;
                    --left;
                    cur.push_back(members[best]);
                    for (int k = 0; k < d; ++k) cnt[(size_t)k * ncol + col[best * d + k]]++;
                }
                out.push_back(std::move(cur));
            }
        };
        auto degree_classes = [](int count, const std::vector<int> &ptr) {   // widest first
            std::vector<std::vector<int>> cls;
            std::vector<int> order(count);
            for (int i = 0; i < count; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return ptr[x + 1] - ptr[x] > ptr[y + 1] - ptr[y]; });
            for (int i = 0; i < count; ++i) {
                if (i == 0 || ptr[order[i] + 1] - ptr[order[i]] != ptr[order[i - 1] + 1] - ptr[order[i - 1]]) cls.emplace_back();
                cls.back().push_back(order[i]);
            }
            return cls;
        };
        // check phase: groups of 32 rows
        for (const auto &cls : degree_classes(m, rp)) {
            std::vector<std::vector<int>> groups;
            pack(cls, rp, col_idx, 32, 32, groups, nullptr);
            const int dc = rp[cls[0] + 1] - rp[cls[0]], blocks = (dc + 3) / 4;
            for (const auto &gr : groups) {
                oc_cn_ginfo.push_back(make_int2((int)oc_cnT.size(), dc));
                sp_cn_moff.push_back(sp_msg_words);
                for (size_t l = 0; l < gr.size(); ++l) {
                    row_word0[gr[l]] = sp_msg_words + (int)l;
                    row_lane[gr[l]] = (int)l;
                }
                sp_msg_words += dc * 32;
                for (int l = 0; l < 32; ++l) oc_cn_row.push_back(l < (int)gr.size() ? (uint16_t)slot0[gr[l]] : (uint16_t)rec_slots);
                for (int kb = 0; kb < blocks; ++kb)
                    for (int l = 0; l < 32; ++l) {
                        uint32_t c[4] = {0, 0, 0, 0};
                        if (l < (int)gr.size())
                            for (int j = 0; j < 4 && kb * 4 + j < dc; ++j) c[j] = (uint32_t)col_idx[rp[gr[l]] + kb * 4 + j];
                        oc_cnT.push_back(make_uint2(c[0] | (c[1] << 16), c[2] | (c[3] << 16)));
                    }
            }
        }
        // sum-product variable phase: groups of 32 bits of one degree; a 4-byte gather of the lanes' k-th messages is
        // conflict-free when the 32 rows sit in 32 different lanes of their check groups (bank = word mod 32 = lane)
        // shared-memory words: L[l_slots] first, then the messages, then the always-zero padding word
        const int l_slots = (int)qk::onchip_l_slots(n), zero_word = l_slots + sp_msg_words;
        sp_ok = sp_msg_words > 0 &&
                qk::onchip_spa_smem_bytes(n, sp_msg_words, (int)oc_cn_ginfo.size(), (n + 31) / 32) <= qk::kOnchipSmemMax;   // lower bound
        if (sp_ok) {
            for (const auto &cls : degree_classes(n, col_ptr)) {
                std::vector<std::vector<int>> groups;
                pack(cls, col_ptr, csc_row.data(), 32, 32, groups, row_lane.data());
                const int dv = col_ptr[cls[0] + 1] - col_ptr[cls[0]], blocks = (dv + 3) / 4;
                for (const auto &gr : groups) {
                    sp_group_item0.push_back((int)(sp_items.size() / 32));
                    for (int kb = 0; kb < blocks; ++kb)
                        for (int l = 0; l < 32; ++l) {
                            uint32_t e[4] = {(uint32_t)zero_word, (uint32_t)zero_word, (uint32_t)zero_word, (uint32_t)zero_word};
                            if (l < (int)gr.size())
                                for (int j = 0; j < 4 && kb * 4 + j < dv; ++j) {
                                    const int p = col_ptr[gr[l]] + kb * 4 + j, r = csc_row[p];
                                    e[j] = (uint32_t)(l_slots + row_word0[r] + 32 * (csc_edge[p] - rp[r]));
                                }
                            const uint32_t bit = l < (int)gr.size() ? (uint32_t)gr[l] : (uint32_t)n;
                            const uint32_t flags = (kb == 0 ? 0x10000u : 0u) | (kb == blocks - 1 ? 0x20000u : 0u);
                            sp_items.push_back(make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), bit | flags,
                                                          (uint32_t)sp_group_item0.size() - 1u));
                        }
                }
            }
            sp_group_item0.push_back((int)(sp_items.size() / 32));
            // the final word: every shared-memory word index must fit 16 bits, and the true group count enters the size
            sp_ok = qk::onchip_spa_smem_bytes(n, sp_msg_words, (int)oc_cn_ginfo.size(), (int)sp_group_item0.size() - 1) <= qk::kOnchipSmemMax &&
                    zero_word <= 65535;
        }
    }

    if (oc_ok) {
        // Self-check of the tables the on-chip kernel indexes shared memory with (the kernel itself does no bounds tests):
        // every bit index < n, every record slot <= m (m = scratch), every shift in [0, 31], every edge present once.
        size_t edges_cn = 0;
        for (size_t g = 0; g < oc_cn_ginfo.size(); ++g) {
            const int dc = oc_cn_ginfo[g].y, blocks = (dc + 3) / 4;
            if (dc < 1 || dc > 64 || (size_t)oc_cn_ginfo[g].x + (size_t)blocks * 32 > oc_cnT.size())
                return fail(QKDLDPC_ERR_STATE, "on-chip check table: bad group header %zu", g);
            for (int l = 0; l < 32; ++l) {
                const int slot = oc_cn_row[g * 32 + l];
                if (slot > rec_slots) return fail(QKDLDPC_ERR_STATE, "on-chip check table: record slot %d out of range", slot);
                const int row = slot == rec_slots ? m : (int)(std::upper_bound(slot0.begin(), slot0.end(), slot) - slot0.begin()) - 1;
                if (row < m && (slot0[row] != slot || rp[row + 1] - rp[row] != dc))
                    return fail(QKDLDPC_ERR_STATE, "on-chip check table: slot %d is not the first record of a row of %d edges", slot, dc);
                for (int k = 0; k < dc; ++k) {
                    const uint2 w = oc_cnT[oc_cn_ginfo[g].x + (k / 4) * 32 + l];
                    const uint32_t c = (k % 4 == 0) ? (w.x & 0xFFFFu) : (k % 4 == 1) ? (w.x >> 16) : (k % 4 == 2) ? (w.y & 0xFFFFu) : (w.y >> 16);
                    if ((int)c >= n || (row < m && (int)c != col_idx[rp[row] + k]))
                        return fail(QKDLDPC_ERR_STATE, "on-chip check table: wrong bit index in row %d", row);
                    edges_cn += row < m;
                }
            }
        }
        if (edges_cn != (size_t)nnz) return fail(QKDLDPC_ERR_STATE, "on-chip check table covers %zu of %lld edges", edges_cn, (long long)nnz);
    }
    if (sp_ok) {
        // the sum-product tables: every message word <= msg_words (the zero word), every edge owns exactly one word and
        // appears once, in ascending check order, between the first and the last item of its bit
        const int l_slots = (int)qk::onchip_l_slots(n), zero_word = l_slots + sp_msg_words;
        std::vector<char> seen((size_t)zero_word + 1, 0);
        std::vector<int> next_k((size_t)n, 0);
        size_t edges_sv = 0;
        if (sp_cn_moff.size() != oc_cn_ginfo.size()) return fail(QKDLDPC_ERR_STATE, "sum-product table: group count");
        for (size_t g = 0; g < oc_cn_ginfo.size(); ++g)
            if (sp_cn_moff[g] % 32 != 0 || sp_cn_moff[g] + oc_cn_ginfo[g].y * 32 > sp_msg_words)
                return fail(QKDLDPC_ERR_STATE, "sum-product table: bad message offset of group %zu", g);
        for (size_t i = 0; i < sp_items.size(); ++i) {
            const uint4 it = sp_items[i];
            const int bit = (int)(it.z & 0xFFFFu);
            const bool first = (it.z & 0x10000u) != 0, last = (it.z & 0x20000u) != 0;
            if (bit > n) return fail(QKDLDPC_ERR_STATE, "sum-product variable table: bit %d out of range", bit);
            const int grp = (int)it.w;
            if (grp < 0 || grp + 1 >= (int)sp_group_item0.size() || (int)(i / 32) < sp_group_item0[grp] || (int)(i / 32) >= sp_group_item0[grp + 1])
                return fail(QKDLDPC_ERR_STATE, "sum-product variable table: group of item %zu", i / 32);
            const int word[4] = {(int)(it.x & 0xFFFFu), (int)(it.x >> 16), (int)(it.y & 0xFFFFu), (int)(it.y >> 16)};
            if (bit < n && first != (next_k[bit] == 0)) return fail(QKDLDPC_ERR_STATE, "sum-product variable table: first flag of bit %d", bit);
            for (int j = 0; j < 4; ++j) {
                if (word[j] < l_slots || word[j] > zero_word) return fail(QKDLDPC_ERR_STATE, "sum-product variable table: word %d out of range", word[j]);
                if (bit == n) continue;
                const int k = next_k[bit], dv = col_ptr[bit + 1] - col_ptr[bit];
                if (k >= dv) {
                    if (word[j] != zero_word) return fail(QKDLDPC_ERR_STATE, "sum-product variable table: padding of bit %d", bit);
                    continue;
                }
                const int p = col_ptr[bit] + k, row = csc_row[p];
                if (word[j] != l_slots + row_word0[row] + 32 * (csc_edge[p] - rp[row]) || seen[word[j]])
                    return fail(QKDLDPC_ERR_STATE, "sum-product variable table: wrong word for bit %d", bit);
                seen[word[j]] = 1;
                next_k[bit] = k + 1;
                ++edges_sv;
            }
            if (bit < n && last != (next_k[bit] == col_ptr[bit + 1] - col_ptr[bit]))
                return fail(QKDLDPC_ERR_STATE, "sum-product variable table: last flag of bit %d", bit);
        }
        if (edges_sv != (size_t)nnz) return fail(QKDLDPC_ERR_STATE, "sum-product tables cover %zu of %lld edges", edges_sv, (long long)nnz);
    }
    // min-sum kernels: their own layout (storage order = processing order, conflict-aware lanes and edge order), once for
    // 16-byte records (float32 / float64 state) and once for the float32 kernel's 8-byte records
    build_oc2_layout(n, m, nnz, row_ptr, col_idx, col_ptr.data(), csc_edge.data(), csc_row.data(), 1, T.oc2);
    if (const char *err = check_oc2_layout(n, m, nnz, row_ptr, col_idx, col_ptr.data(), csc_edge.data(), csc_row.data(), T.oc2))
        return fail(QKDLDPC_ERR_STATE, "on-chip min-sum layout: %s", err);
    build_oc2_layout(n, m, nnz, row_ptr, col_idx, col_ptr.data(), csc_edge.data(), csc_row.data(), 1, T.oc2r8, Oc2Params::rec8());
    if (const char *err = check_oc2_layout(n, m, nnz, row_ptr, col_idx, col_ptr.data(), csc_edge.data(), csc_row.data(), T.oc2r8))
        return fail(QKDLDPC_ERR_STATE, "on-chip min-sum layout (8-byte records): %s", err);
    return QKDLDPC_OK;
}

}  // namespace qkhost

extern "C" int qkdldpc_onchip_layout_model(int32_t n, int32_t m, int64_t nnz, const int32_t *row_ptr, const int32_t *col_idx, int64_t *out) {
    if (!row_ptr || !col_idx || !out || n < 1 || m < 1 || nnz < 1 || row_ptr[0] != 0 || row_ptr[m] != nnz)
        return fail(QKDLDPC_ERR_INVALID, "bad graph");
    std::vector<int> col_ptr(n + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) {
        if (col_idx[e] < 0 || col_idx[e] >= n) return fail(QKDLDPC_ERR_INVALID, "column index out of range");
        col_ptr[col_idx[e] + 1]++;
    }
    for (int i = 0; i < n; ++i) col_ptr[i + 1] += col_ptr[i];
    std::vector<int> csc_edge(nnz), csc_row(nnz), cur(col_ptr.begin(), col_ptr.end() - 1);
    for (int j = 0; j < m; ++j)
        for (int e = row_ptr[j]; e < row_ptr[j + 1]; ++e) {
            const int p = cur[col_idx[e]]++;
            csc_edge[p] = e;
            csc_row[p] = j;
        }
    for (int r8 = 0; r8 < 2; ++r8) {   // 16-byte records, then the float32 kernel's 8-byte records
        qkhost::Oc2Tables T;
        qkhost::build_oc2_layout(n, m, nnz, row_ptr, col_idx, col_ptr.data(), csc_edge.data(), csc_row.data(), 1, T,
                                 r8 ? qkhost::Oc2Params::rec8() : qkhost::Oc2Params());
        if (const char *err = qkhost::check_oc2_layout(n, m, nnz, row_ptr, col_idx, col_ptr.data(), csc_edge.data(), csc_row.data(), T))
            return fail(QKDLDPC_ERR_STATE, "on-chip min-sum layout (%d-byte records): %s", r8 ? 8 : 16, err);
        int64_t *o = out + 6 * r8;
        o[0] = T.ok;
        o[1] = T.ok ? T.cn_gather : 0;
        o[2] = T.ok ? T.cn_gather_min : 0;
        o[3] = T.ok ? T.vn_gather : 0;
        o[4] = T.ok ? T.vn_gather_min : 0;
        o[5] = T.ok && T.vt16_ok;
    }
    return QKDLDPC_OK;
}

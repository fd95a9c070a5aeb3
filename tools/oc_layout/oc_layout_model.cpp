// Bank-conflict model of the on-chip min-sum layout (qkd_ldpc_v_b200/csrc/onchip_layout.hpp) on a code dumped by dump_code.py:
//   g++ -O2 -std=c++17 -o /tmp/oc_layout_model tools/oc_layout/oc_layout_model.cpp && /tmp/oc_layout_model /tmp/I80.bin [effort]
// Prints the modelled shared-memory wavefronts per decoder iteration of the two gathers against their conflict-free minimum.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "../../qkd_ldpc_v_b200/csrc/onchip_layout.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    int hdr[3];
    if (fread(hdr, 4, 3, f) != 3) return 1;
    const int n = hdr[0], m = hdr[1], nnz = hdr[2];
    std::vector<int> rp(m + 1), ci(nnz);
    if (fread(rp.data(), 4, m + 1, f) != (size_t)m + 1 || fread(ci.data(), 4, nnz, f) != (size_t)nnz) return 1;
    fclose(f);
    std::vector<int> col_ptr(n + 1, 0), csc_edge(nnz), csc_row(nnz);
    for (int e = 0; e < nnz; ++e) col_ptr[ci[e] + 1]++;
    for (int i = 0; i < n; ++i) col_ptr[i + 1] += col_ptr[i];
    std::vector<int> cur(col_ptr.begin(), col_ptr.end() - 1);
    for (int j = 0; j < m; ++j)
        for (int e = rp[j]; e < rp[j + 1]; ++e) {
            const int p = cur[ci[e]]++;
            csc_edge[p] = e;
            csc_row[p] = j;
        }
    for (int rec8 = 0; rec8 <= 1; ++rec8) {
        const qkhost::Oc2Params prm = rec8 ? qkhost::Oc2Params::rec8() : qkhost::Oc2Params();
        const int effort = argc > 2 ? atoi(argv[2]) : 1;
        qkhost::Oc2Tables T;
        const auto t0 = std::chrono::steady_clock::now();
        qkhost::build_oc2_layout(n, m, nnz, rp.data(), ci.data(), col_ptr.data(), csc_edge.data(), csc_row.data(), effort, T, prm);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (!T.ok) { printf("%s: not eligible\n", rec8 ? "8-byte records" : "16-byte records"); continue; }
        const char *err = qkhost::check_oc2_layout(n, m, nnz, rp.data(), ci.data(), col_ptr.data(), csc_edge.data(), csc_row.data(), T);
        const double units = nnz / 32.0;
        // a record gather of vn_w lanes moves vn_w * rec_bytes = 128 bytes: one wavefront when conflict-free
        printf("%s, effort %d: %.0f ms  check %s | CN gather %lld (min %lld) = %.3f per 32 edges | VN record gather %lld (min %lld) = %.3f wavefronts per 32 edges | "
               "groups cn %zu vn %zu, record slots %d, l_slots %d\n",
               rec8 ? "8-byte records" : "16-byte records", effort, ms, err ? err : "ok", T.cn_gather, T.cn_gather_min, T.cn_gather / units, T.vn_gather,
               T.vn_gather_min, T.vn_gather / units, T.cn_g.size(), T.vn_g.size(), T.rec_slots, T.l_slots);
        // per degree class of the variable phase
        {
            const int W = prm.vn_w;
            std::vector<long long> cost(256, 0), mn(256, 0);
            for (const auto &g : T.vn_g)
                for (int k = 0; k < g.deg; ++k)
                    for (int q = 0; q < 32 / W; ++q) {
                        int slots[16];
                        for (int l = 0; l < W; ++l)
                            slots[l] = (int)(((&T.vT[(size_t)g.off + (size_t)(k / 4) * 32 + q * W + l].x)[k % 4] >> 5) / (unsigned)prm.rec_bytes);
                        cost[g.deg] += qkhost::oc2::octet_cost(slots, W, W);
                        mn[g.deg] += 1;
                    }
            for (int d = 0; d < 256; ++d)
                if (mn[d]) printf("   dv %d: %lld / %lld = %.2f\n", d, cost[d], mn[d], (double)cost[d] / mn[d]);
        }
    }
    return 0;
}

#include "rate_adapt.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>

namespace qkdldpc {

H_matrix_params adapt_code_rate(Xoshiro256pp &prng, const H_matrix &matrix, double QBER, double delta, double efficiency,
                                bool untainted, std::string *warn) {
    const double h_b = -QBER * std::log2(QBER) - (1. - QBER) * std::log2(1. - QBER);   // binary entropy
    const double optimal_R = 1. - efficiency * h_b;
    const size_t n = matrix.n(), m = matrix.m();
    const double original_R = 1. - static_cast<double>(m) / static_cast<double>(n);
    const int s = static_cast<int>(std::ceil((original_R - optimal_R * (1. - delta)) * static_cast<double>(n)));
    const int p = static_cast<int>(delta * static_cast<double>(n) - static_cast<double>(s));
    H_matrix_params out{};
    char buf[512];
    if (s <= 0 || p <= 0) {
        if (warn) {
            std::snprintf(buf, sizeof buf, "WARNING: R0 = %.3f, QBER = %.4f, delta = %.3f, f_EC = %.3f. Adapted code rate R = %.3f beyond the "
                          "achievable rate range: Rmin = %.3f, Rmax = %.3f. This parameters will not be used in simulations.",
                          original_R, QBER, delta, efficiency, optimal_R, (original_R - delta) / (1. - delta), original_R / (1. - delta));
            *warn = buf;
        }
        return out;
    }
    std::vector<int> positions(n);
    if (untainted) {
        const auto &u = matrix.punctured_bits_untainted;
        if (static_cast<size_t>(p) > u.size()) {
            if (warn) {
                std::snprintf(buf, sizeof buf, "WARNING: R0 = %.3f, QBER = %.4f, delta = %.3f, f_EC = %.3f. The calculated number of punctured "
                              "bits (%d) exceeds the number of bits produced by untainted algorithm (%zu). These parameters will not be used.",
                              original_R, QBER, delta, efficiency, p, u.size());
                *warn = buf;
            }
            return out;
        }
        out.punctured_bits.assign(u.begin(), u.begin() + p);
    } else {
        std::iota(positions.begin(), positions.end(), 0);
        std::shuffle(positions.begin(), positions.end(), prng);
        out.punctured_bits.assign(positions.begin(), positions.begin() + p);
    }
    std::sort(out.punctured_bits.begin(), out.punctured_bits.end());
    std::iota(positions.begin(), positions.end(), 0);
    std::vector<int> remaining(n - static_cast<size_t>(p));
    std::set_difference(positions.begin(), positions.end(), out.punctured_bits.begin(), out.punctured_bits.end(), remaining.begin());
    std::shuffle(remaining.begin(), remaining.end(), prng);
    out.shortened_bits.assign(remaining.begin(), remaining.begin() + s);
    std::sort(out.shortened_bits.begin(), out.shortened_bits.end());
    out.delta = delta;
    out.efficiency = efficiency;
    out.shortened_fraction = static_cast<double>(s) / static_cast<double>(n);
    out.punctured_fraction = static_cast<double>(p) / static_cast<double>(n);
    out.adapted_code_rate = static_cast<double>(n - m - static_cast<size_t>(s)) / static_cast<double>(n - static_cast<size_t>(p) - static_cast<size_t>(s));
    return out;
}

std::vector<int> select_punctured_bits_untainted(Xoshiro256pp &prng, const H_matrix &matrix) {
    const int n = static_cast<int>(matrix.n());
    // second-order neighbourhoods (get_second_order_neighbors, :975-997), as sorted vectors
    std::vector<std::vector<int>> n2(static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) {
        std::set<int> s;
        for (int c : matrix.bit_nodes[static_cast<size_t>(i)])
            s.insert(matrix.check_nodes[static_cast<size_t>(c)].begin(), matrix.check_nodes[static_cast<size_t>(c)].end());
        s.erase(i);
        n2[static_cast<size_t>(i)].assign(s.begin(), s.end());
    }
    // The reference recounts |N2(i) ∩ X| for every i in X on every round (O(N * |N2|) per pick, minutes for N = 10^5). The
    // counts only change for second-order neighbours of removed nodes, so they are maintained incrementally, and the
    // nodes of X are kept in one ordered set per count value: the candidates of a round are the lowest non-empty
    // bucket, already in ascending node order, and ONE draw of uniform_int_distribution<size_t>(0, candidates-1) is made
    // per pick -- the random stream, and therefore the selection, is identical to the reference's
    // (tests/test_host_surface.py). N = 102400: 16 s with per-round scans, well under a second this way.
    std::vector<char> in_x(static_cast<size_t>(n), 1);
    std::vector<int> cnt(static_cast<size_t>(n));
    int max_cnt = 0;
    for (int i = 0; i < n; ++i) {
        cnt[static_cast<size_t>(i)] = static_cast<int>(n2[static_cast<size_t>(i)].size());
        max_cnt = std::max(max_cnt, cnt[static_cast<size_t>(i)]);
    }
    std::vector<std::set<int>> bucket(static_cast<size_t>(max_cnt) + 1);
    for (int i = 0; i < n; ++i) bucket[static_cast<size_t>(cnt[static_cast<size_t>(i)])].insert(bucket[static_cast<size_t>(cnt[static_cast<size_t>(i)])].end(), i);
    int remaining = n, low = 0;   // low: no bucket below it is non-empty
    std::vector<int> picked;
    auto remove_from_x = [&](int v) {
        if (!in_x[static_cast<size_t>(v)]) return;
        in_x[static_cast<size_t>(v)] = 0;
        --remaining;
        bucket[static_cast<size_t>(cnt[static_cast<size_t>(v)])].erase(v);
        for (int w : n2[static_cast<size_t>(v)])
            if (in_x[static_cast<size_t>(w)]) {
                int &c = cnt[static_cast<size_t>(w)];
                bucket[static_cast<size_t>(c)].erase(w);
                --c;
                bucket[static_cast<size_t>(c)].insert(w);
                low = std::min(low, c);
            }
    };
    while (remaining > 0) {
        while (bucket[static_cast<size_t>(low)].empty()) ++low;
        const std::set<int> &candidates = bucket[static_cast<size_t>(low)];
        std::uniform_int_distribution<size_t> pick(0, candidates.size() - 1);
        auto it = candidates.begin();
        std::advance(it, static_cast<std::ptrdiff_t>(pick(prng)));
        const int chosen = *it;
        picked.push_back(chosen);
        remove_from_x(chosen);
        for (int w : n2[static_cast<size_t>(chosen)]) remove_from_x(w);
    }
    return picked;
}

static std::vector<int> read_untp(const fs::path &p) {
    std::vector<int> v;
    std::ifstream in(p);
    std::string line;
    if (in.is_open() && std::getline(in, line)) {
        std::istringstream iss(line);
        std::copy(std::istream_iterator<int>(iss), std::istream_iterator<int>(), std::back_inserter(v));
    }
    return v;
}

std::vector<int> get_punctured_bits_untainted(const fs::path &matrix_path, Xoshiro256pp &prng, const H_matrix &matrix,
                                              const fs::path &cache_dir) {
    fs::path beside = matrix_path;
    beside.replace_extension(".untp");
    fs::path cached = cache_dir.empty() ? beside : cache_dir / beside.filename();
    std::vector<int> v = read_untp(beside);
    fs::path used = beside;
    if (v.empty() && cached != beside) {
        v = read_untp(cached);
        used = cached;
    }
    for (int x : v)
        if (x < 0 || static_cast<size_t>(x) >= matrix.n())
            throw std::runtime_error("The punctured bit index '" + std::to_string(x) + "' is out of range [0," +
                                     std::to_string(matrix.n() - 1) + "]. File: " + used.string());
    if (v.empty()) {
        v = select_punctured_bits_untainted(prng, matrix);
        if (!cache_dir.empty()) fs::create_directories(cache_dir);
        std::ofstream out(cached);
        if (!out.is_open()) throw std::runtime_error("Unable to open file for writing: " + cached.string());
        std::copy(v.begin(), v.end(), std::ostream_iterator<int>(out, " "));
    }
    return v;
}

namespace {
// find_available_index, :121-137
int first_unmarked(const std::vector<int> &checks, const std::vector<char> &marked) {
    for (int c : checks)
        if (!marked[static_cast<size_t>(c)]) return c;
    return -1;
}
std::vector<int> by_ascending_weight(const H_matrix &matrix, const std::vector<int> &candidates) {
    // The reference std::sort's (index, check list) pairs by list size (:155-159, :226-230). std::sort is not stable,
    // but it is deterministic: its moves depend only on the comparison results and the element count, so sorting
    // (index, weight) pairs with the same comparator yields the same permutation under the same libstdc++.
    std::vector<std::pair<int, size_t>> order;
    order.reserve(candidates.size());
    for (int i : candidates) order.emplace_back(i, matrix.bit_nodes[static_cast<size_t>(i)].size());
    std::sort(order.begin(), order.end(),
              [](const std::pair<int, size_t> &a, const std::pair<int, size_t> &b) { return a.second < b.second; });
    std::vector<int> out;
    out.reserve(order.size());
    for (auto &o : order) out.push_back(o.first);
    return out;
}
}  // namespace

std::vector<int> get_bits_positions_to_remove(const H_matrix &matrix) {
    std::vector<int> all(matrix.n());
    std::iota(all.begin(), all.end(), 0);
    std::vector<char> marked(matrix.m(), 0);
    std::vector<int> out;
    for (int i : by_ascending_weight(matrix, all)) {
        const int c = first_unmarked(matrix.bit_nodes[static_cast<size_t>(i)], marked);
        if (c != -1) {
            out.push_back(i);
            marked[static_cast<size_t>(c)] = 1;
        }
    }
    std::sort(out.begin(), out.end());
    return out;
}

std::vector<int> get_bits_positions_to_remove_rate_adapt(const H_matrix &matrix, const H_matrix_params &p) {
    std::vector<char> marked(matrix.m(), 0);
    std::vector<int> out, rest;
    size_t s = 0, q = 0;
    for (int i = 0; i < static_cast<int>(matrix.n()); ++i) {
        if (s < p.shortened_bits.size() && p.shortened_bits[s] == i) {          // bounds-checked (the reference is not)
            out.push_back(i);
            ++s;
        } else if (q < p.punctured_bits.size() && p.punctured_bits[q] == i) {
            out.push_back(i);
            const int c = first_unmarked(matrix.bit_nodes[static_cast<size_t>(i)], marked);
            if (c != -1) marked[static_cast<size_t>(c)] = 1;
            ++q;
        } else {
            rest.push_back(i);
        }
    }
    for (int i : by_ascending_weight(matrix, rest)) {
        const int c = first_unmarked(matrix.bit_nodes[static_cast<size_t>(i)], marked);
        if (c != -1) {
            out.push_back(i);
            marked[static_cast<size_t>(c)] = 1;
        }
    }
    std::sort(out.begin(), out.end());
    return out;
}

void extend_frame(Xoshiro256pp &prng, const H_matrix_params &mp, const std::vector<int> &alice, const std::vector<int> &bob,
                  std::vector<int> &alice_ext, std::vector<int> &bob_ext) {
    const size_t total = alice.size();
    alice_ext.assign(total, 0);
    bob_ext.assign(total, 0);
    std::uniform_int_distribution<int> coin(0, 1);
    size_t p = 0, s = 0, k = 0;
    for (size_t i = 0; i < total; ++i) {
        if (p < mp.punctured_bits.size() && static_cast<size_t>(mp.punctured_bits[p]) == i) {
            alice_ext[i] = coin(prng);
            bob_ext[i] = coin(prng);
            ++p;
        } else if (s < mp.shortened_bits.size() && static_cast<size_t>(mp.shortened_bits[s]) == i) {
            ++s;   // both zero
        } else {
            alice_ext[i] = alice[k];
            bob_ext[i] = bob[k];
            ++k;
        }
    }
}

}  // namespace qkdldpc

// extern "C" view of the host-side helpers, so that the Python tests / bench can drive the same C++ code the
// qkdldpc_sim binary uses (input generation, loaders, rate adaptation).
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "keygen.hpp"

#define HOST_API extern "C" __attribute__((visibility("default")))

// run_trial's key generation for a list of per-trial seeds; packed output [count][words]; returns accurate QBER.
HOST_API double qkdhost_gen_keys(const std::uint64_t *seeds, std::int64_t count, std::int64_t n, double qber,
                                 std::uint32_t *alice_packed, std::uint32_t *bob_packed, int threads) {
    const std::size_t words = static_cast<std::size_t>((n + 31) / 32);
    if (threads < 1) threads = 1;
    std::vector<double> acc(static_cast<std::size_t>(threads), 0.0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        const std::int64_t lo = count * t / threads, hi = count * (t + 1) / threads;
        pool.emplace_back([=, &acc] {
            std::vector<int> a(static_cast<std::size_t>(n)), b;
            for (std::int64_t i = lo; i < hi; ++i) {
                qkdldpc::Xoshiro256pp prng(seeds[i]);
                qkdldpc::fill_random_bits(prng, a);
                acc[static_cast<std::size_t>(t)] = qkdldpc::inject_errors(prng, a, qber, b);
                qkdldpc::pack_frame(a, alice_packed + static_cast<std::size_t>(i) * words);
                qkdldpc::pack_frame(b, bob_packed + static_cast<std::size_t>(i) * words);
            }
        });
    }
    for (auto &th : pool) th.join();
    return static_cast<double>(static_cast<std::size_t>(static_cast<double>(n) * qber)) / static_cast<double>(n);
}

HOST_API void qkdhost_trial_seeds(std::uint64_t simulation_seed, std::int64_t count, std::uint64_t *seeds_out) {
    auto s = qkdldpc::trial_seeds(simulation_seed, static_cast<std::size_t>(count));
    std::memcpy(seeds_out, s.data(), s.size() * sizeof(std::uint64_t));
}

// ---------------------------------------------------------------------------------------------------------------
// Test-facing views of the loaders, the combination builder, the statistics and the CSV writer.
// ---------------------------------------------------------------------------------------------------------------
#include <cstdio>
#include <sstream>
#include <algorithm>
#include <iterator>

#include "config.hpp"
#include "matrix.hpp"
#include "rate_adapt.hpp"
#include "simulation.hpp"

namespace {
thread_local std::string g_host_err;
template <typename F>
int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_host_err = e.what();
        return -1;
    }
}
int64_t emit(const std::string &s, char *out, int64_t cap) {
    if (out && cap > 0) {
        const size_t k = std::min<size_t>(s.size(), static_cast<size_t>(cap - 1));
        std::memcpy(out, s.data(), k);
        out[k] = 0;
    }
    return static_cast<int64_t>(s.size());
}
}  // namespace

HOST_API const char *qkdhost_last_error() { return g_host_err.c_str(); }

HOST_API void *qkdhost_matrix_load(const char *path, int format) {
    qkdldpc::H_matrix *h = nullptr;
    if (guarded([&] { h = new qkdldpc::H_matrix(qkdldpc::read_matrix(path, format)); }) != 0) return nullptr;
    return h;
}
HOST_API void qkdhost_matrix_free(void *m) { delete static_cast<qkdldpc::H_matrix *>(m); }
HOST_API void qkdhost_matrix_info(void *mv, std::int64_t *n, std::int64_t *m, std::int64_t *nnz, int *is_regular) {
    const auto *h = static_cast<qkdldpc::H_matrix *>(mv);
    *n = static_cast<std::int64_t>(h->n());
    *m = static_cast<std::int64_t>(h->m());
    std::int64_t e = 0;
    for (const auto &r : h->check_nodes) e += static_cast<std::int64_t>(r.size());
    *nnz = e;
    *is_regular = h->is_regular ? 1 : 0;
}
// CSR exactly as qkdldpc_sim hands it to qkdldpc_code_create (validated: ascending, consistent with bit_nodes).
HOST_API int qkdhost_matrix_csr(void *mv, std::int32_t *row_ptr, std::int32_t *col_idx) {
    return guarded([&] {
        const auto g = qkdldpc::to_csr_checked(*static_cast<qkdldpc::H_matrix *>(mv), "matrix");
        std::copy(g.row_ptr.begin(), g.row_ptr.end(), row_ptr);
        std::copy(g.col_idx.begin(), g.col_idx.end(), col_idx);
    });
}
// bit_nodes as loaded (column view), for comparison with the reference loader.
HOST_API void qkdhost_matrix_csc(void *mv, std::int32_t *col_ptr, std::int32_t *row_idx) {
    const auto *h = static_cast<qkdldpc::H_matrix *>(mv);
    std::int32_t e = 0;
    col_ptr[0] = 0;
    for (size_t i = 0; i < h->n(); ++i) {
        for (int r : h->bit_nodes[i]) row_idx[e++] = r;
        col_ptr[i + 1] = e;
    }
}

static std::string describe_config(const qkdldpc::config_data &c) {
    std::ostringstream o;
    o.precision(17);
    o << "threads=" << c.THREADS_NUMBER << " trials=" << c.TRIALS_NUMBER << " seed=" << c.SIMULATION_SEED
      << " privacy=" << c.ENABLE_PRIVACY_MAINTENANCE << " throughput=" << c.ENABLE_THROUGHPUT_MEASUREMENT << " rtt_on=" << c.CONSIDER_RTT
      << " rtt=" << c.RTT << " alg=" << c.DECODING_ALGORITHM << " max_iter=" << c.DECODING_ALG_MAX_ITERATIONS
      << " format=" << c.MATRIX_FORMAT << " thr_on=" << c.ENABLE_DECODING_ALG_MSG_LLR_THRESHOLD << " thr=" << c.DECODING_ALG_MSG_LLR_THRESHOLD
      << " adapt=" << c.ENABLE_CODE_RATE_ADAPTATION << " untainted=" << c.ENABLE_UNTAINTED_PUNCTURING
      << " adapt_ranges=" << c.USE_ADAPTATION_PARAMETERS_RANGES << "\n";
    auto src = [&](const char *name, const qkdldpc::scaling_factor_source &s) {
        o << name << ": use_range=" << s.use_range << " range=" << s.range.begin << ":" << s.range.end << ":" << s.range.step << " maps=";
        for (const auto &m : s.maps) o << m.code_rate << ">" << m.scaling_factor << ",";
        o << "\n";
    };
    src("primary", c.DECODING_ALG_PARAMS.primary);
    src("secondary", c.DECODING_ALG_PARAMS.secondary);
    o << "qber_ranges=";
    for (const auto &r : c.R_QBER_RANGES) o << r.code_rate << ">" << r.QBER_begin << ":" << r.QBER_end << ":" << r.QBER_step << ",";
    o << "\nadapt_param_ranges=";
    for (const auto &r : c.R_ADAPT_PARAMS_RANGES)
        o << r.code_rate << ">" << r.delta_begin << ":" << r.delta_end << ":" << r.delta_step << "/" << r.efficiency_begin << ":" << r.efficiency_end
          << ":" << r.efficiency_step << ",";
    o << "\nadapt_param_maps=";
    for (const auto &r : c.R_QBER_ADAPT_PARAMS_MAPS)
        o << r.code_rate << ">" << r.QBER_adapt_params.QBER << "/" << r.QBER_adapt_params.delta << "/" << r.QBER_adapt_params.efficiency << ",";
    o << "\n";
    return o.str();
}

// Parsed configuration as canonical text (also reports which schema generation the file used).
HOST_API std::int64_t qkdhost_describe_config(const char *config_path, char *out, std::int64_t cap, int *schema_version) {
    std::string s;
    if (guarded([&] {
            const auto c = qkdldpc::parse_config_data(config_path);
            if (schema_version) *schema_version = c.schema_version;
            s = describe_config(c);
        }) != 0)
        return -1;
    return emit(s, out, cap);
}

static std::uint64_t fnv(const std::vector<int> &v) {
    std::uint64_t h = 1469598103934665603ull;
    for (int x : v) {
        h = (h ^ static_cast<std::uint32_t>(x)) * 1099511628211ull;
    }
    return h;
}

// Combination list of prepare_sim_inputs as canonical text: one line per (matrix, combination).
HOST_API std::int64_t qkdhost_describe_inputs(const char *config_path, const char *matrix_dir, const char *untp_cache, char *out,
                                              std::int64_t cap) {
    std::string s;
    if (guarded([&] {
            const auto cfg = qkdldpc::parse_config_data(config_path);
            auto paths = qkdldpc::get_file_paths_in_directory(matrix_dir, ".mtrx");
            const auto inputs = qkdldpc::prepare_sim_inputs(cfg, paths, untp_cache ? untp_cache : "");
            std::ostringstream o;
            o.precision(17);
            size_t k = 0;
            for (const auto &in : inputs)
                for (const auto &c : in.combinations) {
                    const auto &mp = c.matrix_params;
                    o << k++ << " " << in.matrix_path.filename().string() << " n=" << in.matrix.n() << " m=" << in.matrix.m()
                      << " regular=" << in.matrix.is_regular << " qber=" << c.config_QBER << " delta=" << mp.delta << " eff=" << mp.efficiency
                      << " pf=" << mp.punctured_fraction << " sf=" << mp.shortened_fraction << " ra=" << mp.adapted_code_rate
                      << " p=" << mp.punctured_bits.size() << ":" << fnv(mp.punctured_bits) << " s=" << mp.shortened_bits.size() << ":"
                      << fnv(mp.shortened_bits) << " rm=" << mp.bits_to_remove.size() << ":" << fnv(mp.bits_to_remove)
                      << " f1=" << c.scaling_factors.primary << " f2=" << c.scaling_factors.secondary << "\n";
                }
            s = o.str();
        }) != 0)
        return -1;
    return emit(s, out, cap);
}

// Statistics + CSV text (header line + one data line) of one combination from per-trial results, through the tally
// vector exactly as qkdldpc_sim does it. flags: bit0 syndromes_match, bit1 keys_match.
HOST_API std::int64_t qkdhost_csv_from_trials(const char *config_path, const std::int32_t *iters, const std::uint8_t *flags, std::int64_t count,
                                              const char *matrix_name, std::int64_t n, std::int64_t m, int is_regular, double config_qber,
                                              double accurate_qber, double primary, double secondary, const double *adapt5, char *out,
                                              std::int64_t cap) {
    std::string s;
    if (guarded([&] {
            auto cfg = qkdldpc::parse_config_data(config_path);
            cfg.TRIALS_NUMBER = static_cast<size_t>(count);
            std::vector<std::uint64_t> tally(cfg.DECODING_ALG_MAX_ITERATIONS + 5, 0);
            for (std::int64_t i = 0; i < count; ++i) {
                tally[0]++;
                if (flags[i] & 1u) {
                    tally[1]++;
                    if (flags[i] & 2u) tally[2]++;
                    tally[4 + static_cast<size_t>(iters[i])]++;
                }
                tally[3] += static_cast<std::uint64_t>(iters[i]);
            }
            qkdldpc::sim_result r{};
            r.sim_number = 0;
            r.matrix_filename = matrix_name;
            r.is_regular = is_regular != 0;
            r.num_bit_nodes = static_cast<size_t>(n);
            r.num_check_nodes = static_cast<size_t>(m);
            r.config_QBER = config_qber;
            r.accurate_QBER = accurate_qber;
            r.scaling_factors = {primary, secondary};
            if (adapt5) {
                r.delta = adapt5[0]; r.efficiency = adapt5[1]; r.punctured_fraction = adapt5[2];
                r.shortened_fraction = adapt5[3]; r.adapted_code_rate = adapt5[4];
            }
            qkdldpc::process_tally(tally.data(), cfg.DECODING_ALG_MAX_ITERATIONS, cfg.TRIALS_NUMBER, r);
            s = qkdldpc::csv_header(cfg) + "\n" + qkdldpc::csv_line(cfg, r) + "\n";
        }) != 0)
        return -1;
    return emit(s, out, cap);
}

HOST_API std::int64_t qkdhost_format_shortest(double v, char *out, std::int64_t cap) { return emit(qkdldpc::format_shortest_comma(v), out, cap); }

// adapt_code_rate with a fresh generator (same contract as the oracle driver's ref_adapt_code_rate).
HOST_API int qkdhost_adapt_code_rate(void *mv, std::uint64_t seed, int untainted, const std::int32_t *untp, std::int64_t n_untp, double qber,
                                     double delta, double efficiency, int privacy_maintenance, std::int32_t *punct_out, std::int64_t *n_punct,
                                     std::int32_t *short_out, std::int64_t *n_short, std::int32_t *remove_out, std::int64_t *n_remove,
                                     double *fractions /*[3]*/) {
    return guarded([&] {
        qkdldpc::H_matrix h = *static_cast<qkdldpc::H_matrix *>(mv);
        h.punctured_bits_untainted.assign(untp, untp + n_untp);
        qkdldpc::Xoshiro256pp prng(seed);
        qkdldpc::H_matrix_params mp = qkdldpc::adapt_code_rate(prng, h, qber, delta, efficiency, untainted != 0);
        if (!(mp.punctured_bits.empty() && mp.shortened_bits.empty())) {
            if (privacy_maintenance) mp.bits_to_remove = qkdldpc::get_bits_positions_to_remove_rate_adapt(h, mp);
            else std::merge(mp.punctured_bits.begin(), mp.punctured_bits.end(), mp.shortened_bits.begin(), mp.shortened_bits.end(),
                            std::back_inserter(mp.bits_to_remove));
        }
        *n_punct = static_cast<std::int64_t>(mp.punctured_bits.size());
        *n_short = static_cast<std::int64_t>(mp.shortened_bits.size());
        *n_remove = static_cast<std::int64_t>(mp.bits_to_remove.size());
        std::copy(mp.punctured_bits.begin(), mp.punctured_bits.end(), punct_out);
        std::copy(mp.shortened_bits.begin(), mp.shortened_bits.end(), short_out);
        std::copy(mp.bits_to_remove.begin(), mp.bits_to_remove.end(), remove_out);
        fractions[0] = mp.punctured_fraction;
        fractions[1] = mp.shortened_fraction;
        fractions[2] = mp.adapted_code_rate;
    });
}

// select_punctured_bits_untainted with a fresh generator (selection order).
HOST_API std::int64_t qkdhost_untainted(void *mv, std::uint64_t seed, std::int32_t *out) {
    std::int64_t k = -1;
    guarded([&] {
        qkdldpc::Xoshiro256pp prng(seed);
        const auto v = qkdldpc::select_punctured_bits_untainted(prng, *static_cast<qkdldpc::H_matrix *>(mv));
        std::copy(v.begin(), v.end(), out);
        k = static_cast<std::int64_t>(v.size());
    });
    return k;
}

// Rate-adapted trial inputs: run_trial's keys (simulation.cpp:549-555) followed by the frame extension of
// QKD_LDPC_RATE_ADAPT (qkd_ldpc_algorithm.cpp:1136-1174) from the same per-trial generator. Packed output.
HOST_API double qkdhost_gen_keys_rate_adapt(const std::uint64_t *seeds, std::int64_t count, std::int64_t n, double qber,
                                            const std::int32_t *punct, std::int64_t n_punct, const std::int32_t *shortd, std::int64_t n_short,
                                            std::uint32_t *alice_packed, std::uint32_t *bob_packed) {
    const std::size_t words = static_cast<std::size_t>((n + 31) / 32);
    qkdldpc::H_matrix_params mp;
    mp.punctured_bits.assign(punct, punct + n_punct);
    mp.shortened_bits.assign(shortd, shortd + n_short);
    std::vector<int> a(static_cast<std::size_t>(n)), b, ae, be;
    double acc = 0;
    for (std::int64_t i = 0; i < count; ++i) {
        qkdldpc::Xoshiro256pp prng(seeds[i]);
        qkdldpc::fill_random_bits(prng, a);
        acc = qkdldpc::inject_errors(prng, a, qber, b);
        qkdldpc::extend_frame(prng, mp, a, b, ae, be);
        qkdldpc::pack_frame(ae, alice_packed + static_cast<std::size_t>(i) * words);
        qkdldpc::pack_frame(be, bob_packed + static_cast<std::size_t>(i) * words);
    }
    return acc;
}

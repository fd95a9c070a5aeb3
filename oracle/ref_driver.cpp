// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// A thin extern "C" driver around the UNMODIFIED reference translation units
// (compiled where they lie under /root/reference/src by oracle/Makefile into
// oracle/_ref/libqkdref.so). It exposes the reference's own functions so that
// (1) the C restatement in oracle/ldpc_oracle.c can be pinned against the real
// code, (2) golden vectors can be generated (tests/golden/make_golden.py) and
// (3) bench.py --impl reference can time the reference's CPU decoder.
//
// Nothing here re-implements reference arithmetic: every decode goes through
// sum_product_decoding / ... / QKD_LDPC / run_trial of the reference itself.
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "array_and_matrix_operations.hpp"
#include "config.hpp"
#include "qkd_ldpc_algorithm.hpp"
#include "simulation.hpp"

config_data CFG;  // the reference defines this in main.cpp:22, which is not compiled here

namespace {
struct RefMatrix {
    H_matrix h;
};
thread_local std::string g_err;

decoding_result dispatch(int alg, const std::vector<double> &llr, const H_matrix &h, const std::vector<int> &synd,
                         size_t max_iter, double primary, double secondary, double thr, std::vector<int> &out) {
    switch (alg) {
        case 0: return sum_product_decoding(llr, h, synd, max_iter, thr, out);
        case 1: return sum_product_linear_approx_decoding(llr, h, synd, max_iter, thr, out);
        case 2: return min_sum_normalized_decoding(llr, h, synd, max_iter, primary, thr, out);
        case 3: return min_sum_offset_decoding(llr, h, synd, max_iter, primary, thr, out);
        case 4: return adaptive_min_sum_normalized_decoding(llr, h, synd, max_iter, primary, secondary, thr, out);
        default: return adaptive_min_sum_offset_decoding(llr, h, synd, max_iter, primary, secondary, thr, out);
    }
}
}  // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

// format: 0 uncompressed, 1 alist, 2 sparse_1, 3 sparse_2 (config.hpp:202)
void *ref_matrix_load(const char *path, int format) {
    try {
        auto m = std::make_unique<RefMatrix>();
        fs::path p(path);
        if (format == 0) m->h = read_sparse_uncompressed_matrix(p);
        else if (format == 1) m->h = read_sparse_matrix_alist(p);
        else if (format == 2) m->h = read_sparse_matrix_1(p);
        else m->h = read_sparse_matrix_2(p);
        return m.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void ref_matrix_free(void *m) { delete static_cast<RefMatrix *>(m); }

void ref_matrix_info(void *mv, int64_t *n, int64_t *m, int64_t *nnz_rows, int64_t *nnz_cols, int *is_regular) {
    auto &h = static_cast<RefMatrix *>(mv)->h;
    *n = static_cast<int64_t>(h.bit_nodes.size());
    *m = static_cast<int64_t>(h.check_nodes.size());
    int64_t a = 0, b = 0;
    for (auto &r : h.check_nodes) a += static_cast<int64_t>(r.size());
    for (auto &c : h.bit_nodes) b += static_cast<int64_t>(c.size());
    *nnz_rows = a;
    *nnz_cols = b;
    *is_regular = h.is_regular ? 1 : 0;
}

// Adjacency exactly as the reference holds it (NOT re-sorted).
void ref_matrix_csr(void *mv, int32_t *row_ptr, int32_t *col_idx) {
    auto &h = static_cast<RefMatrix *>(mv)->h;
    int32_t k = 0;
    row_ptr[0] = 0;
    for (size_t j = 0; j < h.check_nodes.size(); ++j) {
        for (int c : h.check_nodes[j]) col_idx[k++] = c;
        row_ptr[j + 1] = k;
    }
}
void ref_matrix_csc(void *mv, int32_t *col_ptr, int32_t *row_idx) {
    auto &h = static_cast<RefMatrix *>(mv)->h;
    int32_t k = 0;
    col_ptr[0] = 0;
    for (size_t i = 0; i < h.bit_nodes.size(); ++i) {
        for (int r : h.bit_nodes[i]) row_idx[k++] = r;
        col_ptr[i + 1] = k;
    }
}

// Everything the hot path reads from the global CFG.
void ref_set_cfg(int algorithm, int64_t max_iter, int enable_threshold, double threshold, int privacy_maintenance,
                 int rate_adaptation) {
    CFG.DECODING_ALGORITHM = static_cast<size_t>(algorithm);
    CFG.DECODING_ALG_MAX_ITERATIONS = static_cast<size_t>(max_iter);
    CFG.ENABLE_DECODING_ALG_MSG_LLR_THRESHOLD = enable_threshold != 0;
    CFG.DECODING_ALG_MSG_LLR_THRESHOLD = threshold;
    CFG.ENABLE_PRIVACY_MAINTENANCE = privacy_maintenance != 0;
    CFG.ENABLE_CODE_RATE_ADAPTATION = rate_adaptation != 0;
    CFG.ENABLE_THROUGHPUT_MEASUREMENT = false;
    CFG.TRACE_QKD_LDPC = CFG.TRACE_DECODING_ALG = CFG.TRACE_DECODING_ALG_LLR = false;
}

// run_trial's input generation (simulation.cpp:549-555) with the reference's own helpers.
double ref_gen_keys(uint64_t seed, int64_t n, double qber, int32_t *alice, int32_t *bob) {
    XoshiroCpp::Xoshiro256PlusPlus prng(seed);
    std::vector<int> a(static_cast<size_t>(n)), b(static_cast<size_t>(n));
    fill_random_bits(prng, a);
    double acc = inject_errors(prng, a, qber, b);
    std::memcpy(alice, a.data(), sizeof(int32_t) * n);
    std::memcpy(bob, b.data(), sizeof(int32_t) * n);
    return acc;
}

// seeds[] as QKD_LDPC_batch_simulation draws them (simulation.cpp:713-719).
void ref_trial_seeds(uint64_t simulation_seed, int64_t count, uint64_t *seeds) {
    XoshiroCpp::Xoshiro256PlusPlus prng(simulation_seed);
    std::uniform_int_distribution<size_t> distribution(0, std::numeric_limits<size_t>::max());
    for (int64_t i = 0; i < count; ++i) seeds[i] = distribution(prng);
}

// One of the six reference decoders on caller-provided LLRs / syndrome. Returns iterations_num.
int64_t ref_decode(void *mv, int alg, const double *llr, const int32_t *syndrome, int64_t max_iter, double primary,
                   double secondary, int enable_threshold, double threshold, int32_t *bits_out, int *syndromes_match) {
    auto &h = static_cast<RefMatrix *>(mv)->h;
    CFG.ENABLE_DECODING_ALG_MSG_LLR_THRESHOLD = enable_threshold != 0;
    CFG.TRACE_DECODING_ALG = CFG.TRACE_DECODING_ALG_LLR = false;
    std::vector<double> l(llr, llr + h.bit_nodes.size());
    std::vector<int> s(syndrome, syndrome + h.check_nodes.size());
    std::vector<int> out(h.bit_nodes.size());
    decoding_result r = dispatch(alg, l, h, s, static_cast<size_t>(max_iter), primary, secondary, threshold, out);
    std::memcpy(bits_out, out.data(), sizeof(int32_t) * out.size());
    *syndromes_match = r.syndromes_match ? 1 : 0;
    return static_cast<int64_t>(r.iterations_num);
}

void ref_syndrome(void *mv, const int32_t *bits, int32_t *synd_out) {
    auto &h = static_cast<RefMatrix *>(mv)->h;
    std::vector<int> b(bits, bits + h.bit_nodes.size());
    std::vector<int> s(h.check_nodes.size());
    calculate_syndrome(b, h, s);
    std::memcpy(synd_out, s.data(), sizeof(int32_t) * s.size());
}

// The reference's run_trial (simulation.cpp:540-577) for `count` seeds over `threads` host threads with the
// same static block partition as BS::thread_pool::detach_loop. ref_set_cfg must have been called.
// flags: bit0 syndromes_match, bit1 keys_match. Returns 0, or -1 on exception.
int ref_run_trials(void *mv, double qber, const uint64_t *seeds, int64_t count, double primary, double secondary,
                   const int32_t *punct, int64_t n_punct, const int32_t *shortd, int64_t n_short,
                   const int32_t *bits_to_remove, int64_t n_remove, int threads, int32_t *iters, uint8_t *flags,
                   double *accurate_qber) {
    auto &h = static_cast<RefMatrix *>(mv)->h;
    H_matrix_params mp;
    mp.punctured_bits.assign(punct, punct + n_punct);
    mp.shortened_bits.assign(shortd, shortd + n_short);
    mp.bits_to_remove.assign(bits_to_remove, bits_to_remove + n_remove);
    decoding_scaling_factors sf{primary, secondary};
    if (threads < 1) threads = 1;
    std::vector<std::string> errs(static_cast<size_t>(threads));
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        int64_t lo = count * t / threads, hi = count * (t + 1) / threads;
        pool.emplace_back([&, lo, hi, t] {
            try {
                for (int64_t i = lo; i < hi; ++i) {
                    trial_result r = run_trial(h, qber, static_cast<size_t>(seeds[i]), mp, sf);
                    iters[i] = static_cast<int32_t>(r.ldpc_res.decoding_res.iterations_num);
                    flags[i] = static_cast<uint8_t>((r.ldpc_res.decoding_res.syndromes_match ? 1 : 0) |
                                                    (r.ldpc_res.keys_match ? 2 : 0));
                    if (accurate_qber) accurate_qber[i] = r.accurate_QBER;
                }
            } catch (const std::exception &e) {
                errs[static_cast<size_t>(t)] = e.what();
            }
        });
    }
    for (auto &th : pool) th.join();
    for (auto &e : errs)
        if (!e.empty()) {
            g_err = e;
            return -1;
        }
    return 0;
}

// Decode-only timing entry: QKD_LDPC (qkd_ldpc_algorithm.cpp:1031) on pre-generated keys (int32 per bit,
// frame-major), `threads` host threads, static block partition. Used by bench.py's CPU baseline.
int ref_qkd_ldpc_batch(void *mv, const int32_t *alice, const int32_t *bob, int64_t n_frames, double qber,
                       double primary, double secondary, int threads, int32_t *iters, uint8_t *flags) {
    auto &h = static_cast<RefMatrix *>(mv)->h;
    const size_t n = h.bit_nodes.size();
    decoding_scaling_factors sf{primary, secondary};
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        int64_t lo = n_frames * t / threads, hi = n_frames * (t + 1) / threads;
        pool.emplace_back([&, lo, hi] {
            std::vector<int> a(n), b(n);
            for (int64_t i = lo; i < hi; ++i) {
                a.assign(alice + i * n, alice + (i + 1) * n);
                b.assign(bob + i * n, bob + (i + 1) * n);
                LDPC_result r = QKD_LDPC(h, a, b, qber, sf, H_matrix_params{});
                iters[i] = static_cast<int32_t>(r.decoding_res.iterations_num);
                flags[i] = static_cast<uint8_t>((r.decoding_res.syndromes_match ? 1 : 0) | (r.keys_match ? 2 : 0));
            }
        });
    }
    for (auto &th : pool) th.join();
    return 0;
}

// adapt_code_rate (array_and_matrix_operations.cpp:1129-1223). punctured_untainted may be null/0.
// Outputs are written into caller buffers of capacity n. Returns 0 / -1.
int ref_adapt_code_rate(void *mv, uint64_t seed, int untainted, const int32_t *untp, int64_t n_untp, double qber,
                        double delta, double efficiency, int32_t *punct_out, int64_t *n_punct, int32_t *short_out,
                        int64_t *n_short, int32_t *remove_out, int64_t *n_remove, double *fractions /*[3]*/) {
    try {
        H_matrix h = static_cast<RefMatrix *>(mv)->h;
        h.punctured_bits_untainted.assign(untp, untp + n_untp);
        CFG.ENABLE_UNTAINTED_PUNCTURING = untainted != 0;
        XoshiroCpp::Xoshiro256PlusPlus prng(seed);
        H_matrix_params mp = adapt_code_rate(prng, h, qber, delta, efficiency);
        *n_punct = static_cast<int64_t>(mp.punctured_bits.size());
        *n_short = static_cast<int64_t>(mp.shortened_bits.size());
        *n_remove = static_cast<int64_t>(mp.bits_to_remove.size());
        std::copy(mp.punctured_bits.begin(), mp.punctured_bits.end(), punct_out);
        std::copy(mp.shortened_bits.begin(), mp.shortened_bits.end(), short_out);
        std::copy(mp.bits_to_remove.begin(), mp.bits_to_remove.end(), remove_out);
        fractions[0] = mp.punctured_fraction;
        fractions[1] = mp.shortened_fraction;
        fractions[2] = mp.adapted_code_rate;
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// ---- host-surface views (config parser, combination builder, statistics, CSV writer) of the reference ----------
// Same canonical text as qkdhost_describe_config / qkdhost_describe_inputs / qkdhost_csv_from_trials, produced by the
// reference's own parse_config_data, prepare_sim_inputs, process_trials_results and write_file.
static int64_t emit_text(const std::string &s, char *out, int64_t cap) {
    if (out && cap > 0) {
        const size_t k = std::min<size_t>(s.size(), static_cast<size_t>(cap - 1));
        std::memcpy(out, s.data(), k);
        out[k] = 0;
    }
    return static_cast<int64_t>(s.size());
}

static uint64_t fnv_ints(const std::vector<int> &v) {
    uint64_t h = 1469598103934665603ull;
    for (int x : v) h = (h ^ static_cast<uint32_t>(x)) * 1099511628211ull;
    return h;
}

int64_t ref_describe_config(const char *config_path, char *out, int64_t cap) {
    try {
        const config_data c = parse_config_data(config_path);
        std::ostringstream o;
        o.precision(17);
        o << "threads=" << c.THREADS_NUMBER << " trials=" << c.TRIALS_NUMBER << " seed=" << c.SIMULATION_SEED
          << " privacy=" << c.ENABLE_PRIVACY_MAINTENANCE << " throughput=" << c.ENABLE_THROUGHPUT_MEASUREMENT << " rtt_on=" << c.CONSIDER_RTT
          << " rtt=" << c.RTT << " alg=" << c.DECODING_ALGORITHM << " max_iter=" << c.DECODING_ALG_MAX_ITERATIONS
          << " format=" << c.MATRIX_FORMAT << " thr_on=" << c.ENABLE_DECODING_ALG_MSG_LLR_THRESHOLD << " thr=" << c.DECODING_ALG_MSG_LLR_THRESHOLD
          << " adapt=" << c.ENABLE_CODE_RATE_ADAPTATION << " untainted=" << c.ENABLE_UNTAINTED_PUNCTURING
          << " adapt_ranges=" << c.USE_ADAPTATION_PARAMETERS_RANGES << "\n";
        auto src = [&](const char *name, const auto &sf) {
            o << name << ": use_range=" << sf.use_range << " range=" << sf.range.begin << ":" << sf.range.end << ":" << sf.range.step << " maps=";
            for (const auto &m : sf.maps) o << m.code_rate << ">" << m.scaling_factor << ",";
            o << "\n";
        };
        src("primary", c.DECODING_ALG_PARAMS.primary);
        src("secondary", c.DECODING_ALG_PARAMS.secondary);
        o << "qber_ranges=";
        for (const auto &r : c.R_QBER_RANGES) o << r.code_rate << ">" << r.QBER_begin << ":" << r.QBER_end << ":" << r.QBER_step << ",";
        o << "\nadapt_param_ranges=";
        for (const auto &r : c.R_ADAPT_PARAMS_RANGES)
            o << r.code_rate << ">" << r.delta_begin << ":" << r.delta_end << ":" << r.delta_step << "/" << r.efficiency_begin << ":"
              << r.efficiency_end << ":" << r.efficiency_step << ",";
        o << "\nadapt_param_maps=";
        for (const auto &r : c.R_QBER_ADAPT_PARAMS_MAPS)
            o << r.code_rate << ">" << r.QBER_adapt_params.QBER << "/" << r.QBER_adapt_params.delta << "/" << r.QBER_adapt_params.efficiency << ",";
        o << "\n";
        return emit_text(o.str(), out, cap);
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

int64_t ref_describe_inputs(const char *config_path, const char *matrix_dir, char *out, int64_t cap) {
    try {
        CFG = parse_config_data(config_path);
        std::vector<fs::path> paths = get_file_paths_in_directory(matrix_dir, ".mtrx");
        const std::vector<sim_input> inputs = prepare_sim_inputs(paths);
        std::ostringstream o;
        o.precision(17);
        size_t k = 0;
        for (const auto &in : inputs)
            for (const auto &c : in.combinations) {
                const auto &mp = c.matrix_params;
                o << k++ << " " << in.matrix_path.filename().string() << " n=" << in.matrix.bit_nodes.size() << " m=" << in.matrix.check_nodes.size()
                  << " regular=" << in.matrix.is_regular << " qber=" << c.config_QBER << " delta=" << mp.delta << " eff=" << mp.efficiency
                  << " pf=" << mp.punctured_fraction << " sf=" << mp.shortened_fraction << " ra=" << mp.adapted_code_rate
                  << " p=" << mp.punctured_bits.size() << ":" << fnv_ints(mp.punctured_bits) << " s=" << mp.shortened_bits.size() << ":"
                  << fnv_ints(mp.shortened_bits) << " rm=" << mp.bits_to_remove.size() << ":" << fnv_ints(mp.bits_to_remove)
                  << " f1=" << c.scaling_factors.primary << " f2=" << c.scaling_factors.secondary << "\n";
            }
        return emit_text(o.str(), out, cap);
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// process_trials_results + write_file on caller-provided per-trial results; returns the CSV text.
int64_t ref_csv_from_trials(const char *config_path, const int32_t *iters, const uint8_t *flags, int64_t count, const char *matrix_name,
                            int64_t n, int64_t m, int is_regular, double config_qber, double accurate_qber, double primary, double secondary,
                            const double *adapt5, const char *tmp_dir, char *out, int64_t cap) {
    try {
        CFG = parse_config_data(config_path);
        CFG.TRIALS_NUMBER = static_cast<size_t>(count);
        CFG.ENABLE_THROUGHPUT_MEASUREMENT = false;
        std::vector<trial_result> tr(static_cast<size_t>(count));
        for (int64_t i = 0; i < count; ++i) {
            tr[i].ldpc_res.decoding_res.iterations_num = static_cast<size_t>(iters[i]);
            tr[i].ldpc_res.decoding_res.syndromes_match = (flags[i] & 1u) != 0;
            tr[i].ldpc_res.keys_match = (flags[i] & 2u) != 0;
            tr[i].accurate_QBER = accurate_qber;
        }
        H_matrix h;
        h.bit_nodes.resize(static_cast<size_t>(n));
        h.check_nodes.resize(static_cast<size_t>(m));
        h.is_regular = is_regular != 0;
        H_matrix_params mp{};
        std::vector<sim_result> res(1);
        res[0].sim_number = 0;
        res[0].matrix_filename = matrix_name;
        res[0].is_regular = h.is_regular;
        res[0].num_bit_nodes = static_cast<size_t>(n);
        res[0].num_check_nodes = static_cast<size_t>(m);
        res[0].config_QBER = config_qber;
        res[0].accurate_QBER = accurate_qber;
        res[0].scaling_factors.primary = primary;
        res[0].scaling_factors.secondary = secondary;
        if (adapt5) {
            res[0].delta = adapt5[0]; res[0].efficiency = adapt5[1]; res[0].punctured_fraction = adapt5[2];
            res[0].shortened_fraction = adapt5[3]; res[0].adapted_code_rate = adapt5[4];
        }
        process_trials_results(tr, h, mp, res[0]);
        const fs::path file = write_file(res, "00h-00m-00s", tmp_dir);
        std::ifstream in(file);
        std::stringstream ss;
        ss << in.rdbuf();
        in.close();
        fs::remove(file);
        return emit_text(ss.str(), out, cap);
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// select_punctured_bits_untainted (array_and_matrix_operations.cpp:1002-1068) with a fresh generator.
int64_t ref_untainted(void *mv, uint64_t seed, int32_t *out) {
    try {
        XoshiroCpp::Xoshiro256PlusPlus prng(seed);
        const std::vector<int> v = select_punctured_bits_untainted(prng, static_cast<RefMatrix *>(mv)->h);
        std::copy(v.begin(), v.end(), out);
        return static_cast<int64_t>(v.size());
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

// get_bits_positions_to_remove_rate_adapt (array_and_matrix_operations.cpp:189-256) on given position lists.
// pad > 0 appends `pad` sentinel entries (-1) beyond the logical end of both lists' STORAGE (capacity, not size) so
// that the reference's unchecked reads of shortened_bits[s] / punctured_bits[p] past the end (:215,:220) hit a
// defined value instead of heap garbage.
int64_t ref_bits_to_remove_rate_adapt(void *mv, const int32_t *punct, int64_t n_p, const int32_t *shortd, int64_t n_s, int pad,
                                      int32_t *out) {
    try {
        H_matrix_params mp{};
        mp.punctured_bits.reserve(static_cast<size_t>(n_p + pad));
        mp.shortened_bits.reserve(static_cast<size_t>(n_s + pad));
        mp.punctured_bits.assign(punct, punct + n_p);
        mp.shortened_bits.assign(shortd, shortd + n_s);
        for (int k = 0; k < pad; ++k) {   // write sentinels into the spare capacity, then shrink the size back
            mp.punctured_bits.push_back(-1);
            mp.shortened_bits.push_back(-1);
        }
        mp.punctured_bits.resize(static_cast<size_t>(n_p));
        mp.shortened_bits.resize(static_cast<size_t>(n_s));
        const std::vector<int> v = get_bits_positions_to_remove_rate_adapt(static_cast<RefMatrix *>(mv)->h, mp);
        std::copy(v.begin(), v.end(), out);
        return static_cast<int64_t>(v.size());
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"

#include "simulation.hpp"

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <future>
#include <limits>
#include <stdexcept>
#include <thread>

#include "keygen.hpp"

namespace qkdldpc {

// ---------------------------------------------------------------------------------------------------------------
// Range / map lookups
// ---------------------------------------------------------------------------------------------------------------

// steps = round((end - begin) / step) + 1, value = begin + j * step; begin == end -> one value
// (simulation.cpp:192-204, 325-343).
std::vector<double> expand_range(double begin, double end, double step) {
    std::vector<double> out;
    if (begin == end) {
        out.push_back(begin);
        return out;
    }
    const size_t steps = static_cast<size_t>(std::round((end - begin) / step)) + 1;
    for (size_t j = 0; j < steps; ++j) out.push_back(begin + static_cast<double>(j) * step);
    return out;
}

static std::string rate_error(const char *what, double code_rate) {
    char buf[256];
    std::snprintf(buf, sizeof buf, "An error occurred while %s based on code rate(R). Matrix code rate, R = %g.", what, code_rate);
    return buf;
}

std::vector<double> get_rate_based_QBER_range(double code_rate, const std::vector<R_QBER_range> &ranges) {
    for (const auto &r : ranges)
        if (code_rate <= r.code_rate) {
            auto v = expand_range(r.QBER_begin, r.QBER_end, r.QBER_step);
            if (!v.empty()) return v;
            break;
        }
    throw std::runtime_error(rate_error("generating a QBER range", code_rate));
}

double get_rate_based_scaling_factor_value(double code_rate, const std::vector<R_scaling_factor_map> &maps) {
    for (const auto &m : maps)
        if (code_rate <= m.code_rate) return m.scaling_factor;
    throw std::runtime_error(rate_error("searching scaling factor value", code_rate));
}

namespace {

struct adapt_values { std::vector<double> delta, efficiency; };

adapt_values adaptation_ranges(double code_rate, const std::vector<R_adaptation_parameters_range> &ranges) {
    adapt_values v;
    for (const auto &r : ranges)
        if (code_rate <= r.code_rate) {
            v.delta = expand_range(r.delta_begin, r.delta_end, r.delta_step);
            v.efficiency = expand_range(r.efficiency_begin, r.efficiency_end, r.efficiency_step);
            break;
        }
    if (v.delta.empty()) throw std::runtime_error(rate_error("generating a delta range", code_rate));
    if (v.efficiency.empty()) throw std::runtime_error(rate_error("generating an efficiency(f_EC) range", code_rate));
    return v;
}

// All consecutive entries that share the first code_rate >= R (simulation.cpp:287-320).
std::vector<QBER_adaptation_parameters> adaptation_maps(double code_rate, const std::vector<R_QBER_adaptation_parameters_map> &maps) {
    std::vector<QBER_adaptation_parameters> out;
    double target = -1.;
    for (const auto &e : maps) {
        if (out.empty()) {
            if (code_rate <= e.code_rate) {
                target = e.code_rate;
                out.push_back(e.QBER_adapt_params);
            }
        } else if (e.code_rate == target) {
            out.push_back(e.QBER_adapt_params);
        } else {
            break;
        }
    }
    if (out.empty()) throw std::runtime_error(rate_error("generating a QBER - delta - efficiency(f_EC) maps", code_rate));
    return out;
}

std::vector<double> factor_values(const scaling_factor_source &src, double code_rate) {
    if (src.use_range) {
        auto v = expand_range(src.range.begin, src.range.end, src.range.step);
        if (v.empty()) throw std::runtime_error("An error occurred while generating vector of scaling factor values.");
        return v;
    }
    return {get_rate_based_scaling_factor_value(code_rate, src.maps)};
}

void finish_adapted_params(const config_data &cfg, const H_matrix &matrix, H_matrix_params &mp) {
    if (cfg.ENABLE_PRIVACY_MAINTENANCE) {
        mp.bits_to_remove = get_bits_positions_to_remove_rate_adapt(matrix, mp);
    } else {
        mp.bits_to_remove.clear();
        std::merge(mp.punctured_bits.begin(), mp.punctured_bits.end(), mp.shortened_bits.begin(), mp.shortened_bits.end(),
                   std::back_inserter(mp.bits_to_remove));
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// prepare_sim_inputs (simulation.cpp:371-537). One generator seeded with SIMULATION_SEED runs through all matrices
// (untainted lists and random puncturing / shortening positions draw from it in matrix order).
// ---------------------------------------------------------------------------------------------------------------
std::vector<sim_input> prepare_sim_inputs(const config_data &cfg, const std::vector<fs::path> &matrix_paths,
                                          const fs::path &untp_cache_dir, bool print_warnings) {
    Xoshiro256pp prng(cfg.SIMULATION_SEED);
    std::vector<sim_input> inputs(matrix_paths.size());
    for (size_t i = 0; i < matrix_paths.size(); ++i) {
        sim_input &in = inputs[i];
        in.matrix = read_matrix(matrix_paths[i], static_cast<int>(cfg.MATRIX_FORMAT));
        in.matrix_path = matrix_paths[i];
        const double code_rate = in.matrix.code_rate();

        std::vector<std::pair<double, H_matrix_params>> qber_params;
        size_t skipped = 0;
        if (cfg.ENABLE_CODE_RATE_ADAPTATION) {
            if (cfg.ENABLE_UNTAINTED_PUNCTURING)
                in.matrix.punctured_bits_untainted = get_punctured_bits_untainted(matrix_paths[i], prng, in.matrix, untp_cache_dir);
            auto try_add = [&](double qber, double delta, double efficiency) {
                std::string warn;
                H_matrix_params mp = adapt_code_rate(prng, in.matrix, qber, delta, efficiency, cfg.ENABLE_UNTAINTED_PUNCTURING, &warn);
                if (mp.punctured_bits.empty() && mp.shortened_bits.empty()) {
                    if (!warn.empty() && print_warnings) std::fprintf(stderr, "%s\n", warn.c_str());
                    ++skipped;
                    return;   // combination skipped (simulation.cpp:413-415)
                }
                finish_adapted_params(cfg, in.matrix, mp);
                qber_params.emplace_back(qber, std::move(mp));
            };
            if (cfg.USE_ADAPTATION_PARAMETERS_RANGES) {
                const adapt_values av = adaptation_ranges(code_rate, cfg.R_ADAPT_PARAMS_RANGES);
                for (double qber : get_rate_based_QBER_range(code_rate, cfg.R_QBER_RANGES))
                    for (double delta : av.delta)
                        for (double eff : av.efficiency) try_add(qber, delta, eff);
            } else {
                for (const auto &p : adaptation_maps(code_rate, cfg.R_QBER_ADAPT_PARAMS_MAPS)) try_add(p.QBER, p.delta, p.efficiency);
            }
        } else {
            H_matrix_params mp{};
            if (cfg.ENABLE_PRIVACY_MAINTENANCE) mp.bits_to_remove = get_bits_positions_to_remove(in.matrix);
            for (double qber : get_rate_based_QBER_range(code_rate, cfg.R_QBER_RANGES)) qber_params.emplace_back(qber, mp);
        }

        std::vector<decoding_scaling_factors> factors;
        const size_t alg = cfg.DECODING_ALGORITHM;
        if (alg == DEC_NMSA || alg == DEC_OMSA) {
            for (double p : factor_values(cfg.DECODING_ALG_PARAMS.primary, code_rate)) factors.push_back({p, 0.});
        } else if (alg == DEC_ANMSA || alg == DEC_AOMSA) {
            const auto prim = factor_values(cfg.DECODING_ALG_PARAMS.primary, code_rate);
            const auto sec = factor_values(cfg.DECODING_ALG_PARAMS.secondary, code_rate);
            for (double p : prim)
                for (double s : sec) factors.push_back({p, s});
        } else {
            factors.push_back({});
        }

        if (skipped > 0 && !print_warnings)
            std::fprintf(stderr, "%s: %zu parameter combinations are outside the achievable rate range and will not be used\n",
                         matrix_paths[i].filename().string().c_str(), skipped);
        in.combinations.reserve(qber_params.size() * factors.size());
        for (const auto &qp : qber_params)
            for (const auto &sf : factors) in.combinations.push_back({qp.first, qp.second, sf});
    }
    return inputs;
}

// ---------------------------------------------------------------------------------------------------------------
// Statistics from the tally vector. tally[4 + k] = number of syndrome-matched frames with iterations_num == k.
// ---------------------------------------------------------------------------------------------------------------
void process_tally(const uint64_t *tally, size_t max_iterations, size_t trials_number, sim_result &r) {
    const uint64_t ok_dec = tally[1], ok_ldpc = tally[2];
    size_t it_min = std::numeric_limits<size_t>::max(), it_max = 0;
    double mean = 0., var = 0.;
    for (size_t k = 0; k <= max_iterations; ++k) {
        const uint64_t c = tally[4 + k];
        if (!c) continue;
        it_min = std::min(it_min, k);
        it_max = std::max(it_max, k);
        mean += static_cast<double>(c) * static_cast<double>(k);   // exact: integer-valued partial sums < 2^53
    }
    if (ok_dec > 0) {
        mean /= static_cast<double>(ok_dec);
        for (size_t k = 0; k <= max_iterations; ++k)
            if (tally[4 + k]) var += static_cast<double>(tally[4 + k]) * std::pow(static_cast<double>(k) - mean, 2);
        var /= static_cast<double>(ok_dec);
    }
    r.iter_success_dec_alg_max = it_max;
    r.iter_success_dec_alg_min = (it_min == std::numeric_limits<size_t>::max()) ? 0 : it_min;
    r.iter_success_dec_alg_mean = mean;
    r.iter_success_dec_alg_std_dev = std::sqrt(var);
    r.ratio_trials_success_ldpc = static_cast<double>(ok_ldpc) / static_cast<double>(trials_number);
    r.ratio_trials_success_dec_alg = static_cast<double>(ok_dec) / static_cast<double>(trials_number);
    r.iterations_executed = tally[3];
}

// ---------------------------------------------------------------------------------------------------------------
// Batch simulation: per combination, trials are split into contiguous ranges over the devices (quirk Q16: results
// do not depend on the partition); each device thread generates its frames chunk by chunk with the reference's
// per-trial RNG streams (seed = seeds[n] + combination index, simulation.cpp:743) and hands them to the GPU.
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct chunk_buffers {
    std::vector<uint32_t> alice, bob;
    double accurate_qber = 0.;
};

void generate_chunk(const config_data &cfg, const H_matrix &matrix, const sim_combination &comb, const std::vector<uint64_t> &seeds,
                    size_t curr_sim, size_t first, size_t count, size_t host_threads, chunk_buffers &out) {
    const size_t n = matrix.n(), words = (n + 31) / 32;
    out.alice.assign(count * words, 0u);
    out.bob.assign(count * words, 0u);
    host_threads = std::max<size_t>(1, std::min(host_threads, count));
    std::vector<std::thread> pool;
    std::vector<double> acc(host_threads, 0.);
    std::atomic<bool> zero_qber{false};
    for (size_t t = 0; t < host_threads; ++t) {
        const size_t lo = count * t / host_threads, hi = count * (t + 1) / host_threads;
        pool.emplace_back([&, t, lo, hi] {
            std::vector<int> a(n), b, ae, be;
            for (size_t k = lo; k < hi; ++k) {
                Xoshiro256pp prng(seeds[first + k] + curr_sim);
                fill_random_bits(prng, a);
                acc[t] = inject_errors(prng, a, comb.config_QBER, b);
                if (acc[t] == 0.) zero_qber = true;
                if (cfg.ENABLE_CODE_RATE_ADAPTATION) {
                    extend_frame(prng, comb.matrix_params, a, b, ae, be);
                    pack_frame(ae, out.alice.data() + k * words);
                    pack_frame(be, out.bob.data() + k * words);
                } else {
                    pack_frame(a, out.alice.data() + k * words);
                    pack_frame(b, out.bob.data() + k * words);
                }
            }
        });
    }
    for (auto &th : pool) th.join();
    if (zero_qber) throw std::runtime_error("Key size '" + std::to_string(n) + "' is too small for QBER.");   // simulation.cpp:556-557
    out.accurate_qber = acc[0];
}

}  // namespace

namespace {

// Decodes trials [lo, hi) of one combination on one handle, chunk by chunk: the next chunk's inputs are generated on
// host threads while the GPU decodes the current one. Adds the chunk tallies to `tally`.
struct range_outcome {
    double accurate_qber = 0., ms = 0.;
};

// One entry of the qkdldpc_run_trials_multi table: the combination's parameters, its position lists and -- when the
// protocol removes bits after the decoder -- H_matrix_params.bits_to_remove.
void fill_combination(const config_data &cfg, const sim_combination &comb, size_t curr_sim, qkdldpc_combination &e) {
    const auto &mp = comb.matrix_params;
    const bool ra = cfg.ENABLE_CODE_RATE_ADAPTATION;
    e.qber = comb.config_QBER;
    e.primary = comb.scaling_factors.primary;
    e.secondary = comb.scaling_factors.secondary;
    e.punct_pos = ra ? mp.punctured_bits.data() : nullptr;
    e.n_punct = ra ? static_cast<int32_t>(mp.punctured_bits.size()) : 0;
    e.short_pos = ra ? mp.shortened_bits.data() : nullptr;
    e.n_short = ra ? static_cast<int32_t>(mp.shortened_bits.size()) : 0;
    e.seed_offset = static_cast<uint64_t>(curr_sim);
    const bool removal = (ra || cfg.ENABLE_PRIVACY_MAINTENANCE) && !mp.bits_to_remove.empty();
    e.remove_pos = removal ? mp.bits_to_remove.data() : nullptr;
    e.n_remove = removal ? static_cast<int32_t>(mp.bits_to_remove.size()) : 0;
    e.reserved = 0;
}
range_outcome decode_trial_range(const config_data &cfg, const decoder_api &api, qkdldpc_code *code, const qkdldpc_params &P,
                                 const H_matrix &matrix, const sim_combination &comb, const std::vector<uint64_t> &seeds, size_t curr_sim,
                                 size_t lo, size_t hi, size_t chunk, size_t gen_threads, bool host_keygen, std::vector<uint64_t> &tally) {
    range_outcome out;
    if (!host_keygen && api.run_trials) {
        // inputs are generated on the device from the per-trial seeds (bit-identical to run_trial's, simulation.cpp:549-555):
        // only 8 bytes per trial cross PCIe
        const auto &mp = comb.matrix_params;
        const bool ra = cfg.ENABLE_CODE_RATE_ADAPTATION;
        std::vector<uint64_t> t(tally.size());
        // remove_bits is the last step of QKD_LDPC (privacy maintenance) and of QKD_LDPC_RATE_ADAPT (always)
        // (qkd_ldpc_algorithm.cpp:1089-1092, 1218-1220): the combination entry carries the list and the library builds the
        // final keys on the device inside the timed call, as the reference's chrono region does (simulation.cpp:559-568)
        const bool removal = (ra || cfg.ENABLE_PRIVACY_MAINTENANCE) && !mp.bits_to_remove.empty() && api.run_trials_multi != nullptr;
        for (size_t pos = lo; pos < hi; pos += chunk) {
            const size_t cnt = std::min(chunk, hi - pos);
            double acc = 0.;
            const auto t0 = std::chrono::steady_clock::now();
            int rc;
            if (removal) {
                qkdldpc_combination cb{};
                fill_combination(cfg, comb, curr_sim, cb);
                rc = api.run_trials_multi(code, &P, 1, &cb, static_cast<int64_t>(cnt), seeds.data() + pos, nullptr, nullptr, t.data(), &acc);
            } else {
                rc = api.run_trials(code, &P, static_cast<int64_t>(cnt), seeds.data() + pos, static_cast<uint64_t>(curr_sim), comb.config_QBER,
                                    ra ? mp.punctured_bits.data() : nullptr, ra ? static_cast<int32_t>(mp.punctured_bits.size()) : 0,
                                    ra ? mp.shortened_bits.data() : nullptr, ra ? static_cast<int32_t>(mp.shortened_bits.size()) : 0,
                                    nullptr, nullptr, nullptr, t.data(), &acc);
            }
            out.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (rc != 0) {
                const std::string msg = api.last_error();
                throw std::runtime_error(msg.rfind("Key size", 0) == 0 ? msg : "qkdldpc_run_trials: " + msg);
            }
            for (size_t k = 0; k < tally.size(); ++k) tally[k] += t[k];
            if (pos == lo) out.accurate_qber = acc;
        }
        return out;
    }
    chunk_buffers cur, nxt;
    std::vector<uint64_t> t(tally.size());
    size_t pos = lo;
    if (pos < hi) generate_chunk(cfg, matrix, comb, seeds, curr_sim, pos, std::min(chunk, hi - pos), gen_threads, cur);
    while (pos < hi) {
        const size_t cnt = std::min(chunk, hi - pos), npos = pos + cnt;
        std::future<void> prefetch;
        if (npos < hi)
            prefetch = std::async(std::launch::async, [&, npos] {
                generate_chunk(cfg, matrix, comb, seeds, curr_sim, npos, std::min(chunk, hi - npos), gen_threads, nxt);
            });
        const double q = cur.accurate_qber;
        const auto &mp = comb.matrix_params;
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = api.decode_batch(code, &P, static_cast<int64_t>(cnt), cur.alice.data(), cur.bob.data(), &q, 1, mp.punctured_bits.data(),
                                        static_cast<int32_t>(mp.punctured_bits.size()), mp.shortened_bits.data(),
                                        static_cast<int32_t>(mp.shortened_bits.size()), nullptr, nullptr, nullptr, t.data());
        out.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (prefetch.valid()) prefetch.get();
        if (rc != 0) throw std::runtime_error(std::string("qkdldpc_decode_batch: ") + api.last_error());
        for (size_t k = 0; k < tally.size(); ++k) tally[k] += t[k];
        if (pos == lo) out.accurate_qber = q;
        pos = npos;
        std::swap(cur, nxt);
    }
    return out;
}

}  // namespace

// Two ways to keep the GPUs busy, chosen per config:
//  * many trials per combination (the FER configs: 10^5 .. 10^6): combinations run one after the other, the trials of
//    each are split into contiguous ranges over the devices (the reference's detach_loop over threads);
//  * few trials per combination (the rate-adaptation sweeps: thousands of combinations x 100 trials): a batch of 100
//    frames cannot fill a B200, so several combinations are decoded CONCURRENTLY -- every device gets a few handles
//    (each with its own stream) and worker threads pull combinations from a queue. Results are stored by combination
//    index, so the CSV is the same either way (quirk Q16: a trial's inputs depend only on seeds[n] + combination index).
std::vector<sim_result> QKD_LDPC_batch_simulation(const config_data &cfg, const std::vector<sim_input> &sim_in,
                                                  const decoder_api &api, const device_options &dev) {
    size_t sim_total = 0;
    for (const auto &in : sim_in) sim_total += in.combinations.size();
    std::vector<sim_result> results(sim_total);
    const std::vector<uint64_t> seeds = trial_seeds(cfg.SIMULATION_SEED, cfg.TRIALS_NUMBER);
    const size_t trials = cfg.TRIALS_NUMBER;
    const size_t n_dev = std::max<size_t>(1, dev.devices.size());
    const size_t tally_len = cfg.DECODING_ALG_MAX_ITERATIONS + 5;
    const size_t chunk = static_cast<size_t>(std::max<int64_t>(1, dev.chunk_frames));
    // concurrent combinations per device: enough small batches in flight to cover ~2 frames per SM-resident CTA slot
    size_t per_dev = 1;
    if (dev.concurrent_combinations > 0) per_dev = static_cast<size_t>(dev.concurrent_combinations);
    else if (trials < 2048) per_dev = std::min<size_t>(8, (2048 + trials - 1) / trials);
    const bool sweep_mode = per_dev > 1 || (n_dev > 1 && trials < 2048);
    const size_t lanes = sweep_mode ? n_dev * per_dev : n_dev;
    const size_t gen_threads = std::max<size_t>(1, cfg.THREADS_NUMBER / lanes);

    qkdldpc_params P0{};
    P0.algorithm = static_cast<int32_t>(cfg.DECODING_ALGORITHM);
    P0.max_iterations = static_cast<int32_t>(cfg.DECODING_ALG_MAX_ITERATIONS);
    P0.enable_threshold = cfg.ENABLE_DECODING_ALG_MSG_LLR_THRESHOLD ? 1 : 0;
    P0.threshold = cfg.DECODING_ALG_MSG_LLR_THRESHOLD;
    P0.message_precision = dev.message_precision;

    size_t first_sim = 0;
    for (const auto &in : sim_in) {
        const H_matrix &matrix = in.matrix;
        const std::string name = in.matrix_path.filename().string();
        const CsrGraph g = to_csr_checked(matrix, name);
        if (P0.message_precision == 32 && P0.algorithm <= 1 && matrix.n() > 65536)
            // float32 FORCED where the precision policy would pick float64. Measured on the n = 102400 codes at config
            // 100k.json's operating points (profiles/r01_g_config_parity.md): a few percent of the frames that float64 decodes
            // never converge in float32 (also in a float build of the reference), so the FER is not the reference's
            std::fprintf(stderr, "note: %s: sum-product decoding with float32 messages has an error floor on codes this long; "
                                 "the default (--precision 0) uses float64 here\n", name.c_str());
        const size_t n_comb = in.combinations.size();

        std::vector<qkdldpc_code *> codes(lanes, nullptr);   // lane l runs on device l % n_dev
        auto destroy_all = [&] { for (auto *c : codes) if (c) api.code_destroy(c); };
        qkdldpc_options opt{};
        opt.pool_bytes = dev.pool_bytes;
        {
            // one handle per lane, created concurrently: a handle costs a CUDA context on first use of its device and ~0.2 s
            // of host-side table building, which would otherwise add up over 8 devices (qkdldpc_last_error is thread-local)
            std::vector<std::string> create_err(lanes);
            std::vector<std::thread> creators;
            for (size_t l = 0; l < lanes; ++l)
                creators.emplace_back([&, l] {
                    const int device = dev.devices.empty() ? 0 : dev.devices[l % n_dev];
                    if (api.code_create(&codes[l], g.n, g.m, static_cast<int64_t>(g.col_idx.size()), g.row_ptr.data(), g.col_idx.data(), device, &opt) != 0) {
                        create_err[l] = api.last_error();
                        if (create_err[l].empty()) create_err[l] = "unknown error";
                    }
                });
            for (auto &t : creators) t.join();
            for (size_t l = 0; l < lanes; ++l)
                if (!create_err[l].empty()) {
                    destroy_all();
                    throw std::runtime_error("qkdldpc_code_create failed for " + name + ": " + create_err[l]);
                }
        }

        // Trials of one combination sharded over several DISTINCT devices: the per-device tallies are summed by ONE
        // ncclAllReduce per combination over NVLink (the handles own the communicator, include/qkdldpc.h); the same device
        // listed twice, or a missing NCCL library, falls back to adding the vectors on the host (same integers).
        bool nccl_reduce = false;
        if (!sweep_mode && n_dev > 1 && api.comm_init_all && api.tally_allreduce) {
            std::vector<int> seen;
            bool distinct = true;
            for (size_t d = 0; d < n_dev; ++d) {
                distinct = distinct && std::find(seen.begin(), seen.end(), dev.devices[d]) == seen.end();
                seen.push_back(dev.devices[d]);
            }
            if (distinct) {
                nccl_reduce = api.comm_init_all(codes.data(), static_cast<int32_t>(n_dev)) == 0;
                if (!nccl_reduce && dev.verbose) std::fprintf(stderr, "note: tallies are summed on the host (%s)\n", api.last_error());
            }
        }
        std::vector<std::vector<uint64_t>> totals(n_comb, std::vector<uint64_t>(tally_len, 0));
        std::vector<double> acc_qber(n_comb, 0.), comb_ms(n_comb, 0.);
        std::vector<std::string> errors(lanes);
        auto params_of = [&](const sim_combination &comb) {
            qkdldpc_params P = P0;
            P.primary = comb.scaling_factors.primary;
            P.secondary = comb.scaling_factors.secondary;
            return P;
        };

        if (sweep_mode) {
            // Few trials per combination. With device-side input generation whole batches of combinations go through ONE
            // generate + ONE decode launch (qkdldpc_run_trials_multi); otherwise (host-generated inputs) the lanes pull
            // single combinations from a queue and decode them concurrently.
            const bool multi = !dev.host_keygen && api.run_trials_multi != nullptr;
            const size_t per_call = multi ? std::max<size_t>(1, std::min<size_t>(512, (size_t)65536 / std::max<size_t>(trials, 1))) : 1;
            std::atomic<size_t> next{0};
            std::vector<std::thread> workers;
            for (size_t l = 0; l < lanes; ++l)
                workers.emplace_back([&, l] {
                    try {
                        for (size_t c0 = next.fetch_add(per_call); c0 < n_comb; c0 = next.fetch_add(per_call)) {
                            const size_t cnt = std::min(per_call, n_comb - c0);
                            if (!multi) {
                                const sim_combination &comb = in.combinations[c0];
                                const range_outcome o = decode_trial_range(cfg, api, codes[l], params_of(comb), matrix, comb, seeds, first_sim + c0,
                                                                           0, trials, chunk, gen_threads, dev.host_keygen, totals[c0]);
                                acc_qber[c0] = o.accurate_qber;
                                comb_ms[c0] = o.ms;
                                continue;
                            }
                            std::vector<qkdldpc_combination> table(cnt);
                            for (size_t k = 0; k < cnt; ++k) fill_combination(cfg, in.combinations[c0 + k], first_sim + c0 + k, table[k]);
                            std::vector<uint64_t> tl(cnt * tally_len);
                            std::vector<double> acc(cnt);
                            const auto t0 = std::chrono::steady_clock::now();
                            const int rc = api.run_trials_multi(codes[l], &P0, static_cast<int32_t>(cnt), table.data(), static_cast<int64_t>(trials),
                                                                seeds.data(), nullptr, nullptr, tl.data(), acc.data());
                            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                            if (rc != 0) {
                                const std::string msg = api.last_error();
                                throw std::runtime_error(msg.rfind("Key size", 0) == 0 ? msg : "qkdldpc_run_trials_multi: " + msg);
                            }
                            for (size_t k = 0; k < cnt; ++k) {
                                std::copy(tl.begin() + k * tally_len, tl.begin() + (k + 1) * tally_len, totals[c0 + k].begin());
                                acc_qber[c0 + k] = acc[k];
                                comb_ms[c0 + k] = ms / static_cast<double>(cnt);   // the call's time, shared equally
                            }
                        }
                    } catch (const std::exception &e) {
                        errors[l] = e.what();
                        next.store(n_comb);   // stop the other workers
                    }
                });
            for (auto &w : workers) w.join();
        } else {
            for (size_t ci = 0; ci < n_comb && errors[0].empty(); ++ci) {
                const sim_combination &comb = in.combinations[ci];
                std::vector<std::vector<uint64_t>> part(n_dev, std::vector<uint64_t>(tally_len, 0));
                std::vector<range_outcome> outc(n_dev);
                std::vector<std::thread> workers;
                for (size_t d = 0; d < n_dev; ++d)
                    workers.emplace_back([&, d] {
                        try {
                            const size_t lo = trials * d / n_dev, hi = trials * (d + 1) / n_dev;   // contiguous trial range
                            outc[d] = decode_trial_range(cfg, api, codes[d], params_of(comb), matrix, comb, seeds, first_sim + ci, lo, hi, chunk,
                                                         gen_threads, dev.host_keygen, part[d]);
                        } catch (const std::exception &e) {
                            errors[d] = e.what();
                        }
                        // K5: every device thread enters the collective (also after an error: the others would wait for ever)
                        if (nccl_reduce && api.tally_allreduce(codes[d], part[d].data(), static_cast<int64_t>(tally_len)) != 0 && errors[d].empty())
                            errors[d] = std::string("qkdldpc_tally_allreduce: ") + api.last_error();
                    });
                for (auto &w : workers) w.join();
                if (nccl_reduce) {
                    totals[ci] = part[0];   // every rank holds the sum
                    for (size_t d = 1; d < n_dev; ++d)
                        if (part[d] != part[0] && errors[d].empty()) errors[d] = "tally all-reduce: ranks disagree";
                } else {
                    for (size_t d = 0; d < n_dev; ++d)   // integer sums of the per-device tallies on the host
                        for (size_t k = 0; k < tally_len; ++k) totals[ci][k] += part[d][k];
                }
                for (size_t d = 0; d < n_dev; ++d) {
                    comb_ms[ci] = std::max(comb_ms[ci], outc[d].ms);
                }
                acc_qber[ci] = outc[0].accurate_qber;
                for (size_t d = 1; d < n_dev; ++d)
                    if (!errors[d].empty()) errors[0] = errors[d];
            }
        }
        destroy_all();
        for (const auto &e : errors)
            if (!e.empty()) throw std::runtime_error(e);

        for (size_t ci = 0; ci < n_comb; ++ci) {
            const sim_combination &comb = in.combinations[ci];
            const size_t curr_sim = first_sim + ci;
            sim_result &r = results[curr_sim];
            r.sim_number = curr_sim;
            r.matrix_filename = name;
            r.is_regular = matrix.is_regular;
            r.num_bit_nodes = matrix.n();
            r.num_check_nodes = matrix.m();
            r.delta = comb.matrix_params.delta;
            r.efficiency = comb.matrix_params.efficiency;
            r.punctured_fraction = comb.matrix_params.punctured_fraction;
            r.shortened_fraction = comb.matrix_params.shortened_fraction;
            r.adapted_code_rate = comb.matrix_params.adapted_code_rate;
            r.config_QBER = comb.config_QBER;
            r.accurate_QBER = acc_qber[ci];
            r.scaling_factors = comb.scaling_factors;
            process_tally(totals[ci].data(), cfg.DECODING_ALG_MAX_ITERATIONS, trials, r);

            // Throughput columns: the reference times every single-frame CPU call (simulation.cpp:559-568); a batched
            // GPU call has no per-trial time, so every trial is attributed the batch's mean time per frame.
            r.gpu_ms = comb_ms[ci];
            const double out_key_length = (cfg.ENABLE_CODE_RATE_ADAPTATION || cfg.ENABLE_PRIVACY_MAINTENANCE)
                                              ? static_cast<double>(matrix.n() - comb.matrix_params.bits_to_remove.size())
                                              : static_cast<double>(matrix.n());
            r.out_key_length = static_cast<size_t>(out_key_length);
            r.gpu_gbit_s = r.gpu_ms > 0 ? out_key_length * static_cast<double>(trials) / (r.gpu_ms * 1e-3) / 1e9 : 0.;
            if (cfg.ENABLE_THROUGHPUT_MEASUREMENT && r.gpu_ms > 0) {
                double us_per_frame = r.gpu_ms * 1e3 / static_cast<double>(trials);
                if (cfg.CONSIDER_RTT) us_per_frame += cfg.RTT * 1000.;
                const double thr = out_key_length * 1e6 / us_per_frame;   // bits/s
                r.throughput_mean = r.throughput_min = r.throughput_max = static_cast<size_t>(thr);
                r.throughput_std_dev = 0;
            }
            if (dev.verbose)
                std::fprintf(stderr, "[%zu/%zu] %s QBER=%.4f: FER=%.6f, mean it %.2f, %.1f ms, %.3f Gbit/s\n", curr_sim + 1, sim_total,
                             name.c_str(), comb.config_QBER, 1. - r.ratio_trials_success_ldpc, r.iter_success_dec_alg_mean, r.gpu_ms,
                             r.gpu_gbit_s);
        }
        first_sim += n_comb;
    }
    return results;
}

// ---------------------------------------------------------------------------------------------------------------
// Results file
// ---------------------------------------------------------------------------------------------------------------

static void comma(std::string &s) {
    for (char &c : s)
        if (c == '.') c = ',';
}

// fmt "{:.{d}Lf}": correctly rounded fixed notation (printf does the same), decimal comma.
std::string format_fixed_comma(double value, int decimals) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.*f", decimals, value);
    std::string s = buf;
    comma(s);
    return s;
}

// fmt "{:L}" of a double: shortest round-trip digits; fixed notation while -4 <= exp10 < 16, else d.ddde±XX.
std::string format_shortest_comma(double value) {
    if (std::isnan(value)) return std::signbit(value) ? "-nan" : "nan";
    if (std::isinf(value)) return value < 0 ? "-inf" : "inf";
    if (value == 0.) return std::signbit(value) ? "-0" : "0";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, std::fabs(value), std::chars_format::scientific);
    std::string sci(buf, res.ptr);                          // d[.ddd]e±XX
    const size_t epos = sci.find('e');
    const int exp10 = std::stoi(sci.substr(epos + 1));
    std::string digits;
    for (size_t i = 0; i < epos; ++i)
        if (sci[i] != '.') digits.push_back(sci[i]);
    std::string out = value < 0 ? "-" : "";
    const int nd = static_cast<int>(digits.size());
    if (exp10 < -4 || exp10 >= 16) {
        out += digits[0];
        if (nd > 1) out += "." + digits.substr(1);
        char e[16];
        std::snprintf(e, sizeof e, "e%c%02d", exp10 < 0 ? '-' : '+', std::abs(exp10));
        out += e;
    } else if (exp10 >= 0) {
        if (nd <= exp10 + 1) out += digits + std::string(static_cast<size_t>(exp10 + 1 - nd), '0');
        else out += digits.substr(0, static_cast<size_t>(exp10) + 1) + "." + digits.substr(static_cast<size_t>(exp10) + 1);
    } else {
        out += "0." + std::string(static_cast<size_t>(-exp10 - 1), '0') + digits;
    }
    comma(out);
    return out;
}

static bool has_one_factor(size_t alg) { return alg == DEC_NMSA || alg == DEC_OMSA; }
static bool has_two_factors(size_t alg) { return alg == DEC_ANMSA || alg == DEC_AOMSA; }

std::string results_base_filename(const config_data &cfg, const std::string &sim_duration) {
    std::string rate_adapt = "OFF";
    if (cfg.ENABLE_CODE_RATE_ADAPTATION) rate_adapt = std::string("ON[punct=") + (cfg.ENABLE_UNTAINTED_PUNCTURING ? "untainted]" : "random]");
    std::string rtt;
    if (cfg.ENABLE_THROUGHPUT_MEASUREMENT && cfg.CONSIDER_RTT) {
        char b[64];
        std::snprintf(b, sizeof b, ",RTT=%.3fms", cfg.RTT);
        rtt = b;
    }
    return "ldpc(trial_num=" + std::to_string(cfg.TRIALS_NUMBER) + ",dec_alg=" + decoding_algorithm_name(cfg.DECODING_ALGORITHM) +
           ",max_dec_alg_iters=" + std::to_string(cfg.DECODING_ALG_MAX_ITERATIONS) +
           ",priv_maint=" + (cfg.ENABLE_PRIVACY_MAINTENANCE ? "ON" : "OFF") + ",rate_adapt=" + rate_adapt + rtt +
           ",seed=" + std::to_string(cfg.SIMULATION_SEED) + ",sim_duration=" + sim_duration + ")";
}

std::string csv_header(const config_data &cfg) {
    std::string h = "#;MATRIX_FILENAME;TYPE;R;M;N;CONFIG_QBER;ACCURATE_QBER;ITER_SUCCESS_MEAN;ITER_SUCCESS_STD;ITER_SUCCESS_MIN;"
                    "ITER_SUCCESS_MAX;RATIO_SUCCESS_DEC;RATIO_SUCCESS_LDPC;FER";
    if (cfg.ENABLE_CODE_RATE_ADAPTATION) h += ";DELTA;EFFICIENCY;PUNCT_FRACTION;SHORT_FRACTION;R_ADAPTED";
    if (cfg.ENABLE_THROUGHPUT_MEASUREMENT) h += ";THROUGHPUT_MEAN;THROUGHPUT_STD;THROUGHPUT_MIN;THROUGHPUT_MAX";
    switch (cfg.DECODING_ALGORITHM) {
        case DEC_NMSA: h += ";ALPHA"; break;
        case DEC_OMSA: h += ";BETA"; break;
        case DEC_ANMSA: h += ";ALPHA;NU"; break;
        case DEC_AOMSA: h += ";BETA;SIGMA"; break;
        default: break;
    }
    return h;
}

std::string csv_line(const config_data &cfg, const sim_result &r) {
    const double T = static_cast<double>(cfg.TRIALS_NUMBER);
    const double FER = std::round((1. - r.ratio_trials_success_ldpc) * T) / T;   // simulation.cpp:117-118
    std::string s = std::to_string(r.sim_number) + ";" + r.matrix_filename + ";" + (r.is_regular ? "regular" : "irregular") + ";" +
                    format_fixed_comma(1. - static_cast<double>(r.num_check_nodes) / static_cast<double>(r.num_bit_nodes), 3) + ";" +
                    std::to_string(r.num_check_nodes) + ";" + std::to_string(r.num_bit_nodes) + ";" + format_fixed_comma(r.config_QBER, 4) +
                    ";" + format_fixed_comma(r.accurate_QBER, 4) + ";" + format_fixed_comma(r.iter_success_dec_alg_mean, 2) + ";" +
                    format_fixed_comma(r.iter_success_dec_alg_std_dev, 2) + ";" + std::to_string(r.iter_success_dec_alg_min) + ";" +
                    std::to_string(r.iter_success_dec_alg_max) + ";" + format_shortest_comma(r.ratio_trials_success_dec_alg) + ";" +
                    format_shortest_comma(r.ratio_trials_success_ldpc) + ";" + format_shortest_comma(FER);
    if (cfg.ENABLE_CODE_RATE_ADAPTATION)
        s += ";" + format_fixed_comma(r.delta, 3) + ";" + format_fixed_comma(r.efficiency, 3) + ";" + format_fixed_comma(r.punctured_fraction, 3) +
             ";" + format_fixed_comma(r.shortened_fraction, 3) + ";" + format_fixed_comma(r.adapted_code_rate, 3);
    if (cfg.ENABLE_THROUGHPUT_MEASUREMENT)
        s += ";" + std::to_string(r.throughput_mean) + ";" + std::to_string(r.throughput_std_dev) + ";" + std::to_string(r.throughput_min) +
             ";" + std::to_string(r.throughput_max);
    if (has_one_factor(cfg.DECODING_ALGORITHM) || has_two_factors(cfg.DECODING_ALGORITHM)) s += ";" + format_fixed_comma(r.scaling_factors.primary, 3);
    if (has_two_factors(cfg.DECODING_ALGORITHM)) s += ";" + format_fixed_comma(r.scaling_factors.secondary, 3);
    return s;
}

fs::path write_file(const config_data &cfg, const std::vector<sim_result> &data, const std::string &sim_duration, const fs::path &directory) {
    if (!fs::exists(directory)) fs::create_directories(directory);
    const std::string base = results_base_filename(cfg, sim_duration);
    fs::path path = directory / (base + ".csv");
    for (size_t k = 1; fs::exists(path); ++k) path = directory / (base + "_" + std::to_string(k) + ".csv");   // simulation.cpp:94-101
    std::ofstream out(path, std::ios::out | std::ios::trunc);
    if (!out) throw std::runtime_error("An error occurred while writing to the file: " + path.string());
    out << csv_header(cfg) << "\n";
    for (const auto &r : data) out << csv_line(cfg, r) << "\n";
    return path;
}

}  // namespace qkdldpc

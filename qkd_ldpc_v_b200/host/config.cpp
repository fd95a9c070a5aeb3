#include "config.hpp"

#include <algorithm>
#include <cmath>
#include <ctime>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "json.hpp"

namespace qkdldpc {

namespace {

// config.cpp:3-19
scaling_factor_range parse_scaling_factor_range(const Json &j) {
    scaling_factor_range r{j.at("begin").as_double(), j.at("end").as_double(), j.at("step").as_double()};
    if (r.begin <= 0. || r.end <= 0. || r.step <= 0.) throw std::runtime_error("Scaling factor range begin, end, step must be > 0!");
    if (r.begin > r.end) throw std::runtime_error("Scaling factor range begin cannot be larger than end!");
    if (r.begin != r.end && r.step - EPSILON > r.end - r.begin) throw std::runtime_error("Scaling factor range step is too large!");
    return r;
}

// config.cpp:21-50 (sorted ascending by code rate: lookups take the first entry with code_rate >= R, quirk Q14)
std::vector<R_scaling_factor_map> parse_scaling_factor_maps(const Json &j, const std::string &key) {
    std::vector<R_scaling_factor_map> maps;
    for (const auto &m : j.items()) {
        const double code_rate = m.at("code_rate").as_double(), sf = m.at(key).as_double();
        if (code_rate <= 0. || code_rate >= 1.) throw std::runtime_error("Code rate(R) must be: 0 < R < 1!");
        if (sf <= 0.) throw std::runtime_error("Scaling factor must be > 0!");
        maps.push_back({code_rate, sf});
    }
    if (maps.empty()) throw std::runtime_error("Array with code rate(R) and scaling factor maps is empty!");
    std::sort(maps.begin(), maps.end(), [](const R_scaling_factor_map &a, const R_scaling_factor_map &b) { return a.code_rate < b.code_rate; });
    return maps;
}

void parse_factor(const Json &params, const char *name, scaling_factor_source &dst) {
    const std::string n(name);
    dst.use_range = params.at("use_" + n + "_range").as_bool();
    if (dst.use_range) dst.range = parse_scaling_factor_range(params.at(n + "_range"));
    else dst.maps = parse_scaling_factor_maps(params.at("code_rate_" + n + "_maps"), n);
}

void check_range(double begin, double end, double step, bool unit_interval, const char *what) {
    const std::string w(what);
    if (unit_interval) {
        if (begin <= 0. || begin >= 1. || end <= 0. || end >= 1. || begin > end)
            throw std::runtime_error("Invalid " + w + " begin or end parameters. " + w + " must be: 0 < " + w + " < 1, and begin cannot be larger than end!");
    } else if (begin < 1. || end < 1. || begin > end) {
        throw std::runtime_error("Invalid efficiency begin or end parameters. Efficiency(f_EC) must be: f_EC >= 1, and begin cannot be larger than end!");
    }
    if (step <= 0.) throw std::runtime_error(w + " step must be > 0!");
    if (begin != end && step - EPSILON > end - begin) throw std::runtime_error(w + " step is too large.");
}

}  // namespace

const char *decoding_algorithm_name(size_t alg) {
    static const char *names[] = {"SPA", "SPA-LIN-APPROX", "NMSA", "OMSA", "ANMSA", "AOMSA"};
    return alg <= DEC_AOMSA ? names[alg] : "?";
}

config_data parse_config_text(const std::string &text) {
    const Json config = Json::parse(text);
    if (config.empty()) throw std::runtime_error("Configuration file is empty");
    config_data cfg{};
    cfg.schema_version = 4;

    cfg.THREADS_NUMBER = config.at("threads_number").as_size();
    if (cfg.THREADS_NUMBER < 1) throw std::runtime_error("Number of threads must be >= 1!");
    cfg.TRIALS_NUMBER = config.at("trials_number").as_size();
    if (cfg.TRIALS_NUMBER < 1) throw std::runtime_error("Number of trials must be >= 1!");
    if (config.at("use_config_simulation_seed").as_bool()) cfg.SIMULATION_SEED = config.at("simulation_seed").as_size();
    else cfg.SIMULATION_SEED = static_cast<size_t>(std::time(nullptr));

    cfg.ENABLE_PRIVACY_MAINTENANCE = config.at("enable_privacy_maintenance").as_bool();
    cfg.ENABLE_THROUGHPUT_MEASUREMENT = config.at("enable_throughput_measurement").as_bool();
    if (cfg.ENABLE_THROUGHPUT_MEASUREMENT) {
        const Json &tm = config.at("throughput_measurement_parameters");
        cfg.CONSIDER_RTT = tm.at("consider_RTT").as_bool();
        if (cfg.CONSIDER_RTT) {
            cfg.RTT = tm.at("RTT").as_double();
            if (cfg.RTT < 0.) throw std::runtime_error("Round-Trip Time (RTT) must be >= 0!");
        }
    }

    // config.cpp:138-140; v1 files carry a boolean instead (false -> SPA, true -> NMSA)
    if (config.contains("decoding_algorithm")) {
        cfg.DECODING_ALGORITHM = config.at("decoding_algorithm").as_size();
    } else if (config.contains("use_min_sum_normalized_algorithm")) {
        cfg.DECODING_ALGORITHM = config.at("use_min_sum_normalized_algorithm").as_bool() ? DEC_NMSA : DEC_SPA;
        cfg.schema_version = 1;
    } else {
        throw std::runtime_error("JSON: key 'decoding_algorithm' not found");
    }
    if (cfg.DECODING_ALGORITHM > DEC_AOMSA)
        throw std::runtime_error("Only six options are available: \n0 - SPA;\n1 - SPA (with linear approximation of tanh and atanh);\n2 - NMSA;\n3 - OMSA;\n4 - ANMSA;\n5 - AOMSA.");

    auto &dp = cfg.DECODING_ALG_PARAMS;
    if (cfg.DECODING_ALGORITHM == DEC_NMSA) {
        parse_factor(config.at("min_sum_normalized_parameters"), "alpha", dp.primary);
    } else if (cfg.DECODING_ALGORITHM == DEC_OMSA) {
        parse_factor(config.at("min_sum_offset_parameters"), "beta", dp.primary);
    } else if (cfg.DECODING_ALGORITHM == DEC_ANMSA || cfg.DECODING_ALGORITHM == DEC_AOMSA) {
        const bool norm = cfg.DECODING_ALGORITHM == DEC_ANMSA;
        const Json &p = config.at(norm ? "adaptive_min_sum_normalized_parameters" : "adaptive_min_sum_offset_parameters");
        parse_factor(p, norm ? "alpha" : "beta", dp.primary);
        parse_factor(p, norm ? "nu" : "sigma", dp.secondary);
        if (!(dp.primary.use_range || dp.secondary.use_range)) {   // config.cpp:200-235
            const std::string alg = norm ? "ANMSA" : "AOMSA", a = norm ? "alpha" : "beta", b = norm ? "nu" : "sigma";
            if (dp.primary.maps.size() != dp.secondary.maps.size())
                throw std::runtime_error(alg + ": The sizes of code_rate_" + a + "_maps and code_rate_" + b + "_maps vectors must match! (" +
                                         std::to_string(dp.primary.maps.size()) + " vs " + std::to_string(dp.secondary.maps.size()) + ")");
            for (size_t i = 0; i < dp.primary.maps.size(); ++i)
                if (std::abs(dp.primary.maps[i].code_rate - dp.secondary.maps[i].code_rate) > EPSILON)
                    throw std::runtime_error(alg + ": Mismatch of code_rate in " + a + " and " + b + " maps");
        }
    }

    cfg.DECODING_ALG_MAX_ITERATIONS = config.at("decoding_algorithm_max_iterations").as_size();
    if (cfg.DECODING_ALG_MAX_ITERATIONS < 1) throw std::runtime_error("Minimum number of decoding algorithm iterations must be >= 1!");
    cfg.MATRIX_FORMAT = config.at("matrix_format").as_size();
    if (cfg.MATRIX_FORMAT > 3) throw std::runtime_error("Only four options are available: \n0 - uncompressed;\n1 - sparse alist;\n2 - sparse_1;\n3 - sparse_2.");
    cfg.TRACE_QKD_LDPC = config.at("trace_qkd_ldpc").as_bool();
    cfg.TRACE_DECODING_ALG = config.at("trace_decoding_algorithm").as_bool();
    cfg.TRACE_DECODING_ALG_LLR = config.at("trace_decoding_algorithm_llr").as_bool();
    cfg.ENABLE_DECODING_ALG_MSG_LLR_THRESHOLD = config.at("enable_decoding_algorithm_msg_llr_threshold").as_bool();
    if (cfg.ENABLE_DECODING_ALG_MSG_LLR_THRESHOLD) {
        cfg.DECODING_ALG_MSG_LLR_THRESHOLD = config.at("decoding_algorithm_msg_llr_threshold").as_double();
        if (cfg.DECODING_ALG_MSG_LLR_THRESHOLD <= 0.) throw std::runtime_error("Sum-product message LLR threshold must be > 0!");
    }

    // config.cpp:259-293; legacy spellings: code_rate_QBER_maps, flat (v1/v2) or nested (v3)
    const bool legacy_qber = !config.contains("code_rate_QBER_ranges");
    const Json &qr = config.at(legacy_qber ? "code_rate_QBER_maps" : "code_rate_QBER_ranges");
    if (legacy_qber) cfg.schema_version = std::min(cfg.schema_version, 3);
    for (const auto &r : qr.items()) {
        R_QBER_range x{};
        x.code_rate = r.at("code_rate").as_double();
        if (r.contains("QBER")) {
            const Json &q = r.at("QBER");
            x.QBER_begin = q.at("begin").as_double(); x.QBER_end = q.at("end").as_double(); x.QBER_step = q.at("step").as_double();
        } else {
            x.QBER_begin = r.at("QBER_begin").as_double(); x.QBER_end = r.at("QBER_end").as_double(); x.QBER_step = r.at("QBER_step").as_double();
            cfg.schema_version = std::min(cfg.schema_version, 2);
        }
        cfg.R_QBER_RANGES.push_back(x);
    }
    if (cfg.R_QBER_RANGES.empty()) throw std::runtime_error("Array with code rate(R) and QBER ranges is empty!");
    for (const auto &x : cfg.R_QBER_RANGES) {
        if (x.code_rate <= 0. || x.code_rate >= 1.) throw std::runtime_error("Code rate(R) must be: 0 < R < 1!");
        check_range(x.QBER_begin, x.QBER_end, x.QBER_step, true, "QBER");
    }
    std::sort(cfg.R_QBER_RANGES.begin(), cfg.R_QBER_RANGES.end(), [](const R_QBER_range &a, const R_QBER_range &b) { return a.code_rate < b.code_rate; });

    // config.cpp:295-394; absent in v1/v2 (=> false); v3 keeps the pieces at top level
    cfg.ENABLE_CODE_RATE_ADAPTATION = config.contains("enable_code_rate_adaptation") && config.at("enable_code_rate_adaptation").as_bool();
    if (cfg.ENABLE_CODE_RATE_ADAPTATION) {
        const bool nested = config.contains("code_rate_adaptation_parameters");
        const Json &ra = nested ? config.at("code_rate_adaptation_parameters") : config;
        if (!nested) cfg.schema_version = std::min(cfg.schema_version, 3);
        cfg.ENABLE_UNTAINTED_PUNCTURING = ra.at("enable_untainted_puncturing").as_bool();
        cfg.USE_ADAPTATION_PARAMETERS_RANGES = nested ? ra.at("use_adaptation_parameters_ranges").as_bool() : true;
        if (cfg.USE_ADAPTATION_PARAMETERS_RANGES) {
            const Json &rr = ra.at(nested ? "code_rate_adaptation_parameters_ranges" : "code_rate_adaptation_parameters_maps");
            for (const auto &r : rr.items()) {
                R_adaptation_parameters_range x{};
                x.code_rate = r.at("code_rate").as_double();
                const Json &d = r.at("delta"), &e = r.at("efficiency");
                x.delta_begin = d.at("begin").as_double(); x.delta_end = d.at("end").as_double(); x.delta_step = d.at("step").as_double();
                x.efficiency_begin = e.at("begin").as_double(); x.efficiency_end = e.at("end").as_double(); x.efficiency_step = e.at("step").as_double();
                cfg.R_ADAPT_PARAMS_RANGES.push_back(x);
            }
            if (cfg.R_ADAPT_PARAMS_RANGES.empty()) throw std::runtime_error("Array with code rate(R) and adaptation parameters ranges is empty!");
            for (const auto &x : cfg.R_ADAPT_PARAMS_RANGES) {
                if (x.code_rate <= 0. || x.code_rate >= 1.) throw std::runtime_error("Code rate(R) must be: 0 < R < 1!");
                check_range(x.delta_begin, x.delta_end, x.delta_step, true, "Delta");
                check_range(x.efficiency_begin, x.efficiency_end, x.efficiency_step, false, "Efficiency");
            }
            std::sort(cfg.R_ADAPT_PARAMS_RANGES.begin(), cfg.R_ADAPT_PARAMS_RANGES.end(),
                      [](const R_adaptation_parameters_range &a, const R_adaptation_parameters_range &b) { return a.code_rate < b.code_rate; });
        } else {
            for (const auto &m : ra.at("code_rate_QBER_adaptation_parameters_maps").items()) {
                R_QBER_adaptation_parameters_map x{};
                x.code_rate = m.at("code_rate").as_double();
                x.QBER_adapt_params = {m.at("QBER").as_double(), m.at("delta").as_double(), m.at("efficiency").as_double()};
                cfg.R_QBER_ADAPT_PARAMS_MAPS.push_back(x);
            }
            if (cfg.R_QBER_ADAPT_PARAMS_MAPS.empty()) throw std::runtime_error("Array with code rate(R), QBER and adaptation parameters maps is empty!");
            for (const auto &x : cfg.R_QBER_ADAPT_PARAMS_MAPS) {
                if (x.code_rate <= 0. || x.code_rate >= 1.) throw std::runtime_error("Code rate(R) must be: 0 < R < 1!");
                if (x.QBER_adapt_params.QBER <= 0. || x.QBER_adapt_params.QBER >= 1.) throw std::runtime_error("Invalid QBER parameter. QBER must be: 0 < QBER < 1!");
                if (x.QBER_adapt_params.delta <= 0. || x.QBER_adapt_params.delta >= 1.) throw std::runtime_error("Invalid delta parameter. Delta must be: 0 < delta < 1!");
                if (x.QBER_adapt_params.efficiency < 1.) throw std::runtime_error("Invalid efficiency parameter. Efficiency(f_EC) must be: f_EC >= 1!");
            }
            std::sort(cfg.R_QBER_ADAPT_PARAMS_MAPS.begin(), cfg.R_QBER_ADAPT_PARAMS_MAPS.end(),
                      [](const R_QBER_adaptation_parameters_map &a, const R_QBER_adaptation_parameters_map &b) { return a.code_rate < b.code_rate; });
        }
    }
    return cfg;
}

config_data parse_config_data(const fs::path &config_path) {
    if (!fs::exists(config_path)) throw std::runtime_error("Configuration file not found: " + config_path.string());
    if (config_path.extension() != ".json") throw std::runtime_error("Configuration file must have a .json extension: " + config_path.string());
    std::ifstream f(config_path);
    if (!f.is_open()) throw std::runtime_error("Failed to open configuration file: " + config_path.string());
    std::stringstream ss;
    ss << f.rdbuf();
    try {
        return parse_config_text(ss.str());
    } catch (const std::exception &e) {
        throw std::runtime_error(std::string("An error occurred while reading a configuration parameter of ") + config_path.string() + ": " + e.what());
    }
}

}  // namespace qkdldpc

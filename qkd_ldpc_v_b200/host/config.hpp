// Run configuration (reference: struct config_data, src/config.hpp:103-198; parser src/config.cpp:89-403).
// Unlike the reference there is no global CFG: the struct is passed explicitly, and the fields the decoders read
// cross the C ABI as qkdldpc_params.
//
// The parser accepts the four schema generations found in the reference tree:
//   v1  use_min_sum_normalized_algorithm (false -> SPA, true -> NMSA), flat code_rate_QBER_maps{QBER_begin,...}
//   v2  decoding_algorithm + flat code_rate_QBER_maps
//   v3  code_rate_QBER_maps with nested QBER{begin,end,step}; top-level enable_untainted_puncturing and
//       code_rate_adaptation_parameters_maps (ranges form)
//   v4  what config.cpp parses today: code_rate_QBER_ranges, nested code_rate_adaptation_parameters{...}
#pragma once
#include <cstddef>
#include <filesystem>
#include <string>
#include <vector>

namespace qkdldpc {

namespace fs = std::filesystem;

inline constexpr size_t DEC_SPA = 0, DEC_SPA_APPROX = 1, DEC_NMSA = 2, DEC_OMSA = 3, DEC_ANMSA = 4, DEC_AOMSA = 5;
inline constexpr double EPSILON = 1e-6;

struct scaling_factor_range { double begin{}, end{}, step{}; };
struct R_scaling_factor_map { double code_rate{}, scaling_factor{}; };
struct scaling_factor_source {
    bool use_range{};
    scaling_factor_range range{};
    std::vector<R_scaling_factor_map> maps{};
};
struct decoding_algorithm_params { scaling_factor_source primary, secondary; };
struct decoding_scaling_factors { double primary{}, secondary{}; };
struct R_QBER_range { double code_rate{}, QBER_begin{}, QBER_end{}, QBER_step{}; };
struct R_adaptation_parameters_range {
    double code_rate{}, delta_begin{}, delta_end{}, delta_step{}, efficiency_begin{}, efficiency_end{}, efficiency_step{};
};
struct QBER_adaptation_parameters { double QBER{}, delta{}, efficiency{}; };
struct R_QBER_adaptation_parameters_map { double code_rate{}; QBER_adaptation_parameters QBER_adapt_params{}; };

struct config_data {
    size_t THREADS_NUMBER{};
    size_t TRIALS_NUMBER{};
    size_t SIMULATION_SEED{};
    bool ENABLE_PRIVACY_MAINTENANCE{};
    bool ENABLE_THROUGHPUT_MEASUREMENT{};
    bool CONSIDER_RTT{};
    double RTT{};
    size_t DECODING_ALGORITHM{};
    decoding_algorithm_params DECODING_ALG_PARAMS{};
    size_t DECODING_ALG_MAX_ITERATIONS{};
    size_t MATRIX_FORMAT{};
    bool TRACE_QKD_LDPC{}, TRACE_DECODING_ALG{}, TRACE_DECODING_ALG_LLR{};
    bool ENABLE_DECODING_ALG_MSG_LLR_THRESHOLD{};
    double DECODING_ALG_MSG_LLR_THRESHOLD{};
    std::vector<R_QBER_range> R_QBER_RANGES{};
    bool ENABLE_CODE_RATE_ADAPTATION{};
    bool ENABLE_UNTAINTED_PUNCTURING{};
    bool USE_ADAPTATION_PARAMETERS_RANGES{};
    std::vector<R_adaptation_parameters_range> R_ADAPT_PARAMS_RANGES{};
    std::vector<R_QBER_adaptation_parameters_map> R_QBER_ADAPT_PARAMS_MAPS{};
    int schema_version{};   // 1..4, which generation the file was written in (informational)
};

config_data parse_config_data(const fs::path &config_path);
config_data parse_config_text(const std::string &json_text);
const char *decoding_algorithm_name(size_t alg);   // names used in the results file name (simulation.cpp:29-63)

}  // namespace qkdldpc

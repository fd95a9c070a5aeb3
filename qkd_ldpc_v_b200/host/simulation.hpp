// Simulation driver above the C ABI (reference: src/simulation.{hpp,cpp}).
//
//   prepare_sim_inputs        simulation.cpp:371-537   per matrix: QBER list x (delta, f_EC) x scaling factors
//   QKD_LDPC_batch_simulation simulation.cpp:693-768   per combination: TRIALS frames. The reference's
//                                                      pool.detach_loop over run_trial (:740-746) is ONE batched
//                                                      qkdldpc_decode_batch call per GPU here; host threads only
//                                                      generate the trial inputs (same per-trial RNG streams).
//   process tallies           simulation.cpp:580-690   from the integer tally vector the GPUs return (exact sums)
//   write_file                simulation.cpp:4-176     the ';'-separated CSV with decimal commas
#pragma once
#include <cstdint>
#include <filesystem>
#include <string>
#include <vector>

#include "../../include/qkdldpc.h"
#include "config.hpp"
#include "matrix.hpp"
#include "rate_adapt.hpp"

namespace qkdldpc {

struct sim_combination {
    double config_QBER{};
    H_matrix_params matrix_params{};
    decoding_scaling_factors scaling_factors{};
};

struct sim_input {
    H_matrix matrix{};
    fs::path matrix_path{};
    std::vector<sim_combination> combinations;
};

// simulation.hpp:43-68
struct sim_result {
    size_t sim_number{};
    std::string matrix_filename{};
    bool is_regular{};
    size_t num_bit_nodes{}, num_check_nodes{};
    double delta{}, efficiency{}, punctured_fraction{}, shortened_fraction{}, adapted_code_rate{};
    double config_QBER{}, accurate_QBER{};
    decoding_scaling_factors scaling_factors{};
    size_t iter_success_dec_alg_max{}, iter_success_dec_alg_min{};
    double iter_success_dec_alg_mean{}, iter_success_dec_alg_std_dev{};
    double ratio_trials_success_dec_alg{}, ratio_trials_success_ldpc{};
    size_t throughput_max{}, throughput_min{}, throughput_mean{}, throughput_std_dev{};
    // GPU side-car (not in the reference's CSV): device time of the batch and decoded payload rate
    double gpu_ms{};
    double gpu_gbit_s{};
    uint64_t iterations_executed{};
    size_t out_key_length{};   // bits of the final key: n, or n - bits_to_remove (privacy maintenance / rate adaptation)
};

// Range / map lookups (simulation.cpp:182-368): first entry with code_rate >= R of an ascending list (quirk Q14).
std::vector<double> expand_range(double begin, double end, double step);
std::vector<double> get_rate_based_QBER_range(double code_rate, const std::vector<R_QBER_range> &ranges);
double get_rate_based_scaling_factor_value(double code_rate, const std::vector<R_scaling_factor_map> &maps);

// `untp_cache_dir`: where freshly generated `.untp` lists are written when the matrix directory is read-only.
std::vector<sim_input> prepare_sim_inputs(const config_data &cfg, const std::vector<fs::path> &matrix_paths,
                                          const fs::path &untp_cache_dir = {}, bool print_warnings = true);

// Statistics of one combination from the tally vector of include/qkdldpc.h (length max_iterations + 5).
// Same numbers as process_trials_results (simulation.cpp:580-624,683-689).
void process_tally(const uint64_t *tally, size_t max_iterations, size_t trials_number, sim_result &result);

struct device_options {
    std::vector<int> devices{0};     // CUDA devices to shard the trials over (contiguous trial ranges per device)
    int message_precision = 0;       // 0: the library's precision policy (include/qkdldpc.h); 32 / 64: forced
    int64_t chunk_frames = 65536;    // frames generated / uploaded per qkdldpc_decode_batch call and device
    int64_t pool_bytes = 0;          // 0 = library default
    bool host_keygen = false;        // true: generate the trial inputs on host threads and upload them (cross-check path);
                                     // false: generate them on the device, bit-identical (qkdldpc_run_trials)
    int concurrent_combinations = 0; // combinations decoded at the same time per device (own handle + stream each);
                                     // 0 = automatic: 1 for large trial counts, up to 8 for sweeps of small batches
    bool verbose = true;
};

// The C ABI entry points the driver calls (include/qkdldpc.h). qkdldpc_sim fills this with the functions of
// libqkdldpc_cuda; keeping them as pointers lets libqkdldpc_host (loaders, configs, statistics) load without CUDA.
struct decoder_api {
    decltype(&qkdldpc_code_create) code_create = nullptr;
    decltype(&qkdldpc_code_destroy) code_destroy = nullptr;
    decltype(&qkdldpc_decode_batch) decode_batch = nullptr;
    decltype(&qkdldpc_last_error) last_error = nullptr;
    decltype(&qkdldpc_run_trials) run_trials = nullptr;   // batched run_trial with inputs generated on the device
    decltype(&qkdldpc_run_trials_multi) run_trials_multi = nullptr;   // ... for several combinations in one call
    decltype(&qkdldpc_comm_init_all) comm_init_all = nullptr;         // multi-GPU: the handles' NCCL communicator ...
    decltype(&qkdldpc_tally_allreduce) tally_allreduce = nullptr;     // ... and the tally all-reduce (K5)
};

std::vector<sim_result> QKD_LDPC_batch_simulation(const config_data &cfg, const std::vector<sim_input> &sim_in,
                                                  const decoder_api &api, const device_options &dev);

// Number formatting of the results file: fmt's "{:.3Lf}" / "{:L}" with ',' as the decimal point.
std::string format_fixed_comma(double value, int decimals);
std::string format_shortest_comma(double value);
std::string results_base_filename(const config_data &cfg, const std::string &sim_duration);
std::string csv_header(const config_data &cfg);
std::string csv_line(const config_data &cfg, const sim_result &r);
fs::path write_file(const config_data &cfg, const std::vector<sim_result> &data, const std::string &sim_duration,
                    const fs::path &directory);

}  // namespace qkdldpc

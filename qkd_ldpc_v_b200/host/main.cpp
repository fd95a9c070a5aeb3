// qkdldpc_sim -- the reference's QKD_LDPC command-line driver (src/main.cpp:24-203) on top of libqkdldpc_cuda.
//
// Same directory conventions as the reference (main.cpp:6-20): every configs/*.json is run in directory order
// against every *.mtrx of the directory selected by `matrix_format`, and one CSV per config lands in results/.
// Differences: the root directory is a run-time option (the reference compiles SOURCE_DIR in), the decoder runs on
// the GPU(s), there is no "Press Enter" pause unless --wait is given, and a JSON side-car with the GPU timing of
// every combination is written next to the CSV.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>

#include "../../include/qkdldpc.h"
#include "config.hpp"
#include "simulation.hpp"

using namespace qkdldpc;

namespace {

const char *kHelp =
    "qkdldpc_sim [options]\n"
    "  --root DIR          directory holding configs/, sparse_matrices/ and results/ (default: current directory)\n"
    "  --config FILE       run this one config instead of every configs/*.json\n"
    "  --matrix-dir DIR    read *.mtrx from DIR instead of sparse_matrices/<by matrix_format>\n"
    "  --results-dir DIR   write the CSV here (default: <root>/results)\n"
    "  --untp-cache DIR    where generated .untp untainted-puncturing lists are stored when the matrix directory\n"
    "                      is read-only (default: next to the matrix)\n"
    "  --gpus N            shard the trials of every combination over CUDA devices 0..N-1 (default 1)\n"
    "  --devices a,b,...   explicit device list\n"
    "  --precision 0|32|64 0 (default): the library's policy -- float64 state for OMSA / ANMSA / AOMSA and for sum-product\n"
    "                      decoding of codes longer than 65536 bits, float32 otherwise; 32 / 64 force float32 messages /\n"
    "                      float64 messages (bit-identical to the CPU reference for the min-sum family)\n"
    "  --chunk-frames K    frames per decode call and device (default 65536)\n"
    "  --concurrent K      combinations decoded concurrently per device (own handle and stream each); default: automatic\n"
    "                      (1 for large trial counts, up to 8 for rate-adaptation sweeps of ~100 trials)\n"
    "  --trials T          override trials_number of the config\n"
    "  --host-keygen       generate the trial inputs on host threads and upload them; default: on the device, bit-identical\n"
    "  --quiet             no per-combination progress lines\n"
    "  --wait              wait for Enter before exiting, like the reference\n"
    "\n"
    "Configuration keys (JSON, schema generations v1-v4 of the reference are accepted):\n"
    "  threads_number                  host threads used to generate the trial inputs\n"
    "  trials_number                   frames per parameter combination\n"
    "  use_config_simulation_seed, simulation_seed\n"
    "  enable_privacy_maintenance      remove bits after reconciliation (affects the output key length only)\n"
    "  enable_throughput_measurement, throughput_measurement_parameters{consider_RTT, RTT}\n"
    "  decoding_algorithm              0 SPA, 1 SPA with linear tanh/atanh approximation, 2 NMSA (alpha),\n"
    "                                  3 OMSA (beta), 4 ANMSA (alpha, nu), 5 AOMSA (beta, sigma)\n"
    "  min_sum_normalized_parameters / min_sum_offset_parameters /\n"
    "  adaptive_min_sum_normalized_parameters / adaptive_min_sum_offset_parameters\n"
    "                                  scaling factors: one range for all matrices, or a code-rate map\n"
    "  decoding_algorithm_max_iterations\n"
    "  matrix_format                   0 uncompressed (matrices_uncompressed), 1 alist (matrices_alist),\n"
    "                                  2 MacKay PEG rows, 1-based (matrices_1), 3 rows+columns, 0-based (matrices_2)\n"
    "  enable_decoding_algorithm_msg_llr_threshold, decoding_algorithm_msg_llr_threshold\n"
    "  code_rate_QBER_ranges           per code rate: QBER begin/end/step (first entry with code_rate >= R is used)\n"
    "  enable_code_rate_adaptation, code_rate_adaptation_parameters{enable_untainted_puncturing,\n"
    "      use_adaptation_parameters_ranges, code_rate_adaptation_parameters_ranges,\n"
    "      code_rate_QBER_adaptation_parameters_maps}\n"
    "  trace_* keys are accepted and ignored (the GPU decoder has no trace output).\n";

fs::path matrix_dir_for(const fs::path &root, size_t format) {
    const fs::path base = root / "sparse_matrices";
    switch (format) {
        case MAT_SPARSE_UNCOMPRESSED: return base / "matrices_uncompressed";
        case MAT_SPARSE_ALIST: return base / "matrices_alist";
        case MAT_SPARSE_1: return base / "matrices_1";
        default: return base / "matrices_2";
    }
}

void print_config_info(const config_data &cfg, const std::string &name, size_t number) {
    std::printf("CONFIG #%zu: %s (schema v%d)\n", number, name.c_str(), cfg.schema_version);
    std::printf("  trials %zu, seed %zu, host threads %zu, algorithm %s, max iterations %zu, matrix format %zu\n", cfg.TRIALS_NUMBER,
                cfg.SIMULATION_SEED, cfg.THREADS_NUMBER, decoding_algorithm_name(cfg.DECODING_ALGORITHM), cfg.DECODING_ALG_MAX_ITERATIONS,
                cfg.MATRIX_FORMAT);
    std::printf("  message threshold %s (%.3f), privacy maintenance %s, rate adaptation %s%s\n",
                cfg.ENABLE_DECODING_ALG_MSG_LLR_THRESHOLD ? "on" : "off", cfg.DECODING_ALG_MSG_LLR_THRESHOLD,
                cfg.ENABLE_PRIVACY_MAINTENANCE ? "on" : "off", cfg.ENABLE_CODE_RATE_ADAPTATION ? "on" : "off",
                cfg.ENABLE_CODE_RATE_ADAPTATION ? (cfg.ENABLE_UNTAINTED_PUNCTURING ? " (untainted puncturing)" : " (random puncturing)") : "");
}

void write_sidecar(const fs::path &csv, const std::vector<sim_result> &results, const device_options &dev) {
    fs::path p = csv;
    p.replace_extension(".gpu.json");
    std::ofstream out(p);
    out << "{\"devices\": " << dev.devices.size() << ", \"message_precision\": " << dev.message_precision << ", \"combinations\": [\n";
    for (size_t i = 0; i < results.size(); ++i) {
        const auto &r = results[i];
        char buf[512];
        std::snprintf(buf, sizeof buf,
                      "  {\"sim_number\": %zu, \"matrix\": \"%s\", \"config_QBER\": %.6g, \"batch_ms\": %.6f, \"decoded_gbit_s\": %.6g, "
                      "\"out_key_length\": %zu, \"iterations_executed\": %llu}%s\n",
                      r.sim_number, r.matrix_filename.c_str(), r.config_QBER, r.gpu_ms, r.gpu_gbit_s, r.out_key_length,
                      static_cast<unsigned long long>(r.iterations_executed), i + 1 < results.size() ? "," : "");
        out << buf;
    }
    out << "]}\n";
}

}  // namespace

int main(int argc, char *argv[]) {
    bool wait = false;
    auto finish = [&](int code) {
        if (wait) {
            std::printf("Press Enter to exit...");
            std::fflush(stdout);
            std::cin.get();
        }
        return code;
    };
    try {
        fs::path root = fs::current_path(), config_file, matrix_dir, results_dir, untp_cache;
        device_options dev;
        long trials_override = 0;
        for (int i = 1; i < argc; ++i) {
            const std::string arg = argv[i];
            auto value = [&]() -> std::string {
                if (i + 1 >= argc) throw std::runtime_error("option " + arg + " needs a value");
                return argv[++i];
            };
            if (arg == "-help" || arg == "--help" || arg == "-h") {
                std::fputs(kHelp, stdout);
                return EXIT_SUCCESS;
            } else if (arg == "--root") root = value();
            else if (arg == "--config") config_file = value();
            else if (arg == "--matrix-dir") matrix_dir = value();
            else if (arg == "--results-dir") results_dir = value();
            else if (arg == "--untp-cache") untp_cache = value();
            else if (arg == "--gpus") {
                const int g = std::stoi(value());
                if (g < 1) throw std::runtime_error("--gpus must be >= 1");
                dev.devices.clear();
                for (int d = 0; d < g; ++d) dev.devices.push_back(d);
            } else if (arg == "--devices") {
                dev.devices.clear();
                std::string list = value();
                size_t pos = 0;
                while (pos <= list.size()) {
                    const size_t c = list.find(',', pos);
                    dev.devices.push_back(std::stoi(list.substr(pos, c == std::string::npos ? std::string::npos : c - pos)));
                    if (c == std::string::npos) break;
                    pos = c + 1;
                }
            } else if (arg == "--precision") dev.message_precision = std::stoi(value());
            else if (arg == "--chunk-frames") dev.chunk_frames = std::stoll(value());
            else if (arg == "--concurrent") dev.concurrent_combinations = std::stoi(value());
            else if (arg == "--trials") trials_override = std::stol(value());
            else if (arg == "--host-keygen") dev.host_keygen = true;
            else if (arg == "--quiet") dev.verbose = false;
            else if (arg == "--wait") wait = true;
            else throw std::runtime_error("unknown option " + arg + " (see --help)");
        }
        if (dev.message_precision != 0 && dev.message_precision != 32 && dev.message_precision != 64)
            throw std::runtime_error("--precision must be 0, 32 or 64");
        if (results_dir.empty()) results_dir = root / "results";

        if (qkdldpc_device_count() <= 0) throw std::runtime_error("no CUDA device available: the decoder has no CPU fallback");
        decoder_api api;
        api.code_create = &qkdldpc_code_create;
        api.code_destroy = &qkdldpc_code_destroy;
        api.decode_batch = &qkdldpc_decode_batch;
        api.last_error = &qkdldpc_last_error;
        api.run_trials = &qkdldpc_run_trials;
        api.run_trials_multi = &qkdldpc_run_trials_multi;
        api.comm_init_all = &qkdldpc_comm_init_all;
        api.tally_allreduce = &qkdldpc_tally_allreduce;

        std::vector<fs::path> config_paths;
        if (!config_file.empty()) config_paths.push_back(config_file);
        else config_paths = get_file_paths_in_directory(root / "configs", ".json");

        for (size_t i = 0; i < config_paths.size(); ++i) {
            config_data cfg = parse_config_data(config_paths[i]);
            if (trials_override > 0) cfg.TRIALS_NUMBER = static_cast<size_t>(trials_override);
            print_config_info(cfg, config_paths[i].filename().string(), i + 1);
            const fs::path mdir = matrix_dir.empty() ? matrix_dir_for(root, cfg.MATRIX_FORMAT) : matrix_dir;
            const std::vector<fs::path> matrix_paths = get_file_paths_in_directory(mdir, ".mtrx");
            const std::vector<sim_input> inputs = prepare_sim_inputs(cfg, matrix_paths, untp_cache, dev.verbose);

            const auto t0 = std::chrono::steady_clock::now();
            const std::vector<sim_result> results = QKD_LDPC_batch_simulation(cfg, inputs, api, dev);
            const long total_s = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
            char duration[64];
            std::snprintf(duration, sizeof duration, "%02ldh-%02ldm-%02lds", total_s / 3600, (total_s / 60) % 60, total_s % 60);

            const fs::path csv = write_file(cfg, results, duration, results_dir);
            write_sidecar(csv, results, dev);
            std::printf("The results are written to the file: %s\n\n", csv.string().c_str());
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ERROR: %s\n", e.what());
        return finish(EXIT_FAILURE);
    }
    std::printf("Simulations successfully completed!\n");
    return finish(EXIT_SUCCESS);
}

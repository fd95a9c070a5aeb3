// Parity-check matrix container and the four on-disk formats of the reference
// (H_matrix, array_and_matrix_operations.hpp:60-77; loaders array_and_matrix_operations.cpp:291-886;
//  format numbering config.hpp:202: 0 uncompressed, 1 alist, 2 "sparse_1", 3 "sparse_2").
#pragma once
#include <cstdint>
#include <filesystem>
#include <string>
#include <vector>

namespace qkdldpc {

namespace fs = std::filesystem;

enum MatrixFormat { MAT_SPARSE_UNCOMPRESSED = 0, MAT_SPARSE_ALIST = 1, MAT_SPARSE_1 = 2, MAT_SPARSE_2 = 3 };

struct H_matrix {
    std::vector<std::vector<int>> bit_nodes;     // per bit: its checks
    std::vector<std::vector<int>> check_nodes;   // per check: its bits
    std::vector<int> punctured_bits_untainted;   // maximal untainted puncturing set (selection order)
    bool is_regular = false;

    size_t n() const { return bit_nodes.size(); }
    size_t m() const { return check_nodes.size(); }
    double code_rate() const { return 1. - static_cast<double>(m()) / static_cast<double>(n()); }
};

H_matrix read_sparse_uncompressed_matrix(const fs::path &matrix_path);
H_matrix read_sparse_matrix_alist(const fs::path &matrix_path);
H_matrix read_sparse_matrix_1(const fs::path &matrix_path);
H_matrix read_sparse_matrix_2(const fs::path &matrix_path);
H_matrix read_matrix(const fs::path &matrix_path, int matrix_format);

// CSR of check_nodes for qkdldpc_code_create. The decoders address message slots through running cursors
// (qkd_ldpc_algorithm.cpp:54,67-69,109,116-118), which is only meaningful when both adjacency lists are ascending
// and mutually consistent (quirk Q1; true for every shipped matrix). Throws std::runtime_error otherwise.
struct CsrGraph {
    int32_t n = 0, m = 0;
    std::vector<int32_t> row_ptr, col_idx;
};
CsrGraph to_csr_checked(const H_matrix &matrix, const std::string &name_for_errors);

std::vector<fs::path> get_file_paths_in_directory(const fs::path &directory_path, const std::string &extension);

}  // namespace qkdldpc

#include "matrix.hpp"

#include <algorithm>
#include <fstream>
#include <numeric>
#include <sstream>
#include <stdexcept>

namespace qkdldpc {

namespace {

// Every loader first turns the file into rows of integers (one row per text line).
std::vector<std::vector<int>> read_int_lines(const fs::path &path) {
    std::ifstream file(path);
    if (!file.is_open()) throw std::runtime_error("Failed to open file: " + path.string());
    std::vector<std::vector<int>> rows;
    std::string line;
    while (std::getline(file, line)) {
        std::istringstream iss(line);
        std::vector<int> v;
        int x;
        while (iss >> x) v.push_back(x);
        rows.push_back(std::move(v));
    }
    if (rows.empty()) throw std::runtime_error("File is empty or cannot be read properly: " + path.string());
    return rows;
}

bool all_same_size(const std::vector<std::vector<int>> &lists) {
    for (auto &l : lists)
        if (l.size() != lists[0].size()) return false;
    return true;
}

// get_bit_nodes_from_check_nodes (array_and_matrix_operations.cpp:55-84) in O(E): ascending check order per bit.
std::vector<std::vector<int>> bit_nodes_from_check_nodes(const std::vector<std::vector<int>> &check_nodes) {
    int max_bit = 0;
    for (auto &row : check_nodes)
        for (int c : row) max_bit = std::max(max_bit, c);
    std::vector<std::vector<int>> bit_nodes(static_cast<size_t>(max_bit) + 1);
    for (size_t j = 0; j < check_nodes.size(); ++j)
        for (int c : check_nodes[j]) bit_nodes[static_cast<size_t>(c)].push_back(static_cast<int>(j));
    return bit_nodes;
}

}  // namespace

// alist (https://rptu.de/channel-codes/matrix-file-formats), array_and_matrix_operations.cpp:291-468:
// "N M" / "dv_max dc_max" / N column weights / M row weights / N lines of 1-based checks / M lines of 1-based bits.
// Zero padding is tolerated; both adjacency lists are taken from the file as they are.
H_matrix read_sparse_matrix_alist(const fs::path &path) {
    const auto v = read_int_lines(path);
    if (v.size() < 4) throw std::runtime_error("Insufficient data in the file: " + path.string());
    if (v[0].size() != 2 || v[1].size() != 2) throw std::runtime_error("Wrong sparse alist matrix format: " + path.string());
    const size_t col_num = static_cast<size_t>(v[0][0]), row_num = static_cast<size_t>(v[0][1]);
    const size_t n = v[2].size(), m = v[3].size();
    if (v.size() < 4 + n + m) throw std::runtime_error("Insufficient data in the file: " + path.string());
    if (col_num != n)
        throw std::runtime_error("Number of columns '" + std::to_string(col_num) + "' is not the same as the length of the third line '" + std::to_string(n) + "'. File: " + path.string());
    if (row_num != m)
        throw std::runtime_error("Number of rows '" + std::to_string(row_num) + "' is not the same as the length of the fourth line '" + std::to_string(m) + "'. File: " + path.string());
    auto nonzeros = [](const std::vector<int> &r) { return static_cast<int>(std::count_if(r.begin(), r.end(), [](int x) { return x != 0; })); };
    for (size_t i = 0; i < n; ++i)
        if (nonzeros(v[4 + i]) != v[2][i])
            throw std::runtime_error("Number of non-zero elements '" + std::to_string(nonzeros(v[4 + i])) + "' in the line '" + std::to_string(4 + i + 1) + "' does not match the weight in the third line '" + std::to_string(v[2][i]) + "'. File: " + path.string());
    for (size_t i = 0; i < m; ++i)
        if (nonzeros(v[4 + n + i]) != v[3][i])
            throw std::runtime_error("Number of non-zero elements '" + std::to_string(nonzeros(v[4 + n + i])) + "' in the line '" + std::to_string(4 + n + i + 1) + "' does not match the weight in the fourth line '" + std::to_string(v[3][i]) + "'. File: " + path.string());
    H_matrix h;
    h.bit_nodes.resize(n);
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < v[2][i]; ++k) h.bit_nodes[i].push_back(v[4 + i].at(static_cast<size_t>(k)) - 1);
    h.check_nodes.resize(m);
    for (size_t i = 0; i < m; ++i)
        for (int k = 0; k < v[3][i]; ++k) h.check_nodes[i].push_back(v[4 + n + i].at(static_cast<size_t>(k)) - 1);
    // "regular" only if BOTH weight sequences are constant (:370-388): every shipped alist code reports irregular
    h.is_regular = std::all_of(v[2].begin(), v[2].end(), [&](int w) { return w == v[2][0]; }) &&
                   std::all_of(v[3].begin(), v[3].end(), [&](int w) { return w == v[3][0]; });
    return h;
}

// "sparse_1" (MacKay PEG output), :478-617: N / M / max row weight / M lines of 1-based bit indices, 0 = padding.
H_matrix read_sparse_matrix_1(const fs::path &path) {
    const auto v = read_int_lines(path);
    if (v.size() < 3) throw std::runtime_error("Insufficient data in the file: " + path.string());
    if (v[0].size() != 1 || v[1].size() != 1 || v[2].size() != 1) throw std::runtime_error("Wrong sparse matrix format: " + path.string());
    const size_t col_num = static_cast<size_t>(v[0][0]), row_num = static_cast<size_t>(v[1][0]);
    const size_t max_row_weight = static_cast<size_t>(v[2][0]);
    if (v.size() < 3 + row_num) throw std::runtime_error("Insufficient data in the file: " + path.string());
    H_matrix h;
    h.check_nodes.resize(row_num);
    bool max_weight_seen = false;
    for (size_t i = 0; i < row_num; ++i) {
        const auto &row = v[3 + i];
        if (row.size() > max_row_weight && !row.empty())
            throw std::runtime_error("Actual weight '" + std::to_string(row.size()) + "' of row '" + std::to_string(3 + i) + "' exceeded the maximum specified weight '" + std::to_string(max_row_weight) + "'.");
        for (int b : row) {
            if (b < 0) throw std::runtime_error("Bit node index cannot be less than zero: " + std::to_string(b) + ", row '" + std::to_string(3 + i) + "'.");
            if (b != 0) h.check_nodes[i].push_back(b - 1);
        }
        if (!row.empty() && row.size() == max_row_weight) max_weight_seen = true;
    }
    if (!max_weight_seen)
        throw std::runtime_error("None of the row weights matched the specified maximum weight '" + std::to_string(max_row_weight) + "'. File: " + path.string());
    h.is_regular = all_same_size(h.check_nodes);   // rows only (:589-598)
    h.bit_nodes = bit_nodes_from_check_nodes(h.check_nodes);
    if (h.bit_nodes.size() != col_num)
        throw std::runtime_error("The actual number of bit nodes '" + std::to_string(h.bit_nodes.size()) + "' did not match the specified number '" + std::to_string(col_num) + "' of bit nodes.");
    return h;
}

// "sparse_2", :626-761: "N M" / M lines of 0-based bit indices / N lines of 0-based check indices.
H_matrix read_sparse_matrix_2(const fs::path &path) {
    const auto v = read_int_lines(path);
    if (v.size() < 2) throw std::runtime_error("Insufficient data in the file: " + path.string());
    if (v[0].size() != 2) throw std::runtime_error("Wrong sparse matrix format: " + path.string());
    const size_t col_num = static_cast<size_t>(v[0][0]), row_num = static_cast<size_t>(v[0][1]);
    if (v.size() < 1 + col_num + row_num) throw std::runtime_error("Insufficient data in the file: " + path.string());
    H_matrix h;
    h.check_nodes.resize(row_num);
    for (size_t i = 0; i < row_num; ++i)
        for (int b : v[1 + i]) {
            if (b < 0) throw std::runtime_error("Bit node index cannot be less than zero: " + std::to_string(b) + ", row '" + std::to_string(1 + i) + "'.");
            h.check_nodes[i].push_back(b);
        }
    h.bit_nodes.resize(col_num);
    for (size_t i = 0; i < col_num; ++i)
        for (int c : v[1 + row_num + i]) {
            if (c < 0) throw std::runtime_error("Check node index cannot be less than zero: " + std::to_string(c) + ", row '" + std::to_string(1 + row_num + i) + "'.");
            h.bit_nodes[i].push_back(c);
        }
    h.is_regular = all_same_size(h.check_nodes) && all_same_size(h.bit_nodes);
    return h;
}

// Dense 0/1 text, :764-886: M lines x N entries; rejects other values, ragged rows, all-zero rows / columns.
H_matrix read_sparse_uncompressed_matrix(const fs::path &path) {
    const auto v = read_int_lines(path);
    for (auto &row : v) {
        for (int x : row)
            if (x != 0 && x != 1) throw std::runtime_error("Parity check matrix can only take values 0 or 1.");
        if (row.size() != v[0].size()) throw std::runtime_error("Different lengths of rows in a matrix. File: " + path.string());
    }
    const size_t n = v[0].size(), m = v.size();
    H_matrix h;
    h.bit_nodes.resize(n);
    h.check_nodes.resize(m);
    for (size_t j = 0; j < m; ++j)
        for (size_t i = 0; i < n; ++i)
            if (v[j][i]) {
                h.check_nodes[j].push_back(static_cast<int>(i));
                h.bit_nodes[i].push_back(static_cast<int>(j));
            }
    for (size_t i = 0; i < n; ++i)
        if (h.bit_nodes[i].empty()) throw std::runtime_error("Column '" + std::to_string(i + 1) + "' weight cannot be equal to zero. File: " + path.string());
    for (size_t j = 0; j < m; ++j)
        if (h.check_nodes[j].empty()) throw std::runtime_error("Row '" + std::to_string(j + 1) + "' weight cannot be equal to zero. File: " + path.string());
    h.is_regular = all_same_size(h.bit_nodes) && all_same_size(h.check_nodes);
    return h;
}

H_matrix read_matrix(const fs::path &path, int matrix_format) {
    switch (matrix_format) {
        case MAT_SPARSE_UNCOMPRESSED: return read_sparse_uncompressed_matrix(path);
        case MAT_SPARSE_ALIST: return read_sparse_matrix_alist(path);
        case MAT_SPARSE_1: return read_sparse_matrix_1(path);
        case MAT_SPARSE_2: return read_sparse_matrix_2(path);
        default: throw std::runtime_error("Only four options are available: \n0 - uncompressed;\n1 - sparse alist;\n2 - sparse_1;\n3 - sparse_2.");
    }
}

CsrGraph to_csr_checked(const H_matrix &h, const std::string &name) {
    CsrGraph g;
    g.n = static_cast<int32_t>(h.n());
    g.m = static_cast<int32_t>(h.m());
    g.row_ptr.assign(1, 0);
    std::vector<std::vector<int>> derived(h.n());
    for (size_t j = 0; j < h.m(); ++j) {
        const auto &row = h.check_nodes[j];
        for (size_t k = 0; k < row.size(); ++k) {
            if (row[k] < 0 || static_cast<size_t>(row[k]) >= h.n())
                throw std::runtime_error(name + ": bit index " + std::to_string(row[k]) + " out of range in check " + std::to_string(j));
            if (k > 0 && row[k - 1] >= row[k])
                throw std::runtime_error(name + ": check " + std::to_string(j) + " does not list its bits in strictly ascending order; "
                                         "the reference decoders pair message slots by position and give undefined pairings for such input");
            g.col_idx.push_back(row[k]);
            derived[static_cast<size_t>(row[k])].push_back(static_cast<int>(j));
        }
        g.row_ptr.push_back(static_cast<int32_t>(g.col_idx.size()));
    }
    if (derived != h.bit_nodes)
        throw std::runtime_error(name + ": bit_nodes is not the ascending transpose of check_nodes (inconsistent or unsorted column lists)");
    return g;
}

// utils.cpp:20-41: raw directory-iteration order (quirk Q15: unspecified, filesystem dependent -- kept on purpose).
std::vector<fs::path> get_file_paths_in_directory(const fs::path &dir, const std::string &extension) {
    std::vector<fs::path> out;
    if (!fs::exists(dir) || !fs::is_directory(dir)) throw std::runtime_error("Directory doesn't exist: " + dir.string());
    for (const auto &e : fs::directory_iterator(dir))
        if (e.is_regular_file() && e.path().extension() == extension) out.push_back(e.path());
    if (out.empty()) throw std::runtime_error("No files with extension '" + extension + "' in directory: " + dir.string());
    return out;
}

}  // namespace qkdldpc

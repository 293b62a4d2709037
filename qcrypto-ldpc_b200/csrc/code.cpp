// Parity-check matrix readers and table builders (host side).
//
// Replaces AFF3CT's tools::LDPC_matrix_handler::read as called by the reference drivers
// ("main.cpp (alist)":333,340; "main.cpp (5g-qc)":389) and MATLAB's `load base_matrices/NR_*.txt`
// (ML/BPSK_nrldpc_sim_FP.m:8-11).  Formats: SURVEY.md Appendix A.
#include <algorithm>
#include <fstream>
#include <numeric>
#include <sstream>

#include "qldpc_internal.hpp"

namespace qldpc {

namespace {

// next line that holds at least one integer, parsed into `vals`
bool next_int_line(std::istream &in, std::vector<long> &vals)
{
    std::string line;
    while (std::getline(in, line)) {
        vals.clear();
        std::istringstream ss(line);
        long v;
        while (ss >> v) vals.push_back(v);
        if (!vals.empty()) return true;
        if (!ss.eof() && ss.fail()) {
            // non-numeric garbage on a non-empty line
            bool blank = std::all_of(line.begin(), line.end(), [](char c) { return c == ' ' || c == '\t' || c == '\r'; });
            if (!blank) return false;
        }
    }
    return false;
}

}  // namespace

void HostCode::finalize_from_csr()
{
    edges = row_ptr.empty() ? 0 : row_ptr.back();
    max_chk_degree = 0;
    for (int c = 0; c < m; ++c) {
        std::sort(col_idx.begin() + row_ptr[c], col_idx.begin() + row_ptr[c + 1]);
        max_chk_degree = std::max(max_chk_degree, row_ptr[c + 1] - row_ptr[c]);
    }
    std::vector<int32_t> deg(n, 0);
    for (int e = 0; e < edges; ++e) deg[col_idx[e]]++;
    var_ptr.assign(n + 1, 0);
    for (int v = 0; v < n; ++v) var_ptr[v + 1] = var_ptr[v] + deg[v];
    max_var_degree = n ? *std::max_element(deg.begin(), deg.end()) : 0;
    var_edge.assign(edges, 0);
    std::fill(deg.begin(), deg.end(), 0);
    for (int c = 0; c < m; ++c)
        for (int e = row_ptr[c]; e < row_ptr[c + 1]; ++e) {
            const int v = col_idx[e];
            var_edge[var_ptr[v] + deg[v]++] = e;
        }
    k = n - m;
    if (k < 0) k = 0;
    info_pos.resize(k);
    // QC codes: systematic bits first (ML/nrldpc_encode.m:15); alist codes: parity first
    // (BOOT/matrices/G/PEGReg504x1008.alist puts the identity in columns n-k..n-1).
    std::iota(info_pos.begin(), info_pos.end(), z > 0 ? 0 : n - k);
}

bool HostCode::has_nr_core() const
{
    if (z <= 0 || base_rows < 4 || base_cols <= base_rows) return false;
    const int kb = base_cols - base_rows;
    auto B = [&](int r, int c) { return base[r * base_cols + c]; };
    // double diagonal on rows 0..3 / columns kb+1..kb+3, and a weight-3 first parity column
    for (int i = 0; i < 3; ++i)
        if (B(i, kb + 1 + i) != 0 || B(i + 1, kb + 1 + i) != 0) return false;
    if (B(0, kb) < 0 || B(3, kb) < 0) return false;
    if ((B(1, kb) < 0) == (B(2, kb) < 0)) return false;
    // extension part: identity diagonal
    for (int r = 4; r < base_rows; ++r)
        for (int c = kb + 4; c < base_cols; ++c)
            if ((B(r, c) >= 0) != (c - kb == r) || (c - kb == r && B(r, c) != 0)) return false;
    for (int r = 0; r < 4; ++r)
        for (int c = kb + 4; c < base_cols; ++c)
            if (B(r, c) >= 0) return false;
    return true;
}

int parse_alist(const std::string &path, HostCode &out)
{
    std::ifstream in(path);
    if (!in) return QLDPC_ERR_IO;
    std::vector<long> v;
    if (!next_int_line(in, v) || v.size() < 2) return QLDPC_ERR_FORMAT;
    const long n = v[0], m = v[1];
    if (n <= 0 || m <= 0 || n > (1 << 24) || m > (1 << 24)) return QLDPC_ERR_FORMAT;
    if (!next_int_line(in, v) || v.size() < 2 || v[0] <= 0 || v[1] <= 0) return QLDPC_ERR_FORMAT;
    std::vector<long> vdeg, cdeg;
    if (!next_int_line(in, vdeg) || (long)vdeg.size() != n) return QLDPC_ERR_FORMAT;
    if (!next_int_line(in, cdeg) || (long)cdeg.size() != m) return QLDPC_ERR_FORMAT;

    std::vector<std::vector<int32_t>> rows(m);
    for (long var = 0; var < n; ++var) {
        if (!next_int_line(in, v)) return QLDPC_ERR_FORMAT;
        long d = 0;
        for (long c1 : v) {
            if (c1 == 0) continue;  // zero padding up to the maximum degree
            if (c1 < 1 || c1 > m) return QLDPC_ERR_FORMAT;
            rows[c1 - 1].push_back((int32_t)var);
            ++d;
        }
        if (d != vdeg[var]) return QLDPC_ERR_FORMAT;
    }
    // the check-major half of the file must describe the same graph
    for (long c = 0; c < m; ++c) {
        if (!next_int_line(in, v)) return QLDPC_ERR_FORMAT;
        std::vector<int32_t> lst;
        for (long v1 : v) {
            if (v1 == 0) continue;
            if (v1 < 1 || v1 > n) return QLDPC_ERR_FORMAT;
            lst.push_back((int32_t)(v1 - 1));
        }
        std::sort(lst.begin(), lst.end());
        std::vector<int32_t> mine = rows[c];
        std::sort(mine.begin(), mine.end());
        if ((long)lst.size() != cdeg[c] || lst != mine) return QLDPC_ERR_FORMAT;
    }
    out = HostCode();
    out.n = (int)n;
    out.m = (int)m;
    out.row_ptr.assign(m + 1, 0);
    for (long c = 0; c < m; ++c) out.row_ptr[c + 1] = out.row_ptr[c] + (int32_t)rows[c].size();
    out.col_idx.reserve(out.row_ptr.back());
    for (auto &r : rows) out.col_idx.insert(out.col_idx.end(), r.begin(), r.end());
    out.finalize_from_csr();
    return QLDPC_OK;
}

// Expansion convention (ML/mul_sh.m:9, ML/check_cword.m:12): check lane i of block-row r is
// connected to variable lane (i + shift) mod z of block-column c.
int build_qc(const int32_t *base, int rows, int cols, int z, HostCode &out)
{
    if (!base || rows <= 0 || cols <= 0 || z <= 0) return QLDPC_ERR_ARG;
    if ((long long)cols * z > (1 << 24) || (long long)rows * z > (1 << 24)) return QLDPC_ERR_ARG;
    out = HostCode();
    out.z = z;
    out.base_rows = rows;
    out.base_cols = cols;
    out.base.resize((size_t)rows * cols);
    for (int i = 0; i < rows * cols; ++i) out.base[i] = base[i] < 0 ? -1 : base[i] % z;  // test2.qc holds shifts >= z
    out.n = cols * z;
    out.m = rows * z;
    out.row_ptr.reserve((size_t)out.m + 1);
    out.row_ptr.push_back(0);
    for (int r = 0; r < rows; ++r)
        for (int i = 0; i < z; ++i) {
            for (int c = 0; c < cols; ++c) {
                const int s = out.base[r * cols + c];
                if (s >= 0) out.col_idx.push_back(c * z + (i + s) % z);
            }
            out.row_ptr.push_back((int32_t)out.col_idx.size());
        }
    out.finalize_from_csr();
    return QLDPC_OK;
}

// ".qc": first line "cols rows z", then `rows` lines of `cols` shifts (BOOT/matrices/H/NR_1_1_192.qc:1-3)
int parse_qc(const std::string &path, HostCode &out)
{
    std::ifstream in(path);
    if (!in) return QLDPC_ERR_IO;
    std::vector<long> v;
    if (!next_int_line(in, v) || v.size() != 3) return QLDPC_ERR_FORMAT;
    const long cols = v[0], rows = v[1], z = v[2];
    if (cols <= 0 || rows <= 0 || z <= 0 || cols > 4096 || rows > 4096) return QLDPC_ERR_FORMAT;
    std::vector<int32_t> base;
    base.reserve(rows * cols);
    for (long r = 0; r < rows; ++r) {
        if (!next_int_line(in, v) || (long)v.size() != cols) return QLDPC_ERR_FORMAT;
        for (long x : v) base.push_back((int32_t)x);
    }
    return build_qc(base.data(), (int)rows, (int)cols, (int)z, out);
}

}  // namespace qldpc

// Internal host-side types of libqldpc_b200 (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/qldpc.h"

namespace qldpc {

// ---- parity-check matrix on the host ------------------------------------------------------
struct HostCode {
    int n = 0, m = 0, k = 0, edges = 0;
    // CSR by check, variables ascending inside a check (edge id = position in col_idx)
    std::vector<int32_t> row_ptr, col_idx;
    // CSC by variable: var_ptr[n+1], var_edge[e] = CSR edge id, checks ascending inside a variable
    std::vector<int32_t> var_ptr, var_edge;
    // quasi-cyclic description (z == 0 for alist / CSR codes)
    int z = 0, base_rows = 0, base_cols = 0;
    std::vector<int32_t> base;  // base_rows*base_cols, -1 = zero block, else shift in [0,z)
    std::vector<int32_t> info_pos;
    int max_chk_degree = 0, max_var_degree = 0;

    void finalize_from_csr();                   // sorts rows, builds CSC + degree stats, default info_pos
    bool has_nr_core() const;                   // NR double-diagonal parity structure (for the encoder)
};

int parse_alist(const std::string &path, HostCode &out);
int parse_qc(const std::string &path, HostCode &out);
int build_qc(const int32_t *base, int rows, int cols, int z, HostCode &out);

// ---- device-side tables ---------------------------------------------------------------------
// One entry per non-zero block of the base matrix, row-major, columns ascending: the layer
// schedule of ML/BPSK_nrldpc_sim_FP.m:45-47.
// layered int8 kernel tables ------------------------------------------------------------------
struct Li8Edge {         // 32 bytes, read as two int4 broadcasts
    int32_t off0;        // byte offset of the belief word for a lane that does not wrap
    int32_t off1;        // ... for a lane that wraps past the end of the column
    int32_t thresh;      // lane index (in words) from which the wrap happens
    int32_t hdw;         // word offset of the column's hard-decision bit vector (col * ZW32)
    int32_t selA0;       // PRMT selector (bytes -> half2 pair A) without wrap
    int32_t selA1;       // ... with wrap
    int32_t selW0;       // PRMT selector packing two half2 pairs back into a belief word, no wrap
    int32_t selW1;       // ... with wrap
};
struct Li8Layer {        // 24 bytes
    int32_t edge_begin;  // first edge of the block row (edges are row-major, columns ascending)
    int16_t degree;      // all edges of the row
    int16_t n_core;      // edges whose messages are stored (degree - has_ext)
    int32_t r_off;       // index of the row's first stored message (units of W words), -1 for register rows
    int16_t reg_idx;     // 0..3: messages live in registers, -1 otherwise
    int16_t has_ext;     // last edge goes to a weight-1, shift-0 column: no stored message, no belief update
    int32_t g_off;       // streamed mode: word offset of the row's messages in the per-frame scratch
    int16_t st;          // streamed mode: words per thread in that row (n_core rounded up to 4)
    int16_t pad;
};
// streamed layered int8 kernel (layered_i8s.cu): one table blob copied to shared memory
struct Li8sRow {         // 32 bytes, read as two int4 broadcasts
    int32_t e_off;       // byte offset (in the blob) of the row's first core-edge entry pair (2 x 16 bytes per edge:
                         //   {belief byte offset, selA, selB, selW} without and with wrap)
    int32_t thr_off;     // byte offset of the row's wrap thresholds (int32 each, 16-byte aligned)
    int32_t g_off;       // BYTE offset of the row's messages in the per-frame scratch ([nv][W] uint4)
    int32_t pack;        // nc | nv << 8 | variant << 16: core edges, uint4 per thread = ceil(nc / 4), code variant
    int32_t ext_src;     // byte offset of the row's extension column in a frame of LLRs, -1 if none
    int32_t ext_hd;      // BYTE offset of that column's (doubled) hard-decision vector
    int32_t syn_off;     // byte offset of the row's syndrome-phase entries (one u32 per edge incl. the extension edge)
    int32_t deg;         // all edges
};
struct Li8sCol {         // 8 bytes: a core block column
    int32_t llr_off;     // byte offset of the column in a frame of LLRs (col * Z)
    int32_t hd_off;      // word offset of its (doubled) hard-decision vector (col * 2 * ZW32)
};
struct Li8sGeo {         // launch geometry of the streamed kernel; index 0: no syndrome input, 1: with syndrome rows
    int tab_bytes = 0, off_rows = 0, off_pcols = 0, n_pack = 0;
    int rg_u4 = 0, stage_bytes = 0, off_ring = 0, off_ext = 0, off_hd = 0, off_syn = 0;
    int slot_bytes[2] = {0, 0}, slots[2] = {0, 0}, off_mbar[2] = {0, 0}, off_stg[2] = {-1, -1};
    // bit input (LLR synthesis inside the kernel): the staging buffer holds N/8 bytes of packed bits instead of the core
    // LLRs, and one N-byte magnitude table per CTA follows the slots; bits_ok: that layout fits with the same slot count
    int bits_slot_bytes[2] = {0, 0}, bits_off_stg[2] = {-1, -1}, bits_off_magtab[2] = {0, 0}, bits_smem[2] = {0, 0};
    bool bits_ok[2] = {false, false};
};
// generic QC tables (syndrome phase of the int8 kernel, generic layered kernel)
struct QcEdgeAux {       // 8 bytes
    int32_t hdw;         // col * ZW32
    int16_t col;         // block column
    int16_t shift;       // shift in [0,z)
};
struct QcLayer {
    int32_t edge_begin;
    int32_t degree;
};

struct CudaCheck {
    static thread_local std::string last;
    static int fail(cudaError_t e, const char *what);
};
#define QLDPC_CUDA(expr)                                                         \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) return ::qldpc::CudaCheck::fail(_e, #expr);       \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    int ensure(size_t count) {
        if (count <= n) return QLDPC_OK;
        release();
        if (cudaMalloc(&p, count * sizeof(T)) != cudaSuccess) { p = nullptr; return QLDPC_ERR_NOMEM; }
        n = count;
        return QLDPC_OK;
    }
    int upload(const std::vector<T> &v) {
        if (int r = ensure(v.size() ? v.size() : 1)) return r;
        if (!v.empty() && cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
            return QLDPC_ERR_CUDA;
        return QLDPC_OK;
    }
};

// device statistics block, layout shared with the kernels
struct DevStats {
    unsigned long long frames, failures, iter_sum;
    unsigned long long hist[QLDPC_ITER_HIST_BINS];
};

}  // namespace qldpc

// ---- opaque handles of the C ABI -------------------------------------------------------------
struct qldpc_code {
    qldpc::HostCode h;
};

struct qldpc_decoder {
    qldpc::HostCode code;
    qldpc_decoder_config cfg{};
    int kernel_family = 0;            // see KF_* in api.cpp
    const char *kernel_name = "none";
    int sm_count = 0;
    int max_smem_optin = 0;
    int norm_eighths = 8;             // integer NMS factor k/8
    int offset_int = 0;
    int out_words = 0, syn_words = 0, cw_words = 0, out_bits = 0;
    bool info_is_prefix = true;       // info bits are [0,k): kernels write them directly
    uint64_t kernel_launches = 0;

    // tables
    qldpc::DevBuf<qldpc::Li8Edge> d_li8_edges;
    qldpc::DevBuf<qldpc::Li8Layer> d_li8_layers;
    qldpc::DevBuf<uint16_t> d_li8_pack_cols;
    qldpc::DevBuf<qldpc::QcEdgeAux> d_qc_aux;
    qldpc::DevBuf<qldpc::QcLayer> d_qc_layers;
    qldpc::DevBuf<int32_t> d_qc_col_ptr;          // QC flooding: per block column, its edges in ascending block-row order
    qldpc::DevBuf<int2> d_qc_col_edges;
    bool flood_qc = false;
    qldpc::DevBuf<int32_t> d_row_ptr, d_col_idx, d_var_ptr, d_var_edge, d_info_pos;
    qldpc::DevBuf<qldpc::DevStats> d_stats;
    // scratch (grown on demand)
    qldpc::DevBuf<uint8_t> d_scratch;       // messages / beliefs that do not fit on chip
    qldpc::DevBuf<uint32_t> d_allbits;      // all-n hard decisions when info bits must be gathered
    // staging for the host-pointer entry points
    qldpc::DevBuf<uint8_t> d_in, d_out;
    qldpc::DevBuf<uint32_t> d_syn;
    // layered int8 launch geometry
    int li8_slots = 0, li8_tpg = 0, li8_smem = 0, li8_regdc = 0, li8_n_store = 0, li8_n_pack = 0;
    bool li8_ext = false;
    bool li8_stream = false;          // messages streamed through an L2-resident scratch (4 frames per SM)
    int li8_rg_words = 0;             // scratch words per frame slot
    qldpc::DevBuf<uint32_t> d_li8_rg;
    // streamed layered int8 kernel (layered_i8s.cu)
    bool li8s = false;
    qldpc::Li8sGeo li8s_geo;
    qldpc::DevBuf<uint8_t> d_li8s_tab;
    qldpc::DevBuf<uint4> d_li8s_rg;
    size_t l2_persist_bytes = 0, l2_window_max = 0;   // L2 set-aside for the message scratch (0: off)
};

// C ABI of libqldpc_b200 (see include/qldpc.h for the reference interface each entry point replaces).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <vector>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>

#include "kernels.hpp"

namespace qldpc {

thread_local std::string CudaCheck::last;

int CudaCheck::fail(cudaError_t e, const char *what)
{
    last = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();   // clear the sticky-less error state
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? QLDPC_ERR_NO_DEVICE : QLDPC_ERR_CUDA;
}

namespace {

enum { KF_NONE = 0, KF_LAYERED_I8 = 1, KF_LAYERED_GENERIC = 2, KF_FLOODING = 3, KF_LAYERED_I8S = 4, KF_LAYERED_CSR = 5 };

int dtype_size(int dtype) { return dtype == QLDPC_DTYPE_F32 ? 4 : (dtype == QLDPC_DTYPE_I16 ? 2 : 1); }

// selector that rotates the packing bytes {0,2,4,6} left by rho byte positions
uint16_t pack_selector(int rho)
{
    static const int src[4] = {0, 2, 4, 6};
    unsigned sel = 0;
    for (int m = 0; m < 4; ++m) sel |= (unsigned)src[(m - rho) & 3] << (4 * m);
    return (uint16_t)sel;
}
uint16_t unpack_selector(int rho)   // bytes rho, rho+1 -> low bytes of the two halves; high bytes zero
{
    return (uint16_t)((rho & 3) | (4 << 4) | (((rho + 1) & 3) << 8) | (4 << 12));
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// generic QC tables (generic layered kernel, syndrome phase of the int8 kernel)
int build_qc_tables(qldpc_decoder *d)
{
    const HostCode &c = d->code;
    std::vector<QcEdgeAux> aux;
    std::vector<QcLayer> layers;
    const int ZW32 = (c.z + 31) / 32;
    for (int r = 0; r < c.base_rows; ++r) {
        QcLayer ly{(int32_t)aux.size(), 0};
        for (int col = 0; col < c.base_cols; ++col) {
            const int s = c.base[r * c.base_cols + col];
            if (s < 0) continue;
            aux.push_back(QcEdgeAux{col * ZW32, (int16_t)col, (int16_t)s});
            ++ly.degree;
        }
        layers.push_back(ly);
    }
    if (int r = d->d_qc_aux.upload(aux)) return r;
    if (int r = d->d_qc_layers.upload(layers)) return r;
    // column view for the QC flooding kernel: edges of a block column in ascending block-row order
    std::vector<int32_t> col_ptr(c.base_cols + 1, 0);
    std::vector<int2> col_edges;
    for (int col = 0; col < c.base_cols; ++col) {
        for (int r = 0, e = 0; r < c.base_rows; ++r)
            for (int cc = 0; cc < c.base_cols; ++cc) {
                const int s = c.base[r * c.base_cols + cc];
                if (s < 0) continue;
                if (cc == col) col_edges.push_back(make_int2(e, s % c.z));
                ++e;
            }
        col_ptr[col + 1] = (int32_t)col_edges.size();
    }
    if (int r = d->d_qc_col_ptr.upload(col_ptr)) return r;
    if (int r = d->d_qc_col_edges.upload(col_edges)) return r;
    return QLDPC_OK;
}

// Tables and shared-memory geometry of the layered int8 kernel; false when the code does not fit
// the kernel's assumptions (the caller then uses the generic kernel).
bool plan_layered_i8(qldpc_decoder *d, LayeredI8Params &p, bool upload)
{
    // QLDPC_FLAG_LI8_RESIDENT / _STREAM override the automatic choice of the message placement (parity tests of both)
    const bool force_stream = (d->cfg.flags & QLDPC_FLAG_LI8_STREAM) != 0;
    const bool force_resident = (d->cfg.flags & QLDPC_FLAG_LI8_RESIDENT) != 0;
    const HostCode &c = d->code;
    if (c.z <= 0 || c.z % 4 != 0 || c.max_chk_degree > 20 || d->cfg.max_iter < 1) return false;
    const int W = c.z / 4, ZW32 = (c.z + 31) / 32;
    const int nnz = c.edges / c.z;
    const int tpg = round_up(W, 32);
    if (tpg > layered_i8_max_threads_stream()) return false;
    const int R = c.base_rows, C = c.base_cols;
    auto B = [&](int r, int col) { return c.base[r * C + col]; };
    const bool fast_bits = (W % 32 == 0) && (c.z % 32 == 0);

    // column weights; extension columns = weight 1, shift 0, and the LAST edge of their row
    std::vector<int> colw(C, 0), deg(R, 0), last_col(R, -1);
    for (int r = 0; r < R; ++r)
        for (int col = 0; col < C; ++col)
            if (B(r, col) >= 0) { colw[col]++; deg[r]++; last_col[r] = col; }
    std::vector<int> has_ext(R, 0);
    std::vector<char> is_ext_col(C, 0);
    if (fast_bits)
        for (int r = 0; r < R; ++r)
            if (deg[r] >= 2 && colw[last_col[r]] == 1 && B(r, last_col[r]) == 0) { has_ext[r] = 1; is_ext_col[last_col[r]] = 1; }

    // register rows: the kRegRows heaviest rows, if they all have 19 or 20 edges and no extension edge
    std::vector<int> reg_idx(R, -1);
    int regdc = 0;
    if (R >= layered_i8_reg_rows()) {
        std::vector<int> order(R);
        for (int r = 0; r < R; ++r) order[r] = r;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return deg[a] > deg[b]; });
        bool ok = true;
        for (int k = 0; k < layered_i8_reg_rows(); ++k) ok = ok && !has_ext[order[k]] && (deg[order[k]] == 19 || deg[order[k]] == 20);
        for (int k = layered_i8_reg_rows(); k < R; ++k) ok = ok && (deg[order[k]] - has_ext[order[k]] <= 10);
        if (ok) {
            regdc = 20;
            for (int k = 0; k < layered_i8_reg_rows(); ++k) reg_idx[order[k]] = k;
        }
    }

    std::vector<Li8Edge> edges;
    std::vector<Li8Layer> layers;
    int n_store = 0;
    for (int r = 0; r < R; ++r) {
        Li8Layer ly{};
        ly.edge_begin = (int32_t)edges.size();
        ly.degree = (int16_t)deg[r];
        ly.has_ext = (int16_t)has_ext[r];
        ly.n_core = (int16_t)(deg[r] - has_ext[r]);
        ly.reg_idx = (int16_t)reg_idx[r];
        ly.r_off = reg_idx[r] >= 0 ? -1 : n_store;
        if (reg_idx[r] < 0) n_store += ly.n_core;
        for (int col = 0; col < C; ++col) {
            const int s = B(r, col);
            if (s < 0) continue;
            const int q = s / W, rr = s % W;
            Li8Edge e{};
            e.off0 = (col * W + rr) * 4;
            e.off1 = (col * W + rr - W) * 4;
            e.thresh = W - rr;
            e.hdw = col * ZW32;
            e.selA0 = unpack_selector(q);
            e.selA1 = unpack_selector(q + 1);
            e.selW0 = pack_selector(q);
            e.selW1 = pack_selector(q + 1);
            edges.push_back(e);
        }
        layers.push_back(ly);
    }
    std::vector<uint16_t> pack_cols;
    for (int col = 0; col < C; ++col)
        if (!is_ext_col[col]) pack_cols.push_back((uint16_t)col);

    p.Z = c.z; p.W = W; p.ZW32 = ZW32;
    p.brows = R; p.bcols = C; p.nnz = nnz; p.N = c.n;
    p.n_pack = (int)pack_cols.size();
    p.tab_bytes = round_up(nnz * (int)(sizeof(Li8Edge) + sizeof(QcEdgeAux)) + R * (int)sizeof(Li8Layer) + round_up(C, 8) * 2 + 16, 16) + 128;
    const int L_bytes = C * W * 4;
    const int hd_bytes = round_up(C * ZW32 * 4, 16);
    const int syn_bytes = round_up(R * ZW32 * 4, 16);
    const int avail = d->max_smem_optin - p.tab_bytes;

    // geometry A: messages resident in shared memory (+ register rows)
    const int R_bytes = n_store * W * 4;
    const int slot_res = round_up(L_bytes, 16) + round_up(R_bytes, 16) + hd_bytes + syn_bytes;
    int slots_res = std::min(std::min(avail / slot_res, layered_i8_max_threads() / tpg), 15);

    // geometry B: messages streamed through an L2-resident scratch, 2-deep cp.async ring per frame
    int g_off = 0, st_max = 4;
    for (auto &ly : layers) {
        ly.st = (int16_t)round_up(std::max<int>(ly.n_core, 1), 4);
        ly.g_off = g_off;
        g_off += W * ly.st;
        st_max = std::max<int>(st_max, ly.st);
    }
    const int stage_words = W * st_max;
    const int slot_str = round_up(L_bytes, 16) + 2 * stage_words * 4 + hd_bytes + syn_bytes;
    int slots_str = std::min(std::min(avail / slot_str, layered_i8_max_threads_stream() / tpg), 15);
    if (tpg > layered_i8_max_threads_stream()) slots_str = 0;

    bool stream = slots_str > slots_res;
    if (force_stream && slots_str >= 1) stream = true;
    if (force_resident && slots_res >= 1) stream = false;
    const int slots = stream ? slots_str : slots_res;
    if (slots < 1) return false;
    if (stream) {   // no register rows in streamed mode: every row's messages go through the ring
        regdc = 0;
        int n = 0;
        for (auto &ly : layers) { ly.reg_idx = -1; ly.r_off = n; n += ly.n_core; }
        n_store = n;
    }
    p.stream = stream ? 1 : 0;
    p.rg_words = g_off;
    p.stage_words = stage_words;
    p.n_store = n_store;
    p.regdc = regdc;
    p.off_R = round_up(L_bytes, 16);
    p.off_hd = p.off_R + (stream ? 2 * stage_words * 4 : round_up(R_bytes, 16));
    p.off_syn = p.off_hd + hd_bytes;
    p.slot_bytes = p.off_syn + syn_bytes;
    p.slots = slots;
    p.tpg = tpg;
    d->li8_slots = slots;
    d->li8_tpg = tpg;
    d->li8_smem = p.tab_bytes + slots * p.slot_bytes;
    d->li8_regdc = regdc;
    d->li8_n_store = n_store;
    d->li8_n_pack = p.n_pack;
    d->li8_stream = stream;
    d->li8_rg_words = g_off;
    if (upload) {
        if (d->d_li8_edges.upload(edges) || d->d_li8_layers.upload(layers) || d->d_li8_pack_cols.upload(pack_cols)) return false;
    }
    return true;
}

// Tables and shared-memory geometry of the streamed layered int8 kernel (layered_i8s.cu); false when the code
// does not fit its assumptions (the caller then tries the older layered_i8 kernel).
bool plan_layered_i8s(qldpc_decoder *d)
{
    if (d->cfg.flags & (QLDPC_FLAG_LI8_RESIDENT | QLDPC_FLAG_LI8_STREAM)) return false;
    const HostCode &c = d->code;
    if (c.z <= 0 || c.z % 128 != 0 || d->cfg.max_iter < 1) return false;
    const int Z = c.z, W = Z / 4, ZW32 = Z / 32;
    const int R = c.base_rows, C = c.base_cols;
    if (R < 3 || W > layered_i8s_max_threads()) return false;
    auto B = [&](int r, int col) { return c.base[r * C + col]; };

    // extension columns = weight 1, shift 0, and the LAST edge of a row with at least one other edge
    std::vector<int> colw(C, 0), deg(R, 0), last_col(R, -1), has_ext(R, 0), lpos(C, -1);
    for (int r = 0; r < R; ++r)
        for (int col = 0; col < C; ++col)
            if (B(r, col) >= 0) { colw[col]++; deg[r]++; last_col[r] = col; }
    std::vector<char> is_ext_col(C, 0);
    for (int r = 0; r < R; ++r) {
        if (deg[r] < 1) return false;
        if (deg[r] >= 2 && colw[last_col[r]] == 1 && B(r, last_col[r]) == 0) { has_ext[r] = 1; is_ext_col[last_col[r]] = 1; }
        if (deg[r] - has_ext[r] > 20) return false;
    }
    std::vector<Li8sCol> pcols;
    for (int col = 0; col < C; ++col)
        if (!is_ext_col[col]) { lpos[col] = (int)pcols.size(); pcols.push_back(Li8sCol{col * Z, col * 2 * ZW32}); }

    Li8sGeo geo;
    geo.n_pack = (int)pcols.size();
    geo.off_rows = 0;
    int off = R * (int)sizeof(Li8sRow);
    std::vector<Li8sRow> rows(R);
    std::vector<uint32_t> etab, thr, synt;
    int g_off = 0, nv_max = 1;
    // pass 1: sizes and offsets
    const int off_etab = off;
    int n_core_total = 0, thr_total = 0;
    for (int r = 0; r < R; ++r) { n_core_total += deg[r] - has_ext[r]; thr_total += round_up(deg[r] - has_ext[r], 4); }
    const int off_thr = off_etab + n_core_total * 32;
    const int off_synt = off_thr + thr_total * 4;
    int n_edges_total = 0;
    for (int r = 0; r < R; ++r) n_edges_total += deg[r];
    geo.off_pcols = round_up(off_synt + n_edges_total * 4, 16);
    geo.tab_bytes = round_up(geo.off_pcols + geo.n_pack * (int)sizeof(Li8sCol), 16);

    for (int r = 0; r < R; ++r) {
        Li8sRow &ly = rows[r];
        const int nc = deg[r] - has_ext[r];
        ly.e_off = off_etab + (int)etab.size() * 4;
        ly.thr_off = off_thr + (int)thr.size() * 4;
        ly.syn_off = off_synt + (int)synt.size() * 4;
        const int nv = (nc + 3) / 4;
        const int variant = nc <= 10 ? (nc - 1) * 2 + has_ext[r] : 20 + ((nc + 1) / 2 * 2 - 12) + has_ext[r];
        ly.pack = nc | (nv << 8) | (variant << 16);
        ly.g_off = g_off * 16;
        g_off += nv * W;
        nv_max = std::max<int>(nv_max, nv);
        ly.deg = deg[r];
        ly.ext_src = has_ext[r] ? last_col[r] * Z : -1;
        ly.ext_hd = has_ext[r] ? last_col[r] * 2 * ZW32 * 4 : 0;
        int n_thr = 0;
        for (int col = 0; col < C; ++col) {
            const int s = B(r, col);
            if (s < 0) continue;
            synt.push_back(((uint32_t)((col * 2 * ZW32 + (s >> 5)) * 4) << 5) | (uint32_t)(s & 31));
            if (has_ext[r] && col == last_col[r]) continue;
            const int q = s / W, rr = s % W;
            const uint32_t off0 = (uint32_t)((lpos[col] * W + rr) * 4);
            const uint32_t a0 = unpack_selector(q), a1 = unpack_selector(q + 1);
            const uint32_t e[8] = {off0, a0, a0 ^ 0x0202u, pack_selector(q),
                                   off0 - 4u * (uint32_t)W, a1, a1 ^ 0x0202u, pack_selector(q + 1)};
            etab.insert(etab.end(), e, e + 8);
            thr.push_back((uint32_t)(W - rr));
            ++n_thr;
        }
        for (; n_thr % 4; ++n_thr) thr.push_back((uint32_t)W);
    }
    geo.rg_u4 = g_off;
    geo.stage_bytes = nv_max * W * 16;
    const int L_bytes = geo.n_pack * W * 4;
    geo.off_ring = L_bytes;
    geo.off_ext = geo.off_ring + 2 * geo.stage_bytes;
    geo.off_hd = geo.off_ext + 2 * Z;
    geo.off_syn = geo.off_hd + C * 2 * ZW32 * 4;
    // three mbarriers (32 bytes) close the slot; without syndrome input the syndrome rows are not allocated
    geo.off_mbar[0] = geo.off_syn;
    geo.off_mbar[1] = geo.off_syn + R * ZW32 * 4;
    geo.slot_bytes[0] = geo.off_mbar[0] + 32;
    geo.slot_bytes[1] = geo.off_mbar[1] + 32;
    const int avail = d->max_smem_optin - geo.tab_bytes - kLi8sSlotBase;
    for (int k = 0; k < 2; ++k) {
        const int slots = std::min(std::min(avail / geo.slot_bytes[k], layered_i8s_max_threads() / W), 15);
        geo.slots[k] = slots;
        // bit-input layout: packed bits staged instead of core LLRs, magnitude table behind the slots
        geo.bits_off_stg[k] = geo.slot_bytes[k];
        geo.bits_slot_bytes[k] = geo.slot_bytes[k] + round_up(c.n / 8, 16);
        geo.bits_off_magtab[k] = geo.tab_bytes + kLi8sSlotBase + slots * geo.bits_slot_bytes[k];
        geo.bits_smem[k] = geo.bits_off_magtab[k] + round_up(c.n, 16);
        geo.bits_ok[k] = slots >= 1 && c.n % 128 == 0 && geo.bits_smem[k] <= d->max_smem_optin;
        // frame-prefetch staging buffer, if it fits without giving up a frame slot
        const int with_stg = geo.slot_bytes[k] + L_bytes;
        if (slots >= 1 && avail / with_stg >= slots) {
            geo.off_stg[k] = geo.slot_bytes[k];
            geo.slot_bytes[k] = with_stg;
        }
    }
    if (geo.slots[1] < 1) return false;

    std::vector<uint8_t> blob(geo.tab_bytes, 0);
    std::memcpy(blob.data() + geo.off_rows, rows.data(), rows.size() * sizeof(Li8sRow));
    std::memcpy(blob.data() + off_etab, etab.data(), etab.size() * 4);
    std::memcpy(blob.data() + off_thr, thr.data(), thr.size() * 4);
    std::memcpy(blob.data() + off_synt, synt.data(), synt.size() * 4);
    std::memcpy(blob.data() + geo.off_pcols, pcols.data(), pcols.size() * sizeof(Li8sCol));
    if (d->d_li8s_tab.upload(blob)) return false;
    d->li8s_geo = geo;
    d->li8s = true;
    d->li8_slots = geo.slots[0];   // chunking of the host-pointer entry points
    return true;
}

struct Lane {   // per-stream staging of the host-pointer entry points
    cudaStream_t st = nullptr;
    DevBuf<uint8_t> in, post;
    DevBuf<uint32_t> syn, out, bits;
    DevBuf<uint8_t> ok;
    DevBuf<uint16_t> iters;
};

}  // namespace
}  // namespace qldpc

using namespace qldpc;

struct qldpc_decoder_lanes {
    Lane lane[2];
};

// ---------------------------------------------------------------------------------------- codes

extern "C" int qldpc_code_from_alist_file(const char *path, qldpc_code **out)
{
    if (!path || !out) return QLDPC_ERR_ARG;
    qldpc_code *c = new (std::nothrow) qldpc_code();
    if (!c) return QLDPC_ERR_NOMEM;
    const int r = parse_alist(path, c->h);
    if (r) { delete c; return r; }
    *out = c;
    return QLDPC_OK;
}

extern "C" int qldpc_code_from_qc_file(const char *path, qldpc_code **out)
{
    if (!path || !out) return QLDPC_ERR_ARG;
    qldpc_code *c = new (std::nothrow) qldpc_code();
    if (!c) return QLDPC_ERR_NOMEM;
    const int r = parse_qc(path, c->h);
    if (r) { delete c; return r; }
    *out = c;
    return QLDPC_OK;
}

extern "C" int qldpc_code_from_qc(const int32_t *base, int32_t rows, int32_t cols, int32_t z, qldpc_code **out)
{
    if (!base || !out) return QLDPC_ERR_ARG;
    qldpc_code *c = new (std::nothrow) qldpc_code();
    if (!c) return QLDPC_ERR_NOMEM;
    const int r = build_qc(base, rows, cols, z, c->h);
    if (r) { delete c; return r; }
    *out = c;
    return QLDPC_OK;
}

extern "C" int qldpc_code_from_csr(int32_t n, int32_t m, const int32_t *row_ptr, const int32_t *col_idx, qldpc_code **out)
{
    if (!row_ptr || !col_idx || !out || n <= 0 || m <= 0) return QLDPC_ERR_ARG;
    if (row_ptr[0] != 0) return QLDPC_ERR_ARG;
    for (int c = 0; c < m; ++c)
        if (row_ptr[c + 1] < row_ptr[c]) return QLDPC_ERR_ARG;
    for (int e = 0; e < row_ptr[m]; ++e)
        if (col_idx[e] < 0 || col_idx[e] >= n) return QLDPC_ERR_ARG;
    qldpc_code *c = new (std::nothrow) qldpc_code();
    if (!c) return QLDPC_ERR_NOMEM;
    c->h.n = n;
    c->h.m = m;
    c->h.row_ptr.assign(row_ptr, row_ptr + m + 1);
    c->h.col_idx.assign(col_idx, col_idx + row_ptr[m]);
    c->h.finalize_from_csr();
    *out = c;
    return QLDPC_OK;
}

extern "C" int qldpc_code_set_info_bits_pos(qldpc_code *code, const int32_t *pos, int32_t k)
{
    if (!code || !pos || k <= 0 || k > code->h.n) return QLDPC_ERR_ARG;
    for (int i = 0; i < k; ++i)
        if (pos[i] < 0 || pos[i] >= code->h.n) return QLDPC_ERR_ARG;
    code->h.info_pos.assign(pos, pos + k);
    code->h.k = k;
    return QLDPC_OK;
}

extern "C" int qldpc_code_get_info(const qldpc_code *code, qldpc_code_info *info)
{
    if (!code || !info) return QLDPC_ERR_ARG;
    const HostCode &h = code->h;
    info->n = h.n; info->m = h.m; info->k = h.k; info->edges = h.edges; info->z = h.z;
    info->base_rows = h.base_rows; info->base_cols = h.base_cols;
    info->max_chk_degree = h.max_chk_degree; info->max_var_degree = h.max_var_degree;
    return QLDPC_OK;
}

extern "C" void qldpc_code_free(qldpc_code *code) { delete code; }

// ------------------------------------------------------------------------------------- decoders

extern "C" void qldpc_decoder_config_default(qldpc_decoder_config *cfg)
{
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->schedule = QLDPC_SCHED_FLOODING;   // the reference's live configuration: flooding SPA,
    cfg->rule = QLDPC_RULE_SPA;             // B=int, Q=float (BOOT/src/main.cpp:113,193)
    cfg->dtype = QLDPC_DTYPE_F32;
    cfg->max_iter = 100;                    // BOOT/src/main.cpp:100
    cfg->early_stop = 1;                    // :101
    cfg->syndrome_depth = 1;                // :102
    cfg->norm_factor = 1.0f;                // :98
    cfg->offset = 0.0f;                     // :99
    cfg->out_mode = QLDPC_OUT_INFO;
    cfg->device = 0;
}

struct qldpc_decoder_full : qldpc_decoder {
    // n_devices > 1: this object only holds one complete single-device decoder per device; the host-pointer entry
    // points shard the frames over them, one host thread each
    std::vector<qldpc_decoder *> children;
    qldpc_decoder_lanes lanes;
    DevBuf<uint8_t> d_llr_tmp;          // LLRs synthesised by qldpc_decode_bits_device
    DevBuf<uint8_t> d_mag;              // bit input of the streamed int8 kernel: 2 x n magnitudes (host pipeline / device call)
    DevBuf<int8_t> d_ext;               // ... and the slots' extension-column LLRs, 2 lanes x SMs x slots x n bytes
    DevBuf<int32_t> d_base;
    DevBuf<uint32_t> d_mask_known, d_mask_punct, d_tmp_bits;
    size_t scratch_msg_bytes = 0, scratch_app_bytes = 0;
    DevBuf<uint8_t> d_scratch2;
    int gen_grid = 0, gen_beliefs_global = 0, gen_compressed = 0;
    int flood_block = 0, flood_smem = 0, flood_use_smem = 0;
    // clustered QC flooding kernel (flooding_qcx.cu): blocks per cluster, co-resident clusters, table bytes
    int qcx_cl = 0, qcx_clusters = 0, qcx_smem = 0, qcx_lanes = 0;
};

static qldpc_decoder_full *full(qldpc_decoder *d) { return static_cast<qldpc_decoder_full *>(d); }

// Frame sharding of a multi-device decoder: child g takes frames [g*F/G, (g+1)*F/G), each call on its own host thread
// (every child owns its device, streams and staging buffers, so the calls share nothing).  `fn(child, first, count)`.
template <typename Fn>
static int shard_frames(qldpc_decoder_full *d, int n_frames, Fn fn)
{
    const int G = (int)d->children.size();
    std::vector<int> rc(G, QLDPC_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        const int f0 = (int)((long long)n_frames * g / G), f1 = (int)((long long)n_frames * (g + 1) / G);
        if (f1 > f0) th.emplace_back([&, g, f0, f1] { rc[g] = fn(d->children[g], f0, f1 - f0); });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (rc[g]) return rc[g];
    return QLDPC_OK;
}

// the kernels of the int8 layered family write the information bits themselves when they are whole leading block columns
static bool li8_direct_out(const qldpc_decoder *d)
{
    return d->cfg.out_mode == QLDPC_OUT_ALL || (d->info_is_prefix && d->code.z > 0 && d->code.k % d->code.z == 0);
}

// integer LLR magnitudes of the bit-input entry points: rounded to the nearest integer, inside the dtype's range
static bool llr_mags_ok(int dtype, float noisy, float known)
{
    if (dtype == QLDPC_DTYPE_F32) return true;
    const float hi = dtype == QLDPC_DTYPE_I8 ? 127.0f : 32767.0f;
    return noisy >= 0.0f && known >= 0.0f && lrintf(noisy) <= hi && lrintf(known) <= hi;
}

extern "C" int qldpc_decoder_create(const qldpc_code *code, const qldpc_decoder_config *cfg, qldpc_decoder **out)
{
    if (!code || !cfg || !out) return QLDPC_ERR_ARG;
    if (cfg->max_iter < 0 || cfg->max_iter > 65535) return QLDPC_ERR_ARG;
    if (cfg->schedule != QLDPC_SCHED_FLOODING && cfg->schedule != QLDPC_SCHED_LAYERED) return QLDPC_ERR_ARG;
    if (cfg->rule < QLDPC_RULE_SPA || cfg->rule > QLDPC_RULE_OMS) return QLDPC_ERR_ARG;
    if (cfg->dtype < QLDPC_DTYPE_F32 || cfg->dtype > QLDPC_DTYPE_I8) return QLDPC_ERR_ARG;
    if (cfg->out_mode != QLDPC_OUT_INFO && cfg->out_mode != QLDPC_OUT_ALL) return QLDPC_ERR_ARG;
    const bool is_int = cfg->dtype != QLDPC_DTYPE_F32;
    if (is_int && cfg->rule == QLDPC_RULE_SPA) return QLDPC_ERR_UNSUPPORTED;   // SPA is float only (as in AFF3CT)
    // layered decoding of a non-quasi-cyclic H: float only (the integer layered arithmetic is defined on circulants,
    // ML/BPSK_nrldpc_sim_FP.m), check degree bounded by the kernel's per-thread store
    if (cfg->schedule == QLDPC_SCHED_LAYERED && code->h.z <= 0 &&
        (is_int || code->h.max_chk_degree > layered_csr_max_degree())) return QLDPC_ERR_UNSUPPORTED;
    if (cfg->n_devices < 0 || cfg->n_devices > QLDPC_MAX_DEVICES) return QLDPC_ERR_ARG;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return QLDPC_ERR_NO_DEVICE; }
    if (cfg->n_devices > 1) {   // one complete decoder per device; this object shards the host-pointer calls over them
        for (int g = 0; g < cfg->n_devices; ++g) {
            if (cfg->devices[g] < 0 || cfg->devices[g] >= ndev) return QLDPC_ERR_ARG;
            for (int h = 0; h < g; ++h)
                if (cfg->devices[h] == cfg->devices[g]) return QLDPC_ERR_ARG;
        }
        qldpc_decoder_full *par = new (std::nothrow) qldpc_decoder_full();
        if (!par) return QLDPC_ERR_NOMEM;
        for (int g = 0; g < cfg->n_devices; ++g) {
            qldpc_decoder_config cc = *cfg;
            cc.n_devices = 0;
            cc.device = cfg->devices[g];
            qldpc_decoder *child = nullptr;
            const int rc = qldpc_decoder_create(code, &cc, &child);
            if (rc) { qldpc_decoder_free(par); return rc; }
            par->children.push_back(child);
        }
        const qldpc_decoder *c0 = par->children[0];
        par->code = c0->code;
        par->cfg = *cfg;
        par->kernel_family = c0->kernel_family;
        par->kernel_name = c0->kernel_name;
        par->out_words = c0->out_words; par->syn_words = c0->syn_words; par->cw_words = c0->cw_words; par->out_bits = c0->out_bits;
        par->info_is_prefix = c0->info_is_prefix;
        *out = par;
        return QLDPC_OK;
    }
    const int device = cfg->n_devices == 1 ? cfg->devices[0] : cfg->device;
    if (device < 0 || device >= ndev) return QLDPC_ERR_ARG;
    QLDPC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    QLDPC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return QLDPC_ERR_NO_DEVICE;   // kernels are built for sm_100a only

    qldpc_decoder_full *d = new (std::nothrow) qldpc_decoder_full();
    if (!d) return QLDPC_ERR_NOMEM;
    d->code = code->h;
    d->cfg = *cfg;
    d->cfg.device = device;
    d->cfg.n_devices = 0;
    if (d->cfg.syndrome_depth < 1) d->cfg.syndrome_depth = 1;
    d->sm_count = prop.multiProcessorCount;
    d->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    const HostCode &c = d->code;

    int rc = QLDPC_OK;
    auto bail = [&](int r) { delete d; return r; };

    if (is_int) {
        if (cfg->rule == QLDPC_RULE_NMS) {
            const float k8 = cfg->norm_factor * 8.0f;
            const int k = (int)lrintf(k8);
            if (k < 1 || k > 8 || std::fabs(k8 - (float)k) > 1e-4f) return bail(QLDPC_ERR_ARG);
            d->norm_eighths = k;
        } else {
            const int o = (int)lrintf(cfg->offset);
            if (o < 0 || std::fabs(cfg->offset - (float)o) > 1e-4f) return bail(QLDPC_ERR_ARG);
            d->offset_int = o;
        }
        if (d->cfg.msg_max <= 0) d->cfg.msg_max = cfg->dtype == QLDPC_DTYPE_I8 ? 31 : 511;
        if (d->cfg.app_max <= 0) d->cfg.app_max = cfg->dtype == QLDPC_DTYPE_I8 ? 127 : 8191;
        if (cfg->dtype == QLDPC_DTYPE_I8 && (d->cfg.msg_max > 126 || d->cfg.app_max > 127)) return bail(QLDPC_ERR_ARG);
        // int16 tier: beliefs are held as int16 (shared memory), messages as int16 / 16-bit check-node constants
        if (d->cfg.msg_max > 32766 || d->cfg.app_max > 32767) return bail(QLDPC_ERR_ARG);
    }

    d->cw_words = (c.n + 31) / 32;
    d->syn_words = (c.m + 31) / 32;
    d->out_bits = cfg->out_mode == QLDPC_OUT_ALL ? c.n : c.k;
    d->out_words = (d->out_bits + 31) / 32;
    d->info_is_prefix = true;
    for (int i = 0; i < c.k; ++i)
        if (c.info_pos[i] != i) { d->info_is_prefix = false; break; }

    if ((rc = d->d_row_ptr.upload(c.row_ptr))) return bail(rc);
    if ((rc = d->d_col_idx.upload(c.col_idx))) return bail(rc);
    if ((rc = d->d_var_ptr.upload(c.var_ptr))) return bail(rc);
    if ((rc = d->d_var_edge.upload(c.var_edge))) return bail(rc);
    if ((rc = d->d_info_pos.upload(c.info_pos))) return bail(rc);
    if (c.z > 0) {
        if ((rc = build_qc_tables(d))) return bail(rc);
        if ((rc = d->d_base.upload(c.base))) return bail(rc);
    }
    if ((rc = d->d_stats.ensure(1))) return bail(rc);
    if (cudaMemset(d->d_stats.p, 0, sizeof(DevStats)) != cudaSuccess) return bail(QLDPC_ERR_CUDA);

    if (cfg->schedule == QLDPC_SCHED_LAYERED && c.z <= 0) {
        // one thread per frame, frame-minor state (layered_csr.cu); the thread count bounds the scratch to 512 MB
        d->kernel_family = KF_LAYERED_CSR;
        d->kernel_name = "layered_csr";
        const size_t per_thread = (size_t)(c.n + c.edges) * 4;
        const size_t cap = std::max<size_t>(128, ((512u << 20) / per_thread) / 128 * 128);
        d->gen_grid = (int)std::min<size_t>(cap, (size_t)d->sm_count * 8 * 128);
        d->scratch_msg_bytes = (size_t)d->gen_grid * c.edges * 4;
        d->scratch_app_bytes = (size_t)d->gen_grid * c.n * 4;
    } else if (cfg->schedule == QLDPC_SCHED_LAYERED) {
        LayeredI8Params p{};
        const bool i8_ok = cfg->dtype == QLDPC_DTYPE_I8 && d->cfg.app_max == 127 && d->cfg.msg_max <= 63;
        const bool fast_s = i8_ok && plan_layered_i8s(d);
        const bool fast = !fast_s && i8_ok && plan_layered_i8(d, p, true);
        if (fast_s) {
            if (cfg->flags & QLDPC_FLAG_L2_PERSIST) {
                // L2 set-aside for the message scratch: the streamed LLRs would otherwise evict it (ncu: DRAM traffic halves).
                // This is a device-wide limit of the CUDA context, hence opt-in.
                const size_t want = (size_t)d->sm_count * d->li8s_geo.slots[0] * d->li8s_geo.rg_u4 * 16;
                const size_t lim = std::min<size_t>(want, (size_t)prop.persistingL2CacheMaxSize);
                if (lim > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim) == cudaSuccess) {
                    d->l2_persist_bytes = lim;
                    d->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
                }
                cudaGetLastError();
            }
            d->kernel_family = KF_LAYERED_I8S;
            d->kernel_name = "layered_i8s_zpack4";
        } else if (fast) {
            d->kernel_family = KF_LAYERED_I8;
            d->kernel_name = "layered_i8_zpack4";
            if (d->li8_stream && (rc = d->d_li8_rg.ensure(2 * (size_t)d->sm_count * d->li8_slots * d->li8_rg_words))) return bail(rc);
        } else {
            d->kernel_family = KF_LAYERED_GENERIC;
            d->kernel_name = "layered_generic";
        }
        // the generic kernel also serves posterior requests of the fast family: beliefs in shared memory (in a global scratch
        // when a frame's beliefs do not fit), messages in an L2-resident scratch, as many CTAs per SM as fit
        {
            const int nnz = c.edges / c.z;
            int smem = layered_generic_smem_bytes(c.base_rows, nnz, c.n, cfg->dtype);
            d->gen_beliefs_global = smem > d->max_smem_optin;
            if (d->gen_beliefs_global) smem = layered_generic_smem_bytes(c.base_rows, nnz, 0, cfg->dtype);
            const int per_sm = layered_generic_blocks_per_sm(cfg->dtype, c.z, smem, d->gen_beliefs_global);
            if (per_sm < 1) return bail(QLDPC_ERR_UNSUPPORTED);
            d->gen_grid = d->sm_count * per_sm;
            int max_deg = 0;
            for (int r = 0; r < c.base_rows; ++r) {
                int deg = 0;
                for (int j = 0; j < c.base_cols; ++j) deg += c.base[(size_t)r * c.base_cols + j] >= 0;
                max_deg = std::max(max_deg, deg);
            }
            // compressed check-node state costs a few selects per edge and pays when the per-edge messages of all frames in
            // flight would not stay in L2 (BG1 Z=384 float: 144 MB): measured +29 % there, -8 % where they fit
            d->gen_compressed = cfg->rule != QLDPC_RULE_SPA && max_deg <= layered_generic_max_compiled_degree() &&
                                (size_t)d->gen_grid * layered_generic_msg_scratch_bytes(cfg->dtype, 0, c.base_rows, nnz, c.z) >
                                    (size_t)prop.l2CacheSize / 10 * 6;
            d->scratch_msg_bytes = (size_t)d->gen_grid * layered_generic_msg_scratch_bytes(cfg->dtype, d->gen_compressed, c.base_rows, nnz, c.z);
            d->scratch_app_bytes = d->gen_beliefs_global ? (size_t)d->gen_grid * c.n * layered_generic_belief_bytes(cfg->dtype) : 0;
        }
    } else {
        d->kernel_family = KF_FLOODING;
        d->kernel_name = "flooding_csr";
        // quasi-cyclic codes, float min-sum: the circulant-aware kernel (lanes on consecutive threads, no index gathers, the
        // early-termination test fused into the check phase): 6.9 vs 6.1 Gbit/s on the N=65536 code.  SPA stays on the CSR
        // kernel: it is bound by the double-precision tanh/atanh, and the fused test costs it one extra check phase.
        d->flood_qc = c.z > 0 && cfg->dtype == QLDPC_DTYPE_F32 && cfg->rule != QLDPC_RULE_SPA;
        if (d->flood_qc) d->kernel_name = "flooding_qc";
        const size_t need = (size_t)(c.edges + c.n) * 4;
        d->flood_use_smem = need + 1024 <= (size_t)d->max_smem_optin;
        d->flood_smem = d->flood_use_smem ? (int)need : 0;
        const int work = std::max(c.n, c.m);
        d->flood_block = std::min(1024, std::max(128, round_up(work / 2, 32)));
        d->gen_grid = d->flood_use_smem ? d->sm_count * std::max(1, std::min(8, d->max_smem_optin / (int)(need + 1024)))
                                        : d->sm_count * 2;
        if (!d->flood_use_smem) {
            d->scratch_msg_bytes = (size_t)d->gen_grid * c.edges * 4;
            d->scratch_app_bytes = (size_t)d->gen_grid * c.n * 4;
        }
        // Large quasi-cyclic codes (messages do not fit in shared memory, whole warps per circulant): one frame per cluster
        // of CL thread blocks, CL chosen so that the state of all frames in flight stays inside L2.
        if (c.z > 0 && c.z % 32 == 0 && !d->flood_use_smem) {
            const int nnz = c.edges / c.z, zp = flooding_qcx_run_lanes(c.z);
            const int smem = flooding_qcx_smem_bytes(c.base_rows, c.base_cols, nnz, cfg->dtype);
            const size_t msg_bytes = (size_t)nnz * zp * flooding_qcx_msg_bytes(cfg->dtype);
            const size_t post_bytes = (size_t)c.base_cols * zp * flooding_qcx_post_bytes(cfg->dtype);
            const size_t l2_budget = (size_t)prop.l2CacheSize / 10 * 7;
            int max_deg = 0;
            for (int r = 0; r < c.m; ++r) max_deg = std::max(max_deg, c.row_ptr[r + 1] - c.row_ptr[r]);
            d->qcx_lanes = flooding_qcx_lanes_per_thread(max_deg);
            if (c.z % (32 * d->qcx_lanes) != 0) d->qcx_lanes = 2;
            for (int cl = 1; cl <= 8 && c.z % (32 * d->qcx_lanes * cl) == 0; cl *= 2) {
                const int n = flooding_qcx_max_clusters(cfg->dtype, d->qcx_lanes, cl, smem);
                if (n < 1) continue;
                d->qcx_cl = cl; d->qcx_clusters = n; d->qcx_smem = smem;
                if ((size_t)n * (msg_bytes + post_bytes) <= l2_budget) break;
            }
            if (d->qcx_cl > 0) {
                d->kernel_name = "flooding_qc_cluster";
                d->flood_qc = false;
                d->scratch_msg_bytes = (size_t)d->qcx_clusters * msg_bytes;
                d->scratch_app_bytes = (size_t)d->qcx_clusters * post_bytes;
            }
        }
    }
    for (auto &ln : d->lanes.lane)
        if (cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking) != cudaSuccess) return bail(QLDPC_ERR_CUDA);
    *out = d;
    return QLDPC_OK;
}

extern "C" void qldpc_decoder_free(qldpc_decoder *dec)
{
    if (!dec) return;
    qldpc_decoder_full *d = full(dec);
    if (!d->children.empty() || d->cfg.n_devices > 1) {
        for (qldpc_decoder *c : d->children) qldpc_decoder_free(c);
        delete d;
        return;
    }
    cudaSetDevice(d->cfg.device);
    cudaDeviceSynchronize();
    for (auto &ln : d->lanes.lane)
        if (ln.st) cudaStreamDestroy(ln.st);
    delete d;
}

extern "C" int32_t qldpc_out_words(const qldpc_decoder *dec) { return dec ? dec->out_words : 0; }
extern "C" int32_t qldpc_syndrome_words(const qldpc_decoder *dec) { return dec ? dec->syn_words : 0; }
extern "C" int32_t qldpc_codeword_words(const qldpc_decoder *dec) { return dec ? dec->cw_words : 0; }
extern "C" const char *qldpc_decoder_kernel_name(const qldpc_decoder *dec) { return dec ? dec->kernel_name : "none"; }

// scratch for kernels whose state does not fit on chip (lazy: the tuned int8 path never needs it)
static int ensure_scratch(qldpc_decoder_full *d)
{
    if (d->scratch_msg_bytes && d->d_scratch.n < d->scratch_msg_bytes)
        if (int r = d->d_scratch.ensure(d->scratch_msg_bytes)) return r;
    if (d->scratch_app_bytes && d->d_scratch2.n < d->scratch_app_bytes)
        if (int r = d->d_scratch2.ensure(d->scratch_app_bytes)) return r;
    return QLDPC_OK;
}

// Bit input of the streamed int8 kernel (LLR synthesis inside the decoder): needs room for its shared-memory layout (packed
// bits staged per slot, one magnitude table per CTA) and 16-byte aligned bits (bulk copies).
static bool fused_bits_ok(const qldpc_decoder_full *d, const uint32_t *d_bits, bool with_syndrome)
{
    return d->kernel_family == KF_LAYERED_I8S && !(d->cfg.flags & QLDPC_FLAG_NO_FUSED_BITS) &&
           d->li8s_geo.bits_ok[with_syndrome ? 1 : 0] && (reinterpret_cast<uintptr_t>(d_bits) & 15) == 0;
}

// Device-side alias of a pinned host buffer (cudaHostAlloc / cudaHostRegister under unified addressing), or null when
// the pointer is pageable host memory / not known to the driver.
static void *pinned_alias(const void *p)
{
    if (!p) return nullptr;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// scratch_lane: which copy of the streamed-message scratch to use. Launches that may overlap in time (the two
// streams of the host-pointer pipelines) MUST use different copies; the public device entry point uses copy 0,
// so a decoder supports one qldpc_decode_device call in flight at a time.
// d_bits / d_mag (both or neither): bit input of the streamed int8 kernel instead of d_llr (see fused_bits_ok).
static int decode_device_impl(qldpc_decoder *dec, const void *d_llr, const uint32_t *d_syndrome, int32_t n_frames,
                              uint32_t *d_out_bits, uint8_t *d_ok, uint16_t *d_iters, void *d_posterior,
                              void *cuda_stream, int scratch_lane, const uint32_t *d_bits = nullptr,
                              const uint8_t *d_mag = nullptr)
{
    if (!dec || (!d_llr && !d_bits) || !d_out_bits || n_frames < 0) return QLDPC_ERR_ARG;
    if (n_frames == 0) return QLDPC_OK;
    qldpc_decoder_full *d = full(dec);
    if (!d->children.empty()) return QLDPC_ERR_UNSUPPORTED;   // device pointers belong to one device
    const HostCode &c = d->code;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    const qldpc_decoder_config &cfg = d->cfg;
    int rc;

    int family = d->kernel_family;
    if ((family == KF_LAYERED_I8 || family == KF_LAYERED_I8S) && d_posterior) family = KF_LAYERED_GENERIC;
    if (family == KF_LAYERED_I8S && (reinterpret_cast<uintptr_t>(d_llr) & 15) != 0) family = KF_LAYERED_GENERIC;   // bulk-copy staging
    if (d_bits && family != KF_LAYERED_I8S) return QLDPC_ERR_ARG;

    if (family == KF_LAYERED_I8S) {
        const Li8sGeo &geo = d->li8s_geo;
        const int k = d_syndrome ? 1 : 0;
        LayeredI8sParams p{};
        const bool direct = li8_direct_out(d);
        if (!direct)
            if ((rc = d->d_allbits.ensure((size_t)n_frames * d->cw_words))) return rc;
        p.llr = (const int8_t *)d_llr;
        if (d_bits) {
            const size_t lane_ext = (size_t)d->sm_count * std::max(geo.slots[0], geo.slots[1]) * c.n;
            if ((rc = d->d_ext.ensure(2 * lane_ext))) return rc;
            p.llr = nullptr;
            p.bits = d_bits;
            p.mag = d_mag;
            p.ext_scratch = d->d_ext.p + (size_t)scratch_lane * lane_ext;
        }
        p.syn = d_syndrome;
        p.out = direct ? d_out_bits : d->d_allbits.p;
        p.ok = d_ok; p.iters = d_iters; p.stats = d->d_stats.p;
        p.tab = d->d_li8s_tab.p;
        p.F = n_frames;
        p.Z = c.z; p.W = c.z / 4; p.ZW32 = c.z / 32;
        p.brows = c.base_rows; p.bcols = c.base_cols; p.N = c.n;
        p.n_pack = geo.n_pack;
        p.out_cols = (direct && cfg.out_mode == QLDPC_OUT_INFO) ? c.k / c.z : c.base_cols;
        p.out_words = p.out_cols * p.ZW32;
        p.syn_words = d->syn_words;
        p.max_iter = cfg.max_iter; p.early_stop = cfg.early_stop;
        p.rule = cfg.rule; p.norm_eighths = d->norm_eighths;
        {
            auto dup = [](uint32_t v) { return (v & 0xffffu) | (v << 16); };
            p.h2_lo = dup(0x8000u | (uint32_t)(cfg.msg_max + 1));
            p.h2_hi = dup((uint32_t)cfg.msg_max);
            p.h2_cap = dup((uint32_t)(cfg.msg_max + 1));
            p.h2_span = dup((uint32_t)(2 * cfg.msg_max + 1));
            p.h2_negoff = dup(0x8000u | (uint32_t)d->offset_int);
        }
        p.slots = geo.slots[k];
        p.tab_bytes = geo.tab_bytes; p.off_rows = geo.off_rows; p.off_pcols = geo.off_pcols;
        p.slot_bytes = geo.slot_bytes[k]; p.off_ring = geo.off_ring; p.stage_bytes = geo.stage_bytes;
        p.off_ext = geo.off_ext; p.off_hd = geo.off_hd; p.off_syn = geo.off_syn; p.off_mbar = geo.off_mbar[k]; p.off_stg = geo.off_stg[k];
        int smem_bytes = geo.tab_bytes + kLi8sSlotBase + p.slots * p.slot_bytes;
        if (d_bits) {
            p.slot_bytes = geo.bits_slot_bytes[k]; p.off_stg = geo.bits_off_stg[k]; p.off_magtab = geo.bits_off_magtab[k];
            smem_bytes = geo.bits_smem[k];
        }
        p.rg_u4 = geo.rg_u4;
        const int grid = std::min(d->sm_count, (n_frames + p.slots - 1) / p.slots);
        const size_t lane_u4 = (size_t)d->sm_count * std::max(geo.slots[0], geo.slots[1]) * geo.rg_u4;
        if ((rc = d->d_li8s_rg.ensure(2 * lane_u4 + 2))) return rc;
        p.rg = d->d_li8s_rg.p + (size_t)scratch_lane * lane_u4;
        // frame queue: with early termination the frames of a slot take 1..max_iter iterations each, and a fixed
        // f, f + grid*slots, ... assignment leaves the slots that drew the easy frames idle at the end of the launch
        p.frame_ctr = nullptr;
        if (n_frames > grid * p.slots) {
            p.frame_ctr = reinterpret_cast<unsigned int *>(d->d_li8s_rg.p + 2 * lane_u4 + scratch_lane);
            QLDPC_CUDA(cudaMemsetAsync(p.frame_ctr, 0, sizeof(unsigned int), st));
        }
        if (d->l2_persist_bytes > 0) {
            // keep (a share of) the message scratch resident in L2 while the LLR stream flows through it
            const size_t win = std::min<size_t>((size_t)grid * p.slots * geo.rg_u4 * 16, (size_t)d->l2_window_max);
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = p.rg;
            av.accessPolicyWindow.num_bytes = win;
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)d->l2_persist_bytes / (double)std::max<size_t>(win, 1));
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
            cudaGetLastError();
        }
        p.discard_scratch = (d->cfg.flags & QLDPC_FLAG_DISCARD_SCRATCH) ? 1 : 0;
        if ((rc = launch_layered_i8s(p, grid, smem_bytes, st))) return rc;
        d->kernel_launches++;
        if (!direct) {
            if ((rc = launch_gather_bits(d->d_allbits.p, d_out_bits, n_frames, d->cw_words, d->out_words, c.k,
                                         d->d_info_pos.p, st))) return rc;
            d->kernel_launches++;
        }
        return QLDPC_OK;
    }

    if (family == KF_LAYERED_I8) {
        LayeredI8Params p{};
        if (!plan_layered_i8(d, p, false)) return QLDPC_ERR_UNSUPPORTED;
        const bool direct = li8_direct_out(d);
        if (!direct)
            if ((rc = d->d_allbits.ensure((size_t)n_frames * d->cw_words))) return rc;
        p.llr = (const int8_t *)d_llr;
        p.syn = d_syndrome;
        p.out = direct ? d_out_bits : d->d_allbits.p;
        p.ok = d_ok; p.iters = d_iters; p.stats = d->d_stats.p;
        p.edges = d->d_li8_edges.p; p.layers = d->d_li8_layers.p; p.aux = d->d_qc_aux.p; p.pack_cols = d->d_li8_pack_cols.p;
        p.F = n_frames;
        p.out_cols = (direct && cfg.out_mode == QLDPC_OUT_INFO) ? c.k / c.z : c.base_cols;
        p.out_words = (p.out_cols * c.z + 31) / 32;
        p.syn_words = d->syn_words;
        p.max_iter = cfg.max_iter; p.early_stop = cfg.early_stop;
        p.rule = cfg.rule; p.offset = d->offset_int; p.norm_eighths = d->norm_eighths; p.msg_max = cfg.msg_max;
        {
            auto dup = [](uint32_t v) { return (v & 0xffffu) | (v << 16); };
            p.h2_lo = dup(0x8000u | (uint32_t)(cfg.msg_max + 1));
            p.h2_hi = dup((uint32_t)cfg.msg_max);
            p.h2_cap = dup((uint32_t)(cfg.msg_max + 1));
            p.h2_negoff = dup(0x8000u | (uint32_t)d->offset_int);
        }
        const int grid = std::min(d->sm_count, (n_frames + p.slots - 1) / p.slots);
        if (p.stream) {
            const size_t lane_words = (size_t)d->sm_count * p.slots * p.rg_words;
            if ((rc = d->d_li8_rg.ensure(2 * lane_words))) return rc;
            p.rg = d->d_li8_rg.p + (size_t)scratch_lane * lane_words;
        }
        if ((rc = launch_layered_i8(p, grid, d->li8_smem, st))) return rc;
        d->kernel_launches++;
        if (!direct) {
            if ((rc = launch_gather_bits(d->d_allbits.p, d_out_bits, n_frames, d->cw_words, d->out_words, c.k,
                                         d->d_info_pos.p, st))) return rc;
            d->kernel_launches++;
        }
        return QLDPC_OK;
    }

    // families that write all n hard decisions and may need an info-bit gather
    if ((rc = ensure_scratch(d))) return rc;
    const bool direct = cfg.out_mode == QLDPC_OUT_ALL;
    uint32_t *allbits = d_out_bits;
    if (!direct) {
        if ((rc = d->d_allbits.ensure((size_t)n_frames * d->cw_words))) return rc;
        allbits = d->d_allbits.p;
    }
    if (family == KF_LAYERED_CSR) {
        LayeredCsrParams p{};
        p.llr = (const float *)d_llr; p.syn = d_syndrome; p.allbits = allbits; p.ok = d_ok; p.iters = d_iters;
        p.posterior = (float *)d_posterior; p.stats = d->d_stats.p;
        p.row_ptr = d->d_row_ptr.p; p.col_idx = d->d_col_idx.p;
        p.branch = (float *)d->d_scratch.p; p.var = (float *)d->d_scratch2.p;
        p.F = n_frames; p.N = c.n; p.M = c.m; p.E = c.edges; p.cw_words = d->cw_words; p.syn_words = d->syn_words;
        p.max_iter = cfg.max_iter; p.early_stop = cfg.early_stop; p.syndrome_depth = cfg.syndrome_depth;
        p.rule = cfg.rule; p.norm = cfg.norm_factor; p.offset = cfg.offset;
        p.fast_spa = (cfg.flags & QLDPC_FLAG_FAST_SPA) ? 1 : 0;
        const int threads = std::min(d->gen_grid, (n_frames + 127) / 128 * 128);
        // the state is indexed with the launch's thread count as stride: every launch lays it out afresh
        if ((rc = launch_layered_csr(p, threads, st))) return rc;
    } else if (family == KF_LAYERED_GENERIC) {
        LayeredGenParams p{};
        p.llr = d_llr; p.syn = d_syndrome; p.allbits = allbits; p.ok = d_ok; p.iters = d_iters;
        p.posterior = d_posterior; p.stats = d->d_stats.p;
        p.aux = d->d_qc_aux.p; p.layers = d->d_qc_layers.p;
        p.msg = d->d_scratch.p; p.app = d->gen_beliefs_global ? d->d_scratch2.p : nullptr;
        p.compressed = d->gen_compressed;
        p.fast_spa = (cfg.flags & QLDPC_FLAG_FAST_SPA) ? 1 : 0;
        p.F = n_frames; p.Z = c.z; p.brows = c.base_rows; p.bcols = c.base_cols; p.nnz = c.edges / c.z;
        p.N = c.n; p.M = c.m; p.cw_words = d->cw_words; p.syn_words = d->syn_words;
        p.max_iter = cfg.max_iter; p.early_stop = cfg.early_stop; p.syndrome_depth = cfg.syndrome_depth;
        p.rule = cfg.rule; p.dtype = cfg.dtype; p.norm = cfg.norm_factor; p.offset = cfg.offset;
        p.offset_int = d->offset_int; p.norm_eighths = d->norm_eighths; p.msg_max = cfg.msg_max; p.app_max = cfg.app_max;
        if ((rc = launch_layered_generic(p, std::min(d->gen_grid, n_frames), st))) return rc;
    } else {
        FloodParams p{};
        p.llr = d_llr; p.syn = d_syndrome; p.allbits = allbits; p.ok = d_ok; p.iters = d_iters;
        p.posterior = d_posterior; p.stats = d->d_stats.p;
        p.row_ptr = d->d_row_ptr.p; p.col_idx = d->d_col_idx.p; p.var_ptr = d->d_var_ptr.p; p.var_edge = d->d_var_edge.p;
        p.c2v = d->d_scratch.p; p.post = d->d_scratch2.p;
        p.F = n_frames; p.N = c.n; p.M = c.m; p.E = c.edges; p.cw_words = d->cw_words; p.syn_words = d->syn_words;
        p.max_iter = cfg.max_iter; p.early_stop = cfg.early_stop; p.syndrome_depth = cfg.syndrome_depth;
        p.rule = cfg.rule; p.dtype = cfg.dtype; p.norm = cfg.norm_factor; p.offset = cfg.offset;
        p.offset_int = d->offset_int; p.norm_eighths = d->norm_eighths;
        p.vmax = cfg.dtype == QLDPC_DTYPE_I16 ? 32767 : 127;
        p.use_smem = d->flood_use_smem;
        p.fast_spa = (cfg.flags & QLDPC_FLAG_FAST_SPA) ? 1 : 0;
        if (d->qcx_cl > 0) {
            FloodQcxParams xp{};
            xp.llr = d_llr; xp.syn = d_syndrome; xp.allbits = allbits; xp.ok = d_ok; xp.iters = d_iters;
            xp.posterior = d_posterior; xp.stats = d->d_stats.p;
            xp.aux = d->d_qc_aux.p; xp.layers = d->d_qc_layers.p; xp.col_ptr = d->d_qc_col_ptr.p; xp.col_edges = d->d_qc_col_edges.p;
            xp.c2v = d->d_scratch.p; xp.post = d->d_scratch2.p;
            xp.F = n_frames; xp.Z = c.z; xp.nnz = c.edges / c.z; xp.N = c.n; xp.M = c.m; xp.brows = c.base_rows; xp.bcols = c.base_cols;
            xp.cw_words = d->cw_words; xp.syn_words = d->syn_words;
            xp.max_iter = cfg.max_iter; xp.early_stop = cfg.early_stop; xp.syndrome_depth = cfg.syndrome_depth;
            xp.rule = cfg.rule; xp.dtype = cfg.dtype; xp.norm = cfg.norm_factor; xp.offset = cfg.offset;
            xp.offset_int = d->offset_int; xp.norm_eighths = d->norm_eighths; xp.vmax = p.vmax;
            xp.lanes = d->qcx_lanes;
            xp.fast_spa = (cfg.flags & QLDPC_FLAG_FAST_SPA) ? 1 : 0;
            if ((rc = launch_flooding_qcx(xp, std::min(d->qcx_clusters, n_frames), d->qcx_cl, d->qcx_smem, st))) return rc;
        } else if (d->flood_qc) {
            FloodQcParams qp{};
            qp.llr = (const float *)d_llr; qp.syn = d_syndrome; qp.allbits = allbits; qp.ok = d_ok; qp.iters = d_iters;
            qp.posterior = (float *)d_posterior; qp.stats = d->d_stats.p;
            qp.aux = d->d_qc_aux.p; qp.layers = d->d_qc_layers.p; qp.col_ptr = d->d_qc_col_ptr.p; qp.col_edges = d->d_qc_col_edges.p;
            qp.c2v = (float *)d->d_scratch.p; qp.post = (float *)d->d_scratch2.p;
            qp.F = n_frames; qp.Z = c.z; qp.nnz = c.edges / c.z; qp.N = c.n; qp.M = c.m;
            qp.cw_words = d->cw_words; qp.syn_words = d->syn_words;
            qp.max_iter = cfg.max_iter; qp.early_stop = cfg.early_stop; qp.syndrome_depth = cfg.syndrome_depth;
            qp.rule = cfg.rule; qp.norm = cfg.norm_factor; qp.offset = cfg.offset; qp.use_smem = d->flood_use_smem;
            if ((rc = launch_flooding_qc(qp, std::min(d->gen_grid, n_frames), std::min(d->flood_block, layered_flood_qc_threads()), d->flood_smem, st))) return rc;
        } else
        {
            int smem_bytes = d->flood_smem;
            if (!d->flood_use_smem && cfg.dtype == QLDPC_DTYPE_F32 && cfg.rule == QLDPC_RULE_SPA) {
                p.tanh_cache = 1;
                smem_bytes = 8 * d->flood_block * 4;
            }
            if ((rc = launch_flooding(p, std::min(d->gen_grid, n_frames), d->flood_block, smem_bytes, st))) return rc;
        }
    }
    d->kernel_launches++;
    if (!direct) {
        if ((rc = launch_gather_bits(allbits, d_out_bits, n_frames, d->cw_words, d->out_words, c.k, d->d_info_pos.p, st)))
            return rc;
        d->kernel_launches++;
    }
    return QLDPC_OK;
}

extern "C" int qldpc_decode_device(qldpc_decoder *dec, const void *d_llr, const uint32_t *d_syndrome, int32_t n_frames,
                                   uint32_t *d_out_bits, uint8_t *d_ok, uint16_t *d_iters, void *d_posterior,
                                   void *cuda_stream)
{
    return decode_device_impl(dec, d_llr, d_syndrome, n_frames, d_out_bits, d_ok, d_iters, d_posterior, cuda_stream, 0);
}

// Chunk schedule of the host-pointer entry points.  Chunks are whole waves of the persistent grid; they start small
// (two waves: the first kernel starts after a 5 MB copy) and double up to about `target_bytes` of device input, so that
// a large batch needs few launches (every launch ends with a partly idle wave) without paying for it in pipeline fill.
struct ChunkPlan {
    int wave = 1, first = 1, max = 1;
    int at(int idx) const
    {
        long long c = first;
        for (int k = 0; k < idx && c < max; ++k) c *= 2;
        return (int)std::min<long long>(c, max);
    }
};
static ChunkPlan pick_chunk(const qldpc_decoder_full *d, size_t frame_bytes, size_t target_bytes, int n_frames)
{
    ChunkPlan cp;
    cp.wave = d->sm_count * std::max(1, d->li8_slots);
    long long chunk = (long long)std::max<size_t>(1, target_bytes / frame_bytes);
    chunk = std::max<long long>(cp.wave, (chunk + cp.wave / 2) / cp.wave * cp.wave);
    cp.max = (int)std::min<long long>(chunk, n_frames);
    cp.first = std::min(cp.max, 2 * cp.wave);
    return cp;
}

// Host-pointer entry point: frames are cut into chunks that ping-pong over two streams so the
// H2D copy of chunk i+1 overlaps the decode of chunk i and the D2H of chunk i-1.
extern "C" int qldpc_decode(qldpc_decoder *dec, const void *llr, const uint32_t *syndrome, int32_t n_frames,
                            uint32_t *out_bits, uint8_t *ok, uint16_t *iters, void *posterior)
{
    if (!dec || !llr || !out_bits || n_frames < 0) return QLDPC_ERR_ARG;
    if (n_frames == 0) return QLDPC_OK;
    qldpc_decoder_full *d = full(dec);
    const HostCode &c = d->code;
    const size_t esz = dtype_size(d->cfg.dtype);
    const size_t frame_in = (size_t)c.n * esz;
    if (!d->children.empty())
        return shard_frames(d, n_frames, [&](qldpc_decoder *ch, int f0, int nf) {
            return qldpc_decode(ch, (const char *)llr + (size_t)f0 * frame_in, syndrome ? syndrome + (size_t)f0 * d->syn_words : nullptr,
                                nf, out_bits + (size_t)f0 * d->out_words, ok ? ok + f0 : nullptr, iters ? iters + f0 : nullptr,
                                posterior ? (char *)posterior + (size_t)f0 * c.n * 4 : nullptr);
        });
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    // chunk: ~320 MiB of LLRs, at least one wave of the persistent grid
    const ChunkPlan plan = pick_chunk(d, frame_in, 320u << 20, n_frames);
    const int chunk = plan.max;
    // the scratch of the gather path is shared: those configurations run on one lane only
    const bool shared_scratch = !((d->kernel_family == KF_LAYERED_I8 || d->kernel_family == KF_LAYERED_I8S) && !posterior &&
                                  li8_direct_out(d));
    int rc = QLDPC_OK;
    for (auto &ln : d->lanes.lane) {
        if ((rc = ln.in.ensure((size_t)chunk * frame_in))) return rc;
        if ((rc = ln.out.ensure((size_t)chunk * d->out_words))) return rc;
        if ((rc = ln.ok.ensure(chunk))) return rc;
        if ((rc = ln.iters.ensure(chunk))) return rc;
        if (syndrome && (rc = ln.syn.ensure((size_t)chunk * d->syn_words))) return rc;
        if (posterior && (rc = ln.post.ensure((size_t)chunk * c.n * 4))) return rc;
    }
    int idx = 0;
    for (int f0 = 0, nf = 0; f0 < n_frames; f0 += nf, ++idx) {
        Lane &ln = d->lanes.lane[shared_scratch ? 0 : (idx & 1)];
        nf = std::min(plan.at(idx), n_frames - f0);
        QLDPC_CUDA(cudaMemcpyAsync(ln.in.p, (const char *)llr + (size_t)f0 * frame_in, (size_t)nf * frame_in,
                                   cudaMemcpyHostToDevice, ln.st));
        if (syndrome)
            QLDPC_CUDA(cudaMemcpyAsync(ln.syn.p, syndrome + (size_t)f0 * d->syn_words, (size_t)nf * d->syn_words * 4,
                                       cudaMemcpyHostToDevice, ln.st));
        rc = decode_device_impl(dec, ln.in.p, syndrome ? ln.syn.p : nullptr, nf, ln.out.p, ln.ok.p, ln.iters.p,
                                posterior ? ln.post.p : nullptr, ln.st, shared_scratch ? 0 : (idx & 1));
        if (rc) break;
        QLDPC_CUDA(cudaMemcpyAsync(out_bits + (size_t)f0 * d->out_words, ln.out.p, (size_t)nf * d->out_words * 4,
                                   cudaMemcpyDeviceToHost, ln.st));
        if (ok) QLDPC_CUDA(cudaMemcpyAsync(ok + f0, ln.ok.p, nf, cudaMemcpyDeviceToHost, ln.st));
        if (iters) QLDPC_CUDA(cudaMemcpyAsync(iters + f0, ln.iters.p, (size_t)nf * 2, cudaMemcpyDeviceToHost, ln.st));
        if (posterior)
            QLDPC_CUDA(cudaMemcpyAsync((char *)posterior + (size_t)f0 * c.n * 4, ln.post.p, (size_t)nf * c.n * 4,
                                       cudaMemcpyDeviceToHost, ln.st));
    }
    for (auto &ln : d->lanes.lane) {
        cudaError_t e = cudaStreamSynchronize(ln.st);
        if (e != cudaSuccess && rc == QLDPC_OK) rc = CudaCheck::fail(e, "cudaStreamSynchronize");
    }
    return rc;
}

// ---------------------------------------------------------------------------- bit-level helpers

extern "C" int qldpc_syndrome_device(qldpc_decoder *dec, const uint32_t *d_bits, int32_t n_frames, uint32_t *d_syndrome,
                                     void *cuda_stream)
{
    if (!dec || !d_bits || !d_syndrome || n_frames < 0) return QLDPC_ERR_ARG;
    qldpc_decoder_full *d = full(dec);
    if (!d->children.empty()) return QLDPC_ERR_UNSUPPORTED;
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    const HostCode &c = d->code;
    // quasi-cyclic codes with whole words per circulant: rotate + XOR of packed words; everything else: CSR bit gather
    const int rc = (c.z > 0 && c.z % 32 == 0)
                       ? launch_syndrome_qc(d_bits, d_syndrome, n_frames, c.base_rows, c.z, d->cw_words, d->syn_words,
                                            d->d_qc_layers.p, d->d_qc_aux.p, (cudaStream_t)cuda_stream)
                       : launch_syndrome_csr(d_bits, d_syndrome, n_frames, c.n, c.m, d->cw_words, d->syn_words, d->d_row_ptr.p,
                                             d->d_col_idx.p, (cudaStream_t)cuda_stream);
    if (!rc && n_frames) d->kernel_launches++;
    return rc;
}

extern "C" int qldpc_syndrome(qldpc_decoder *dec, const uint32_t *bits, int32_t n_frames, uint32_t *syndrome)
{
    if (!dec || !bits || !syndrome || n_frames < 0) return QLDPC_ERR_ARG;
    if (n_frames == 0) return QLDPC_OK;
    qldpc_decoder_full *d = full(dec);
    if (!d->children.empty())
        return shard_frames(d, n_frames, [&](qldpc_decoder *ch, int f0, int nf) {
            return qldpc_syndrome(ch, bits + (size_t)f0 * d->cw_words, nf, syndrome + (size_t)f0 * d->syn_words);
        });
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    Lane &ln = d->lanes.lane[0];
    int rc;
    if ((rc = d->d_tmp_bits.ensure((size_t)n_frames * d->cw_words))) return rc;
    if ((rc = ln.syn.ensure((size_t)n_frames * d->syn_words))) return rc;
    QLDPC_CUDA(cudaMemcpyAsync(d->d_tmp_bits.p, bits, (size_t)n_frames * d->cw_words * 4, cudaMemcpyHostToDevice, ln.st));
    if ((rc = qldpc_syndrome_device(dec, d->d_tmp_bits.p, n_frames, ln.syn.p, ln.st))) return rc;
    QLDPC_CUDA(cudaMemcpyAsync(syndrome, ln.syn.p, (size_t)n_frames * d->syn_words * 4, cudaMemcpyDeviceToHost, ln.st));
    QLDPC_CUDA(cudaStreamSynchronize(ln.st));
    return QLDPC_OK;
}

extern "C" int qldpc_make_llr_device(qldpc_decoder *dec, const uint32_t *d_bits, const uint32_t *d_known_mask,
                                     const uint32_t *d_punct_mask, float llr_noisy, float llr_known, int32_t n_frames,
                                     void *d_llr_out, void *cuda_stream)
{
    if (!dec || !d_bits || !d_llr_out || n_frames < 0) return QLDPC_ERR_ARG;
    qldpc_decoder_full *d = full(dec);
    if (!d->children.empty()) return QLDPC_ERR_UNSUPPORTED;
    if (!llr_mags_ok(d->cfg.dtype, llr_noisy, llr_known)) return QLDPC_ERR_ARG;
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    const int rc = launch_make_llr(d_bits, d_known_mask, d_punct_mask, llr_noisy, llr_known, n_frames, d->code.n,
                                   d->cw_words, d->cfg.dtype, d_llr_out, (cudaStream_t)cuda_stream);
    if (!rc && n_frames) d->kernel_launches++;
    return rc;
}

extern "C" int qldpc_make_llr(qldpc_decoder *dec, const uint32_t *bits, const uint32_t *known_mask,
                              const uint32_t *punct_mask, float llr_noisy, float llr_known, int32_t n_frames, void *llr_out)
{
    if (!dec || !bits || !llr_out || n_frames < 0) return QLDPC_ERR_ARG;
    if (n_frames == 0) return QLDPC_OK;
    qldpc_decoder_full *d = full(dec);
    const size_t esz = dtype_size(d->cfg.dtype);
    if (!d->children.empty())
        return shard_frames(d, n_frames, [&](qldpc_decoder *ch, int f0, int nf) {
            return qldpc_make_llr(ch, bits + (size_t)f0 * d->cw_words, known_mask, punct_mask, llr_noisy, llr_known, nf,
                                  (char *)llr_out + (size_t)f0 * d->code.n * esz);
        });
    if (!llr_mags_ok(d->cfg.dtype, llr_noisy, llr_known)) return QLDPC_ERR_ARG;
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    Lane &ln = d->lanes.lane[0];
    int rc;
    if ((rc = d->d_tmp_bits.ensure((size_t)n_frames * d->cw_words))) return rc;
    if ((rc = ln.in.ensure((size_t)n_frames * d->code.n * esz))) return rc;
    QLDPC_CUDA(cudaMemcpyAsync(d->d_tmp_bits.p, bits, (size_t)n_frames * d->cw_words * 4, cudaMemcpyHostToDevice, ln.st));
    const uint32_t *dk = nullptr, *dp = nullptr;
    if (known_mask) {
        if ((rc = d->d_mask_known.ensure(d->cw_words))) return rc;
        QLDPC_CUDA(cudaMemcpyAsync(d->d_mask_known.p, known_mask, (size_t)d->cw_words * 4, cudaMemcpyHostToDevice, ln.st));
        dk = d->d_mask_known.p;
    }
    if (punct_mask) {
        if ((rc = d->d_mask_punct.ensure(d->cw_words))) return rc;
        QLDPC_CUDA(cudaMemcpyAsync(d->d_mask_punct.p, punct_mask, (size_t)d->cw_words * 4, cudaMemcpyHostToDevice, ln.st));
        dp = d->d_mask_punct.p;
    }
    if ((rc = qldpc_make_llr_device(dec, d->d_tmp_bits.p, dk, dp, llr_noisy, llr_known, n_frames, ln.in.p, ln.st))) return rc;
    QLDPC_CUDA(cudaMemcpyAsync(llr_out, ln.in.p, (size_t)n_frames * d->code.n * esz, cudaMemcpyDeviceToHost, ln.st));
    QLDPC_CUDA(cudaStreamSynchronize(ln.st));
    return QLDPC_OK;
}

extern "C" int qldpc_decode_bits_device(qldpc_decoder *dec, const uint32_t *d_bits, const uint32_t *d_known_mask,
                                        const uint32_t *d_punct_mask, float llr_noisy, float llr_known,
                                        const uint32_t *d_syndrome, int32_t n_frames, uint32_t *d_out_bits, uint8_t *d_ok,
                                        uint16_t *d_iters, void *cuda_stream)
{
    if (!dec || !d_bits || !d_out_bits || n_frames < 0) return QLDPC_ERR_ARG;
    if (n_frames == 0) return QLDPC_OK;
    qldpc_decoder_full *d = full(dec);
    if (!d->children.empty()) return QLDPC_ERR_UNSUPPORTED;
    if (!llr_mags_ok(d->cfg.dtype, llr_noisy, llr_known)) return QLDPC_ERR_ARG;
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    int rc;
    if (fused_bits_ok(d, d_bits, d_syndrome != nullptr)) {   // LLR synthesis inside the decoder kernel: no LLR array at all
        const int n = d->code.n;
        if ((rc = d->d_mag.ensure(2 * (size_t)n))) return rc;
        uint8_t *mag = d->d_mag.p + n;                       // second copy: the first belongs to the host-pointer pipeline
        if ((rc = launch_make_mag_i8(d_known_mask, d_punct_mask, (int)lrintf(llr_noisy), (int)lrintf(llr_known), n, mag,
                                     (cudaStream_t)cuda_stream))) return rc;
        d->kernel_launches++;
        return decode_device_impl(dec, nullptr, d_syndrome, n_frames, d_out_bits, d_ok, d_iters, nullptr, cuda_stream, 0, d_bits, mag);
    }
    if ((rc = d->d_llr_tmp.ensure((size_t)n_frames * d->code.n * dtype_size(d->cfg.dtype)))) return rc;
    if ((rc = qldpc_make_llr_device(dec, d_bits, d_known_mask, d_punct_mask, llr_noisy, llr_known, n_frames, d->d_llr_tmp.p,
                                    cuda_stream))) return rc;
    return qldpc_decode_device(dec, d->d_llr_tmp.p, d_syndrome, n_frames, d_out_bits, d_ok, d_iters, nullptr, cuda_stream);
}

// Host-pointer version: packed bits in, packed bits out, chunks ping-pong over two streams.  With the streamed int8
// kernel the chunks carry bits only (n/8 bytes per frame) and the kernel synthesises its LLRs itself.
extern "C" int qldpc_decode_bits(qldpc_decoder *dec, const uint32_t *bits, const uint32_t *known_mask,
                                 const uint32_t *punct_mask, float llr_noisy, float llr_known, const uint32_t *syndrome,
                                 int32_t n_frames, uint32_t *out_bits, uint8_t *ok, uint16_t *iters)
{
    if (!dec || !bits || !out_bits || n_frames < 0) return QLDPC_ERR_ARG;
    if (n_frames == 0) return QLDPC_OK;
    qldpc_decoder_full *d = full(dec);
    const HostCode &c = d->code;
    if (!d->children.empty())
        return shard_frames(d, n_frames, [&](qldpc_decoder *ch, int f0, int nf) {
            return qldpc_decode_bits(ch, bits + (size_t)f0 * d->cw_words, known_mask, punct_mask, llr_noisy, llr_known,
                                     syndrome ? syndrome + (size_t)f0 * d->syn_words : nullptr, nf,
                                     out_bits + (size_t)f0 * d->out_words, ok ? ok + f0 : nullptr, iters ? iters + f0 : nullptr);
        });
    if (!llr_mags_ok(d->cfg.dtype, llr_noisy, llr_known)) return QLDPC_ERR_ARG;
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    const size_t esz = dtype_size(d->cfg.dtype);
    const size_t frame_llr = (size_t)c.n * esz;
    const ChunkPlan plan = pick_chunk(d, frame_llr, 320u << 20, n_frames);
    const int chunk = plan.max;
    const bool shared_scratch = !((d->kernel_family == KF_LAYERED_I8 || d->kernel_family == KF_LAYERED_I8S) && li8_direct_out(d));
    int rc = QLDPC_OK;
    Lane &l0 = d->lanes.lane[0];
    const uint32_t *dk = nullptr, *dp = nullptr;
    if (known_mask) {
        if ((rc = d->d_mask_known.ensure(d->cw_words))) return rc;
        QLDPC_CUDA(cudaMemcpyAsync(d->d_mask_known.p, known_mask, (size_t)d->cw_words * 4, cudaMemcpyHostToDevice, l0.st));
        dk = d->d_mask_known.p;
    }
    if (punct_mask) {
        if ((rc = d->d_mask_punct.ensure(d->cw_words))) return rc;
        QLDPC_CUDA(cudaMemcpyAsync(d->d_mask_punct.p, punct_mask, (size_t)d->cw_words * 4, cudaMemcpyHostToDevice, l0.st));
        dp = d->d_mask_punct.p;
    }
    // lane buffers are 16-byte aligned (cudaMalloc), chunks are whole frames of n/8 bytes with n % 128 == 0
    const bool fused = fused_bits_ok(d, nullptr, syndrome != nullptr);
    uint8_t *mag = nullptr;
    if (fused) {
        if ((rc = d->d_mag.ensure(2 * (size_t)c.n))) return rc;
        mag = d->d_mag.p;
        if ((rc = launch_make_mag_i8(dk, dp, (int)lrintf(llr_noisy), (int)lrintf(llr_known), c.n, mag, l0.st))) return rc;
        d->kernel_launches++;
    }
    QLDPC_CUDA(cudaStreamSynchronize(l0.st));   // masks / magnitudes are read by both lanes
    // Zero-copy: with bit input a frame costs n/8 bytes in and k/8 + 3 bytes out.  When the caller's frame buffers are
    // pinned host memory the decoder kernel fetches the bits (its per-slot bulk copy of the NEXT frame, one frame ahead of
    // the decode) and stores its results through PCIe itself: ONE launch for the whole batch, no staging copies, no chunk
    // pipeline whose every launch ends with a partly idle wave.  Pageable buffers take the chunked copy pipeline below.
    if (fused && !(d->cfg.flags & QLDPC_FLAG_NO_ZERO_COPY)) {
        const uint32_t *zb = (const uint32_t *)pinned_alias(bits), *zs = (const uint32_t *)pinned_alias(syndrome);
        uint32_t *zo = (uint32_t *)pinned_alias(out_bits);
        uint8_t *zk = (uint8_t *)pinned_alias(ok);
        uint16_t *zi = (uint16_t *)pinned_alias(iters);
        if (zb && zo && (!syndrome || zs) && (!ok || zk) && (!iters || zi) && (reinterpret_cast<uintptr_t>(zb) & 15) == 0) {
            if ((rc = decode_device_impl(dec, nullptr, zs, n_frames, zo, zk, zi, nullptr, l0.st, 0, zb, mag))) return rc;
            QLDPC_CUDA(cudaStreamSynchronize(l0.st));
            return QLDPC_OK;
        }
    }
    for (auto &ln : d->lanes.lane) {
        if (!fused && (rc = ln.in.ensure((size_t)chunk * frame_llr))) return rc;
        if ((rc = ln.bits.ensure((size_t)chunk * d->cw_words))) return rc;
        if ((rc = ln.out.ensure((size_t)chunk * d->out_words))) return rc;
        if ((rc = ln.ok.ensure(chunk))) return rc;
        if ((rc = ln.iters.ensure(chunk))) return rc;
        if (syndrome && (rc = ln.syn.ensure((size_t)chunk * d->syn_words))) return rc;
    }
    // QLDPC_TIMELINE=1 (diagnostics): per chunk, when its H2D copy, LLR synthesis, decode and D2H copy ended, on stderr
    static const bool timeline = std::getenv("QLDPC_TIMELINE") != nullptr;
    std::vector<cudaEvent_t> tl;
    auto mark = [&](cudaStream_t st) {
        if (!timeline) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tl.push_back(e);
    };
    const auto host_t0 = std::chrono::steady_clock::now();
    mark(l0.st);
    int idx = 0;
    for (int f0 = 0, nf = 0; f0 < n_frames; f0 += nf, ++idx) {
        Lane &ln = d->lanes.lane[shared_scratch ? 0 : (idx & 1)];
        nf = std::min(plan.at(idx), n_frames - f0);
        mark(ln.st);
        QLDPC_CUDA(cudaMemcpyAsync(ln.bits.p, bits + (size_t)f0 * d->cw_words, (size_t)nf * d->cw_words * 4,
                                   cudaMemcpyHostToDevice, ln.st));
        mark(ln.st);
        if (syndrome)
            QLDPC_CUDA(cudaMemcpyAsync(ln.syn.p, syndrome + (size_t)f0 * d->syn_words, (size_t)nf * d->syn_words * 4,
                                       cudaMemcpyHostToDevice, ln.st));
        if (!fused && (rc = qldpc_make_llr_device(dec, ln.bits.p, dk, dp, llr_noisy, llr_known, nf, ln.in.p, ln.st))) break;
        mark(ln.st);
        if ((rc = decode_device_impl(dec, fused ? nullptr : ln.in.p, syndrome ? ln.syn.p : nullptr, nf, ln.out.p, ln.ok.p,
                                     ln.iters.p, nullptr, ln.st, shared_scratch ? 0 : (idx & 1), fused ? ln.bits.p : nullptr,
                                     mag))) break;
        mark(ln.st);
        QLDPC_CUDA(cudaMemcpyAsync(out_bits + (size_t)f0 * d->out_words, ln.out.p, (size_t)nf * d->out_words * 4,
                                   cudaMemcpyDeviceToHost, ln.st));
        if (ok) QLDPC_CUDA(cudaMemcpyAsync(ok + f0, ln.ok.p, nf, cudaMemcpyDeviceToHost, ln.st));
        if (iters) QLDPC_CUDA(cudaMemcpyAsync(iters + f0, ln.iters.p, (size_t)nf * 2, cudaMemcpyDeviceToHost, ln.st));
        mark(ln.st);
    }
    const double host_issue_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    for (auto &ln : d->lanes.lane) {
        cudaError_t e = cudaStreamSynchronize(ln.st);
        if (e != cudaSuccess && rc == QLDPC_OK) rc = CudaCheck::fail(e, "cudaStreamSynchronize");
    }
    if (timeline && rc == QLDPC_OK && tl.size() == (size_t)(1 + 5 * idx)) {
        const double host_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
        std::fprintf(stderr, "[qldpc timeline] %d frames, %d chunks, host issue %.3f ms, host total %.3f ms\n", n_frames, idx,
                     host_issue_ms, host_total_ms);
        for (int k = 0, f0 = 0; k < idx; ++k) {
            float t[5];
            for (int j = 0; j < 5; ++j) cudaEventElapsedTime(&t[j], tl[0], tl[1 + 5 * k + j]);
            const int nfk = std::min(plan.at(k), n_frames - f0);
            std::fprintf(stderr, "[qldpc timeline] chunk %2d lane %d frames %6d: start %7.3f h2d %7.3f llr %7.3f decode %7.3f d2h %7.3f ms\n",
                         k, k & 1, nfk, t[0], t[1], t[2], t[3], t[4]);
            f0 += nfk;
        }
    }
    for (auto e : tl) cudaEventDestroy(e);
    return rc;
}

extern "C" int qldpc_encode_nr_device(qldpc_decoder *dec, const uint32_t *d_msg, int32_t n_frames, uint32_t *d_cword,
                                      void *cuda_stream)
{
    if (!dec || !d_msg || !d_cword || n_frames < 0) return QLDPC_ERR_ARG;
    qldpc_decoder_full *d = full(dec);
    if (!d->children.empty()) return QLDPC_ERR_UNSUPPORTED;
    if (!d->code.has_nr_core()) return QLDPC_ERR_UNSUPPORTED;
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    const HostCode &c = d->code;
    const int kbits = (c.base_cols - c.base_rows) * c.z;
    int rc;
    if (c.z % 32 == 0 && d->d_qc_layers.p && d->d_qc_aux.p && c.z / 32 <= 1024) {   // packed words
        const int kb = c.base_cols - c.base_rows;
        const int p1_sh = c.base[1 * c.base_cols + kb] == -1 ? c.base[2 * c.base_cols + kb] : c.base[1 * c.base_cols + kb];
        rc = launch_encode_nr_packed(d_msg, d_cword, n_frames, c.z, c.base_rows, c.base_cols, (c.z - p1_sh % c.z) % c.z, kbits / 32,
                                     d->cw_words, d->d_qc_layers.p, d->d_qc_aux.p, (cudaStream_t)cuda_stream);
    } else {
        rc = launch_encode_nr(d_msg, d_cword, n_frames, c.z, c.base_rows, c.base_cols, d->d_base.p, (kbits + 31) / 32, d->cw_words,
                              (cudaStream_t)cuda_stream);
    }
    if (!rc && n_frames) d->kernel_launches++;
    return rc;
}

extern "C" int qldpc_encode_nr(qldpc_decoder *dec, const uint32_t *msg, int32_t n_frames, uint32_t *cword)
{
    if (!dec || !msg || !cword || n_frames < 0) return QLDPC_ERR_ARG;
    if (n_frames == 0) return QLDPC_OK;
    qldpc_decoder_full *d = full(dec);
    if (!d->code.has_nr_core()) return QLDPC_ERR_UNSUPPORTED;
    const HostCode &c = d->code;
    const int msg_words = ((c.base_cols - c.base_rows) * c.z + 31) / 32;
    if (!d->children.empty())
        return shard_frames(d, n_frames, [&](qldpc_decoder *ch, int f0, int nf) {
            return qldpc_encode_nr(ch, msg + (size_t)f0 * msg_words, nf, cword + (size_t)f0 * d->cw_words);
        });
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    Lane &ln = d->lanes.lane[0];
    int rc;
    if ((rc = d->d_tmp_bits.ensure((size_t)n_frames * msg_words))) return rc;
    if ((rc = ln.out.ensure((size_t)n_frames * d->cw_words))) return rc;
    QLDPC_CUDA(cudaMemcpyAsync(d->d_tmp_bits.p, msg, (size_t)n_frames * msg_words * 4, cudaMemcpyHostToDevice, ln.st));
    if ((rc = qldpc_encode_nr_device(dec, d->d_tmp_bits.p, n_frames, ln.out.p, ln.st))) return rc;
    QLDPC_CUDA(cudaMemcpyAsync(cword, ln.out.p, (size_t)n_frames * d->cw_words * 4, cudaMemcpyDeviceToHost, ln.st));
    QLDPC_CUDA(cudaStreamSynchronize(ln.st));
    return QLDPC_OK;
}

// ----------------------------------------------------------------------------------- statistics

extern "C" int qldpc_get_stats(qldpc_decoder *dec, qldpc_stats *out)
{
    if (!dec || !out) return QLDPC_ERR_ARG;
    qldpc_decoder_full *d = full(dec);
    if (!d->children.empty()) {   // the only reduction of the multi-device path: on the host
        std::memset(out, 0, sizeof(*out));
        for (qldpc_decoder *ch : d->children) {
            qldpc_stats s;
            if (int rc = qldpc_get_stats(ch, &s)) return rc;
            out->frames += s.frames; out->failures += s.failures; out->iter_sum += s.iter_sum;
            out->kernel_launches += s.kernel_launches;
            for (int i = 0; i < QLDPC_ITER_HIST_BINS; ++i) out->iter_hist[i] += s.iter_hist[i];
        }
        return QLDPC_OK;
    }
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    QLDPC_CUDA(cudaDeviceSynchronize());
    DevStats h;
    QLDPC_CUDA(cudaMemcpy(&h, d->d_stats.p, sizeof(h), cudaMemcpyDeviceToHost));
    out->frames = h.frames;
    out->failures = h.failures;
    out->iter_sum = h.iter_sum;
    for (int i = 0; i < QLDPC_ITER_HIST_BINS; ++i) out->iter_hist[i] = h.hist[i];
    out->kernel_launches = d->kernel_launches;
    return QLDPC_OK;
}

extern "C" int qldpc_reset_stats(qldpc_decoder *dec)
{
    if (!dec) return QLDPC_ERR_ARG;
    qldpc_decoder_full *d = full(dec);
    if (!d->children.empty()) {
        for (qldpc_decoder *ch : d->children)
            if (int rc = qldpc_reset_stats(ch)) return rc;
        return QLDPC_OK;
    }
    QLDPC_CUDA(cudaSetDevice(d->cfg.device));
    QLDPC_CUDA(cudaDeviceSynchronize());
    QLDPC_CUDA(cudaMemset(d->d_stats.p, 0, sizeof(DevStats)));
    d->kernel_launches = 0;
    return QLDPC_OK;
}

extern "C" const char *qldpc_strerror(int code)
{
    switch (code) {
    case QLDPC_OK: return "ok";
    case QLDPC_ERR_ARG: return "invalid argument";
    case QLDPC_ERR_IO: return "cannot open matrix file";
    case QLDPC_ERR_FORMAT: return "malformed matrix file";
    case QLDPC_ERR_NOMEM: return "out of memory";
    case QLDPC_ERR_CUDA: return "CUDA error";
    case QLDPC_ERR_UNSUPPORTED: return "unsupported schedule / rule / dtype / code combination";
    case QLDPC_ERR_NO_DEVICE: return "no sm_100 CUDA device";
    default: return "unknown error";
    }
}

extern "C" const char *qldpc_last_cuda_error(void) { return CudaCheck::last.c_str(); }
extern "C" int qldpc_version(void) { return QLDPC_VERSION; }

// ----------------------------------------------------------------------- after reconciliation

static int select_sm100_device(int device)
{
    // the verdict per device is remembered: cudaGetDeviceProperties costs milliseconds, and the post-processing entry
    // points are called once per key block
    static std::atomic<int> verdict[64];                 // 0 unknown, 1 ok, 2 not an sm_100 device
    if (device < 0) return QLDPC_ERR_ARG;
    if (device < 64 && verdict[device].load(std::memory_order_relaxed) == 1) {
        QLDPC_CUDA(cudaSetDevice(device));
        return QLDPC_OK;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return QLDPC_ERR_NO_DEVICE; }
    if (device >= ndev) return QLDPC_ERR_ARG;
    QLDPC_CUDA(cudaSetDevice(device));
    int major = 0;
    QLDPC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (device < 64) verdict[device].store(major == 10 ? 1 : 2, std::memory_order_relaxed);
    if (major != 10) return QLDPC_ERR_NO_DEVICE;
    return QLDPC_OK;
}

extern "C" int qldpc_privacy_amplify(int32_t device, const uint32_t *key, int32_t key_stride_words, const int32_t *workbits,
                                     const int32_t *final_bits, const uint32_t *seeds, int32_t n_blocks, uint32_t *final_key,
                                     int32_t out_stride_words)
{
    if (!key || !workbits || !final_bits || !seeds || !final_key || n_blocks < 0 || key_stride_words <= 0 || out_stride_words < 0)
        return QLDPC_ERR_ARG;
    if (n_blocks == 0) return QLDPC_OK;
    int max_wb = 0, max_fb = 0;
    for (int b = 0; b < n_blocks; ++b) {
        if (workbits[b] < 0 || workbits[b] > 32 * key_stride_words || final_bits[b] < 0 || final_bits[b] > 32 * out_stride_words)
            return QLDPC_ERR_ARG;
        max_wb = std::max(max_wb, workbits[b]);
        max_fb = std::max(max_fb, final_bits[b]);
    }
    int rc;
    if ((rc = select_sm100_device(device))) return rc;
    static thread_local int tables_on = -1;
    if (tables_on != device) {
        if ((rc = pa_upload_jump_tables())) return rc;
        tables_on = device;
    }
    // device staging buffers are kept per host thread: an ecd2 handler calls this once per block
    static thread_local DevBuf<uint32_t> d_key, d_seed, d_out;
    static thread_local DevBuf<int32_t> d_wb, d_fb;
    static thread_local int buf_device = -1;
    if (buf_device != device) { d_key.release(); d_seed.release(); d_out.release(); d_wb.release(); d_fb.release(); buf_device = device; }
    const size_t nk = (size_t)n_blocks * key_stride_words, no = (size_t)n_blocks * std::max(1, out_stride_words);
    if ((rc = d_key.ensure(nk)) || (rc = d_seed.ensure(n_blocks)) || (rc = d_out.ensure(no)) || (rc = d_wb.ensure(n_blocks)) ||
        (rc = d_fb.ensure(n_blocks))) return rc;
    QLDPC_CUDA(cudaMemcpy(d_key.p, key, nk * 4, cudaMemcpyHostToDevice));
    QLDPC_CUDA(cudaMemcpy(d_seed.p, seeds, (size_t)n_blocks * 4, cudaMemcpyHostToDevice));
    QLDPC_CUDA(cudaMemcpy(d_wb.p, workbits, (size_t)n_blocks * 4, cudaMemcpyHostToDevice));
    QLDPC_CUDA(cudaMemcpy(d_fb.p, final_bits, (size_t)n_blocks * 4, cudaMemcpyHostToDevice));
    if (out_stride_words > 0) QLDPC_CUDA(cudaMemcpy(d_out.p, final_key, no * 4, cudaMemcpyHostToDevice));   // untouched words survive
    if ((rc = launch_privacy_amplify(d_key.p, d_wb.p, d_fb.p, d_seed.p, n_blocks, key_stride_words, max_wb, max_fb, d_out.p,
                                     out_stride_words, nullptr))) return rc;
    QLDPC_CUDA(cudaDeviceSynchronize());
    if (out_stride_words > 0) QLDPC_CUDA(cudaMemcpy(final_key, d_out.p, no * 4, cudaMemcpyDeviceToHost));
    return QLDPC_OK;
}

extern "C" int qldpc_crc32_frames(int32_t device, const uint32_t *bits, int32_t n_frames, int32_t words_per_frame,
                                  int32_t stride_words, uint32_t *crc_out)
{
    if (!bits || !crc_out || n_frames < 0 || words_per_frame < 0 || stride_words < words_per_frame) return QLDPC_ERR_ARG;
    if (n_frames == 0) return QLDPC_OK;
    int rc;
    if ((rc = select_sm100_device(device))) return rc;
    // device staging buffers are kept per host thread (the confirmation step calls this once per key block)
    static thread_local DevBuf<uint32_t> d_bits, d_crc;
    static thread_local int buf_device = -1;
    if (buf_device != device) { d_bits.release(); d_crc.release(); buf_device = device; }
    if ((rc = d_bits.ensure((size_t)n_frames * stride_words)) || (rc = d_crc.ensure(n_frames))) return rc;
    QLDPC_CUDA(cudaMemcpy(d_bits.p, bits, (size_t)n_frames * stride_words * 4, cudaMemcpyHostToDevice));
    if ((rc = launch_crc32_frames(d_bits.p, n_frames, words_per_frame, stride_words, d_crc.p, nullptr))) return rc;
    QLDPC_CUDA(cudaDeviceSynchronize());
    QLDPC_CUDA(cudaMemcpy(crc_out, d_crc.p, (size_t)n_frames * 4, cudaMemcpyDeviceToHost));
    return QLDPC_OK;
}

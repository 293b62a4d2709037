// Alice's side for codes WITHOUT the NR structure: systematic encoders for arbitrary parity-check matrices.
//
//   qldpc_encoder_from_h            module::Encoder_LDPC_from_H<B>(K, N, H, G_method, ...)   "main.cpp (alist-v1.0.1)":144
//   qldpc_encoder_from_g_alist_file module::Encoder_LDPC<B>(K, N, G, n_frames)               "main.cpp (alist)":143
//   qldpc_encode                    m.encoder->encode(b.ref_bits, b.enc_bits)                "main.cpp (alist)":417
// Both reduce to one dense GF(2) matrix T (N rows of ceil(K/32) words): codeword bit j = parity(T_j & u).  From H it is
// found by Gauss-Jordan elimination on the host (pivot columns become parity positions; the code's info_bits_pos are
// kept as information positions wherever H allows it, as AFF3CT's "IDENTITY" method does for BOOT/matrices/H/
// PEGReg504x1008.alist, whose parity part is columns 0..503); from a generator matrix file T is G transposed.
// The product u*G itself runs on the GPU, bit-packed: one thread per 32 codeword bits, AND + popc over the message words.
#include <algorithm>
#include <new>

#include "kernels.hpp"

struct qldpc_encoder {
    int n = 0, k = 0, kw = 0, cw = 0, device = 0;
    std::vector<int32_t> info_pos;
    qldpc::DevBuf<uint32_t> d_T, d_msg, d_cw;
    cudaStream_t st = nullptr;
};

namespace qldpc {
namespace {

// T: N rows of kw words (bit i of the message = word i/32, mask 1 << (31 - i%32), as everywhere in this library)
__global__ void encode_dense_kernel(const uint32_t *__restrict__ T, const uint32_t *__restrict__ msg, int F, int N, int kw, int cw,
                                    uint32_t *__restrict__ cword)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)F * cw) return;
    const int f = (int)(t / cw), w = (int)(t - (long long)f * cw);
    const uint32_t *u = msg + (size_t)f * kw;
    uint32_t out = 0;
    for (int b = 0; b < 32; ++b) {
        const int j = 32 * w + b;
        if (j >= N) break;
        const uint32_t *row = T + (size_t)j * kw;
        uint32_t acc = 0;
        for (int i = 0; i < kw; ++i) acc ^= row[i] & u[i];
        out |= (uint32_t)(__popc(acc) & 1) << (31 - b);
    }
    cword[t] = out;
}

int finish(qldpc_encoder *e, const std::vector<uint32_t> &T)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return QLDPC_ERR_NO_DEVICE; }
    if (e->device < 0 || e->device >= ndev) return QLDPC_ERR_ARG;
    QLDPC_CUDA(cudaSetDevice(e->device));
    cudaDeviceProp prop;
    QLDPC_CUDA(cudaGetDeviceProperties(&prop, e->device));
    if (prop.major != 10) return QLDPC_ERR_NO_DEVICE;
    if (int rc = e->d_T.upload(T)) return rc;
    QLDPC_CUDA(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
    return QLDPC_OK;
}

}  // namespace

// Gauss-Jordan on H (m x n, CSR) over GF(2).  Pivots are searched first among the columns that are NOT information
// positions (ascending), then among the others: with a full-rank parity part the code's info_bits_pos survive unchanged.
// Returns the dense map T and the final information positions.
int systematic_from_H(const HostCode &c, std::vector<uint32_t> &T, std::vector<int32_t> &info_pos)
{
    const int n = c.n, m = c.m;
    if ((long long)n * m > (1ll << 31)) return QLDPC_ERR_UNSUPPORTED;   // dense elimination: codes up to ~46k x 46k
    const int nw = (n + 63) / 64;
    std::vector<uint64_t> A((size_t)m * nw, 0);
    for (int r = 0; r < m; ++r)
        for (int e = c.row_ptr[r]; e < c.row_ptr[r + 1]; ++e) A[(size_t)r * nw + c.col_idx[e] / 64] ^= 1ull << (c.col_idx[e] % 64);
    std::vector<char> is_info(n, 0);
    for (int p : c.info_pos) is_info[p] = 1;
    std::vector<int> order;
    for (int j = 0; j < n; ++j) if (!is_info[j]) order.push_back(j);
    for (int j = 0; j < n; ++j) if (is_info[j]) order.push_back(j);
    std::vector<int> pivot_col;   // pivot_col[row]
    int rank = 0;
    for (int j : order) {
        if (rank == m) break;
        int pr = -1;
        for (int r = rank; r < m; ++r)
            if (A[(size_t)r * nw + j / 64] >> (j % 64) & 1) { pr = r; break; }
        if (pr < 0) continue;
        if (pr != rank) std::swap_ranges(A.begin() + (size_t)pr * nw, A.begin() + (size_t)(pr + 1) * nw, A.begin() + (size_t)rank * nw);
        const uint64_t *prow = &A[(size_t)rank * nw];
        for (int r = 0; r < m; ++r)
            if (r != rank && (A[(size_t)r * nw + j / 64] >> (j % 64) & 1)) {
                uint64_t *row = &A[(size_t)r * nw];
                for (int w = 0; w < nw; ++w) row[w] ^= prow[w];
            }
        pivot_col.push_back(j);
        ++rank;
    }
    std::vector<char> is_pivot(n, 0);
    for (int j : pivot_col) is_pivot[j] = 1;
    info_pos.clear();
    for (int j = 0; j < n; ++j) if (!is_pivot[j]) info_pos.push_back(j);
    const int k = (int)info_pos.size(), kw = (k + 31) / 32;
    if (k == 0) return QLDPC_ERR_UNSUPPORTED;
    if ((long long)n * kw > (1ll << 28)) return QLDPC_ERR_UNSUPPORTED;   // 1 GiB of dense map
    T.assign((size_t)n * kw, 0u);
    for (int i = 0; i < k; ++i) T[(size_t)info_pos[i] * kw + i / 32] = 1u << (31 - i % 32);   // systematic positions
    for (int r = 0; r < rank; ++r) {   // x[pivot] = XOR of the information bits of its row
        uint32_t *trow = &T[(size_t)pivot_col[r] * kw];
        const uint64_t *row = &A[(size_t)r * nw];
        for (int i = 0; i < k; ++i)
            if (row[info_pos[i] / 64] >> (info_pos[i] % 64) & 1) trow[i / 32] |= 1u << (31 - i % 32);
    }
    return QLDPC_OK;
}

}  // namespace qldpc

using namespace qldpc;

extern "C" int qldpc_encoder_from_h(const qldpc_code *code, int32_t device, qldpc_encoder **out)
{
    if (!code || !out) return QLDPC_ERR_ARG;
    qldpc_encoder *e = new (std::nothrow) qldpc_encoder();
    if (!e) return QLDPC_ERR_NOMEM;
    std::vector<uint32_t> T;
    int rc = systematic_from_H(code->h, T, e->info_pos);
    e->n = code->h.n; e->k = (int)e->info_pos.size(); e->kw = (e->k + 31) / 32; e->cw = (e->n + 31) / 32; e->device = device;
    if (!rc) rc = finish(e, T);
    if (rc) { delete e; return rc; }
    *out = e;
    return QLDPC_OK;
}

extern "C" int qldpc_encoder_from_g_alist_file(const char *path, int32_t device, qldpc_encoder **out)
{
    if (!path || !out) return QLDPC_ERR_ARG;
    HostCode g;   // the file's "variables" are the N codeword bits, its "checks" the K rows of G
    if (int rc = parse_alist(path, g)) return rc;
    qldpc_encoder *e = new (std::nothrow) qldpc_encoder();
    if (!e) return QLDPC_ERR_NOMEM;
    e->n = g.n; e->k = g.m; e->kw = (e->k + 31) / 32; e->cw = (e->n + 31) / 32; e->device = device;
    std::vector<uint32_t> T((size_t)e->n * e->kw, 0u);
    for (int r = 0; r < g.m; ++r)
        for (int x = g.row_ptr[r]; x < g.row_ptr[r + 1]; ++x) T[(size_t)g.col_idx[x] * e->kw + r / 32] |= 1u << (31 - r % 32);
    // information positions: the columns of G that are unit vectors, one per row (tools::LDPC_matrix_handler reports them
    // with the G file, "main.cpp (alist)":333); a non-systematic G has none
    std::vector<int32_t> pos(e->k, -1);
    for (int j = 0; j < g.n; ++j)
        if (g.var_ptr[j + 1] - g.var_ptr[j] == 1) {
            const int edge = g.var_edge[g.var_ptr[j]];
            const int r = (int)(std::upper_bound(g.row_ptr.begin(), g.row_ptr.end(), edge) - g.row_ptr.begin()) - 1;
            if (pos[r] < 0) pos[r] = j;
        }
    if (std::all_of(pos.begin(), pos.end(), [](int p) { return p >= 0; })) e->info_pos = pos;
    const int rc = finish(e, T);
    if (rc) { delete e; return rc; }
    *out = e;
    return QLDPC_OK;
}

extern "C" void qldpc_encoder_free(qldpc_encoder *enc)
{
    if (!enc) return;
    cudaSetDevice(enc->device);
    if (enc->st) { cudaStreamSynchronize(enc->st); cudaStreamDestroy(enc->st); }
    delete enc;
}

extern "C" int qldpc_encoder_get_info(const qldpc_encoder *enc, int32_t *k, int32_t *n)
{
    if (!enc) return QLDPC_ERR_ARG;
    if (k) *k = enc->k;
    if (n) *n = enc->n;
    return QLDPC_OK;
}

extern "C" int qldpc_encoder_info_bits_pos(const qldpc_encoder *enc, int32_t *pos)
{
    if (!enc || !pos) return QLDPC_ERR_ARG;
    if ((int)enc->info_pos.size() != enc->k) return QLDPC_ERR_UNSUPPORTED;   // non-systematic generator matrix
    std::copy(enc->info_pos.begin(), enc->info_pos.end(), pos);
    return QLDPC_OK;
}

extern "C" int qldpc_encode_device(qldpc_encoder *enc, const uint32_t *d_msg, int32_t n_frames, uint32_t *d_cword, void *cuda_stream)
{
    if (!enc || !d_msg || !d_cword || n_frames < 0) return QLDPC_ERR_ARG;
    if (n_frames == 0) return QLDPC_OK;
    QLDPC_CUDA(cudaSetDevice(enc->device));
    const long long items = (long long)n_frames * enc->cw;
    encode_dense_kernel<<<(unsigned)((items + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>(enc->d_T.p, d_msg, n_frames, enc->n, enc->kw,
                                                                                                  enc->cw, d_cword);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

extern "C" int qldpc_encode(qldpc_encoder *enc, const uint32_t *msg, int32_t n_frames, uint32_t *cword)
{
    if (!enc || !msg || !cword || n_frames < 0) return QLDPC_ERR_ARG;
    if (n_frames == 0) return QLDPC_OK;
    QLDPC_CUDA(cudaSetDevice(enc->device));
    int rc;
    if ((rc = enc->d_msg.ensure((size_t)n_frames * enc->kw)) || (rc = enc->d_cw.ensure((size_t)n_frames * enc->cw))) return rc;
    QLDPC_CUDA(cudaMemcpyAsync(enc->d_msg.p, msg, (size_t)n_frames * enc->kw * 4, cudaMemcpyHostToDevice, enc->st));
    if ((rc = qldpc_encode_device(enc, enc->d_msg.p, n_frames, enc->d_cw.p, enc->st))) return rc;
    QLDPC_CUDA(cudaMemcpyAsync(cword, enc->d_cw.p, (size_t)n_frames * enc->cw * 4, cudaMemcpyDeviceToHost, enc->st));
    QLDPC_CUDA(cudaStreamSynchronize(enc->st));
    return QLDPC_OK;
}

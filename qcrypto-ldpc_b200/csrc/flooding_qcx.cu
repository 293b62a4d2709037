// Flooding belief propagation for LARGE quasi-cyclic codes: host-side dispatch of the clustered kernel.  The kernel itself is
// flooding_qcx_impl.cuh; flooding_qcx_{f32_exact,f32_fast,f32_ms,i16,i8}.cu instantiate it per tier and flavour.
#include "kernels.hpp"

namespace qldpc {

#define QL_QCX_DECLARE(TAG)                                                                                         \
    int flooding_qcx_launch_##TAG(const FloodQcxParams &p, int n_clusters, int cl, int smem_bytes, cudaStream_t st); \
    int flooding_qcx_clusters_##TAG(int lanes, int cl, int smem_bytes);
QL_QCX_DECLARE(f32_exact)
QL_QCX_DECLARE(f32_fast)
QL_QCX_DECLARE(f32_ms)
QL_QCX_DECLARE(i16)
QL_QCX_DECLARE(i8)
#undef QL_QCX_DECLARE

int flooding_qcx_table_bytes(int brows, int bcols, int nnz)
{
    return (brows * 8 + nnz * 8 + (bcols + 2) * 4 + nnz * 8 + 16 * 4 + 16 + 15) / 16 * 16;
}
int flooding_qcx_smem_bytes(int brows, int bcols, int nnz, int dtype)
{
    (void)dtype;
    return flooding_qcx_table_bytes(brows, bcols, nnz);
}

int flooding_qcx_msg_bytes(int dtype) { return dtype == QLDPC_DTYPE_F32 ? 4 : (dtype == QLDPC_DTYPE_I16 ? 2 : 1); }
int flooding_qcx_post_bytes(int dtype) { return dtype == QLDPC_DTYPE_I8 ? 2 : 4; }
int flooding_qcx_run_lanes(int Z) { return Z; }
int flooding_qcx_lanes_per_thread(int max_row_degree) { return max_row_degree <= 8 ? 4 : 2; }   // kMaxDcV4 of the kernel

int flooding_qcx_max_clusters(int dtype, int lanes, int cl, int smem_bytes)
{
    // register count and shared memory are the same for all flavours of a type: ask for the min-sum one
    switch (dtype) {
    case QLDPC_DTYPE_F32: return flooding_qcx_clusters_f32_ms(lanes, cl, smem_bytes);
    case QLDPC_DTYPE_I16: return flooding_qcx_clusters_i16(lanes, cl, smem_bytes);
    case QLDPC_DTYPE_I8: return flooding_qcx_clusters_i8(lanes, cl, smem_bytes);
    default: return 0;
    }
}

int launch_flooding_qcx(const FloodQcxParams &p, int n_clusters, int cl, int smem_bytes, cudaStream_t st)
{
    switch (p.dtype) {
    case QLDPC_DTYPE_F32:
        if (p.rule != QLDPC_RULE_SPA) return flooding_qcx_launch_f32_ms(p, n_clusters, cl, smem_bytes, st);
        return p.fast_spa ? flooding_qcx_launch_f32_fast(p, n_clusters, cl, smem_bytes, st)
                          : flooding_qcx_launch_f32_exact(p, n_clusters, cl, smem_bytes, st);
    case QLDPC_DTYPE_I16: return flooding_qcx_launch_i16(p, n_clusters, cl, smem_bytes, st);
    case QLDPC_DTYPE_I8: return flooding_qcx_launch_i8(p, n_clusters, cl, smem_bytes, st);
    default: return QLDPC_ERR_UNSUPPORTED;
    }
}

}  // namespace qldpc

// Flooding belief propagation for QUASI-CYCLIC codes, float (SPA / normalised / offset min-sum).
//
// Same arithmetic and the same order of every sum as flooding.cu (AFF3CT Decoder_LDPC_BP_flooding as the reference
// instantiates it, BOOT/src/main.cpp:193,365; oracle/qldpc_oracle.c:ora_decode_flooding_f32), so the results are bit
// identical to it and to the oracle -- but the circulant structure is used instead of CSR gathers: the Z lanes of a
// circulant map onto consecutive threads, a check-to-variable message lives at c2v[edge][check lane], and both phases
// touch memory in runs of consecutive lanes (a cyclic shift is a rotated, still contiguous, run):
//   variable phase  (block column c, variable lane m):  post = llr + sum over the column's edges in ascending block-row
//                   order of c2v[e][(m - s_e) mod Z]      (ascending block row == ascending check index, as the oracle sums)
//   check phase     (block row r, check lane l):         x_e = post[c_e][(l + s_e) mod Z] - c2v[e][l] for the row's edges
//                   in ascending block-column order, each x_e kept in a register between the two passes of the update.
// No col_idx / var_edge index loads, no scattered 4-byte gathers.  One CTA per frame in flight; messages and posteriors
// in shared memory when they fit, else in a per-CTA global scratch.
#include "kernels.hpp"
#include "spa_math.cuh"

#ifndef QL_FQ_THREADS
#define QL_FQ_THREADS 1024
#endif
#ifndef QL_FQ_CACHE
#define QL_FQ_CACHE 0    // keep the var-to-check values of a check in registers between the two passes (needs > 64 registers)
#endif

namespace qldpc {

namespace {

__device__ __forceinline__ float spa_t(float x)
{
    const float t = tanh_half_exact(fabsf(x));
    return (t != 0.0f) ? t : 1e-12f;
}

// Check-node update of check (row ly, lane l).  MAXD > 0: the row has at most MAXD edges and their var-to-check values
// stay in registers between the two passes (fully unrolled, predicated); MAXD == 0: any degree, values are re-read.
// Returns the parity of (syndrome bit, hard decisions of the check's variables): the early-termination test of the sweep
// rides on the a-posteriori values the update reads anyway.
template <int MAXD>
__device__ __forceinline__ int check_row(const FloodQcParams &p, const QcLayer ly, const float *post, float *cm, int l, int Z,
                                         int synbit)
{
    constexpr int NX = MAXD > 0 ? MAXD : 1;
    const int deg = ly.degree;
    float x[NX];
    int sign = synbit, hard = synbit;
    auto v2c = [&](int j) {
        const QcEdgeAux ax = p.aux[ly.edge_begin + j];
        int vl = l + ax.shift;
        if (vl >= Z) vl -= Z;
        const float pv = post[ax.col * Z + vl];
        hard ^= (pv < 0.0f) ? 1 : 0;
        return pv - cm[(size_t)j * Z];
    };
    const int trip = MAXD > 0 ? MAXD : deg;
    if (p.rule == QLDPC_RULE_SPA) {
        float product = 1.0f;
#pragma unroll
        for (int j = 0; j < trip; ++j) {
            if (MAXD == 0 || j < deg) {
                const float xv = v2c(j);
                if (MAXD > 0) x[j] = xv;
                product *= spa_t(xv);
                sign ^= signbit(xv) ? 1 : 0;
            }
        }
        const int hard1 = hard;
#pragma unroll
        for (int j = 0; j < trip; ++j) {
            if (MAXD == 0 || j < deg) {
                const float xv = MAXD > 0 ? x[j] : v2c(j);
                float rr = product / spa_t(xv);
                rr = (rr < 1.0f) ? rr : 1.0f - 1.1920929e-07f;
                const float mag = two_atanh_exact(rr);
                cm[(size_t)j * Z] = (sign ^ (signbit(xv) ? 1 : 0)) ? -mag : mag;
            }
        }
        return hard1;
    } else {
        float min1 = 3.402823466e+38f, min2 = 3.402823466e+38f;
#pragma unroll
        for (int j = 0; j < trip; ++j) {
            if (MAXD == 0 || j < deg) {
                const float xv = v2c(j);
                if (MAXD > 0) x[j] = xv;
                const float a = fabsf(xv);
                sign ^= signbit(xv) ? 1 : 0;
                min2 = fminf(min2, fmaxf(a, min1));
                min1 = fminf(min1, a);
            }
        }
        const int hard1 = hard;
        float cst1, cst2;
        if (p.rule == QLDPC_RULE_NMS) {
            cst1 = min2 * p.norm;
            cst2 = min1 * p.norm;
        } else {
            cst1 = fmaxf(0.0f, min2 - p.offset);
            cst2 = fmaxf(0.0f, min1 - p.offset);
        }
#pragma unroll
        for (int j = 0; j < trip; ++j) {
            if (MAXD == 0 || j < deg) {
                const float xv = MAXD > 0 ? x[j] : v2c(j);
                const float mag = (fabsf(xv) == min1) ? cst1 : cst2;
                cm[(size_t)j * Z] = (sign ^ (signbit(xv) ? 1 : 0)) ? -mag : mag;
            }
        }
        return hard1;
    }
}

__global__ void __launch_bounds__(QL_FQ_THREADS, 1) flooding_qc_kernel(const FloodQcParams p)
{
    extern __shared__ __align__(16) char smem[];
    float *c2v, *post;
    if (p.use_smem) {
        c2v = reinterpret_cast<float *>(smem);
        post = c2v + (size_t)p.nnz * p.Z;
    } else {
        c2v = p.c2v + (size_t)blockIdx.x * p.nnz * p.Z;
        post = p.post + (size_t)blockIdx.x * p.N;
    }
    const int tid = threadIdx.x, nt = blockDim.x, Z = p.Z;
    const int E = p.nnz * Z;

    for (int f = blockIdx.x; f < p.F; f += gridDim.x) {
        const float *llr = p.llr + (size_t)f * p.N;
        const uint32_t *syn = p.syn ? p.syn + (size_t)f * p.syn_words : nullptr;
        for (int e = tid; e < E; e += nt) c2v[e] = 0.0f;     // decoder.reset(), BOOT/src/main.cpp:389
        __syncthreads();

        int it = 0, depth = 0;
        bool ok = false;
        for (;;) {
            // ---- variable phase
            for (int v = tid; v < p.N; v += nt) {
                const int c = v / Z, m = v - c * Z;
                float sum = 0.0f;
                for (int k = p.col_ptr[c]; k < p.col_ptr[c + 1]; ++k) {
                    const int2 ce = p.col_edges[k];            // edge id, shift
                    int l = m - ce.y;
                    if (l < 0) l += Z;
                    sum += c2v[ce.x * Z + l];
                }
                post[v] = llr[v] + sum;
            }
            __syncthreads();
            const bool last = it >= p.max_iter;
            if (last) {   // final verdict after the last sweep: syndrome of the hard decisions, no update
                int bad = 0;
                for (int mi = tid; mi < p.M; mi += nt) {
                    const int r = mi / Z, l = mi - r * Z;
                    unsigned s = syn ? (syn[mi >> 5] >> (31 - (mi & 31))) & 1u : 0u;
                    const QcLayer ly = p.layers[r];
                    for (int e = ly.edge_begin; e < ly.edge_begin + ly.degree; ++e) {
                        const QcEdgeAux ax = p.aux[e];
                        int vl = l + ax.shift;
                        if (vl >= Z) vl -= Z;
                        s ^= (unsigned)(post[ax.col * Z + vl] < 0.0f);
                    }
                    bad |= (int)(s & 1u);
                }
                ok = __syncthreads_or(bad) == 0;
                break;
            }
            // ---- check phase; the early-termination test (enable_syndrome) of this sweep is computed on the way.
            // If it passes the decoder stops here: the messages just written are never used, `it` is not advanced.
            const bool want_check = p.early_stop && it > 0;
            int bad = 0;
            for (int mi = tid; mi < p.M; mi += nt) {
                const int r = mi / Z, l = mi - r * Z;
                const int synbit = syn ? (int)((syn[mi >> 5] >> (31 - (mi & 31))) & 1u) : 0;
                const QcLayer ly = p.layers[r];
                [[maybe_unused]] const int deg = ly.degree;
                float *cm = c2v + (size_t)ly.edge_begin * Z + l;      // the row's messages of this lane, Z apart
                int hard;
#if QL_FQ_CACHE
                if (deg <= 8) hard = check_row<8>(p, ly, post, cm, l, Z, synbit);
                else if (deg <= 16) hard = check_row<16>(p, ly, post, cm, l, Z, synbit);
                else
#endif
                hard = check_row<0>(p, ly, post, cm, l, Z, synbit);
                bad |= hard;
            }
            if (want_check) {
                ok = __syncthreads_or(bad) == 0;
                if (ok) { if (++depth >= p.syndrome_depth) break; }
                else depth = 0;
            } else {
                __syncthreads();
            }
            ++it;
        }

        // ---- outputs
        uint32_t *ab = p.allbits + (size_t)f * p.cw_words;
        // one variable per thread, one ballot per warp and output word (was: 32 serial loads and compares per word on the
        // first cw_words threads)
        for (int v0 = 0; v0 < p.cw_words * 32; v0 += nt) {
            const int v = v0 + tid;
            const unsigned b = __ballot_sync(0xffffffffu, v < p.N && post[v] < 0.0f);
            if ((tid & 31) == 0 && (v >> 5) < p.cw_words) ab[v >> 5] = __brev(b);   // bit 31 - k of word w = variable 32 w + k
        }
        if (p.posterior) {
            float *po = p.posterior + (size_t)f * p.N;
            for (int v = tid; v < p.N; v += nt) po[v] = post[v];
        }
        if (tid == 0) {
            if (p.ok) p.ok[f] = ok ? 1 : 0;
            if (p.iters) p.iters[f] = (uint16_t)it;
            if (p.stats) {
                atomicAdd(&p.stats->frames, 1ull);
                if (!ok) atomicAdd(&p.stats->failures, 1ull);
                atomicAdd(&p.stats->iter_sum, (unsigned long long)it);
                atomicAdd(&p.stats->hist[min(it, QLDPC_ITER_HIST_BINS - 1)], 1ull);
            }
        }
        __syncthreads();
    }
}

}  // namespace

int layered_flood_qc_threads() { return QL_FQ_THREADS; }

int launch_flooding_qc(const FloodQcParams &p, int grid, int block, int smem_bytes, cudaStream_t st)
{
    QLDPC_CUDA(cudaFuncSetAttribute(flooding_qc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    flooding_qc_kernel<<<grid, block, smem_bytes, st>>>(p);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

}  // namespace qldpc

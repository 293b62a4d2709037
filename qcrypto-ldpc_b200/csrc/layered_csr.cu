// Horizontal-layered belief propagation on an ARBITRARY parity-check matrix (CSR), float, SPA / NMS / OMS.
//
// Reference: module::Decoder_LDPC_BP_horizontal_layered<B,Q,Update_rule_*> takes any tools::Sparse_matrix
// ("main.cpp (5g-qc)":256-270; logged once on a DVB-S2 matrix, BOOT/data_dvb/data2/DVB_S2_N_16200_K_14400_CR_0.888889.txt:26);
// arithmetic restated in oracle/qldpc_oracle.c:ora_decode_layered_f32 (row-serial: contrib = var - branch,
// branch = rule(contribs of the check), var = contrib + branch; early stop after every iteration but the last).
//
// Inside one frame the rows of a general H are strictly serial (consecutive checks share variables), so the parallelism
// is ACROSS frames: one thread per frame, all threads of a warp walk the same row at the same time.  The state is laid out
// frame-minor -- var[v][t], branch[e][t] with t the thread's column -- so that every access of a warp is one coalesced
// 128-byte line and the row tables are warp-uniform broadcast loads.  A thread leaves the iteration loop when its frame
// has converged.  tanh / atanh in double, rounded once, as in the other bit-exact float kernels.
#include "kernels.hpp"
#include "spa_math.cuh"

namespace qldpc {

namespace {

constexpr int kMaxDeg = 64;
constexpr int kUnrolledDeg = 16;     // rows up to this degree run the register-resident row code below

// NR consecutive checks of degree D for this thread's frame: the column indices (warp-uniform loads), then all 2 NR D state
// loads in flight together, the rows in registers, 2 NR D stores -- one global round trip per row instead of D dependent
// ones, and no local-memory arrays.  Same operation order inside a row as the general loop (and the oracle).  NR = 2 (two
// consecutive rows of equal degree with disjoint columns commute) was measured: PEGReg504x1008 3.24 -> 2.74 Gbit/s (the
// uniform disjointness test and 128 registers with spills cost more than the second row in flight gains); NR = 1 is used.
template <int D, int NR>
__device__ __forceinline__ void csr_rows(const LayeredCsrParams &p, float *var, float *br, int e0, size_t T, const int (&sign0)[NR], bool first)
{
    int vo[NR][D];
    float x[NR][D], vals[NR][D];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < D; ++j) vo[r][j] = __ldg(p.col_idx + e0 + r * D + j);
    float *bre = br + (size_t)e0 * T;
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < D; ++j) x[r][j] = var[(size_t)vo[r][j] * T];
    if (!first) {   // first iteration: every message is still zero and the scratch has not been written yet
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int j = 0; j < D; ++j) x[r][j] -= bre[(size_t)(r * D + j) * T];
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        int sign = sign0[r];
        float product = 1.0f, min1 = 3.402823466e+38f, min2 = 3.402823466e+38f;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const float a = fabsf(x[r][j]);
            sign ^= signbit(x[r][j]) ? 1 : 0;
            if (p.rule == QLDPC_RULE_SPA) {
                const float th = p.fast_spa ? tanh_half_fast(a) : tanh_half_exact(a);
                vals[r][j] = (th != 0.0f) ? th : 1e-12f;
                product *= vals[r][j];
            } else {
                min2 = fminf(min2, fmaxf(a, min1));
                min1 = fminf(min1, a);
            }
        }
        float cst1 = 0.0f, cst2 = 0.0f;
        if (p.rule == QLDPC_RULE_NMS) { cst1 = min2 * p.norm; cst2 = min1 * p.norm; }
        else if (p.rule == QLDPC_RULE_OMS) { cst1 = fmaxf(0.0f, min2 - p.offset); cst2 = fmaxf(0.0f, min1 - p.offset); }
#pragma unroll
        for (int j = 0; j < D; ++j) {
            float mag;
            if (p.rule == QLDPC_RULE_SPA) {
                float q = product / vals[r][j];
                q = (q < 1.0f) ? q : 1.0f - 1.1920929e-07f;
                mag = p.fast_spa ? two_atanh_fast(q) : two_atanh_exact(q);
            } else {
                mag = (fabsf(x[r][j]) == min1) ? cst1 : cst2;
            }
            const float out = (sign ^ (signbit(x[r][j]) ? 1 : 0)) ? -mag : mag;
            bre[(size_t)(r * D + j) * T] = out;
            var[(size_t)vo[r][j] * T] = x[r][j] + out;
        }
    }
}

// parity of the hard decisions of one check of degree D: the D loads are in flight together
template <int D>
__device__ __forceinline__ unsigned csr_row_parity(const LayeredCsrParams &p, const float *var, int e0, size_t T)
{
    float v[D];
#pragma unroll
    for (int j = 0; j < D; ++j) v[j] = var[(size_t)__ldg(p.col_idx + e0 + j) * T];
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < D; ++j) s ^= (unsigned)(v[j] < 0.0f);
    return s;
}

__global__ void __launch_bounds__(128, 4) layered_csr_kernel(const LayeredCsrParams p)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, T = gridDim.x * blockDim.x;
    float *var = p.var + t, *br = p.branch + t;     // column t of the frame-minor state, stride T
    float contrib[kMaxDeg], vals[kMaxDeg];
    for (int f = t; f < p.F; f += T) {
        const float *llr = p.llr + (size_t)f * p.N;
        const uint32_t *syn = p.syn ? p.syn + (size_t)f * p.syn_words : nullptr;
        for (int v = 0; v < p.N; ++v) var[(size_t)v * T] = llr[v];
        // branch[] is not zero-filled: the first iteration takes every old message as zero and writes all of them
        auto syndrome_ok = [&]() {
            unsigned bad = 0;
            for (int m = 0; m < p.M; ++m) {
                unsigned s = syn ? (syn[m >> 5] >> (31 - (m & 31))) & 1u : 0u;
                const int e0 = p.row_ptr[m], d = p.row_ptr[m + 1] - e0;
#define QL_D(D) case D: s ^= csr_row_parity<D>(p, var, e0, (size_t)T); break;
                switch (d) {   // warp-uniform
                    QL_D(1) QL_D(2) QL_D(3) QL_D(4) QL_D(5) QL_D(6) QL_D(7) QL_D(8)
                    QL_D(9) QL_D(10) QL_D(11) QL_D(12) QL_D(13) QL_D(14) QL_D(15) QL_D(16)
                default:
                    for (int e = e0; e < e0 + d; ++e) s ^= (unsigned)(var[(size_t)p.col_idx[e] * T] < 0.0f);
                    break;
                }
#undef QL_D
                bad |= s;
            }
            return bad == 0;
        };
        int executed = 0, depth = 0;
        for (int ite = 0; ite < p.max_iter; ++ite) {
            for (int m = 0; m < p.M; ++m) {
                const int e0 = p.row_ptr[m], d = p.row_ptr[m + 1] - e0;
                int sign = syn ? (int)((syn[m >> 5] >> (31 - (m & 31))) & 1u) : 0;
                if (d <= kUnrolledDeg) {
                    const int sg[1] = {sign};
#define QL_D(D) case D: csr_rows<D, 1>(p, var, br, e0, (size_t)T, sg, ite == 0); break;
                    switch (d) {   // warp-uniform
                        QL_D(1) QL_D(2) QL_D(3) QL_D(4) QL_D(5) QL_D(6) QL_D(7) QL_D(8)
                        QL_D(9) QL_D(10) QL_D(11) QL_D(12) QL_D(13) QL_D(14) QL_D(15) QL_D(16)
                    default: break;
                    }
#undef QL_D
                    continue;
                }
                float product = 1.0f, min1 = 3.402823466e+38f, min2 = 3.402823466e+38f;
                for (int j = 0; j < d; ++j) {
                    const float x = var[(size_t)p.col_idx[e0 + j] * T] - (ite == 0 ? 0.0f : br[(size_t)(e0 + j) * T]);
                    contrib[j] = x;
                    const float a = fabsf(x);
                    sign ^= signbit(x) ? 1 : 0;
                    if (p.rule == QLDPC_RULE_SPA) {
                        const float th = p.fast_spa ? tanh_half_fast(a) : tanh_half_exact(a);
                        const float r = (th != 0.0f) ? th : 1e-12f;
                        product *= r;
                        vals[j] = r;
                    } else {
                        min2 = fminf(min2, fmaxf(a, min1));
                        min1 = fminf(min1, a);
                    }
                }
                float cst1 = 0.0f, cst2 = 0.0f;
                if (p.rule == QLDPC_RULE_NMS) { cst1 = min2 * p.norm; cst2 = min1 * p.norm; }
                else if (p.rule == QLDPC_RULE_OMS) { cst1 = fmaxf(0.0f, min2 - p.offset); cst2 = fmaxf(0.0f, min1 - p.offset); }
                for (int j = 0; j < d; ++j) {
                    const float x = contrib[j];
                    float mag;
                    if (p.rule == QLDPC_RULE_SPA) {
                        float r = product / vals[j];
                        r = (r < 1.0f) ? r : 1.0f - 1.1920929e-07f;
                        mag = p.fast_spa ? two_atanh_fast(r) : two_atanh_exact(r);
                    } else {
                        mag = (fabsf(x) == min1) ? cst1 : cst2;
                    }
                    const float out = (sign ^ (signbit(x) ? 1 : 0)) ? -mag : mag;
                    br[(size_t)(e0 + j) * T] = out;
                    var[(size_t)p.col_idx[e0 + j] * T] = x + out;
                }
            }
            ++executed;
            if (p.early_stop && ite != p.max_iter - 1) {
                if (syndrome_ok()) { if (++depth == p.syndrome_depth) break; }
                else depth = 0;
            }
        }
        const bool ok = syndrome_ok();
        uint32_t *ab = p.allbits + (size_t)f * p.cw_words;
        for (int w = 0; w < p.cw_words; ++w) {
            uint32_t bits = 0;
            for (int b = 0; b < 32; ++b) {
                const int v = 32 * w + b;
                if (v < p.N && var[(size_t)v * T] < 0.0f) bits |= 1u << (31 - b);
            }
            ab[w] = bits;
        }
        if (p.posterior)
            for (int v = 0; v < p.N; ++v) p.posterior[(size_t)f * p.N + v] = var[(size_t)v * T];
        if (p.ok) p.ok[f] = ok ? 1 : 0;
        if (p.iters) p.iters[f] = (uint16_t)executed;
        if (p.stats) {
            atomicAdd(&p.stats->frames, 1ull);
            if (!ok) atomicAdd(&p.stats->failures, 1ull);
            atomicAdd(&p.stats->iter_sum, (unsigned long long)executed);
            atomicAdd(&p.stats->hist[min(executed, QLDPC_ITER_HIST_BINS - 1)], 1ull);
        }
    }
}

}  // namespace

int layered_csr_max_degree() { return kMaxDeg; }

int launch_layered_csr(const LayeredCsrParams &p, int threads, cudaStream_t st)
{
    layered_csr_kernel<<<threads / 128, 128, 0, st>>>(p);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

}  // namespace qldpc

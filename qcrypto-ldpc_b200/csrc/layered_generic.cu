// Generic layered decoder for quasi-cyclic codes: any lifting size, f32 / i16 / i8 messages,
// optional posterior output.  One thread per circulant lane, one CTA per frame in flight;
// beliefs and messages in a per-CTA global scratch (L2-resident).  This is the reference-shaped
// fallback; the tuned int8 path is layered_i8.cu.
//
// Integer arithmetic = ML/BPSK_nrldpc_sim_FP.m:35-94 (oracle: ora_decode_layered_fixed):
//   contrib = L - R_old (:51); t = clip(contrib, -(msg_max+1), msg_max) (:53-56);
//   min1/min2/sign product over the row (:59-63); offset or k/8 normalisation (:65-72);
//   R_new = parity*sign*mag (:73-75); L = clip(contrib + R_new, -(app_max+1), app_max) (:88-91).
// Float arithmetic = AFF3CT Decoder_LDPC_BP_horizontal_layered ("main.cpp (5g-qc)":256-270;
//   oracle: ora_decode_layered_f32): contrib = var - branch; branch = rule(contrib); var = contrib + branch.
// A QC block row is one layer: its Z checks touch disjoint variables, so processing them in
// parallel equals the row-serial order of both references.
#include <type_traits>

#include "kernels.hpp"

namespace qldpc {

namespace {

__device__ __forceinline__ int norm8(int v, int k)
{
    switch (k) {
    case 1: return v >> 3;
    case 2: return v >> 2;
    case 3: return (v >> 2) + (v >> 3);
    case 4: return v >> 1;
    case 5: return (v >> 1) + (v >> 3);
    case 6: return (v >> 1) + (v >> 2);
    case 7: return (v >> 1) + (v >> 2) + (v >> 3);
    default: return v;
    }
}
__device__ __forceinline__ int clipi(int x, int lo, int hi) { return min(max(x, lo), hi); }

__device__ __forceinline__ int syn_bit(const uint32_t *syn, int m) { return syn ? (int)((syn[m >> 5] >> (31 - (m & 31))) & 1u) : 0; }

template <typename T>
__device__ __forceinline__ void layer_lane(const LayeredGenParams &p, T *L, T *R, const QcEdgeAux *aux, int e0, int dc,
                                           int i, int synbit)
{
    const int Z = p.Z;
    if constexpr (std::is_floating_point<T>::value) {
        int sign = synbit;
        float product = 1.0f, min1 = 3.402823466e+38f, min2 = 3.402823466e+38f;
        for (int j = 0; j < dc; ++j) {
            const QcEdgeAux ax = aux[e0 + j];
            int l = i + ax.shift; if (l >= Z) l -= Z;
            const float x = L[ax.col * Z + l] - R[(e0 + j) * Z + i];
            const float a = fabsf(x);
            sign ^= signbit(x) ? 1 : 0;
            if (p.rule == QLDPC_RULE_SPA) {
                const float t = (float)tanh((double)(a * 0.5f));
                product *= (t != 0.0f) ? t : 1e-12f;
            } else {
                min2 = fminf(min2, fmaxf(a, min1));
                min1 = fminf(min1, a);
            }
        }
        float cst1 = 0.f, cst2 = 0.f;
        if (p.rule == QLDPC_RULE_NMS) { cst1 = min2 * p.norm; cst2 = min1 * p.norm; }
        else if (p.rule == QLDPC_RULE_OMS) { cst1 = fmaxf(0.0f, min2 - p.offset); cst2 = fmaxf(0.0f, min1 - p.offset); }
        for (int j = 0; j < dc; ++j) {
            const QcEdgeAux ax = aux[e0 + j];
            int l = i + ax.shift; if (l >= Z) l -= Z;
            const float x = L[ax.col * Z + l] - R[(e0 + j) * Z + i];
            float mag;
            if (p.rule == QLDPC_RULE_SPA) {
                const float t = (float)tanh((double)(fabsf(x) * 0.5f));
                float r = product / ((t != 0.0f) ? t : 1e-12f);
                r = (r < 1.0f) ? r : 1.0f - 1.1920929e-07f;
                mag = 2.0f * (float)atanh((double)r);
            } else {
                mag = (fabsf(x) == min1) ? cst1 : cst2;
            }
            const float out = (sign ^ (signbit(x) ? 1 : 0)) ? -mag : mag;
            R[(e0 + j) * Z + i] = out;
            L[ax.col * Z + l] = x + out;
        }
    } else {
        const int lo = -(p.msg_max + 1), hi = p.msg_max;
        int sign = synbit, min1 = 1 << 30, min2 = 1 << 30;
        for (int j = 0; j < dc; ++j) {
            const QcEdgeAux ax = aux[e0 + j];
            int l = i + ax.shift; if (l >= Z) l -= Z;
            const int t = clipi(L[ax.col * Z + l] - R[(e0 + j) * Z + i], lo, hi);
            const int a = abs(t);
            sign ^= (t < 0);
            min2 = min(min2, max(a, min1));
            min1 = min(min1, a);
        }
        min2 = min(min2, p.msg_max + 1);   // degree-1 row
        int c1, c2;
        if (p.rule == QLDPC_RULE_OMS) { c1 = max(min2 - p.offset_int, 0); c2 = max(min1 - p.offset_int, 0); }
        else { c1 = norm8(min2, p.norm_eighths); c2 = norm8(min1, p.norm_eighths); }
        for (int j = 0; j < dc; ++j) {
            const QcEdgeAux ax = aux[e0 + j];
            int l = i + ax.shift; if (l >= Z) l -= Z;
            const int contrib = L[ax.col * Z + l] - R[(e0 + j) * Z + i];
            const int t = clipi(contrib, lo, hi);
            const int mag = (abs(t) == min1) ? c1 : c2;
            const int out = (sign ^ (t < 0)) ? -mag : mag;
            R[(e0 + j) * Z + i] = out;
            L[ax.col * Z + l] = clipi(contrib + out, -(p.app_max + 1), p.app_max);
        }
    }
}

template <typename T, typename IN>
__global__ void __launch_bounds__(1024, 1) layered_generic_kernel(const LayeredGenParams p)
{
    const int tid = threadIdx.x, nt = blockDim.x, Z = p.Z;
    T *L = reinterpret_cast<T *>(p.app) + (size_t)blockIdx.x * p.N;
    T *R = reinterpret_cast<T *>(p.msg) + (size_t)blockIdx.x * p.nnz * Z;
    constexpr bool kFloat = std::is_floating_point<T>::value;

    for (int f = blockIdx.x; f < p.F; f += gridDim.x) {
        const IN *llr = reinterpret_cast<const IN *>(p.llr) + (size_t)f * p.N;
        const uint32_t *syn = p.syn ? p.syn + (size_t)f * p.syn_words : nullptr;
        for (int v = tid; v < p.N; v += nt) L[v] = (T)llr[v];
        for (int e = tid; e < p.nnz * Z; e += nt) R[e] = (T)0;
        __syncthreads();

        int it = 0, depth = 0;
        bool ok = false, checked = false;
        while (it < p.max_iter) {
            for (int r = 0; r < p.brows; ++r) {
                const QcLayer ly = p.layers[r];
                for (int i = tid; i < Z; i += nt)
                    layer_lane<T>(p, L, R, p.aux, ly.edge_begin, ly.degree, i, syn_bit(syn, r * Z + i));
                __syncthreads();
            }
            ++it;
            checked = false;
            // int tiers: check after every iteration (oracle ora_decode_layered_fixed);
            // float tiers: AFF3CT skips the check after the last iteration and honours syndrome_depth
            const bool want = p.early_stop && (kFloat ? it != p.max_iter : true);
            if (want) {
                int bad = 0;
                for (int m = tid; m < p.M; m += nt) {
                    const int r = m / Z, i = m - r * Z;
                    const QcLayer ly = p.layers[r];
                    unsigned s = (unsigned)syn_bit(syn, m);
                    for (int j = 0; j < ly.degree; ++j) {
                        const QcEdgeAux ax = p.aux[ly.edge_begin + j];
                        int l = i + ax.shift; if (l >= Z) l -= Z;
                        s ^= (unsigned)(L[ax.col * Z + l] < (T)0);
                    }
                    bad |= (int)(s & 1u);
                }
                ok = __syncthreads_or(bad) == 0;
                checked = true;
                if (ok) { if (!kFloat || ++depth >= p.syndrome_depth) break; }
                else depth = 0;
            }
        }
        if (!checked) {
            int bad = 0;
            for (int m = tid; m < p.M; m += nt) {
                const int r = m / Z, i = m - r * Z;
                const QcLayer ly = p.layers[r];
                unsigned s = (unsigned)syn_bit(syn, m);
                for (int j = 0; j < ly.degree; ++j) {
                    const QcEdgeAux ax = p.aux[ly.edge_begin + j];
                    int l = i + ax.shift; if (l >= Z) l -= Z;
                    s ^= (unsigned)(L[ax.col * Z + l] < (T)0);
                }
                bad |= (int)(s & 1u);
            }
            ok = __syncthreads_or(bad) == 0;
        }

        uint32_t *ab = p.allbits + (size_t)f * p.cw_words;
        for (int w = tid; w < p.cw_words; w += nt) {
            uint32_t v = 0;
            for (int b = 0; b < 32; ++b) {
                const int idx = 32 * w + b;
                if (idx < p.N && L[idx] < (T)0) v |= 1u << (31 - b);
            }
            ab[w] = v;
        }
        if (p.posterior) {
            if constexpr (kFloat) {
                float *po = reinterpret_cast<float *>(p.posterior) + (size_t)f * p.N;
                for (int v = tid; v < p.N; v += nt) po[v] = L[v];
            } else {
                int *po = reinterpret_cast<int *>(p.posterior) + (size_t)f * p.N;
                for (int v = tid; v < p.N; v += nt) po[v] = L[v];
            }
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

int launch_layered_generic(const LayeredGenParams &p, int grid, cudaStream_t st)
{
    int block = ((p.Z + 31) / 32) * 32;
    if (block > 1024) block = 1024;
    if (block < 64) block = 64;
    switch (p.dtype) {
    case QLDPC_DTYPE_F32: layered_generic_kernel<float, float><<<grid, block, 0, st>>>(p); break;
    case QLDPC_DTYPE_I16: layered_generic_kernel<int, int16_t><<<grid, block, 0, st>>>(p); break;
    case QLDPC_DTYPE_I8: layered_generic_kernel<int, int8_t><<<grid, block, 0, st>>>(p); break;
    default: return QLDPC_ERR_UNSUPPORTED;
    }
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

}  // namespace qldpc

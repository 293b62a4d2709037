// Flooding belief propagation for an arbitrary parity-check matrix (CSR + CSC edge tables).
//
// Arithmetic = AFF3CT Decoder_LDPC_BP_flooding<B,Q,Update_rule_{SPA,NMS,OMS}> as the reference
// instantiates it (BOOT/src/main.cpp:193, decode_siho at :365), restated in
// oracle/qldpc_oracle.c:ora_decode_flooding_f32 / ora_decode_flooding_fixed:
//   per sweep: post[v] = Y[v] + sum_e c2v[e]   (edges of v in ascending check order)
//              v2c[e]  = post[v] - c2v[e]       (recomputed on the fly, never stored)
//              c2v[e]  = rule(v2c of the check, syndrome bit folded into the sign)
//   early stop (enable_syndrome, syndrome_depth) after every sweep but the last.
// One CTA per frame in flight: thread-per-variable then thread-per-check phases; messages and
// posteriors live in shared memory when (E+N)*4 bytes fit, else in a per-CTA global scratch that
// stays L2-resident.  Float sums / products run in the oracle's order with FMA contraction off,
// and tanh / atanh are evaluated in double and rounded once, so results match the oracle bitwise.
#include <type_traits>

#include "kernels.hpp"
#include "spa_math.cuh"

namespace qldpc {

namespace {

__device__ __forceinline__ float sgn_apply(float mag, int s) { return s ? -mag : mag; }

template <typename T> struct Acc;
template <> struct Acc<float> { typedef float post_t; };
template <> struct Acc<int> { typedef int post_t; };

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

// ---- float check-node update (one thread owns the whole check)
// tcache: when not null, kTanhCache slots of shared memory per thread ([slot][thread]): the tanh of the first pass is kept
// for the second one (the double-precision tanh is the most expensive operation of the SPA update); the sign travels with it.
constexpr int kTanhCache = 8;
__device__ __forceinline__ void check_update_f32(const FloodParams &p, const float *post, float *c2v, int e0, int e1,
                                                 int synbit, float *tcache, int tstride)
{
    int sign = synbit;
    // exact: tanh / atanh in double, rounded once (bit-exact with the oracle); QLDPC_FLAG_FAST_SPA: fp32 on the SFUs
    auto th = [&](float a) { return p.fast_spa ? tanh_half_fast(a) : tanh_half_exact(a); };
    auto ath2 = [&](float r) { return p.fast_spa ? two_atanh_fast(r) : two_atanh_exact(r); };
    if (p.rule == QLDPC_RULE_SPA) {
        float product = 1.0f;
        for (int e = e0; e < e1; ++e) {
            const float x = post[p.col_idx[e]] - c2v[e];
            const float t = th(fabsf(x));
            if (tcache && e - e0 < kTanhCache) tcache[(e - e0) * tstride] = signbit(x) ? -t : t;   // t >= 0; -0.0f keeps the sign
            product *= (t != 0.0f) ? t : 1e-12f;
            sign ^= signbit(x) ? 1 : 0;
        }
        for (int e = e0; e < e1; ++e) {
            float t;
            int sx;
            if (tcache && e - e0 < kTanhCache) {
                const float ts = tcache[(e - e0) * tstride];
                t = fabsf(ts);
                sx = signbit(ts) ? 1 : 0;
            } else {
                const float x = post[p.col_idx[e]] - c2v[e];
                t = th(fabsf(x));
                sx = signbit(x) ? 1 : 0;
            }
            float r = product / ((t != 0.0f) ? t : 1e-12f);
            r = (r < 1.0f) ? r : 1.0f - 1.1920929e-07f;
            const float mag = ath2(r);
            c2v[e] = sgn_apply(mag, sign ^ sx);
        }
    } else {
        float min1 = 3.402823466e+38f, min2 = 3.402823466e+38f;
        for (int e = e0; e < e1; ++e) {
            const float x = post[p.col_idx[e]] - c2v[e];
            const float a = fabsf(x);
            sign ^= signbit(x) ? 1 : 0;
            min2 = fminf(min2, fmaxf(a, min1));
            min1 = fminf(min1, a);
        }
        float cst1, cst2;
        if (p.rule == QLDPC_RULE_NMS) {
            cst1 = min2 * p.norm;
            cst2 = min1 * p.norm;
        } else {
            cst1 = fmaxf(0.0f, min2 - p.offset);
            cst2 = fmaxf(0.0f, min1 - p.offset);
        }
        for (int e = e0; e < e1; ++e) {
            const float x = post[p.col_idx[e]] - c2v[e];
            const float mag = (fabsf(x) == min1) ? cst1 : cst2;
            c2v[e] = sgn_apply(mag, sign ^ (signbit(x) ? 1 : 0));
        }
    }
}

__device__ __forceinline__ int clipi(int x, int lo, int hi) { return min(max(x, lo), hi); }

__device__ __forceinline__ void check_update_int(const FloodParams &p, const int *post, int *c2v, int e0, int e1,
                                                 int synbit)
{
    int sign = synbit, min1 = p.vmax, min2 = p.vmax;
    for (int e = e0; e < e1; ++e) {
        const int x = clipi(post[p.col_idx[e]] - c2v[e], -p.vmax, p.vmax);
        const int a = abs(x);
        sign ^= (x < 0);
        min2 = min(min2, max(a, min1));
        min1 = min(min1, a);
    }
    int cst1, cst2;
    if (p.rule == QLDPC_RULE_OMS) {
        cst1 = max(min2 - p.offset_int, 0);
        cst2 = max(min1 - p.offset_int, 0);
    } else {
        cst1 = norm8(min2, p.norm_eighths);
        cst2 = norm8(min1, p.norm_eighths);
    }
    for (int e = e0; e < e1; ++e) {
        const int x = clipi(post[p.col_idx[e]] - c2v[e], -p.vmax, p.vmax);
        const int mag = (abs(x) == min1) ? cst1 : cst2;
        c2v[e] = (sign ^ (x < 0)) ? -mag : mag;
    }
}

template <typename T, typename IN>
__global__ void __launch_bounds__(1024, 1) flooding_kernel(const FloodParams p)
{
    extern __shared__ __align__(16) char smem[];
    T *c2v, *post;
    if (p.use_smem) {
        c2v = reinterpret_cast<T *>(smem);
        post = c2v + p.E;
    } else {
        c2v = reinterpret_cast<T *>(p.c2v) + (size_t)blockIdx.x * p.E;
        post = reinterpret_cast<T *>(p.post) + (size_t)blockIdx.x * p.N;
    }
    const int tid = threadIdx.x, nt = blockDim.x;
    // SPA with the messages in global scratch: the (otherwise unused) shared memory caches tanh values between the passes
    float *tcache = (!p.use_smem && p.tanh_cache) ? reinterpret_cast<float *>(smem) + tid : nullptr;

    for (int f = blockIdx.x; f < p.F; f += gridDim.x) {
        const IN *llr = reinterpret_cast<const IN *>(p.llr) + (size_t)f * p.N;
        const uint32_t *syn = p.syn ? p.syn + (size_t)f * p.syn_words : nullptr;
        for (int e = tid; e < p.E; e += nt) c2v[e] = (T)0;     // decoder.reset(), BOOT/src/main.cpp:389
        __syncthreads();

        int it = 0, depth = 0;
        bool ok = false;
        for (;;) {
            // variable phase: a-posteriori values (also next sweep's var-to-check base)
            for (int v = tid; v < p.N; v += nt) {
                T sum = (T)0;
                for (int k = p.var_ptr[v]; k < p.var_ptr[v + 1]; ++k) sum += c2v[p.var_edge[k]];
                post[v] = (T)llr[v] + sum;
            }
            __syncthreads();
            const bool last = it >= p.max_iter;
            const bool want_check = last || (p.early_stop && it > 0);
            if (want_check) {
                int bad = 0;
                for (int m = tid; m < p.M; m += nt) {
                    unsigned s = syn ? (syn[m >> 5] >> (31 - (m & 31))) & 1u : 0u;
                    for (int e = p.row_ptr[m]; e < p.row_ptr[m + 1]; ++e) s ^= (unsigned)(post[p.col_idx[e]] < (T)0);
                    bad |= (int)(s & 1u);
                }
                ok = __syncthreads_or(bad) == 0;
                if (last) break;
                if (ok) { if (++depth >= p.syndrome_depth) break; }
                else depth = 0;
            }
            // check phase
            for (int m = tid; m < p.M; m += nt) {
                const int synbit = syn ? (int)((syn[m >> 5] >> (31 - (m & 31))) & 1u) : 0;
                if constexpr (std::is_floating_point<T>::value) check_update_f32(p, (const float *)post, (float *)c2v, p.row_ptr[m], p.row_ptr[m + 1], synbit, tcache, nt);
                else check_update_int(p, (const int *)post, (int *)c2v, p.row_ptr[m], p.row_ptr[m + 1], synbit);
            }
            __syncthreads();
            ++it;
        }

        // outputs
        uint32_t *ab = p.allbits + (size_t)f * p.cw_words;
        // one variable per thread, one ballot per warp and output word (was: 32 serial loads and compares per word on the
        // first cw_words threads)
        for (int v0 = 0; v0 < p.cw_words * 32; v0 += nt) {
            const int v = v0 + tid;
            const unsigned b = __ballot_sync(0xffffffffu, v < p.N && post[v] < (T)0);
            if ((tid & 31) == 0 && (v >> 5) < p.cw_words) ab[v >> 5] = __brev(b);   // bit 31 - k of word w = variable 32 w + k
        }
        if (p.posterior) {
            typename Acc<T>::post_t *po = reinterpret_cast<typename Acc<T>::post_t *>(p.posterior) + (size_t)f * p.N;
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

int launch_flooding(const FloodParams &p, int grid, int block, int smem_bytes, cudaStream_t st)
{
#define QLDPC_FLOOD_LAUNCH(T, IN)                                                                              \
    do {                                                                                                       \
        QLDPC_CUDA(cudaFuncSetAttribute(flooding_kernel<T, IN>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                        smem_bytes));                                                          \
        flooding_kernel<T, IN><<<grid, block, smem_bytes, st>>>(p);                                            \
    } while (0)
    switch (p.dtype) {
    case QLDPC_DTYPE_F32: QLDPC_FLOOD_LAUNCH(float, float); break;
    case QLDPC_DTYPE_I16: QLDPC_FLOOD_LAUNCH(int, int16_t); break;
    case QLDPC_DTYPE_I8: QLDPC_FLOOD_LAUNCH(int, int8_t); break;
    default: return QLDPC_ERR_UNSUPPORTED;
    }
#undef QLDPC_FLOOD_LAUNCH
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

}  // namespace qldpc

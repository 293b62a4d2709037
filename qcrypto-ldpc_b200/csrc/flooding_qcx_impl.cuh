// Flooding belief propagation for LARGE quasi-cyclic codes (the long QKD blocks of BASELINE config 3: N = 65 536, Z = 2 048
// or 1 024): one frame per thread-block CLUSTER, state in an L2-resident scratch, float SPA / min-sum and int8 / int16 min-sum.
//
// Arithmetic: AFF3CT Decoder_LDPC_BP_flooding<B,Q,Update_rule_{SPA,NMS,OMS}> as the reference instantiates it
// (BOOT/src/main.cpp:193, decode_siho at :365; "main.cpp (5g-qc)":236-251), restated in
// oracle/qldpc_oracle.c:ora_decode_flooding_f32 / ora_decode_flooding_fixed -- same sweep structure, same order of every
// sum and product, same early-stop rule as flooding.cu / flooding_qc.cu.  What differs from those kernels:
//   * a frame is worked on by a cluster of CL thread blocks (CL SMs), each owning Z / CL lanes of every circulant, so only
//     SMs / CL frames are in flight and their messages + posteriors (1.05 MB per frame) stay resident in the 126 MB L2:
//     the 16 E bytes per sweep are L2 traffic, HBM sees the channel LLRs.  The phases are separated by the
//     hardware cluster barrier; the early-termination vote crosses the cluster through distributed shared memory;
//   * the Z lanes of a circulant map onto warp lanes: a work item is a chunk of 32 V consecutive lanes (V = 4 for row
//     degrees <= 8, else 2) of one block row, worked on by one warp; thread `lane` has lanes chunk + lane + 32 k, so every
//     load and store of the warp is 32 consecutive values (one or two 128-byte lines) whatever the cyclic shift, and the
//     table look-up, the shift and the address arithmetic of an edge are paid once per V lanes.  A shifted chunk wraps
//     past the end of the circulant for one chunk per edge only (a warp-uniform branch);
//   * the circulant tables sit in shared memory; a work item issues all its loads before the first use (2 V dc values
//     in flight per thread); the variable phase works on four block columns side by side for the same reason;
//   * the check update is compiled per row degree (no predicated slots), rule and SPA flavour are template parameters;
//     the first ncu capture of the generic version (profiles/r2_flooding_qcx_v1_ncu_summary.txt) showed the kernel bound
//     by instruction issue, not by memory: 220 thread instructions per edge and sweep;
//   * integer tiers keep their messages in 8 / 16 bits (posteriors in 16 / 32), not in 32-bit words;
//   * sweep 0 reads no messages at all (they are zero), so the scratch is never cleared.
#include <cooperative_groups.h>

#include <type_traits>

#pragma once
#include "kernels.hpp"
#include "spa_math.cuh"

namespace cg = cooperative_groups;

namespace qldpc {

namespace {

constexpr int kThreads = 512;
constexpr int kMaxDcV4 = 8;          // rows up to this degree: 4 lanes per thread (codes whose heaviest row is heavier: 2)
constexpr int kMaxDcV2 = 20;         // compiled row degrees; heavier rows take the two-pass loop

// ---- SPA kernels of the check update --------------------------------------------------------------------------------
// Two flavours (template parameter FLAVOUR, QLDPC_FLAG_FAST_SPA):
//   exact (default): tanh / atanh in double, rounded once -- what the oracle and the other float kernels do.  The float
//     product / t_j of a SATURATED check leaves 1 - r as a small multiple of 2^-24, so a last-bit difference in one tanh
//     moves 2 atanh(r) by ln 2, ln 3/2, ...: only a correctly rounded tanh keeps such messages (|m| > 14, posteriors > 30)
//     within 1e-3 of the reference.
//   fast: fp32 on the special-function units.  Decoded bits and iteration counts equal the exact flavour's on every test
//     batch; 99.99 % of the posteriors are within 1e-3, the rest (saturated messages) within ln 2.
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

struct RowMeta { int edge_begin, degree; };
// the update rule's parameters, by value (a reference to the kernel parameters handed to a non-inlined function would
// make the compiler copy them to local memory)
struct Upd { int rule, offset_int, norm_eighths, vmax; float norm, offset; };
// decoder flavour of a kernel instantiation
enum { kSpaExact = 0, kSpaFast = 1, kMinSum = 2 };

__device__ __forceinline__ unsigned fbits(float x) { return __float_as_uint(x); }

// ---- one work item of the check phase: block row starting at edge e0 with exactly DC edges, for a chunk of 32 V check
// lanes starting at cb (a multiple of 32 V): this thread works on lanes cb + lane + 32 k, k < V, so that every load and
// store of the warp is 32 consecutive values.  Reads the posteriors of the variables and the old messages, writes the new
// messages.  A cyclic shift turns the chunk into a run of 32 V variable lanes that starts anywhere; the run wraps past the
// end of the circulant for one chunk per edge only, a warp-uniform case.  FIRST: sweep 0, the old messages are zero and
// not read.  synw[k]: syndrome word of lanes cb + 32 k .. + 31 (bit 31 = first lane).  Returns the OR of the
// hard-decision parities (early-stop test).  Offsets are 32-bit (nnz * Z and N are far below 2^31).
template <typename MsgT, typename PostT, int FLAVOUR, bool FIRST, int DC, int V>
__device__ __forceinline__ int check_item(const Upd p, int e0, const int2 *edges, const PostT *post, MsgT *c2v, int cb, int lane,
                                          int Z, const unsigned (&synw)[V])
{
    constexpr bool kFloat = sizeof(PostT) == 4 && sizeof(MsgT) == 4;
    typedef typename std::conditional<kFloat, float, int>::type XT;
    XT x[DC][V];                            // variable-to-check messages: posterior - old message
    unsigned hs[V];                         // XOR of the posteriors' sign bits (hard-decision parity)
#pragma unroll
    for (int k = 0; k < V; ++k) hs[k] = 0;
    MsgT *cm = c2v + (e0 * Z + cb + lane);
#pragma unroll
    for (int j = 0; j < DC; ++j) {          // the loads are independent of everything else: up to 2 V DC requests in flight
        const int2 e = edges[e0 + j];       // (block column * Z, shift)
        int vb = cb + e.y;                  // variable lane of the chunk's first check lane
        if (vb >= Z) vb -= Z;
        const PostT *pp = post + (e.x + vb + lane);
        PostT pv[V];
        if (vb + 32 * V <= Z) {             // warp-uniform: the run does not wrap
#pragma unroll
            for (int k = 0; k < V; ++k) pv[k] = pp[32 * k];
        } else {
#pragma unroll
            for (int k = 0; k < V; ++k) pv[k] = pp[32 * k - ((vb + lane + 32 * k >= Z) ? Z : 0)];
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
            // the variable phase never writes -0 (y + (+0 + m ...)): the sign bit of a float posterior is its hard decision
            if constexpr (kFloat) hs[k] ^= fbits((float)pv[k]);
            else hs[k] ^= (unsigned)(int)pv[k];
            const XT o = FIRST ? (XT)0 : (XT)cm[j * Z + 32 * k];
            x[j][k] = (XT)pv[k] - o;
        }
    }
    int bad = 0;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int synbit = (int)((synw[i] >> (31 - lane)) & 1u);
        bad |= synbit ^ (int)(hs[i] >> 31);
        if constexpr (kFloat) {
            unsigned sign = (unsigned)synbit << 31;
            if constexpr (FLAVOUR != kMinSum) {
                float ts[DC];               // tanh(|x| / 2) with the sign of x
                float product = 1.0f;
#pragma unroll
                for (int j = 0; j < DC; ++j) {
                    const float xv = x[j][i];
                    const float tj = FLAVOUR == kSpaFast ? tanh_half_fast(fabsf(xv)) : tanh_half_exact(fabsf(xv));
                    const float t = (tj != 0.0f) ? tj : 1e-12f;
                    product *= t;
                    sign ^= fbits(xv);
                    ts[j] = __uint_as_float(fbits(t) | (fbits(xv) & 0x80000000u));
                }
#pragma unroll
                for (int j = 0; j < DC; ++j) {
                    float rr = product / fabsf(ts[j]);               // IEEE division, as the oracle
                    rr = (rr < 1.0f) ? rr : 1.0f - 1.1920929e-07f;
                    const float mag = FLAVOUR == kSpaFast ? two_atanh_fast(rr) : two_atanh_exact(rr);
                    cm[j * Z + 32 * i] = (MsgT)__uint_as_float(fbits(mag) ^ ((sign ^ fbits(ts[j])) & 0x80000000u));   // mag >= 0
                }
            } else {
                float min1 = 3.402823466e+38f, min2 = 3.402823466e+38f;
#pragma unroll
                for (int j = 0; j < DC; ++j) {
                    const float a = fabsf(x[j][i]);
                    sign ^= fbits(x[j][i]);
                    min2 = fminf(min2, fmaxf(a, min1));
                    min1 = fminf(min1, a);
                }
                const float cst1 = p.rule == QLDPC_RULE_NMS ? min2 * p.norm : fmaxf(0.0f, min2 - p.offset);
                const float cst2 = p.rule == QLDPC_RULE_NMS ? min1 * p.norm : fmaxf(0.0f, min1 - p.offset);
#pragma unroll
                for (int j = 0; j < DC; ++j) {
                    const float mag = (fabsf(x[j][i]) == min1) ? cst1 : cst2;
                    cm[j * Z + 32 * i] = (MsgT)__uint_as_float(fbits(mag) ^ ((sign ^ fbits(x[j][i])) & 0x80000000u));
                }
            }
        } else {
            int sign = synbit, min1 = p.vmax, min2 = p.vmax;
#pragma unroll
            for (int j = 0; j < DC; ++j) {
                x[j][i] = min(max(x[j][i], -p.vmax), p.vmax);
                const int a = abs(x[j][i]);
                sign ^= x[j][i] < 0;
                min2 = min(min2, max(a, min1));
                min1 = min(min1, a);
            }
            const int cst1 = p.rule == QLDPC_RULE_OMS ? max(min2 - p.offset_int, 0) : norm8(min2, p.norm_eighths);
            const int cst2 = p.rule == QLDPC_RULE_OMS ? max(min1 - p.offset_int, 0) : norm8(min1, p.norm_eighths);
#pragma unroll
            for (int j = 0; j < DC; ++j) {
                const int mag = (abs(x[j][i]) == min1) ? cst1 : cst2;
                cm[j * Z + 32 * i] = (MsgT)((sign ^ (x[j][i] < 0)) ? -mag : mag);
            }
        }
    }
    return bad;
}

// any degree: two passes over memory (rows heavier than the compiled degrees)
template <typename MsgT, typename PostT, int FLAVOUR, bool FIRST, int V>
__device__ __noinline__ int check_item_any(const Upd p, int e0, int deg, const int2 *edges, const PostT *post, MsgT *c2v, int cb,
                                           int lane, int Z, const unsigned (&synw)[V])
{
    constexpr bool kFloat = sizeof(PostT) == 4 && sizeof(MsgT) == 4;
    int bad = 0;
    for (int i = 0; i < V; ++i) {
        const int l = cb + lane + 32 * i, synbit = (int)((synw[i] >> (31 - lane)) & 1u);
        MsgT *cm = c2v + (e0 * Z + l);
        int sign = synbit, hard = synbit;
        auto v2c = [&](int j, bool count_hard) {
            const int2 e = edges[e0 + j];
            int vl = l + e.y;
            if (vl >= Z) vl -= Z;
            const PostT pvj = post[e.x + vl];
            if (count_hard) hard ^= pvj < (PostT)0;
            const MsgT o = FIRST ? (MsgT)0 : cm[j * Z];
            if constexpr (kFloat) return (float)pvj - (float)o;
            else return (float)min(max((int)pvj - (int)o, -p.vmax), p.vmax);   // integers up to 2^24 are exact in a float
        };
        auto th = [&](float a) { return FLAVOUR == kSpaFast ? tanh_half_fast(a) : tanh_half_exact(a); };
        if (kFloat && FLAVOUR != kMinSum) {
            float product = 1.0f;
            for (int j = 0; j < deg; ++j) {
                const float xv = v2c(j, true);
                const float tj = th(fabsf(xv));
                product *= (tj != 0.0f) ? tj : 1e-12f;
                sign ^= signbit(xv) ? 1 : 0;
            }
            bad |= hard;
            for (int j = 0; j < deg; ++j) {
                const float xv = v2c(j, false);
                const float tj = th(fabsf(xv));
                float rr = product / ((tj != 0.0f) ? tj : 1e-12f);
                rr = (rr < 1.0f) ? rr : 1.0f - 1.1920929e-07f;
                const float mag = FLAVOUR == kSpaFast ? two_atanh_fast(rr) : two_atanh_exact(rr);
                cm[j * Z] = (MsgT)((sign ^ (signbit(xv) ? 1 : 0)) ? -mag : mag);
            }
            continue;
        }
        float min1 = kFloat ? 3.402823466e+38f : (float)p.vmax, min2 = min1;
        for (int j = 0; j < deg; ++j) {
            const float xv = v2c(j, true);
            const float a = fabsf(xv);
            sign ^= signbit(xv) ? 1 : 0;
            min2 = fminf(min2, fmaxf(a, min1));
            min1 = fminf(min1, a);
        }
        bad |= hard;
        float cst1, cst2;
        if constexpr (kFloat) {
            cst1 = p.rule == QLDPC_RULE_NMS ? min2 * p.norm : fmaxf(0.0f, min2 - p.offset);
            cst2 = p.rule == QLDPC_RULE_NMS ? min1 * p.norm : fmaxf(0.0f, min1 - p.offset);
        } else {
            cst1 = (float)(p.rule == QLDPC_RULE_OMS ? max((int)min2 - p.offset_int, 0) : norm8((int)min2, p.norm_eighths));
            cst2 = (float)(p.rule == QLDPC_RULE_OMS ? max((int)min1 - p.offset_int, 0) : norm8((int)min1, p.norm_eighths));
        }
        for (int j = 0; j < deg; ++j) {
            const float xv = v2c(j, false);
            const float mag = (fabsf(xv) == min1) ? cst1 : cst2;
            const float o = (sign ^ (signbit(xv) ? 1 : 0)) ? -mag : mag;
            if constexpr (kFloat) cm[j * Z] = o;
            else cm[j * Z] = (MsgT)(int)o;
        }
    }
    return bad;
}

// the row's degree is the same for all threads of a warp (a work item belongs to one warp)
template <typename MsgT, typename PostT, int FLAVOUR, bool FIRST, int V>
__device__ __forceinline__ int check_dispatch(const Upd p, const RowMeta ly, const int2 *edges, const PostT *post, MsgT *c2v,
                                              int cb, int lane, int Z, const unsigned (&synw)[V])
{
#define QL_DC(D) case D: return check_item<MsgT, PostT, FLAVOUR, FIRST, D, V>(p, ly.edge_begin, edges, post, c2v, cb, lane, Z, synw);
    if constexpr (V == 4) {
        switch (ly.degree) {
            QL_DC(2) QL_DC(3) QL_DC(4) QL_DC(5) QL_DC(6) QL_DC(7) QL_DC(8)
        default: break;
        }
    } else {
        switch (ly.degree) {
            QL_DC(2) QL_DC(3) QL_DC(4) QL_DC(5) QL_DC(6) QL_DC(7) QL_DC(8) QL_DC(9) QL_DC(10)
            QL_DC(11) QL_DC(12) QL_DC(13) QL_DC(14) QL_DC(15) QL_DC(16) QL_DC(17) QL_DC(18) QL_DC(19) QL_DC(20)
        default: break;
        }
    }
#undef QL_DC
    return check_item_any<MsgT, PostT, FLAVOUR, FIRST, V>(p, ly.edge_begin, ly.degree, edges, post, c2v, cb, lane, Z, synw);
}

// Shared-memory layout: RowMeta rows[R]; int2 edges[nnz] (block column * Z, shift); int col_ptr[C + 2]; int2 col_edges[nnz]
// (edge id * Z, shift); int vote[2][8].
// Scratch layout (per cluster): messages c2v[edge][Z] indexed by CHECK lane, posteriors post[block column][Z].
// A work item is a chunk of 32 V consecutive lanes of one block row (check phase) or of kCols block columns (variable
// phase) and belongs to one warp; thread `lane` of the warp works on lanes chunk + lane + 32 k, k < V.
template <typename MsgT, typename PostT, typename InT, int FLAVOUR, int V>
__global__ void __launch_bounds__(kThreads, 1) flooding_qcx_kernel(const FloodQcxParams p)
{
    extern __shared__ __align__(16) char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks(), q = (int)cluster.block_rank();
    const int cid = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const int tid = threadIdx.x, Z = p.Z, ZL = Z / CL, lane0 = q * ZL;
    const int R = p.brows, C = p.bcols;
    constexpr int kWarps = kThreads / 32, kChunk = 32 * V;
    const int lane = tid & 31, wid = (int)__shfl_sync(0xffffffffu, tid >> 5, 0);   // warp index, on the uniform datapath
    const int NCH = ZL / kChunk;                                              // chunks per block row / column in this block
    const int step_r = kWarps / NCH, step_c = kWarps - step_r * NCH;          // a warp's next work item: kWarps items further

    RowMeta *rows = reinterpret_cast<RowMeta *>(smem);
    int2 *edges = reinterpret_cast<int2 *>(rows + R);
    int *col_ptr = reinterpret_cast<int *>(edges + p.nnz);
    int2 *col_edges = reinterpret_cast<int2 *>(col_ptr + C + 1 + ((C + 1) & 1));   // 8-byte aligned
    int *vote = reinterpret_cast<int *>(col_edges + p.nnz);
    for (int r = tid; r < R; r += kThreads) rows[r] = RowMeta{p.layers[r].edge_begin, p.layers[r].degree};
    for (int e = tid; e < p.nnz; e += kThreads) {
        edges[e] = make_int2(p.aux[e].col * Z, p.aux[e].shift);
        col_edges[e] = make_int2(p.col_edges[e].x * Z, p.col_edges[e].y);
    }
    for (int c = tid; c <= C; c += kThreads) col_ptr[c] = p.col_ptr[c];
    if (tid < 16) vote[tid] = 0;
    cluster.sync();

    MsgT *c2v = reinterpret_cast<MsgT *>(p.c2v) + (size_t)cid * p.nnz * Z;
    PostT *post = reinterpret_cast<PostT *>(p.post) + (size_t)cid * p.N;
    unsigned vpar = 0;
    const Upd upd{p.rule, p.offset_int, p.norm_eighths, p.vmax, p.norm, p.offset};

    // cluster-wide OR of a per-thread flag: block vote, then every block writes its result into every block's table
    auto cluster_any = [&](int flag) {
        const int mine = __syncthreads_or(flag);
        if (CL == 1) return mine != 0;
        if (tid < CL) *cluster.map_shared_rank(&vote[vpar * 8 + q], tid) = mine;
        cluster.sync();
        int any = 0;
        for (int k = 0; k < CL; ++k) any |= vote[vpar * 8 + k];
        vpar ^= 1u;
        return any != 0;
    };

    for (int f = cid; f < p.F; f += n_clusters) {
        const InT *llr = reinterpret_cast<const InT *>(p.llr) + (size_t)f * p.N;
        const uint32_t *syn = p.syn ? p.syn + (size_t)f * p.syn_words : nullptr;
        int it = 0, depth = 0;
        bool ok = false;
        for (;;) {
            // ---- variable phase: post[v] = llr[v] + sum of the column's messages in ascending block-row order.
            // kCols block columns side by side: a column of weight 3 alone keeps only 3 V loads in flight per thread and
            // the phase waits on L2 latency.
            {
                constexpr int kCols = 4, kDv = V == 4 ? 3 : 4;   // columns per item, edges of each column in flight at a time
                const int n_groups = (C + kCols - 1) / kCols;
                int cgp = wid / NCH, ch = wid - cgp * NCH;
                for (; cgp < n_groups; ch += step_c, cgp += step_r + (ch >= NCH ? 1 : 0), ch -= (ch >= NCH ? NCH : 0)) {
                    const int mb = lane0 + ch * kChunk;          // first variable lane of the chunk
                    InT y[kCols][V];
                    PostT sum[kCols][V];
                    int ka[kCols], kn[kCols], kmax = 0;
#pragma unroll
                    for (int u = 0; u < kCols; ++u) {
                        const int c = cgp * kCols + u;
                        const bool valid = c < C;
                        ka[u] = valid ? col_ptr[c] : 0;
                        kn[u] = (valid && it > 0) ? col_ptr[c + 1] - ka[u] : 0;
                        kmax = max(kmax, kn[u]);
#pragma unroll
                        for (int i = 0; i < V; ++i) {
                            y[u][i] = valid ? llr[c * Z + mb + lane + 32 * i] : (InT)0;
                            sum[u][i] = (PostT)0;
                        }
                    }
                    for (int k = 0; k < kmax; k += kDv) {
                        MsgT m[kCols][kDv][V];
#pragma unroll
                        for (int u = 0; u < kCols; ++u) {
#pragma unroll
                            for (int d = 0; d < kDv; ++d) {
                                if (k + d < kn[u]) {
                                    const int2 ce = col_edges[ka[u] + k + d];            // edge id * Z, shift
                                    int lb = mb - ce.y;                                  // check lane of the chunk's first variable lane
                                    if (lb < 0) lb += Z;
                                    const MsgT *mp = c2v + (ce.x + lb + lane);
                                    if (lb + kChunk <= Z) {
#pragma unroll
                                        for (int i = 0; i < V; ++i) m[u][d][i] = mp[32 * i];
                                    } else {
#pragma unroll
                                        for (int i = 0; i < V; ++i) m[u][d][i] = mp[32 * i - ((lb + lane + 32 * i >= Z) ? Z : 0)];
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kCols; ++u) {
#pragma unroll
                            for (int d = 0; d < kDv; ++d) {      // ascending block-row order
                                if (k + d < kn[u]) {
#pragma unroll
                                    for (int i = 0; i < V; ++i) sum[u][i] += (PostT)m[u][d][i];
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kCols; ++u) {
                        const int c = cgp * kCols + u;
                        if (c < C) {
#pragma unroll
                            for (int i = 0; i < V; ++i) post[c * Z + mb + lane + 32 * i] = (PostT)y[u][i] + sum[u][i];
                        }
                    }
                }
            }
            cluster.sync();
            const bool last = it >= p.max_iter;
            if (last) {   // final verdict after the last sweep: syndrome of the hard decisions, no update
                int bad = 0;
                int r = wid / NCH, ch = wid - r * NCH;
                for (; r < R; ch += step_c, r += step_r + (ch >= NCH ? 1 : 0), ch -= (ch >= NCH ? NCH : 0)) {
                    const int cb = lane0 + ch * kChunk;
                    const RowMeta ly = rows[r];
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const int l = cb + lane + 32 * i;
                        unsigned s = syn ? (syn[(r * Z + cb) / 32 + i] >> (31 - lane)) & 1u : 0u;
                        for (int j = 0; j < ly.degree; ++j) {
                            const int2 e = edges[ly.edge_begin + j];
                            int vl = l + e.y;
                            if (vl >= Z) vl -= Z;
                            s ^= (unsigned)(post[e.x + vl] < (PostT)0);
                        }
                        bad |= (int)s;
                    }
                }
                ok = !cluster_any(bad);
                break;
            }
            // ---- check phase; the early-termination test (enable_syndrome) of this sweep is computed on the way.
            // If it passes the decoder stops here: the messages just written are never used, `it` is not advanced.
            const bool want_check = p.early_stop && it > 0;
            int bad = 0;
            int r = wid / NCH, ch = wid - r * NCH;                        // work item: (block row, chunk of 32 V check lanes)
            for (; r < R; ch += step_c, r += step_r + (ch >= NCH ? 1 : 0), ch -= (ch >= NCH ? NCH : 0)) {
                const int cb = lane0 + ch * kChunk;
                unsigned synw[V];
#pragma unroll
                for (int i = 0; i < V; ++i) synw[i] = syn ? syn[(r * Z + cb) / 32 + i] : 0u;
                bad |= it == 0 ? check_dispatch<MsgT, PostT, FLAVOUR, true, V>(upd, rows[r], edges, post, c2v, cb, lane, Z, synw)
                               : check_dispatch<MsgT, PostT, FLAVOUR, false, V>(upd, rows[r], edges, post, c2v, cb, lane, Z, synw);
            }
            if (want_check) {
                ok = !cluster_any(bad);
                if (ok) { if (++depth >= p.syndrome_depth) break; }
                else depth = 0;
            } else {
                cluster.sync();
            }
            ++it;
        }

        // ---- outputs: a warp packs kOutW words (32 consecutive variables each, Z % (32 CL) == 0) at a time
        uint32_t *ab = p.allbits + (size_t)f * p.cw_words;
        {
            constexpr int kOutW = 4;
            const int wpc = ZL >> 5, n_words = C * wpc;
            for (int w0 = wid * kOutW; w0 < n_words; w0 += kWarps * kOutW) {
                PostT pv[kOutW];
                int vv[kOutW];
#pragma unroll
                for (int k = 0; k < kOutW; ++k) {
                    const int w = min(w0 + k, n_words - 1), c = w / wpc, m = lane0 + ((w - c * wpc) << 5) + lane;
                    vv[k] = c * Z + m;
                    pv[k] = post[vv[k]];
                }
#pragma unroll
                for (int k = 0; k < kOutW; ++k) {
                    const unsigned b = __ballot_sync(0xffffffffu, pv[k] < (PostT)0);
                    if (w0 + k < n_words) {
                        if (lane == 0) ab[vv[k] >> 5] = __brev(b);
                        if (p.posterior) {
                            if constexpr (sizeof(PostT) == 4 && sizeof(MsgT) == 4) reinterpret_cast<float *>(p.posterior)[(size_t)f * p.N + vv[k]] = (float)pv[k];
                            else reinterpret_cast<int *>(p.posterior)[(size_t)f * p.N + vv[k]] = (int)pv[k];
                        }
                    }
                }
            }
        }
        if (q == 0 && tid == 0) {
            if (p.ok) p.ok[f] = ok ? 1 : 0;
            if (p.iters) p.iters[f] = (uint16_t)it;
            if (p.stats) {
                atomicAdd(&p.stats->frames, 1ull);
                if (!ok) atomicAdd(&p.stats->failures, 1ull);
                atomicAdd(&p.stats->iter_sum, (unsigned long long)it);
                atomicAdd(&p.stats->hist[min(it, QLDPC_ITER_HIST_BINS - 1)], 1ull);
            }
        }
        cluster.sync();   // the scratch is reused by the cluster's next frame
    }
}

template <typename K>
int launch_k(K kern, const FloodQcxParams &p, int n_clusters, int cl, int smem_bytes, cudaStream_t st)
{
    QLDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(n_clusters * cl));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    QLDPC_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return QLDPC_OK;
}

template <typename K>
int max_clusters_k(K kern, int cl, int smem_bytes)
{
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)cl);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

}  // namespace

}  // namespace qldpc

// One translation unit per (tier, flavour) instantiates the kernel for 4 and 2 lanes per thread and exports its launcher and
// its occupancy query (the per-degree check code of all tiers in one file takes ptxas many minutes).
#define QL_QCX_DEFINE(TAG, MSG, POST, IN, FLAVOUR)                                                                              \
    namespace qldpc {                                                                                                          \
    int flooding_qcx_launch_##TAG(const FloodQcxParams &p, int n_clusters, int cl, int smem_bytes, cudaStream_t st)            \
    {                                                                                                                          \
        if (p.lanes == 4) return launch_k(flooding_qcx_kernel<MSG, POST, IN, FLAVOUR, 4>, p, n_clusters, cl, smem_bytes, st);  \
        return launch_k(flooding_qcx_kernel<MSG, POST, IN, FLAVOUR, 2>, p, n_clusters, cl, smem_bytes, st);                    \
    }                                                                                                                          \
    int flooding_qcx_clusters_##TAG(int lanes, int cl, int smem_bytes)                                                         \
    {                                                                                                                          \
        if (lanes == 4) return max_clusters_k(flooding_qcx_kernel<MSG, POST, IN, FLAVOUR, 4>, cl, smem_bytes);                 \
        return max_clusters_k(flooding_qcx_kernel<MSG, POST, IN, FLAVOUR, 2>, cl, smem_bytes);                                 \
    }                                                                                                                          \
    }

// Generic layered decoder for quasi-cyclic codes: any lifting size, f32 / i16 / i8 tiers, optional posterior output.
// This is the kernel behind Decoder_LDPC_BP_horizontal_layered for everything the packed int8 kernels (layered_i8s.cu,
// layered_i8.cu) do not take: float layered SPA / NMS / OMS, the int16 tier, Z not a multiple of 4 (the 802.11n codes),
// posterior requests.
//
// Integer arithmetic = ML/BPSK_nrldpc_sim_FP.m:35-94 (oracle: ora_decode_layered_fixed):
//   contrib = L - R_old (:51); t = clip(contrib, -(msg_max+1), msg_max) (:53-56);
//   min1/min2/sign product over the row (:59-63); offset or k/8 normalisation (:65-72);
//   R_new = parity*sign*mag (:73-75); L = clip(contrib + R_new, -(app_max+1), app_max) (:88-91).
// Float arithmetic = AFF3CT Decoder_LDPC_BP_horizontal_layered ("main.cpp (5g-qc)":256-270;
//   oracle: ora_decode_layered_f32): contrib = var - branch; branch = rule(contrib); var = contrib + branch.
// A QC block row is one layer: its Z checks touch disjoint variables, so processing them in
// parallel equals the row-serial order of both references.
//
// Placement: one CTA per frame in flight, one thread per circulant lane (the Z lanes of a circulant on consecutive
// threads; a cyclic shift is a rotated, conflict-free run of shared-memory words).
//   shared memory: the row / edge tables and the frame's N beliefs (float, or int16 for the integer tiers);
//   global scratch, L2-resident: the check-to-variable messages R[edge][lane] (float / int16) -- element (e, lane) is
//     only ever touched by thread `lane`, so no synchronisation covers it and every access of a warp is one line;
//   registers: L - R_old of the row in flight (and tanh for SPA), the row code is compiled per row degree.
// A row costs one global round trip (its old messages, all dc loads in flight together) instead of the 2 dc dependent ones
// of the first version (tables, beliefs and messages all behind L2); several CTAs per SM overlap their round trips.
// Measured and dropped: loading a row's old messages one row ahead into a fixed 20-slot register array (predicated slots
// + spills at 80 registers: BG1 i16 8.5 -> 5.3 Gbit/s), and capping the CTAs per SM so that the scratch fits in L2.
#include <type_traits>

#include "kernels.hpp"
#include "spa_math.cuh"

namespace qldpc {

namespace {

constexpr int kMaxDc = 20;           // compiled row degrees; heavier rows take the two-pass loop

struct RowMeta { int edge_begin, degree; };
// the update rule's parameters, by value
struct Upd { int rule, offset_int, norm_eighths, msg_max, app_max; float norm, offset; int fast_spa; };
// the SPA transcendentals (spa_math.cuh): double rounded once, or fp32 on the SFUs with QLDPC_FLAG_FAST_SPA
__device__ __forceinline__ float tanh_half(const Upd &u, float a) { return u.fast_spa ? tanh_half_fast(a) : tanh_half_exact(a); }
__device__ __forceinline__ float two_atanh(const Upd &u, float r) { return u.fast_spa ? two_atanh_fast(r) : two_atanh_exact(r); }

// AFF3CT's integer normalize: k/8 as a sum of floor(v/2), floor(v/4), floor(v/8); selects, no branch table (k is uniform,
// the switch cost ~20 instructions per call, two calls per row)
__device__ __forceinline__ int norm8(int v, int k)
{
    if (k <= 0 || k >= 8) return v;
    return ((k & 4) ? v >> 1 : 0) + ((k & 2) ? v >> 2 : 0) + ((k & 1) ? v >> 3 : 0);
}
__device__ __forceinline__ int clipi(int x, int lo, int hi) { return min(max(x, lo), hi); }

// One check (block row with DC edges starting at `ed`, check lane `lane`): beliefs in shared memory, old messages at
// Rl[j * Z] (not read when `first`), new messages and beliefs written back.
// Belief addressing: the edge table holds, per edge, the BYTE offset of belief (column, lane 0 + shift) from the belief base
// `Lb` and the wrap threshold Z - shift; check lane `lane` reads the belief at  e.x + lane_b - (lane >= e.y ? wrap_b : 0)
// (lane_b = lane * sizeof(LT), wrap_b = Z * sizeof(LT)): compare, select, one 3-input add -- no index scaling, no modulo.
struct Lane { int lane, lane_b, wrap_b; };
__device__ __forceinline__ int bel_off(const int2 e, const Lane &t) { return e.x + t.lane_b - (t.lane >= e.y ? t.wrap_b : 0); }
// GL = false: the beliefs sit in dynamic shared memory and the table offsets are absolute shared-space addresses (the
// window base is folded into the table), so that an access is LDS / STS [register] with no pointer arithmetic;
// GL = true: global scratch at Lb
extern __shared__ __align__(16) char qldpc_lg_smem[];
template <typename LT, bool GL> __device__ __forceinline__ LT &bel(char *Lb, int off)
{
    if constexpr (GL) return *reinterpret_cast<LT *>(Lb + off);
    else return *reinterpret_cast<LT *>(__cvta_shared_to_generic((size_t)(unsigned)off));   // `off` is a shared-space address
}

template <typename LT, typename MT, int DC, bool GL>
__device__ __forceinline__ void row_lane(const Upd u, char *Lb, MT *Rl, const int2 *ed, const Lane t, int Z, int synbit, bool first)
{
    constexpr bool kFloat = std::is_floating_point<LT>::value;
    typedef typename std::conditional<kFloat, float, int>::type XT;
    XT x[DC];
    int idx[DC];
#pragma unroll
    for (int j = 0; j < DC; ++j) {
        idx[j] = bel_off(ed[j], t);
        const XT ro = first ? (XT)0 : (XT)Rl[j * Z];
        x[j] = (XT)bel<LT, GL>(Lb, idx[j]) - ro;
    }
    if constexpr (kFloat) {
        int sign = synbit;
        if (u.rule == QLDPC_RULE_SPA) {
            float t[DC];
            float product = 1.0f;
#pragma unroll
            for (int j = 0; j < DC; ++j) {
                const float tj = tanh_half(u, fabsf(x[j]));
                t[j] = (tj != 0.0f) ? tj : 1e-12f;
                product *= t[j];
                sign ^= signbit(x[j]) ? 1 : 0;
            }
#pragma unroll
            for (int j = 0; j < DC; ++j) {
                float r = product / t[j];
                r = (r < 1.0f) ? r : 1.0f - 1.1920929e-07f;
                const float mag = two_atanh(u, r);
                const float out = (sign ^ (signbit(x[j]) ? 1 : 0)) ? -mag : mag;
                Rl[j * Z] = out;
                bel<LT, GL>(Lb, idx[j]) = x[j] + out;
            }
        } else {
            // min-sum on the raw bits: the sign product is the XOR of the words (bit 31), the new message is the
            // magnitude's bits with that bit set -- one logic instruction each, no shifts, compares or selects
            uint32_t sacc = (uint32_t)synbit << 31;
            float min1 = 3.402823466e+38f, min2 = 3.402823466e+38f;
#pragma unroll
            for (int j = 0; j < DC; ++j) {
                const float a = fabsf(x[j]);
                sacc ^= __float_as_uint(x[j]);
                min2 = fminf(min2, fmaxf(a, min1));
                min1 = fminf(min1, a);
            }
            float cst1 = 0.f, cst2 = 0.f;
            if (u.rule == QLDPC_RULE_NMS) { cst1 = min2 * u.norm; cst2 = min1 * u.norm; }
            else if (u.rule == QLDPC_RULE_OMS) { cst1 = fmaxf(0.0f, min2 - u.offset); cst2 = fmaxf(0.0f, min1 - u.offset); }
            // the row's sign product goes into the two magnitudes once; per edge: select, then flip by the edge's own sign
            const uint32_t c1s = __float_as_uint(cst1) ^ (sacc & 0x80000000u), c2s = __float_as_uint(cst2) ^ (sacc & 0x80000000u);
#pragma unroll
            for (int j = 0; j < DC; ++j) {
                const uint32_t ms = (fabsf(x[j]) == min1) ? c1s : c2s;
                const float out = __uint_as_float(ms ^ (__float_as_uint(x[j]) & 0x80000000u));
                Rl[j * Z] = out;
                bel<LT, GL>(Lb, idx[j]) = x[j] + out;
            }
        }
    } else {
        const int lo = -(u.msg_max + 1), hi = u.msg_max;
        int sign = synbit, min1 = 1 << 30, min2 = 1 << 30;
#pragma unroll
        for (int j = 0; j < DC; ++j) {
            const int t = clipi(x[j], lo, hi);
            const int a = abs(t);
            sign ^= (t < 0);
            min2 = min(min2, max(a, min1));
            min1 = min(min1, a);
        }
        min2 = min(min2, u.msg_max + 1);   // degree-1 row
        int c1, c2;
        if (u.rule == QLDPC_RULE_OMS) { c1 = max(min2 - u.offset_int, 0); c2 = max(min1 - u.offset_int, 0); }
        else { c1 = norm8(min2, u.norm_eighths); c2 = norm8(min1, u.norm_eighths); }
#pragma unroll
        for (int j = 0; j < DC; ++j) {
            const int t = clipi(x[j], lo, hi);
            const int mag = (abs(t) == min1) ? c1 : c2;
            const int out = (sign ^ (t < 0)) ? -mag : mag;
            Rl[j * Z] = (MT)out;
            bel<LT, GL>(Lb, idx[j]) = (LT)clipi(x[j] + out, -(u.app_max + 1), u.app_max);
        }
    }
}

// Min-sum rules with COMPRESSED check-node state: the dc messages of a check take two magnitudes, so what is
// kept per check lane is {c1, c2, index of the minimum, dc sign bits} -- three 32-bit words for the float tier (c1, c2 as
// floats), two for the integer tiers (c1 | c2 << 16) -- instead of dc words.  The old message of edge j is rebuilt as
// +-(j == index ? c1 : c2): exact, because on a tie of the minimum c1 == c2.  Planes: Rc[plane * Z] (one 128-byte line per
// warp and plane).  BG1 Z=384 float: 212 KB of messages per frame instead of 485 KB, 3 loads + 3 stores per row instead of
// 2 dc; the scratch of all frames in flight fits in L2 again.
// Float tier: everything on the raw bits.  meta = imin << 27 | sign bits, the sign of edge j at bit DC-1-j (the word is
// built by funnel-shifting one sign bit in per edge); the old message of edge j is (j == imin ? c1 : c2) with bit 31 taken
// from meta << (32 - DC + j); the new one is the magnitude's bits with the sign product's bit 31.
template <typename LT, int DC, bool GL>
__device__ __forceinline__ void row_lane_cmp(const Upd u, char *Lb, uint32_t *Rc, const int2 *ed, const Lane t, int Z, int synbit, bool first)
{
    constexpr bool kFloat = std::is_floating_point<LT>::value;
    typedef typename std::conditional<kFloat, float, int>::type XT;
    static_assert(DC <= 27, "index (5 bits) + sign bits share one word");
    int idx[DC];
    if constexpr (kFloat) {
        uint32_t c1o = 0, c2o = 0, meta = 0;
        if (!first) { c1o = Rc[0]; c2o = Rc[Z]; meta = Rc[2 * Z]; }
        const int idxo = (int)(meta >> 27);
        float x[DC];
        uint32_t sacc = (uint32_t)synbit << 31;
        float min1 = 3.402823466e+38f, min2 = 3.402823466e+38f;
#pragma unroll
        for (int j = 0; j < DC; ++j) {
            idx[j] = bel_off(ed[j], t);
            const uint32_t m = (j == idxo) ? c1o : c2o;     // first iteration: c1o = c2o = 0, meta = 0 -> +0
            const float ro = __uint_as_float(m ^ ((meta << (32 - DC + j)) & 0x80000000u));
            x[j] = bel<LT, GL>(Lb, idx[j]) - ro;
            const float a = fabsf(x[j]);
            sacc ^= __float_as_uint(x[j]);
            min2 = fminf(min2, fmaxf(a, min1));
            min1 = fminf(min1, a);
        }
        float cst1 = 0.f, cst2 = 0.f;
        if (u.rule == QLDPC_RULE_NMS) { cst1 = min2 * u.norm; cst2 = min1 * u.norm; }
        else if (u.rule == QLDPC_RULE_OMS) { cst1 = fmaxf(0.0f, min2 - u.offset); cst2 = fmaxf(0.0f, min1 - u.offset); }
        uint32_t nmeta = 0;
        int imin = 0;
        // the row's sign product goes into the two magnitudes once; per edge: select, flip by the edge's own sign, and
        // shift the result's bit 31 into the sign word
        const uint32_t c1s = __float_as_uint(cst1) ^ (sacc & 0x80000000u), c2s = __float_as_uint(cst2) ^ (sacc & 0x80000000u);
#pragma unroll
        for (int j = 0; j < DC; ++j) {
            const bool is_min = fabsf(x[j]) == min1;
            const uint32_t ob = (is_min ? c1s : c2s) ^ (__float_as_uint(x[j]) & 0x80000000u);
            imin = is_min ? j : imin;
            nmeta = __funnelshift_l(ob, nmeta, 1);                                 // nmeta << 1 | sign of the new message
            bel<LT, GL>(Lb, idx[j]) = x[j] + __uint_as_float(ob);
        }
        Rc[0] = __float_as_uint(cst1); Rc[Z] = __float_as_uint(cst2); Rc[2 * Z] = nmeta | ((uint32_t)imin << 27);
    } else {
    XT c1o = 0, c2o = 0;
    uint32_t meta = 0;
    if (!first) {
        const uint32_t cc = Rc[0];
        c1o = (int)(cc & 0xffffu); c2o = (int)(cc >> 16); meta = Rc[Z];
    }
    const int idxo = (int)(meta >> 27);
    XT x[DC];
#pragma unroll
    for (int j = 0; j < DC; ++j) {
        idx[j] = bel_off(ed[j], t);
        const XT m = (j == idxo) ? c1o : c2o;
        const XT ro = ((meta >> j) & 1u) ? -m : m;          // first iteration: c1o = c2o = 0, meta = 0
        x[j] = (XT)bel<LT, GL>(Lb, idx[j]) - ro;
    }
    uint32_t nmeta = 0;
    int imin = 0;
    {
        const int lo = -(u.msg_max + 1), hi = u.msg_max;
        int sign = synbit, min1 = 1 << 30, min2 = 1 << 30;
#pragma unroll
        for (int j = 0; j < DC; ++j) {
            const int t = clipi(x[j], lo, hi);
            const int a = abs(t);
            sign ^= (t < 0);
            min2 = min(min2, max(a, min1));
            min1 = min(min1, a);
        }
        min2 = min(min2, u.msg_max + 1);   // degree-1 row
        int c1, c2;
        if (u.rule == QLDPC_RULE_OMS) { c1 = max(min2 - u.offset_int, 0); c2 = max(min1 - u.offset_int, 0); }
        else { c1 = norm8(min2, u.norm_eighths); c2 = norm8(min1, u.norm_eighths); }
#pragma unroll
        for (int j = 0; j < DC; ++j) {
            const int t = clipi(x[j], lo, hi);
            const bool is_min = abs(t) == min1;
            const int mag = is_min ? c1 : c2;
            const int neg = sign ^ (t < 0);
            const int out = neg ? -mag : mag;
            imin = is_min ? j : imin;
            nmeta |= (uint32_t)neg << j;
            bel<LT, GL>(Lb, idx[j]) = (LT)clipi(x[j] + out, -(u.app_max + 1), u.app_max);
        }
        Rc[0] = (uint32_t)c1 | ((uint32_t)c2 << 16); Rc[Z] = nmeta | ((uint32_t)imin << 27);
    }
    }
}

template <typename LT, bool GL>
__device__ __forceinline__ void row_dispatch_cmp(const Upd u, const RowMeta ly, char *Lb, uint32_t *Rc, const int2 *ed, const Lane t, int Z,
                                                 int synbit, bool first)
{
#define QL_DC(D) case D: row_lane_cmp<LT, D, GL>(u, Lb, Rc, ed, t, Z, synbit, first); return;
    switch (ly.degree) {   // block-uniform; the host selects this mode only when every row is within the compiled degrees
        QL_DC(1) QL_DC(2) QL_DC(3) QL_DC(4) QL_DC(5) QL_DC(6) QL_DC(7) QL_DC(8) QL_DC(9) QL_DC(10)
        QL_DC(11) QL_DC(12) QL_DC(13) QL_DC(14) QL_DC(15) QL_DC(16) QL_DC(17) QL_DC(18) QL_DC(19) QL_DC(20)
    default: break;
    }
#undef QL_DC
}

// any degree: two passes over memory (rows heavier than the compiled degrees)
template <typename LT, typename MT, bool GL>
__device__ __noinline__ void row_lane_any(const Upd u, char *Lb, MT *Rl, const int2 *ed, int dc, const Lane t, int Z, int synbit, bool first)
{
    constexpr bool kFloat = std::is_floating_point<LT>::value;
    auto contrib = [&](int j, int &at) {
        at = bel_off(ed[j], t);
        if constexpr (kFloat) return (float)bel<LT, GL>(Lb, at) - (first ? 0.0f : (float)Rl[j * Z]);
        else return (int)bel<LT, GL>(Lb, at) - (first ? 0 : (int)Rl[j * Z]);
    };
    int at;
    if constexpr (kFloat) {
        int sign = synbit;
        float product = 1.0f, min1 = 3.402823466e+38f, min2 = 3.402823466e+38f;
        for (int j = 0; j < dc; ++j) {
            const float x = contrib(j, at);
            const float a = fabsf(x);
            sign ^= signbit(x) ? 1 : 0;
            if (u.rule == QLDPC_RULE_SPA) {
                const float t = tanh_half(u, a);
                product *= (t != 0.0f) ? t : 1e-12f;
            } else {
                min2 = fminf(min2, fmaxf(a, min1));
                min1 = fminf(min1, a);
            }
        }
        float cst1 = 0.f, cst2 = 0.f;
        if (u.rule == QLDPC_RULE_NMS) { cst1 = min2 * u.norm; cst2 = min1 * u.norm; }
        else if (u.rule == QLDPC_RULE_OMS) { cst1 = fmaxf(0.0f, min2 - u.offset); cst2 = fmaxf(0.0f, min1 - u.offset); }
        for (int j = 0; j < dc; ++j) {
            const float x = contrib(j, at);
            float mag;
            if (u.rule == QLDPC_RULE_SPA) {
                const float t = tanh_half(u, fabsf(x));
                float r = product / ((t != 0.0f) ? t : 1e-12f);
                r = (r < 1.0f) ? r : 1.0f - 1.1920929e-07f;
                mag = two_atanh(u, r);
            } else {
                mag = (fabsf(x) == min1) ? cst1 : cst2;
            }
            const float out = (sign ^ (signbit(x) ? 1 : 0)) ? -mag : mag;
            Rl[j * Z] = out;
            bel<LT, GL>(Lb, at) = x + out;
        }
    } else {
        const int lo = -(u.msg_max + 1), hi = u.msg_max;
        int sign = synbit, min1 = 1 << 30, min2 = 1 << 30;
        for (int j = 0; j < dc; ++j) {
            const int t = clipi(contrib(j, at), lo, hi);
            const int a = abs(t);
            sign ^= (t < 0);
            min2 = min(min2, max(a, min1));
            min1 = min(min1, a);
        }
        min2 = min(min2, u.msg_max + 1);
        int c1, c2;
        if (u.rule == QLDPC_RULE_OMS) { c1 = max(min2 - u.offset_int, 0); c2 = max(min1 - u.offset_int, 0); }
        else { c1 = norm8(min2, u.norm_eighths); c2 = norm8(min1, u.norm_eighths); }
        for (int j = 0; j < dc; ++j) {
            const int c = contrib(j, at);
            const int t = clipi(c, lo, hi);
            const int mag = (abs(t) == min1) ? c1 : c2;
            const int out = (sign ^ (t < 0)) ? -mag : mag;
            Rl[j * Z] = (MT)out;
            bel<LT, GL>(Lb, at) = (LT)clipi(c + out, -(u.app_max + 1), u.app_max);
        }
    }
}

template <typename LT, typename MT, bool GL>
__device__ __forceinline__ void row_dispatch(const Upd u, const RowMeta ly, char *Lb, MT *Rl, const int2 *ed, const Lane t, int Z, int synbit,
                                             bool first)
{
#define QL_DC(D) case D: row_lane<LT, MT, D, GL>(u, Lb, Rl, ed, t, Z, synbit, first); return;
    switch (ly.degree) {   // block-uniform: no divergence
        QL_DC(1) QL_DC(2) QL_DC(3) QL_DC(4) QL_DC(5) QL_DC(6) QL_DC(7) QL_DC(8) QL_DC(9) QL_DC(10)
        QL_DC(11) QL_DC(12) QL_DC(13) QL_DC(14) QL_DC(15) QL_DC(16) QL_DC(17) QL_DC(18) QL_DC(19) QL_DC(20)
    default: break;
    }
#undef QL_DC
    static_assert(kMaxDc == 20, "row_dispatch lists the compiled degrees");
    row_lane_any<LT, MT, GL>(u, Lb, Rl, ed, ly.degree, t, Z, synbit, first);
}

__device__ __forceinline__ int syn_bit(const uint32_t *syn, int m) { return syn ? (int)((syn[m >> 5] >> (31 - (m & 31))) & 1u) : 0; }

// Shared memory: RowMeta rows[brows]; int2 edges[nnz]; LT L[N].
// MAXT / MINB: launch bounds of the instantiation (threads per CTA = Z rounded up to a warp, at most MAXT).
// GL: the beliefs of a frame do not fit in shared memory and live in a global scratch (p.app) instead.
template <typename LT, typename MT, typename IN, int MAXT, int MINB, bool GL>
__global__ void __launch_bounds__(MAXT, MINB) layered_generic_kernel(const LayeredGenParams p)
{
    char *smem = qldpc_lg_smem;
    constexpr bool kFloat = std::is_floating_point<LT>::value;
    const int tid = threadIdx.x, nt = blockDim.x, Z = p.Z, R = p.brows;
    RowMeta *rows = reinterpret_cast<RowMeta *>(smem);
    int2 *edges = reinterpret_cast<int2 *>(rows + R);
    LT *L;
    if constexpr (GL) L = reinterpret_cast<LT *>(p.app) + (size_t)blockIdx.x * p.N;
    else L = reinterpret_cast<LT *>(edges + p.nnz);     // no select with a global pointer here: the accesses stay LDS / STS
    // belief base and edge table for bel_off(): with the beliefs in shared memory the base is the start of the dynamic
    // shared memory, so that every access is LDS / STS [register + immediate]
    char *Lb = GL ? reinterpret_cast<char *>(L) : smem;
    const int l_off = GL ? 0 : (int)(unsigned)__cvta_generic_to_shared(L);
    for (int r = tid; r < R; r += nt) rows[r] = RowMeta{p.layers[r].edge_begin, p.layers[r].degree};
    for (int e = tid; e < p.nnz; e += nt)
        edges[e] = make_int2(l_off + (int)sizeof(LT) * (p.aux[e].col * Z + p.aux[e].shift), Z - p.aux[e].shift);
    MT *Rg = reinterpret_cast<MT *>(p.msg) + (size_t)blockIdx.x * p.nnz * Z;
    const Upd upd{p.rule, p.offset_int, p.norm_eighths, p.msg_max, p.app_max, p.norm, p.offset, p.fast_spa};

    // syndrome of the hard decisions (beliefs in shared memory), OR over the block
    auto syndrome_bad = [&](const uint32_t *syn) {
        int bad = 0;
        for (int r = 0; r < R; ++r) {
            const RowMeta ly = rows[r];
            for (int i = tid; i < Z; i += nt) {
                unsigned s = (unsigned)syn_bit(syn, r * Z + i);
                const Lane t{i, i * (int)sizeof(LT), Z * (int)sizeof(LT)};
                for (int j = 0; j < ly.degree; ++j) s ^= (unsigned)(bel<LT, GL>(Lb, bel_off(edges[ly.edge_begin + j], t)) < (LT)0);
                bad |= (int)(s & 1u);
            }
        }
        return __syncthreads_or(bad) != 0;
    };

    // The same test on bit vectors when Z is a multiple of 32 and a thread owns one lane (every 5G lifting size >= 32): the
    // hard decisions of a block column are balloted into Z-bit vectors behind the beliefs (N / 8 bytes), and a syndrome word
    // (row, 32 check lanes) is the XOR of the rotated windows of its columns' vectors -- two loads and a funnel shift per
    // edge and WORD instead of an address computation, a load and a compare per edge and LANE.
    uint32_t *hd = reinterpret_cast<uint32_t *>(L + p.N);
    const bool packed_syn = !GL && (Z & 31) == 0 && nt >= Z;
    auto syndrome_bad_packed = [&](const uint32_t *syn) {
        const int ZW = Z >> 5;
        for (int c = 0; c < p.bcols; ++c) {
            const bool neg = tid < Z && L[c * Z + tid] < (LT)0;
            const unsigned b = __ballot_sync(0xffffffffu, neg);
            if ((tid & 31) == 0 && tid < Z) hd[c * ZW + (tid >> 5)] = b;      // bit k of word w = lane 32 w + k
        }
        __syncthreads();
        int bad = 0;
        for (int item = tid; item < R * ZW; item += nt) {
            const int r = item / ZW, w = item - r * ZW;
            const RowMeta ly = rows[r];
            uint32_t acc = syn ? __brev(syn[r * ZW + w]) : 0u;                  // the input packs check m at bit 31 - m % 32
            for (int j = 0; j < ly.degree; ++j) {
                const int col = p.aux[ly.edge_begin + j].col, shift = p.aux[ly.edge_begin + j].shift;
                int q = 32 * w + shift;                                         // check lane l reads variable lane (l + shift) mod Z
                if (q >= Z) q -= Z;
                const int qw = q >> 5, qn = qw + 1 == ZW ? 0 : qw + 1;
                const uint32_t *h = hd + col * ZW;
                acc ^= __funnelshift_r(h[qw], h[qn], q & 31);
            }
            bad |= (int)(acc != 0u);
        }
        return __syncthreads_or(bad) != 0;
    };
    auto frame_bad = [&](const uint32_t *syn) { return packed_syn ? syndrome_bad_packed(syn) : syndrome_bad(syn); };

    for (int f = blockIdx.x; f < p.F; f += gridDim.x) {
        const IN *llr = reinterpret_cast<const IN *>(p.llr) + (size_t)f * p.N;
        const uint32_t *syn = p.syn ? p.syn + (size_t)f * p.syn_words : nullptr;
        __syncthreads();                                   // tables written / previous frame's beliefs no longer read
        if (std::is_same<IN, LT>::value && !GL && ((p.N * (int)sizeof(LT)) & 15) == 0 && (reinterpret_cast<size_t>(llr) & 15) == 0 &&
            (((R + p.nnz) & 1) == 0)) {   // the beliefs start 8 (R + nnz) bytes into the 16-byte aligned shared memory
            // input and belief type agree (float, int16): 128-bit copies, a quarter / an eighth of the dependent global round
            // trips of the frame load (with early termination a frame lasts ~2 iterations and its load was a third of it)
            const uint4 *src = reinterpret_cast<const uint4 *>(llr);
            uint4 *dst = reinterpret_cast<uint4 *>(L);
            for (int v = tid; v < ((p.N * (int)sizeof(LT)) >> 4); v += nt) dst[v] = src[v];
        } else {
            for (int v = tid; v < p.N; v += nt) L[v] = (LT)llr[v];
        }
        __syncthreads();

        int it = 0, depth = 0;
        bool ok = false, checked = false;
        // Instantiations with fewer than 1024 threads serve Z <= blockDim.x only (block_threads): one lane per thread, and the
        // thread's lane constants, message pointers and -- when the frame has a syndrome and at most 64 block rows -- its R
        // syndrome bits are set up once per frame, not once per row (the syndrome bit was a dependent global load at the
        // head of every row).  The 1024-thread instantiations keep the general loop (a thread may own several lanes).
        // (float tier only: at the 80 registers of the integer instantiations the values kept across the rows spill,
        // BG1 Z=384 i16 9.8 -> 9.0 Gbit/s)
        constexpr bool kOnePass = MAXT < 1024 && kFloat;
        const bool act = tid < Z;
        const Lane tl{tid, tid * (int)sizeof(LT), Z * (int)sizeof(LT)};
        const bool syn_in_regs = kOnePass && syn != nullptr && R <= 64;
        unsigned long long synmask = 0;
        if (syn_in_regs && act)
            for (int r = 0; r < R; ++r) synmask |= (unsigned long long)syn_bit(syn, r * Z + tid) << r;
        while (it < p.max_iter) {
            if (p.compressed) {
                constexpr int kPlanes = kFloat ? 3 : 2;
                uint32_t *Rc = reinterpret_cast<uint32_t *>(p.msg) + (size_t)blockIdx.x * R * kPlanes * Z;
                if constexpr (kOnePass) {
                    uint32_t *Rt = Rc + tid;
                    const int rstep = kPlanes * Z;
                    for (int r = 0; r < R; ++r) {
                        const RowMeta ly = rows[r];
                        if (act)
                            row_dispatch_cmp<LT, GL>(upd, ly, Lb, Rt, edges + ly.edge_begin, tl, Z,
                                                     syn_in_regs ? (int)((synmask >> r) & 1ull) : syn_bit(syn, r * Z + tid), it == 0);
                        Rt += rstep;
                        __syncthreads();
                    }
                } else {
                    for (int r = 0; r < R; ++r) {
                        const RowMeta ly = rows[r];
                        for (int i = tid; i < Z; i += nt)
                            row_dispatch_cmp<LT, GL>(upd, ly, Lb, Rc + ((size_t)r * kPlanes * Z + i), edges + ly.edge_begin,
                                                     Lane{i, i * (int)sizeof(LT), Z * (int)sizeof(LT)}, Z, syn_bit(syn, r * Z + i), it == 0);
                        __syncthreads();
                    }
                }
            } else if constexpr (kOnePass) {
                MT *Rt = Rg + tid;
                for (int r = 0; r < R; ++r) {
                    const RowMeta ly = rows[r];
                    if (act)
                        row_dispatch<LT, MT, GL>(upd, ly, Lb, Rt + (size_t)ly.edge_begin * Z, edges + ly.edge_begin, tl, Z,
                                                 syn_in_regs ? (int)((synmask >> r) & 1ull) : syn_bit(syn, r * Z + tid), it == 0);
                    __syncthreads();
                }
            } else {
                for (int r = 0; r < R; ++r) {
                    const RowMeta ly = rows[r];
                    for (int i = tid; i < Z; i += nt)
                        row_dispatch<LT, MT, GL>(upd, ly, Lb, Rg + ((size_t)ly.edge_begin * Z + i), edges + ly.edge_begin,
                                                 Lane{i, i * (int)sizeof(LT), Z * (int)sizeof(LT)}, Z, syn_bit(syn, r * Z + i), it == 0);
                    __syncthreads();
                }
            }
            ++it;
            checked = false;
            // int tiers: check after every iteration (oracle ora_decode_layered_fixed);
            // float tiers: AFF3CT skips the check after the last iteration and honours syndrome_depth
            const bool want = p.early_stop && (kFloat ? it != p.max_iter : true);
            if (want) {
                ok = !frame_bad(syn);
                checked = true;
                if (ok) { if (!kFloat || ++depth >= p.syndrome_depth) break; }
                else depth = 0;
            }
        }
        if (!checked) ok = !frame_bad(syn);

        uint32_t *ab = p.allbits + (size_t)f * p.cw_words;
        if (packed_syn) {
            // the bit vectors of the last syndrome test ARE the hard decisions of the final beliefs: word w of the output
            // (MSB first) is the bit reversal of vector word w (N is a multiple of 32 here)
            for (int w = tid; w < p.cw_words; w += nt) ab[w] = __brev(hd[w]);
        } else {
            for (int w = tid; w < p.cw_words; w += nt) {
                uint32_t v = 0;
                for (int b = 0; b < 32; ++b) {
                    const int idx = 32 * w + b;
                    if (idx < p.N && L[idx] < (LT)0) v |= 1u << (31 - b);
                }
                ab[w] = v;
            }
        }
        if (p.posterior) {
            if constexpr (kFloat) {
                float *po = reinterpret_cast<float *>(p.posterior) + (size_t)f * p.N;
                for (int v = tid; v < p.N; v += nt) po[v] = L[v];
            } else {
                int *po = reinterpret_cast<int *>(p.posterior) + (size_t)f * p.N;
                for (int v = tid; v < p.N; v += nt) po[v] = (int)L[v];
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
    }
}

int block_threads(int Z) { return std::min(1024, std::max(64, (Z + 31) / 32 * 32)); }

// the instantiation for a (tier, block size): 128 threads x 4 blocks, 384 x 2, or 1024 x 1 per SM as the register budget
#define QL_GEN_KERNELS(X, LT, MT, IN)                                          \
    X(128, false, (layered_generic_kernel<LT, MT, IN, 128, 4, false>))        \
    X(384, false, (layered_generic_kernel<LT, MT, IN, 384, 2, false>))        \
    X(1024, false, (layered_generic_kernel<LT, MT, IN, 1024, 1, false>))      \
    X(1024, true, (layered_generic_kernel<LT, MT, IN, 1024, 1, true>))

template <typename K>
int occupancy_k(K kern, int block, int smem_bytes)
{
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) { cudaGetLastError(); return 0; }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, block, (size_t)smem_bytes) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
template <typename K>
int launch_k(K kern, const LayeredGenParams &p, int grid, int block, int smem_bytes, cudaStream_t st)
{
    QLDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    kern<<<grid, block, smem_bytes, st>>>(p);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

}  // namespace

int layered_generic_max_compiled_degree() { return kMaxDc; }
// message scratch of one frame in flight: compressed check-node state (min-sum rules) or one value per edge
size_t layered_generic_msg_scratch_bytes(int dtype, int compressed, int brows, int nnz, int Z)
{
    if (compressed) return (size_t)brows * (dtype == QLDPC_DTYPE_F32 ? 3 : 2) * Z * 4;
    return (size_t)nnz * Z * (dtype == QLDPC_DTYPE_F32 ? 4 : 2);
}
int layered_generic_belief_bytes(int dtype) { return dtype == QLDPC_DTYPE_F32 ? 4 : 2; }
int layered_generic_msg_bytes(int dtype) { return dtype == QLDPC_DTYPE_F32 ? 4 : 2; }
int layered_generic_smem_bytes(int brows, int nnz, int N, int dtype)
{
    // tables, beliefs, and N / 8 bytes of hard-decision bit vectors for the packed syndrome test
    return (brows * 8 + nnz * 8 + N * layered_generic_belief_bytes(dtype) + (N + 31) / 32 * 4 + 15) / 16 * 16;
}

// co-resident CTAs per SM of the instantiation that serves (dtype, Z); 0: the beliefs do not fit in shared memory
int layered_generic_blocks_per_sm(int dtype, int Z, int smem_bytes, int beliefs_global)
{
    const int block = block_threads(Z);
    const bool gl = beliefs_global != 0;
#define QL_X(T, G, K) if (block <= T && gl == G) return occupancy_k(K, block, smem_bytes);
    switch (dtype) {
    case QLDPC_DTYPE_F32: QL_GEN_KERNELS(QL_X, float, float, float) break;
    case QLDPC_DTYPE_I16: QL_GEN_KERNELS(QL_X, int16_t, int16_t, int16_t) break;
    case QLDPC_DTYPE_I8: QL_GEN_KERNELS(QL_X, int16_t, int16_t, int8_t) break;
    default: break;
    }
#undef QL_X
    return 0;
}

int launch_layered_generic(const LayeredGenParams &p, int grid, cudaStream_t st)
{
    const int block = block_threads(p.Z);
    const int smem = layered_generic_smem_bytes(p.brows, p.nnz, p.app ? 0 : p.N, p.dtype);
    const bool gl = p.app != nullptr;
#define QL_X(T, G, K) if (block <= T && gl == G) return launch_k(K, p, grid, block, smem, st);
    switch (p.dtype) {
    case QLDPC_DTYPE_F32: QL_GEN_KERNELS(QL_X, float, float, float) break;
    case QLDPC_DTYPE_I16: QL_GEN_KERNELS(QL_X, int16_t, int16_t, int16_t) break;
    case QLDPC_DTYPE_I8: QL_GEN_KERNELS(QL_X, int16_t, int16_t, int8_t) break;
    default: break;
    }
#undef QL_X
    return QLDPC_ERR_UNSUPPORTED;
}

}  // namespace qldpc

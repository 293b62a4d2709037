// Layered int8 min-sum decoder for quasi-cyclic codes, Z % 4 == 0 ("zpack4" family).
//
// Arithmetic = ML/BPSK_nrldpc_sim_FP.m:35-94 (layered offset min-sum, messages in
// [-(maxqr+1), maxqr], beliefs in [-128,127]) generalised by syndrome input, early stop and the
// shift-normalised rule; bit-exact with oracle/qldpc_oracle.c:ora_decode_layered_fixed.
//
// Data layout (shared memory, one "slot" per frame in flight):
//   belief word i of block column c holds the four lanes {i, i+W, i+2W, i+3W} (W = Z/4) as
//   biased bytes (L+128).  A circulant shift s = q*W + r then maps the four check lanes of
//   thread i onto ONE aligned word (i+r) mod W, rotated by q (+1 when i+r wraps) bytes: the
//   cyclic shift is an address offset plus a PRMT selector, never an unaligned access.
//   check-to-variable messages are stored per edge in check-lane order, biased bytes (R+128).
// Arithmetic runs on half2 pairs: a byte b placed in the low byte of an fp16 is the subnormal
// b * 2^-24, and add / sub / min / max of such values are exact integer operations for
// |x| < 2048.  That gives free |x|, a separate sign bit (min.xorsign.abs accumulates min1 and
// the sign product -- with the syndrome bit folded into its initial sign -- in one instruction)
// and 2 lanes per instruction on the native HADD2/HMNMX2/HFMA2 pipes; the byte-SIMD integer
// intrinsics (__vsubss4 ...) are emulated with 5-10 instructions each on sm_100a.
// Early termination: hard decisions are packed with warp ballots into Z-bit vectors and the
// syndrome is evaluated word-wise with funnel shifts + XOR (a few % of an iteration).
#include <cuda_fp16.h>

#include "kernels.hpp"

namespace qldpc {

namespace {

typedef unsigned int u32;

constexpr u32 kSignMask = 0x80008000u;
constexpr u32 kInf2 = 0x03ff03ffu;     // 1023 ulp: larger than any message magnitude
constexpr u32 kOne2 = 0x3c003c00u;     // 1.0h, 1.0h
constexpr int kMaxBlock = 256;

__device__ __forceinline__ __half2 h2(u32 x) { return *reinterpret_cast<__half2 *>(&x); }
__device__ __forceinline__ u32 bits(__half2 h) { return *reinterpret_cast<u32 *>(&h); }
__device__ __forceinline__ u32 dup16(u32 v) { return (v & 0xffffu) | (v << 16); }

__device__ __forceinline__ u32 hsub(u32 a, u32 b) { return bits(__hsub2(h2(a), h2(b))); }
__device__ __forceinline__ u32 hadd(u32 a, u32 b) { return bits(__hadd2(h2(a), h2(b))); }
__device__ __forceinline__ u32 hmin(u32 a, u32 b) { return bits(__hmin2(h2(a), h2(b))); }
__device__ __forceinline__ u32 hmax(u32 a, u32 b) { return bits(__hmax2(h2(a), h2(b))); }
__device__ __forceinline__ u32 habs(u32 a) { return bits(__habs2(h2(a))); }
// PRMT with a register selector (no 0x7777 masking as __byte_perm would add)
__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 sel)
{
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// relu(a + b)
__device__ __forceinline__ u32 hadd_relu(u32 a, u32 b) { return bits(__hfma2_relu(h2(a), h2(kOne2), h2(b))); }
__device__ __forceinline__ u32 heq_mask(u32 a, u32 b) { return __heq2_mask(h2(a), h2(b)); }
// |d| = min(|a|,|b|), sign(d) = sign(a) ^ sign(b)
__device__ __forceinline__ u32 min_xorsign_abs(u32 a, u32 b)
{
    u32 d;
    asm("min.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

__device__ __forceinline__ void bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ bool bar_red_or(int id, int nthreads, bool pred)
{
    u32 r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %3, 0;\n\tbar.red.or.pred q, %1, %2, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(id), "r"(nthreads), "r"((u32)pred)
        : "memory");
    return r != 0;
}

// integer k/8 normalisation on two packed non-negative 16-bit fields (AFF3CT integer NMS).
// NK is the compile-time factor (1..8); NK < 0 reads the factor at run time.
template <int NK>
__device__ __forceinline__ u32 norm_eighths2(u32 x, int k_rt)
{
    const int k = NK < 0 ? k_rt : NK;
    if (k >= 8) return x;
    const u32 s1 = (x >> 1) & 0x7fff7fffu, s2 = (x >> 2) & 0x3fff3fffu, s3 = (x >> 3) & 0x1fff1fffu;
    u32 r = 0;
    if (k & 4) r += s1;
    if (k & 2) r += s2;
    if (k & 1) r += s3;
    return r;
}

struct LayerCtx {
    u32 cLo, cHi;         // message clip [-(msg_max+1), msg_max] as half2 ulps
    u32 cM2cap;           // msg_max+1: second minimum of a degree-1 check
    u32 c128, c255;
    int norm_eighths;
    u32 negOff;           // -offset
};

// One layer (block row) for the four check lanes of this thread.
//   NK    rule: 0 = offset min-sum, 1..8 = normalised by NK/8, -1 = normalised, factor at run time
//   DC    unrolled edge slots; EXACT: the row has exactly DC edges, else DC-1 or DC
//   Li    this thread's belief base address (slot beliefs + 4*i)
//   Rrow  this thread's message word of the row's first edge; consecutive edges are W words apart
//   etab  two int4 per edge: {off0, off1, thresh, -} {selA0, selA1, selW0, selW1}
template <int NK, int DC, bool EXACT>
__device__ __forceinline__ void process_layer(const LayerCtx &cx, char *Li, u32 *Rrow, int W, const int4 *etab, int dc,
                                              int i, u32 synbits)
{
    u32 uA[DC], uB[DC], tA[DC], tB[DC], sw[DC];
    char *ad[DC];
    // running sign product starts at the syndrome bit of each lane
    u32 m1A = kInf2 ^ (((synbits & 1u) << 15) | ((synbits & 2u) << 30));
    u32 m1B = kInf2 ^ (((synbits & 4u) << 13) | ((synbits & 8u) << 28));
    u32 m2A = kInf2, m2B = kInf2;

#pragma unroll
    for (int j = 0; j < DC; ++j) {
        if (EXACT || j < DC - 1 || j < dc) {
            const int4 ed = etab[2 * j];       // off0, off1, thresh, -
            const int4 sl = etab[2 * j + 1];   // selA0, selA1, selW0, selW1
            const bool wrap = i >= ed.z;
            char *a = Li + (wrap ? ed.y : ed.x);
            const u32 selA = wrap ? sl.y : sl.x;
            sw[j] = wrap ? sl.w : sl.z;
            ad[j] = a;
            const u32 X = *reinterpret_cast<const u32 *>(a);
            const u32 Y = Rrow[j * W];
            const u32 xA = prmt(X, 0u, selA), xB = prmt(X, 0u, selA ^ 0x0202u);
            const u32 yA = prmt(Y, 0u, 0x4140u), yB = prmt(Y, 0u, 0x4342u);
            const u32 ua = hsub(xA, yA), ub = hsub(xB, yB);          // L - R_old  (:51)
            const u32 ta = hmin(hmax(ua, cx.cLo), cx.cHi);            // clip to the message range (:54-55)
            const u32 tb = hmin(hmax(ub, cx.cLo), cx.cHi);
            uA[j] = ua; uB[j] = ub; tA[j] = ta; tB[j] = tb;
            m2A = hmax(habs(m1A), hmin(habs(ta), m2A));               // second minimum (:61)
            m2B = hmax(habs(m1B), hmin(habs(tb), m2B));
            m1A = min_xorsign_abs(m1A, ta);                           // first minimum and sign product (:60,:63)
            m1B = min_xorsign_abs(m1B, tb);
        }
    }
    const u32 parA = m1A & kSignMask, parB = m1B & kSignMask;
    const u32 min1A = habs(m1A), min1B = habs(m1B);
    m2A = hmin(m2A, cx.cM2cap);
    m2B = hmin(m2B, cx.cM2cap);
    u32 c1A, c1B, c2A, c2B;   // c1: magnitude sent to the position of the minimum, c2: to all others
    if (NK == 0) {                                                    // :65-72
        c1A = hadd_relu(m2A, cx.negOff); c1B = hadd_relu(m2B, cx.negOff);
        c2A = hadd_relu(min1A, cx.negOff); c2B = hadd_relu(min1B, cx.negOff);
    } else {
        c1A = norm_eighths2<NK>(m2A, cx.norm_eighths); c1B = norm_eighths2<NK>(m2B, cx.norm_eighths);
        c2A = norm_eighths2<NK>(min1A, cx.norm_eighths); c2B = norm_eighths2<NK>(min1B, cx.norm_eighths);
    }
    c1A ^= parA; c2A ^= parA; c1B ^= parB; c2B ^= parB;               // parity folded into both candidates

#pragma unroll
    for (int j = 0; j < DC; ++j) {
        if (EXACT || j < DC - 1 || j < dc) {
            const u32 eA = heq_mask(habs(tA[j]), min1A), eB = heq_mask(habs(tB[j]), min1B);
            // |t| == min1 ? c1 : c2, then the edge's own sign (:73-75)
            const u32 rA = ((eA & c1A) | (~eA & c2A)) ^ (tA[j] & kSignMask);
            const u32 rB = ((eB & c1B) | (~eB & c2B)) ^ (tB[j] & kSignMask);
            const u32 bA = hadd(rA, cx.c128), bB = hadd(rB, cx.c128); // biased new message
            const u32 lA = hmin(hadd_relu(uA[j], bA), cx.c255);       // clip(L - R_old + R_new) biased (:88-91)
            const u32 lB = hmin(hadd_relu(uB[j], bB), cx.c255);
            Rrow[j * W] = prmt(bA, bB, 0x6420u);
            *reinterpret_cast<u32 *>(ad[j]) = prmt(lA, lB, sw[j]);
        }
    }
}

template <int NK>
__device__ __forceinline__ void dispatch_layer(const LayerCtx &cx, char *Li, u32 *Rrow, int W, const int4 *et, int dc,
                                               int i, u32 synbits)
{
    switch (dc) {
    case 1: process_layer<NK, 1, true>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 2: process_layer<NK, 2, true>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 3: process_layer<NK, 3, true>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 4: process_layer<NK, 4, true>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 5: process_layer<NK, 5, true>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 6: process_layer<NK, 6, true>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 7: process_layer<NK, 7, true>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 8: process_layer<NK, 8, true>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 9: process_layer<NK, 9, true>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 10: process_layer<NK, 10, true>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 11: case 12: process_layer<NK, 12, false>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 13: case 14: process_layer<NK, 14, false>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 15: case 16: process_layer<NK, 16, false>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    case 17: case 18: process_layer<NK, 18, false>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    default: process_layer<NK, 20, false>(cx, Li, Rrow, W, et, dc, i, synbits); break;
    }
}

// cnt (<=32) bits of a Z-bit little-endian vector starting at bit pos (pos + cnt <= Z)
__device__ __forceinline__ u32 extract_bits(const u32 *vec, int nwords, int pos, int cnt)
{
    const int w0 = pos >> 5, sh = pos & 31;
    const u32 lo = vec[w0];
    const u32 hi = (w0 + 1 < nwords) ? vec[w0 + 1] : 0u;
    const u32 v = __funnelshift_r(lo, hi, sh);
    return cnt >= 32 ? v : (v & ((1u << cnt) - 1u));
}
// cnt bits starting at `start`, wrapping at Z (cnt <= Z)
__device__ __forceinline__ u32 rotated_bits(const u32 *vec, int Z, int nwords, int start, int cnt)
{
    const int first = min(cnt, Z - start);
    u32 x = extract_bits(vec, nwords, start, first);
    if (first < cnt) x |= extract_bits(vec, nwords, 0, cnt - first) << first;
    return x;
}

template <int NK>
__global__ void __launch_bounds__(kMaxBlock, 1) layered_i8_kernel(const LayeredI8Params p)
{
    extern __shared__ __align__(16) char smem[];
    const int tpg = p.tpg, W = p.W, Z = p.Z, ZW32 = p.ZW32;
    const int g = threadIdx.x / tpg, i = threadIdx.x - g * tpg;
    const int lane = threadIdx.x & 31, wis = i >> 5;
    const int bar_id = 1 + g;
    const bool active = i < W;
    const bool aligned = (W & 31) == 0;

    // ---- shared tables (all slots)
    int4 *etab = reinterpret_cast<int4 *>(smem);   // two int4 per edge
    QcEdgeAux *atab = reinterpret_cast<QcEdgeAux *>(etab + 2 * p.nnz);
    QcLayer *ltab = reinterpret_cast<QcLayer *>(atab + p.nnz);
    for (int e = threadIdx.x; e < p.nnz; e += blockDim.x) {
        const QcEdge ed = p.edges[e];
        const QcEdgeAux ax = p.aux[e];
        etab[2 * e] = make_int4(ed.off0, ed.off1, ed.thresh, 0);
        etab[2 * e + 1] = make_int4(ed.selA0, ed.selA1, ax.selW0, ax.selW1);
        atab[e] = ax;
    }
    for (int r = threadIdx.x; r < p.brows; r += blockDim.x) ltab[r] = p.layers[r];
    __syncthreads();

    char *slot = smem + p.tab_bytes + (size_t)g * p.slot_bytes;
    u32 *Lw = reinterpret_cast<u32 *>(slot);
    u32 *Rw = reinterpret_cast<u32 *>(slot + p.off_R);
    u32 *hd = reinterpret_cast<u32 *>(slot + p.off_hd);
    u32 *synl = reinterpret_cast<u32 *>(slot + p.off_syn);

    LayerCtx cx;
    cx.cLo = dup16(0x8000u | (u32)(p.msg_max + 1));
    cx.cHi = dup16((u32)p.msg_max);
    cx.cM2cap = dup16((u32)(p.msg_max + 1));
    cx.c128 = dup16(128u);
    cx.c255 = dup16(255u);
    cx.norm_eighths = p.norm_eighths;
    cx.negOff = dup16(0x8000u | (u32)p.offset);

    for (int f = blockIdx.x * p.slots + g; f < p.F; f += gridDim.x * p.slots) {
        // ---- load: int8 LLRs -> interleaved biased belief words; messages = 0
        const int8_t *src = p.llr + (size_t)f * p.N;
        if (active) {
            for (int c = 0; c < p.bcols; ++c) {
                const uint8_t *q = reinterpret_cast<const uint8_t *>(src) + c * Z + i;
                const u32 b0 = __ldg(q), b1 = __ldg(q + W), b2 = __ldg(q + 2 * W), b3 = __ldg(q + 3 * W);
                Lw[c * W + i] = (b0 | (b1 << 8) | (b2 << 16) | (b3 << 24)) ^ 0x80808080u;
            }
            for (int e = 0; e < p.nnz; ++e) Rw[e * W + i] = 0x80808080u;
        }
        if (p.syn) {   // syndrome rows, Z-bit little-endian vectors (row r, lane l -> bit l)
            const u32 *sf = p.syn + (size_t)f * p.syn_words;
            for (int idx = i; idx < p.brows * ZW32; idx += tpg) {
                const int r = idx / ZW32, w = idx - r * ZW32;
                u32 v;
                if ((Z & 31) == 0) {
                    v = __brev(__ldg(sf + (r * Z) / 32 + w));
                } else {
                    v = 0;
                    const int nb = min(32, Z - 32 * w);
                    for (int b = 0; b < nb; ++b) {
                        const int gbit = r * Z + 32 * w + b;
                        v |= ((__ldg(sf + (gbit >> 5)) >> (31 - (gbit & 31))) & 1u) << b;
                    }
                }
                synl[idx] = v;
            }
        }
        bar_sync(bar_id, tpg);

        int it = 0;
        bool conv = false;
        for (;;) {
            const bool run_layers = it < p.max_iter;
            if (run_layers) {
                for (int r = 0; r < p.brows; ++r) {
                    const QcLayer ly = ltab[r];
                    if (active) {
                        u32 synbits = 0;
                        if (p.syn) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int l = i + W * k;
                                synbits |= ((synl[r * ZW32 + (l >> 5)] >> (l & 31)) & 1u) << k;
                            }
                        }
                        u32 *Rrow = Rw + ly.edge_begin * W + i;
                        dispatch_layer<NK>(cx, slot + 4 * i, Rrow, W, etab + 2 * ly.edge_begin, ly.degree, i, synbits);
                    }
                    bar_sync(bar_id, tpg);
                }
                ++it;
            }
            // the syndrome / hard-decision phase runs after every iteration when early stop is on,
            // otherwise once after the last iteration
            if (p.early_stop || it >= p.max_iter) {
                if (!aligned) {
                    for (int idx = i; idx < p.bcols * ZW32; idx += tpg) hd[idx] = 0u;
                    bar_sync(bar_id, tpg);
                }
                for (int c = 0; c < p.bcols; ++c) {
                    const u32 X = active ? Lw[c * W + i] : 0x80808080u;   // biased: bit 7 clear <=> L < 0
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const u32 b = __ballot_sync(0xffffffffu, ((X >> (8 * k + 7)) & 1u) == 0u);
                        if (lane == 0) {
                            const int o = 32 * wis + W * k;
                            if (o < Z) {
                                if (aligned) {
                                    hd[c * ZW32 + (o >> 5)] = b;
                                } else {
                                    atomicOr(&hd[c * ZW32 + (o >> 5)], b << (o & 31));
                                    if ((o & 31) && (o >> 5) + 1 < ZW32) atomicOr(&hd[c * ZW32 + (o >> 5) + 1], b >> (32 - (o & 31)));
                                }
                            }
                        }
                    }
                }
                bar_sync(bar_id, tpg);
                u32 bad = 0;
                for (int idx = i; idx < p.brows * ZW32; idx += tpg) {
                    const int r = idx / ZW32, w = idx - r * ZW32;
                    const int nb = min(32, Z - 32 * w);
                    u32 acc = p.syn ? synl[idx] : 0u;
                    const QcLayer ly = ltab[r];
                    for (int e = ly.edge_begin; e < ly.edge_begin + ly.degree; ++e) {
                        const QcEdgeAux ax = atab[e];
                        int start = 32 * w + ax.shift;
                        if (start >= Z) start -= Z;
                        acc ^= rotated_bits(hd + ax.col * ZW32, Z, ZW32, start, nb);
                    }
                    bad |= nb >= 32 ? acc : (acc & ((1u << nb) - 1u));
                }
                conv = !bar_red_or(bar_id, tpg, bad != 0u);
                if (conv || it >= p.max_iter) break;
            }
        }

        // ---- outputs: MSB-first packed hard decisions of the first out_cols block columns
        uint32_t *of = p.out + (size_t)f * p.out_words;
        if ((Z & 31) == 0) {
            for (int j = i; j < p.out_words; j += tpg) of[j] = __brev(hd[j]);
        } else {
            const int nbits = p.out_cols * Z;
            for (int j = i; j < p.out_words; j += tpg) {
                u32 v = 0;
                for (int b = 0; b < 32; ++b) {
                    const int gbit = 32 * j + b;
                    if (gbit < nbits) {
                        const int c = gbit / Z, l = gbit - c * Z;
                        v |= ((hd[c * ZW32 + (l >> 5)] >> (l & 31)) & 1u) << (31 - b);
                    }
                }
                of[j] = v;
            }
        }
        if (i == 0) {
            if (p.ok) p.ok[f] = conv ? 1 : 0;
            if (p.iters) p.iters[f] = (uint16_t)it;
            if (p.stats) {
                atomicAdd(&p.stats->frames, 1ull);
                if (!conv) atomicAdd(&p.stats->failures, 1ull);
                atomicAdd(&p.stats->iter_sum, (unsigned long long)it);
                atomicAdd(&p.stats->hist[min(it, QLDPC_ITER_HIST_BINS - 1)], 1ull);
            }
        }
        bar_sync(bar_id, tpg);   // hd / beliefs are reused by the next frame of this slot
    }
}

}  // namespace

int layered_i8_max_threads() { return kMaxBlock; }

template <int NK>
static int launch_nk(const LayeredI8Params &p, int grid, int smem_bytes, cudaStream_t st)
{
    QLDPC_CUDA(cudaFuncSetAttribute(layered_i8_kernel<NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    layered_i8_kernel<NK><<<grid, p.slots * p.tpg, smem_bytes, st>>>(p);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

int launch_layered_i8(const LayeredI8Params &p, int grid, int smem_bytes, cudaStream_t st)
{
    if (p.rule == QLDPC_RULE_OMS) return launch_nk<0>(p, grid, smem_bytes, st);
    switch (p.norm_eighths) {
    case 8: return launch_nk<8>(p, grid, smem_bytes, st);
    case 6: return launch_nk<6>(p, grid, smem_bytes, st);
    default: return launch_nk<-1>(p, grid, smem_bytes, st);
    }
}

}  // namespace qldpc

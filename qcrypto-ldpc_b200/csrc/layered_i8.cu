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
// Footprint diet so that TWO frames fit in one SM's 227 KB (BG1 Z=384: 105 KB per frame):
//   - an edge into a weight-1, shift-0 column (the NR extension parities) stores no message and
//     never rewrites the belief: belief - message is the channel LLR for ever, so the edge reads
//     the channel byte, takes part in the min-sum, and only its a-posteriori SIGN is kept (ballot
//     straight into the hard-decision vector);
//   - the messages of the four heaviest rows (19 edges each in BG1) live in registers.
#include <cuda_fp16.h>

#include "kernels.hpp"

// Experiment switches (measured on B200, BG1 Z=384, bench.py; see profiles/r1_experiments.md):
//   LOADUNROLL  unrolled frame load + L2 prefetch of the slot's next frame      +1.2 %  (on)
//   EXTEARLY    extension-edge ballots before the core second pass               +0.4 %  (off: noise)
//   PAIRS       degree buckets of two instead of exact-degree code variants      -1.8 %  (off)
//   REGDEDUPE   one code copy for the register rows (messages moved through a working array) -13 % (off)
#ifndef QL_OPT_LOADUNROLL
#define QL_OPT_LOADUNROLL 1
#endif
#ifndef QL_OPT_EXTEARLY
#define QL_OPT_EXTEARLY 1
#endif
#ifndef QL_OPT_PAIRS
#define QL_OPT_PAIRS 0
#endif
#ifndef QL_OPT_REGDEDUPE
#define QL_OPT_REGDEDUPE 0
#endif

namespace qldpc {

namespace {

typedef unsigned int u32;

constexpr u32 kSignMask = 0x80008000u;
constexpr u32 kInf2 = 0x03ff03ffu;     // 1023 ulp: larger than any message magnitude
constexpr u32 kOne2 = 0x3c003c00u;     // 1.0h, 1.0h
constexpr int kMaxBlock = 256;
#ifndef QL_MAXBLOCK_STREAM
#define QL_MAXBLOCK_STREAM 384
#endif
constexpr int kMaxBlockStream = QL_MAXBLOCK_STREAM;
constexpr int kRegRows = 4;

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
// clip to [-(msg_max+1), msg_max]; cap = msg_max+1, hi = msg_max. The lower bound is relu(u+cap)-cap on the
// FMA pipe (HFMA2.RELU + HADD2) so that only the upper bound occupies the half-rate ALU pipe (HMNMX2)
#ifndef QL_OPT_FMACLIP
#define QL_OPT_FMACLIP 0   // measured: 41.5 vs 41.7 Gbit/s, no gain -> keep the shorter sequence
#endif
// relu(a + b)
__device__ __forceinline__ u32 hadd_relu(u32 a, u32 b) { return bits(__hfma2_relu(h2(a), h2(kOne2), h2(b))); }
__device__ __forceinline__ u32 clip_msg(u32 u, u32 lo, u32 hi, u32 cap)
{
#if QL_OPT_FMACLIP
    (void)lo;
    return hmin(hsub(hadd_relu(u, cap), cap), hi);
#else
    (void)cap;
    return hmin(hmax(u, lo), hi);
#endif
}
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

__device__ __forceinline__ u32 vcomp(const uint4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }
__device__ __forceinline__ void vset(uint4 &v, int k, u32 x)
{
    if (k == 0) v.x = x; else if (k == 1) v.y = x; else if (k == 2) v.z = x; else v.w = x;
}

struct LayerCtx {
    u32 cLo, cHi;         // message clip [-(msg_max+1), msg_max] as half2 ulps
    u32 cM2cap;           // msg_max+1: second minimum of a degree-1 check
    u32 c128, c255;
    int norm_eighths;
    u32 negOff;           // -offset
    u32 *hd;              // this slot's hard-decision vectors
    int wis, lane, wq;    // warp in slot, lane, W/32
};

// One layer (block row) for the four check lanes of this thread.
//   NK    rule: 0 = offset min-sum, 1..8 = normalised by NK/8, -1 = normalised, factor at run time
//   DC    unrolled slots of edges with a stored message; MODE 0: exactly DC, 1: DC-1 or DC
//   EXT   one more edge follows into a weight-1 shift-0 column (no stored message, no belief update)
//   RM    where the stored messages live: 0 shared memory (Rrow, W words apart), 1 the register array Rr,
//         2 streamed: this thread's 16-byte blocks in the shared ring (ring_me, filled by cp.async from the
//         L2-resident scratch) in, its blocks in the scratch (rg_me) out
//   Li    this thread's belief base address (slot beliefs + 4*i)
//   Rrow  this thread's message word of the row's first stored edge; consecutive edges are W words apart
//   etab  two int4 per edge: {off0, off1, thresh, hdw} {selA0, selA1, selW0, selW1}
template <int NK, int DC, int MODE, bool EXT, int RM, int RDC>
__device__ __forceinline__ void process_layer(const LayerCtx &cx, char *Li, u32 *Rrow, u32 (&Rr)[RDC], int W,
                                              const int4 *etab, int nc, int i, uint2 m1init,
                                              const uint4 *ring_me = nullptr, uint4 *rg_me = nullptr)
{
    // register rows recompute the belief address / pack selector in the second pass (register budget)
    constexpr bool REG = RM == 1 || (RM == 2 && DC > 10);   // recompute instead of keeping per-edge state live
    constexpr int KEEP = REG ? 1 : DC;
    constexpr int NV = RM == 2 ? (DC + 3) / 4 : 1;
    uint4 Yv[NV];
    if constexpr (RM == 2) {
#pragma unroll
        for (int q = 0; q < NV; ++q) Yv[q] = ring_me[q];
    }
    u32 uA[DC], uB[DC], tA[KEEP], tB[KEEP], sw[KEEP];
    char *ad[KEEP];
    // running sign product starts at the syndrome bit of each lane (folded into m1init by the caller)
    u32 m1A = m1init.x, m1B = m1init.y;
    u32 m2A = kInf2, m2B = kInf2;

#pragma unroll
    for (int j = 0; j < DC; ++j) {
        if (MODE == 0 || j < DC - 1 || j < nc) {
            const int4 ed = etab[2 * j];       // off0, off1, thresh, hdw
            const int4 sl = etab[2 * j + 1];   // selA0, selA1, selW0, selW1
            const bool wrap = i >= ed.z;
            char *a = Li + (wrap ? ed.y : ed.x);
            const u32 selA = wrap ? sl.y : sl.x;
            if constexpr (!REG) { sw[j] = wrap ? sl.w : sl.z; ad[j] = a; }
            const u32 X = *reinterpret_cast<const u32 *>(a);
            u32 Y;
            if constexpr (RM == 1) Y = Rr[j];
            else if constexpr (RM == 2) Y = vcomp(Yv[j >> 2], j & 3);
            else Y = Rrow[j * W];
            const u32 xA = prmt(X, 0u, selA), xB = prmt(X, 0u, selA ^ 0x0202u);
            const u32 yA = prmt(Y, 0u, 0x4140u), yB = prmt(Y, 0u, 0x4342u);
            const u32 ua = hsub(xA, yA), ub = hsub(xB, yB);          // L - R_old  (:51)
            const u32 ta = clip_msg(ua, cx.cLo, cx.cHi, cx.cM2cap);            // clip to the message range (:54-55)
            const u32 tb = clip_msg(ub, cx.cLo, cx.cHi, cx.cM2cap);
            uA[j] = ua; uB[j] = ub;
            if constexpr (!REG) { tA[j] = ta; tB[j] = tb; }
            m2A = hmax(habs(m1A), hmin(habs(ta), m2A));               // second minimum (:61)
            m2B = hmax(habs(m1B), hmin(habs(tb), m2B));
            m1A = min_xorsign_abs(m1A, ta);                           // first minimum and sign product (:60,:63)
            m1B = min_xorsign_abs(m1B, tb);
        }
    }
    u32 ueA = 0, ueB = 0, teA = 0, teB = 0;
    int ehdw = 0;
    if constexpr (EXT) {   // the extension edge: belief - message == channel LLR (see file header)
        const int4 ed = etab[2 * nc];
        const int4 sl = etab[2 * nc + 1];
        const u32 X = *reinterpret_cast<const u32 *>(Li + ed.x);     // shift 0: never wraps
        ehdw = ed.w;
        ueA = hsub(prmt(X, 0u, sl.x), cx.c128);
        ueB = hsub(prmt(X, 0u, sl.x ^ 0x0202u), cx.c128);
        teA = clip_msg(ueA, cx.cLo, cx.cHi, cx.cM2cap);
        teB = clip_msg(ueB, cx.cLo, cx.cHi, cx.cM2cap);
        m2A = hmax(habs(m1A), hmin(habs(teA), m2A));
        m2B = hmax(habs(m1B), hmin(habs(teB), m2B));
        m1A = min_xorsign_abs(m1A, teA);
        m1B = min_xorsign_abs(m1B, teB);
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

#if QL_OPT_EXTEARLY
    if constexpr (EXT) {   // only the sign of channel + new message is observable: ballot it into the hd vector
        const u32 eA = heq_mask(habs(teA), min1A), eB = heq_mask(habs(teB), min1B);
        const u32 rA = ((eA & c1A) | (~eA & c2A)) ^ (teA & kSignMask);
        const u32 rB = ((eB & c1B) | (~eB & c2B)) ^ (teB & kSignMask);
        const u32 aA = hadd(ueA, rA), aB = hadd(ueB, rB);
        const u32 b0 = __ballot_sync(0xffffffffu, (int)(aA << 16) < 0);   // lane i
        const u32 b1 = __ballot_sync(0xffffffffu, (int)aA < 0);           // lane i + W
        const u32 b2 = __ballot_sync(0xffffffffu, (int)(aB << 16) < 0);   // lane i + 2W
        const u32 b3 = __ballot_sync(0xffffffffu, (int)aB < 0);           // lane i + 3W
        if (cx.lane == 0) {   // one lane stores the four words: no lane-dependent select
            u32 *h = cx.hd + ehdw + cx.wis;
            h[0] = b0; h[cx.wq] = b1; h[2 * cx.wq] = b2; h[3 * cx.wq] = b3;
        }
    }
#endif
#pragma unroll
    for (int j = 0; j < DC; ++j) {
        if (MODE == 0 || j < DC - 1 || j < nc) {
            u32 ta, tb;
            if constexpr (REG) {   // register rows re-clip instead of keeping t live
                ta = clip_msg(uA[j], cx.cLo, cx.cHi, cx.cM2cap);
                tb = clip_msg(uB[j], cx.cLo, cx.cHi, cx.cM2cap);
            } else {
                ta = tA[j]; tb = tB[j];
            }
            const u32 eA = heq_mask(habs(ta), min1A), eB = heq_mask(habs(tb), min1B);
            // |t| == min1 ? c1 : c2, then the edge's own sign (:73-75)
            const u32 rA = ((eA & c1A) | (~eA & c2A)) ^ (ta & kSignMask);
            const u32 rB = ((eB & c1B) | (~eB & c2B)) ^ (tb & kSignMask);
            const u32 bA = hadd(rA, cx.c128), bB = hadd(rB, cx.c128); // biased new message
            const u32 lA = hmin(hadd_relu(uA[j], bA), cx.c255);       // clip(L - R_old + R_new) biased (:88-91)
            const u32 lB = hmin(hadd_relu(uB[j], bB), cx.c255);
            const u32 Ynew = prmt(bA, bB, 0x6420u);
            if constexpr (RM == 1) Rr[j] = Ynew;
            else if constexpr (RM == 2) vset(Yv[j >> 2], j & 3, Ynew);
            else Rrow[j * W] = Ynew;
            if constexpr (REG) {
                const int4 ed = etab[2 * j];
                const int4 sl = etab[2 * j + 1];
                const bool wrap = i >= ed.z;
                *reinterpret_cast<u32 *>(Li + (wrap ? ed.y : ed.x)) = prmt(lA, lB, wrap ? sl.w : sl.z);
            } else {
                *reinterpret_cast<u32 *>(ad[j]) = prmt(lA, lB, sw[j]);
            }
        }
    }
    if constexpr (RM == 2) {
#pragma unroll
        for (int q = 0; q < NV; ++q) rg_me[q] = Yv[q];
    }
#if !QL_OPT_EXTEARLY
    if constexpr (EXT) {   // only the sign of channel + new message is observable: ballot it into the hd vector
        const u32 eA = heq_mask(habs(teA), min1A), eB = heq_mask(habs(teB), min1B);
        const u32 rA = ((eA & c1A) | (~eA & c2A)) ^ (teA & kSignMask);
        const u32 rB = ((eB & c1B) | (~eB & c2B)) ^ (teB & kSignMask);
        const u32 aA = hadd(ueA, rA), aB = hadd(ueB, rB);
        const u32 b0 = __ballot_sync(0xffffffffu, (int)(aA << 16) < 0);   // lane i
        const u32 b1 = __ballot_sync(0xffffffffu, (int)aA < 0);           // lane i + W
        const u32 b2 = __ballot_sync(0xffffffffu, (int)(aB << 16) < 0);   // lane i + 2W
        const u32 b3 = __ballot_sync(0xffffffffu, (int)aB < 0);           // lane i + 3W
        if (cx.lane == 0) {   // one lane stores the four words: no lane-dependent select
            u32 *h = cx.hd + ehdw + cx.wis;
            h[0] = b0; h[cx.wq] = b1; h[2 * cx.wq] = b2; h[3 * cx.wq] = b3;
        }
    }
#endif
}

// BIG: rows with more than 10 stored edges exist (never together with register rows: the host
// only enables those when every other row has at most 10 stored edges)
template <int NK, bool EXT, bool BIG, int RM>
__device__ __forceinline__ void dispatch_layer(const LayerCtx &cx, char *Li, u32 *Rrow, int W, const int4 *et, int nc,
                                               int i, uint2 m1init, const uint4 *ring_me = nullptr, uint4 *rg_me = nullptr)
{
    u32 dummy[1];
#define QL_CASE(DCV, MODEV) process_layer<NK, DCV, MODEV, EXT, RM, 1>(cx, Li, Rrow, dummy, W, et, nc, i, m1init, ring_me, rg_me)
#if QL_OPT_PAIRS
    // buckets of two (last slot optional): few code variants keep the instruction working set small
    switch ((nc + 1) >> 1) {
    case 1: QL_CASE(2, 1); break;
    case 2: QL_CASE(4, 1); break;
    case 3: QL_CASE(6, 1); break;
    case 4: QL_CASE(8, 1); break;
    case 5: QL_CASE(10, 1); break;
#else
    switch (nc) {
    case 1: QL_CASE(1, 0); break;
    case 2: QL_CASE(2, 0); break;
    case 3: QL_CASE(3, 0); break;
    case 4: QL_CASE(4, 0); break;
    case 5: QL_CASE(5, 0); break;
    case 6: QL_CASE(6, 0); break;
    case 7: QL_CASE(7, 0); break;
    case 8: QL_CASE(8, 0); break;
    case 9: QL_CASE(9, 0); break;
    case 10: QL_CASE(10, 0); break;
#endif
    default:
        if constexpr (BIG) {
            if (nc <= 12) QL_CASE(12, 1);
            else if (nc <= 14) QL_CASE(14, 1);
            else if (nc <= 16) QL_CASE(16, 1);
            else if (nc <= 18) QL_CASE(18, 1);
            else QL_CASE(20, 1);
        }
        break;
    }
#undef QL_CASE
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

// 4x4 byte transpose: in[k] holds lanes 4j..4j+3 of quarter k; out[m] = belief word of lane 4j+m
__device__ __forceinline__ void transpose4x4(const u32 (&in)[4], u32 (&out)[4])
{
    const u32 t0 = prmt(in[0], in[1], 0x5140u), t1 = prmt(in[2], in[3], 0x5140u);   // bytes 0,1 of the four quarters
    const u32 t2 = prmt(in[0], in[1], 0x7362u), t3 = prmt(in[2], in[3], 0x7362u);   // bytes 2,3
    out[0] = prmt(t0, t1, 0x5410u);
    out[1] = prmt(t0, t1, 0x7632u);
    out[2] = prmt(t2, t3, 0x5410u);
    out[3] = prmt(t2, t3, 0x7632u);
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// REGDC: 0, or the unrolled degree of the kRegRows rows whose messages live in registers,
//        or -1: streamed mode (messages in an L2-resident scratch, 2-deep cp.async ring, up to 384 threads)
template <int NK, int REGDC>
__global__ void __launch_bounds__(REGDC < 0 ? kMaxBlockStream : kMaxBlock, 1) layered_i8_kernel(const LayeredI8Params p)
{
    extern __shared__ __align__(16) char smem[];
    const int tpg = p.tpg, W = p.W, Z = p.Z, ZW32 = p.ZW32;
    const int g = threadIdx.y, i = threadIdx.x;          // block = (tpg, slots); tpg is a multiple of 32
    const int lane = i & 31, wis = i >> 5;
    const int tid = g * tpg + i, nthreads = tpg * blockDim.y;
    const int bar_id = 1 + g;
    const bool active = i < W;
    const bool aligned = (W & 31) == 0;      // every ballot lands on whole words
    const bool fast_bits = aligned && (Z & 31) == 0;

    // ---- shared tables (all slots)
    int4 *etab = reinterpret_cast<int4 *>(smem);   // two int4 per edge
    QcEdgeAux *atab = reinterpret_cast<QcEdgeAux *>(etab + 2 * p.nnz);
    Li8Layer *ltab = reinterpret_cast<Li8Layer *>(atab + p.nnz);
    uint16_t *pcols = reinterpret_cast<uint16_t *>(ltab + p.brows);
    for (int e = tid; e < 2 * p.nnz; e += nthreads) etab[e] = reinterpret_cast<const int4 *>(p.edges)[e];
    for (int e = tid; e < p.nnz; e += nthreads) atab[e] = p.aux[e];
    for (int r = tid; r < p.brows; r += nthreads) ltab[r] = p.layers[r];
    for (int c = tid; c < p.n_pack; c += nthreads) pcols[c] = p.pack_cols[c];
    // 8 x 16 bytes of biased zero messages, 16-byte aligned, right after the tables
    uint4 *zero_blk = reinterpret_cast<uint4 *>(smem + ((reinterpret_cast<char *>(pcols + p.bcols) - smem + 15) & ~15));
    if (tid < 8) zero_blk[tid] = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
    __syncthreads();

    char *slot = smem + p.tab_bytes + (size_t)g * p.slot_bytes;
    u32 *Lw = reinterpret_cast<u32 *>(slot);
    u32 *Rw = reinterpret_cast<u32 *>(slot + p.off_R);
    u32 *hd = reinterpret_cast<u32 *>(slot + p.off_hd);
    u32 *synl = reinterpret_cast<u32 *>(slot + p.off_syn);

    LayerCtx cx;
    cx.cLo = p.h2_lo;           // half2 constants prepared by the host: they stay constant-bank operands
    cx.cHi = p.h2_hi;
    cx.cM2cap = p.h2_cap;
    cx.c128 = 0x00800080u;
    cx.c255 = 0x00ff00ffu;
    cx.norm_eighths = p.norm_eighths;
    cx.negOff = p.h2_negoff;
    cx.hd = hd;
    cx.wis = wis;
    cx.lane = lane;
    cx.wq = W >> 5;

    constexpr bool STREAM = REGDC < 0;
    constexpr int RDC = REGDC > 0 ? REGDC : 1;
    u32 Rreg0[RDC], Rreg1[RDC], Rreg2[RDC], Rreg3[RDC];
    uint4 *ring = reinterpret_cast<uint4 *>(slot + p.off_R);                       // 2 stages of stage_words
    const u32 rg_off = STREAM ? (blockIdx.x * p.slots + g) * (u32)p.rg_words : 0u;   // word offset of this slot's scratch
    const int stage_v = p.stage_words >> 2;

    for (int f = blockIdx.x * p.slots + g; f < p.F; f += gridDim.x * p.slots) {
        // ---- load: int8 LLRs -> interleaved biased belief words; messages = 0
        const int8_t *src = p.llr + (size_t)f * p.N;
        if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
            // 4 aligned 32-bit loads (one per quarter of the column) -> 4x4 byte transpose -> one 128-bit store
            const int wq4 = W >> 2;
#if QL_OPT_LOADUNROLL
            {   // pull this slot's next frame towards L2 while the current one is decoded
                const int fn = f + gridDim.x * p.slots;
                if (fn < p.F) {
                    const char *nx = reinterpret_cast<const char *>(p.llr + (size_t)fn * p.N);
                    for (int o = i * 128; o < p.N; o += tpg * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
                }
            }
#pragma unroll 6
#endif
            for (int it = i; it < p.bcols * wq4; it += tpg) {
                const int c = it / wq4, j = it - c * wq4;
                const u32 *q = reinterpret_cast<const u32 *>(src + c * Z) + j;
                u32 in[4], out[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) in[k] = __ldg(q + k * wq4) ^ 0x80808080u;
                transpose4x4(in, out);
                *reinterpret_cast<uint4 *>(Lw + c * W + 4 * j) = make_uint4(out[0], out[1], out[2], out[3]);
            }
        } else if (active) {
            for (int c = 0; c < p.bcols; ++c) {
                const uint8_t *q = reinterpret_cast<const uint8_t *>(src) + c * Z + i;
                const u32 b0 = __ldg(q), b1 = __ldg(q + W), b2 = __ldg(q + 2 * W), b3 = __ldg(q + 3 * W);
                Lw[c * W + i] = (b0 | (b1 << 8) | (b2 << 16) | (b3 << 24)) ^ 0x80808080u;
            }
        }
        if constexpr (STREAM) {
            cp_async_wait<0>();   // nothing may still be in flight from the previous frame of this slot
        } else if ((W & 3) == 0) {
            for (int idx = i; idx < p.n_store * (W >> 2); idx += tpg)
                reinterpret_cast<uint4 *>(Rw)[idx] = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
        } else {
            for (int idx = i; idx < p.n_store * W; idx += tpg) Rw[idx] = 0x80808080u;
        }
        if constexpr (REGDC > 0) {
#pragma unroll
            for (int j = 0; j < RDC; ++j) Rreg0[j] = Rreg1[j] = Rreg2[j] = Rreg3[j] = 0x80808080u;
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
        int stage = 0;
        bool conv = false;
        for (;;) {
            if (it < p.max_iter) {
                for (int r = 0; r < p.brows; ++r) {
                    const Li8Layer ly = ltab[r];
                    if (active) {
                        uint2 synbits = make_uint2(kInf2, kInf2);   // initial (min1, sign) of the two half2 pairs
                        if (p.syn) {
                            u32 sb = 0;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int l = i + W * k;
                                sb |= ((synl[r * ZW32 + (l >> 5)] >> (l & 31)) & 1u) << k;
                            }
                            synbits.x ^= ((sb & 1u) << 15) | ((sb & 2u) << 30);
                            synbits.y ^= ((sb & 4u) << 13) | ((sb & 8u) << 28);
                        }
                        char *Li = slot + 4 * i;
                        const int4 *et = etab + 2 * ly.edge_begin;
                        if constexpr (STREAM) {
                            // stage the NEXT row's messages while this row is processed; during the first
                            // iteration every row reads a shared block of zero messages instead
                            const int rn = r + 1 < p.brows ? r + 1 : 0;
                            if (it > 0 || rn == 0) {
                                const Li8Layer nx = ltab[rn];
                                const int nvn = nx.st >> 2;
                                uint4 *dst = ring + (stage ^ 1) * stage_v + i * nvn;
                                const uint4 *src = reinterpret_cast<const uint4 *>(p.rg + (rg_off + nx.g_off + i * nx.st));
#pragma unroll
                                for (int q = 0; q < 5; ++q)
                                    if (q < nvn) cp_async16(dst + q, src + q);
                            }
                            cp_async_commit();
                            cp_async_wait<1>();                  // this row's stage has landed
                            const int nv = ly.st >> 2;
                            const uint4 *ring_me = it == 0 ? zero_blk : ring + stage * stage_v + i * nv;
                            uint4 *rg_me = reinterpret_cast<uint4 *>(p.rg + (rg_off + ly.g_off + i * ly.st));
                            if (ly.has_ext) dispatch_layer<NK, true, true, 2>(cx, Li, nullptr, W, et, ly.n_core, i, synbits, ring_me, rg_me);
                            else dispatch_layer<NK, false, true, 2>(cx, Li, nullptr, W, et, ly.n_core, i, synbits, ring_me, rg_me);
                        } else if (REGDC > 0 && ly.reg_idx >= 0) {
                            if constexpr (REGDC > 0) {
#if QL_OPT_REGDEDUPE
                                // one code copy for all register rows: move the row's messages through a working array
                                u32 Rc[RDC];
                                const int ri = ly.reg_idx;
#pragma unroll
                                for (int j = 0; j < RDC; ++j) Rc[j] = ri == 0 ? Rreg0[j] : (ri == 1 ? Rreg1[j] : (ri == 2 ? Rreg2[j] : Rreg3[j]));
                                process_layer<NK, RDC, 1, false, 1, RDC>(cx, Li, nullptr, Rc, W, et, ly.n_core, i, synbits);
#pragma unroll
                                for (int j = 0; j < RDC; ++j) {
                                    if (ri == 0) Rreg0[j] = Rc[j];
                                    else if (ri == 1) Rreg1[j] = Rc[j];
                                    else if (ri == 2) Rreg2[j] = Rc[j];
                                    else Rreg3[j] = Rc[j];
                                }
#else
                                switch (ly.reg_idx) {
                                case 0: process_layer<NK, RDC, 1, false, 1, RDC>(cx, Li, nullptr, Rreg0, W, et, ly.n_core, i, synbits); break;
                                case 1: process_layer<NK, RDC, 1, false, 1, RDC>(cx, Li, nullptr, Rreg1, W, et, ly.n_core, i, synbits); break;
                                case 2: process_layer<NK, RDC, 1, false, 1, RDC>(cx, Li, nullptr, Rreg2, W, et, ly.n_core, i, synbits); break;
                                default: process_layer<NK, RDC, 1, false, 1, RDC>(cx, Li, nullptr, Rreg3, W, et, ly.n_core, i, synbits); break;
                                }
#endif
                            }
                        } else {
                            u32 *Rrow = Rw + ly.r_off * W + i;
                            if (ly.has_ext) dispatch_layer<NK, true, REGDC == 0, 0>(cx, Li, Rrow, W, et, ly.n_core, i, synbits);
                            else dispatch_layer<NK, false, REGDC == 0, 0>(cx, Li, Rrow, W, et, ly.n_core, i, synbits);
                        }
                    }
                    if constexpr (STREAM) stage ^= 1;
                    bar_sync(bar_id, tpg);
                }
                ++it;
            }
            // the syndrome / hard-decision phase runs after every iteration when early stop is on,
            // otherwise once after the last iteration
            if (p.early_stop || it >= p.max_iter) {
                u32 bad = 0;
                if (fast_bits) {
                    // hard decisions of the columns that are not extension columns (those were balloted in the layers)
                    for (int c = 0; c < p.n_pack; ++c) {
                        const int col = pcols[c];
                        const u32 X = Lw[col * W + i];                 // biased: bit 7 clear <=> L < 0
                        const u32 b0 = __ballot_sync(0xffffffffu, (X & 0x80u) == 0u);
                        const u32 b1 = __ballot_sync(0xffffffffu, (X & 0x8000u) == 0u);
                        const u32 b2 = __ballot_sync(0xffffffffu, (X & 0x800000u) == 0u);
                        const u32 b3 = __ballot_sync(0xffffffffu, (int)X >= 0);
                        if (lane == 0) {
                            u32 *h = hd + col * ZW32 + wis;
                            h[0] = b0; h[cx.wq] = b1; h[2 * cx.wq] = b2; h[3 * cx.wq] = b3;
                        }
                    }
                    bar_sync(bar_id, tpg);
                    // syndrome words: thread -> (row r0 + i / ZW32, word i % ZW32)
                    const int rsub = i / ZW32, w = i - rsub * ZW32, rstep = tpg / ZW32;
                    if (rsub < rstep) {
                        for (int r = rsub; r < p.brows; r += rstep) {
                            const Li8Layer ly = ltab[r];
                            u32 acc = p.syn ? synl[r * ZW32 + w] : 0u;
                            for (int e = ly.edge_begin; e < ly.edge_begin + ly.degree; ++e) {
                                const QcEdgeAux ax = atab[e];
                                int w0 = w + (ax.shift >> 5);
                                if (w0 >= ZW32) w0 -= ZW32;
                                const int w1 = (w0 + 1 == ZW32) ? 0 : w0 + 1;
                                acc ^= __funnelshift_r(hd[ax.hdw + w0], hd[ax.hdw + w1], ax.shift & 31);
                            }
                            bad |= acc;
                        }
                    }
                } else {
                    if (!aligned) {
                        for (int idx = i; idx < p.bcols * ZW32; idx += tpg) hd[idx] = 0u;
                        bar_sync(bar_id, tpg);
                    }
                    for (int c = 0; c < p.bcols; ++c) {
                        const u32 X = active ? Lw[c * W + i] : 0x80808080u;
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
                    for (int idx = i; idx < p.brows * ZW32; idx += tpg) {
                        const int r = idx / ZW32, w = idx - r * ZW32;
                        const int nb = min(32, Z - 32 * w);
                        u32 acc = p.syn ? synl[idx] : 0u;
                        const Li8Layer ly = ltab[r];
                        for (int e = ly.edge_begin; e < ly.edge_begin + ly.degree; ++e) {
                            const QcEdgeAux ax = atab[e];
                            int start = 32 * w + ax.shift;
                            if (start >= Z) start -= Z;
                            acc ^= rotated_bits(hd + ax.hdw, Z, ZW32, start, nb);
                        }
                        bad |= nb >= 32 ? acc : (acc & ((1u << nb) - 1u));
                    }
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
int layered_i8_max_threads_stream() { return kMaxBlockStream; }

template <int NK, int REGDC>
static int launch_nk(const LayeredI8Params &p, int grid, int smem_bytes, cudaStream_t st)
{
    QLDPC_CUDA(cudaFuncSetAttribute(layered_i8_kernel<NK, REGDC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    layered_i8_kernel<NK, REGDC><<<grid, dim3(p.tpg, p.slots), smem_bytes, st>>>(p);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

template <int REGDC>
static int launch_reg(const LayeredI8Params &p, int grid, int smem_bytes, cudaStream_t st)
{
#ifdef QL_ONLY_NK6   // experiment builds
    if (p.rule != QLDPC_RULE_OMS && p.norm_eighths == 6) return launch_nk<6, REGDC>(p, grid, smem_bytes, st);
    return QLDPC_ERR_UNSUPPORTED;
#else
    if (p.rule == QLDPC_RULE_OMS) return launch_nk<0, REGDC>(p, grid, smem_bytes, st);
    switch (p.norm_eighths) {
    case 8: return launch_nk<8, REGDC>(p, grid, smem_bytes, st);
    case 6: return launch_nk<6, REGDC>(p, grid, smem_bytes, st);
    default: return launch_nk<-1, REGDC>(p, grid, smem_bytes, st);
    }
#endif
}

int launch_layered_i8(const LayeredI8Params &p, int grid, int smem_bytes, cudaStream_t st)
{
    if (p.stream) return launch_reg<-1>(p, grid, smem_bytes, st);
    if (p.regdc == 20) return launch_reg<20>(p, grid, smem_bytes, st);
    if (p.regdc == 0) return launch_reg<0>(p, grid, smem_bytes, st);
    return QLDPC_ERR_UNSUPPORTED;
}

int layered_i8_reg_rows() { return kRegRows; }

}  // namespace qldpc

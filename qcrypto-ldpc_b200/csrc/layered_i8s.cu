// Layered int8 min-sum decoder, "streamed" kernel (v5): quasi-cyclic codes with Z % 128 == 0
// (BG1/BG2 lifting sizes 128, 256, 384).  Same arithmetic as layered_i8.cu -- ML/BPSK_nrldpc_sim_FP.m:35-94
// generalised by syndrome input, early stop and the shift-normalised rule; bit-exact with
// oracle/qldpc_oracle.c:ora_decode_layered_fixed -- but a different placement of the state, chosen so that
// FIVE frames (BG1 Z=384) are in flight per SM, 15 warps at 120 registers (see kMaxRegs):
//   shared memory, per frame slot:
//     beliefs of the CORE block columns only (columns that are not weight-1/shift-0 extension columns),
//       word i of a column = lanes {i, i+W, i+2W, i+3W} as biased bytes (L+128), W = Z/4;
//     a 2-stage ring of check-to-variable messages (bulk copies from / plain stores to an L2-resident scratch);
//     a 2-deep ring of raw channel LLRs of the extension column of the row being processed / staged
//       (an extension column is read by exactly one row and never rewritten: its belief minus its message is
//       the channel LLR for ever, only the SIGN of its a-posteriori value is observable);
//     hard-decision bit vectors, each stored twice back to back so that a rotated window never wraps;
//     a staging buffer into which the core LLRs of the slot's NEXT frame are bulk-copied during the decode.
//   registers: per-edge state of the row being processed.
// Per edge and thread (4 check lanes = 2 half2 pairs) the instruction budget is ~43:
//   the wrap decision (lane i+r >= W) selects between two precomputed 16-byte table entries (address select + one
//   128-bit shared load, no SEL chain); clips and the k/8 normalisation run on the FMA pipe (relu / multiply forms)
//   and so does the selection of the new message (two FMAs), so that the half-rate ALU pipe only carries PRMT / min-max /
//   logic; ALU and fp16-FMA pipe are both half rate, the target is an even split (profiles/r1_onchip_peaks.md).
// Frames: the first grid*slots are assigned statically, after that a slot draws its next frame from a per-launch counter
// (frame queue): with early termination frames take 1..max_iter iterations and a fixed assignment leaves slots idle at the end.
// All frame groups of a CTA start every iteration together, so that the 255 KB of unrolled row code are fetched once
// per SM (see the comment at the iteration barrier).
// Early termination: hard decisions are balloted into Z-bit vectors; the syndrome is evaluated word-wise
// (funnel shift + XOR) for the first 8 block rows, and for the remaining rows only if those were all zero.
// Alternatives that were measured and lost (looser / tighter group coupling, producer warp, rotating or split staging duty,
// first-iteration specialisation, mbarrier hand-over of the beliefs, compare chain before the dispatch) are logged in
// profiles/r1_experiments.md; this file carries the winning path only.
// Bit input (p.bits != null, qldpc_decode_bits): the frame is given as packed sifted-key bits plus one per-position
// magnitude table shared by all frames; the slot synthesises its int8 LLRs itself at the frame switch (core columns
// straight into belief words, extension columns into a per-slot scratch that the row staging then bulk-copies from), so
// that no LLR array (N bytes per frame, written and read back through HBM by a separate kernel) exists at all.
#include <cuda_fp16.h>

#include "kernels.hpp"

namespace qldpc {

namespace {

typedef unsigned int u32;

constexpr u32 kSignMask = 0x80008000u;
constexpr u32 kInf2 = 0x03ff03ffu;     // 1023 ulp: larger than any message magnitude
constexpr u32 kOne2 = 0x3c003c00u;     // 1.0h, 1.0h
constexpr u32 kMinusOne2 = 0xbc00bc00u;
constexpr u32 k128 = 0x00800080u;
constexpr u32 k255 = 0x00ff00ffu;
constexpr int kLoadUnroll = 4;       // column steps of the frame load in flight per thread (BG1: 26 core columns = 7 steps)
constexpr int kMaxThreads = 480;     // 5 frames x 96 threads = 15 warps (18 warps: 5 on one sub-partition -> 96 registers, spills)
// 120 registers, not the 128 that 15 warps would allow: the warps sit 4/4/4/3 on the four SM sub-partitions, and at 128 the
// three full ones have none left, so no other CTA can start on the SM.  At 120 each keeps 1024 registers free = one
// 32-register warp of a co-resident helper kernel (bitops.cu: LLR synthesis of the host pipeline's next chunk).
constexpr int kMaxRegs = 120;
constexpr int kKeepT = 8;            // rows with more core edges re-clip L - R_old in the second pass instead of keeping t
constexpr int kBigKeepU = 10;        // edges of a big row (> 10 core edges) whose L - R_old stays in registers between the passes

__device__ __forceinline__ __half2 h2(u32 x) { return *reinterpret_cast<__half2 *>(&x); }
__device__ __forceinline__ u32 bits(__half2 h) { return *reinterpret_cast<u32 *>(&h); }
__device__ __forceinline__ u32 hsub(u32 a, u32 b) { return bits(__hsub2(h2(a), h2(b))); }
__device__ __forceinline__ u32 hadd(u32 a, u32 b) { return bits(__hadd2(h2(a), h2(b))); }
__device__ __forceinline__ u32 hmin(u32 a, u32 b) { return bits(__hmin2(h2(a), h2(b))); }
__device__ __forceinline__ u32 hmax(u32 a, u32 b) { return bits(__hmax2(h2(a), h2(b))); }
__device__ __forceinline__ u32 habs(u32 a) { return bits(__habs2(h2(a))); }
__device__ __forceinline__ u32 hadd_relu(u32 a, u32 b) { return bits(__hfma2_relu(h2(a), h2(kOne2), h2(b))); }
// relu(b - a)
__device__ __forceinline__ u32 hrsub_relu(u32 a, u32 b) { return bits(__hfma2_relu(h2(a), h2(kMinusOne2), h2(b))); }
__device__ __forceinline__ u32 heq_one(u32 a, u32 b) { return bits(__heq2(h2(a), h2(b))); }      // 1.0h where equal, else 0
__device__ __forceinline__ u32 hfma(u32 a, u32 b, u32 c) { return bits(__hfma2(h2(a), h2(b), h2(c))); }
__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 sel)
{
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ u32 min_xorsign_abs(u32 a, u32 b)
{
    u32 d;
    asm("min.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// one edge into the running (min1 with sign product, min2) pair
template <bool ADDFORM>
__device__ __forceinline__ void two_min(u32 &m1, u32 &m2, u32 t)
{
    if (ADDFORM) {
        const u32 s = hadd(habs(m1), habs(t));     // exact: both below 2048 ulp
        m1 = min_xorsign_abs(m1, t);
        m2 = hmin(m2, hsub(s, habs(m1)));          // |m1| + |t| - min(|m1|, |t|) = max(|m1|, |t|)
    } else {
        m2 = hmax(habs(m1), hmin(habs(t), m2));
        m1 = min_xorsign_abs(m1, t);
    }
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
__device__ __forceinline__ void mbar_init(u32 mb, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 mb)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(u32 mb, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 mb, u32 parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}"
        ::"r"(mb), "r"(parity) : "memory");
}
// bulk copy global -> shared, completion counted in bytes on an mbarrier (size and addresses multiples of 16)
__device__ __forceinline__ void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 mb)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
}

// The table entry of an edge for this thread: entry 0 when lane i does not wrap past the end of the column,
// entry 1 (16 bytes further) when it does: an address select and one 128-bit shared load, no SEL chain.
struct EdgeEntry { u32 off, selA, selB, selW; };
__device__ __forceinline__ EdgeEntry load_entry(u32 saddr, int i, int thresh)
{
    const uint4 *tp = reinterpret_cast<const uint4 *>(__cvta_shared_to_generic(saddr)) + (i >= thresh ? 1 : 0);
    const uint4 v = *tp;
    EdgeEntry e;
    e.off = v.x; e.selA = v.y; e.selB = v.z; e.selW = v.w;
    return e;
}

// re-read of a message block in the second pass (volatile: must not be merged with the first-pass read)
__device__ __forceinline__ uint4 lds128_volatile(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"((u32)__cvta_generic_to_shared(p)));
    return v;
}

// clip to the message range [-(msg_max+1), msg_max]
struct Cx {
    u32 cLo, cHi, cCap, cSpan;   // -(msg_max+1), msg_max, msg_max+1, 2*msg_max+1 as half2 ulps
    u32 negOff;
    int norm_eighths;
    int lane, wis, wq, ZW32;
};
__device__ __forceinline__ u32 clip_msg(const Cx &cx, u32 u)
{
    // hi - relu(span - relu(u + cap)): three FMA-pipe instructions, none on the ALU pipe
    return hsub(cx.cHi, hrsub_relu(hadd_relu(u, cx.cCap), cx.cSpan));
}
// clip(L - R_old + R_new) in biased form: relu(u + b) capped at 255
__device__ __forceinline__ u32 clip_belief(u32 u, u32 b)
{
    const u32 l = hadd_relu(u, b);
    return hmin(l, k255);
}

// integer k/8 normalisation on two packed non-negative fields: sum of floor(x/2), floor(x/4), floor(x/8) terms
// (AFF3CT's integer normalize).  FMA form: floor(x / 2^s) = (x with its s low bits cleared) * 2^-s, exact in fp16
// for integers held as subnormals; one LOP3 on the ALU pipe per term, the multiply-adds run on the FMA pipe.
template <int NK>
__device__ __forceinline__ u32 norm_eighths2(u32 x, int k_rt)
{
    const int k = NK < 0 ? k_rt : NK;
    if (k >= 8) return x;
    if (NK > 0) {
        u32 r = 0;
        bool first = true;
        if (k & 4) { r = bits(__hmul2(h2(x & 0xfffefffeu), h2(0x38003800u))); first = false; }                      // * 0.5
        if (k & 2) { r = first ? bits(__hmul2(h2(x & 0xfffcfffcu), h2(0x34003400u)))
                               : bits(__hfma2(h2(x & 0xfffcfffcu), h2(0x34003400u), h2(r))); first = false; }     // * 0.25
        if (k & 1) { r = first ? bits(__hmul2(h2(x & 0xfff8fff8u), h2(0x30003000u)))
                               : bits(__hfma2(h2(x & 0xfff8fff8u), h2(0x30003000u), h2(r))); }                    // * 0.125
        return r;
    }
    const u32 s1 = (x >> 1) & 0x7fff7fffu, s2 = (x >> 2) & 0x3fff3fffu, s3 = (x >> 3) & 0x1fff1fffu;
    u32 r = 0;
    if (k & 4) r += s1;
    if (k & 2) r += s2;
    if (k & 1) r += s3;
    return r;
}

__device__ __forceinline__ u32 vcomp(const uint4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }
__device__ __forceinline__ int vcomp(const int4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }
__device__ __forceinline__ void vset(uint4 &v, int k, u32 x)
{
    if (k == 0) v.x = x; else if (k == 1) v.y = x; else if (k == 2) v.z = x; else v.w = x;
}

// One block row for the four check lanes of this thread.
//   NK     rule: 0 = offset min-sum, 1..8 = normalised by NK/8, -1 = normalised, factor at run time
//   DC     unrolled core-edge slots; EXACT: the row has exactly DC core edges, else DC - 1 or DC (nc)
//   EXT    one more edge goes to an extension column (channel bytes in extb, sign balloted into hd_ext)
//   MODE   what stays in registers per edge between the two passes: 0: L - R_old, its clipped value, belief address
//          and pack selector; 1: no clipped value (re-clip); 3 (rows with more than 10 core edges): L - R_old for the
//          first kBigKeepU edges, only the belief word for the others (re-read the table entry and the old message,
//          redo the subtraction and the clip)
//   Li     this thread's belief base (slot beliefs + 4*i); erow: shared address of the row's first table entry
//   thr4   wrap thresholds of the row's edges, four per int4
//   ysrc   this thread's message blocks (ystride uint4 apart); gdst: where the new ones go (W uint4 apart)
template <int NK, int DC, bool EXACT, bool EXT, int MODE>
__device__ __forceinline__ void process_row(const Cx &cx, char *Li, u32 erow, const int4 *thr4, int nc, int i, uint2 m1init,
                                            const uint4 *ysrc, int ystride, uint4 *gdst, int W,
                                            const unsigned char *extb, u32 *hd_ext)
{
    constexpr int KU = MODE <= 2 ? DC : (kBigKeepU < DC ? kBigKeepU : DC);
    constexpr int KEEPT = MODE == 0 ? DC : 1, KEEPA = MODE <= 1 ? DC : 1, KEEPU = KU > 0 ? KU : 1, KEEPX = MODE == 3 ? DC : 1;
    u32 uA[KEEPU], uB[KEEPU], tA[KEEPT], tB[KEEPT], sw[KEEPA], xk[KEEPX];
    char *ad[KEEPA];
    u32 m1A = m1init.x, m1B = m1init.y;   // running (min1, sign product); the sign starts at the syndrome bit
    u32 m2A = kInf2, m2B = kInf2;

    uint4 Yq = make_uint4(0, 0, 0, 0);
    int4 Tq = make_int4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < DC; ++j) {
        if (EXACT || j < DC - 1 || j < nc) {
            if ((j & 3) == 0) {
                Yq = ysrc[(j >> 2) * ystride];
                Tq = thr4[j >> 2];
            }
            const EdgeEntry en = load_entry(erow + 32 * j, i, vcomp(Tq, j & 3));
            char *a = Li + en.off;
            const u32 X = *reinterpret_cast<const u32 *>(a);
            const u32 xA = prmt(X, 0u, en.selA), xB = prmt(X, 0u, en.selB);
            const u32 Y = vcomp(Yq, j & 3);                            // biased old messages (0x80 = zero)
            const u32 yA = prmt(Y, 0u, 0x4140u), yB = prmt(Y, 0u, 0x4342u);
            const u32 ua = hsub(xA, yA), ub = hsub(xB, yB);          // L - R_old  (:51)
            const u32 ta = clip_msg(cx, ua), tb = clip_msg(cx, ub);   // clip to the message range (:54-55)
            if (j < KU) { uA[j] = ua; uB[j] = ub; } else { xk[j] = X; }
            if constexpr (MODE == 0) { tA[j] = ta; tB[j] = tb; }
            if constexpr (MODE <= 1) { sw[j] = en.selW; ad[j] = a; }
            two_min<true>(m1A, m2A, ta);               // first minimum and sign product (:60,:63), second minimum (:61)
            two_min<false>(m1B, m2B, tb);
        }
    }
    u32 ueA = 0, ueB = 0, teA = 0, teB = 0;
    if constexpr (EXT) {   // the extension edge: belief - message == channel LLR
        const u32 b0 = extb[0], b1 = extb[W], b2 = extb[2 * W], b3 = extb[3 * W];
        ueA = hsub((b0 | (b1 << 16)) ^ k128, k128);
        ueB = hsub((b2 | (b3 << 16)) ^ k128, k128);
        teA = clip_msg(cx, ueA);
        teB = clip_msg(cx, ueB);
        two_min<true>(m1A, m2A, teA);
        two_min<false>(m1B, m2B, teB);
    }
    const u32 parA = m1A & kSignMask, parB = m1B & kSignMask;
    const u32 min1A = habs(m1A), min1B = habs(m1B);
    m2A = hmin(m2A, cx.cCap);
    m2B = hmin(m2B, cx.cCap);
    u32 c1A, c1B, c2A, c2B;   // c1: magnitude sent to the position of the minimum, c2: to all others
    if (NK == 0) {                                                    // :65-72
        c1A = hadd_relu(m2A, cx.negOff); c1B = hadd_relu(m2B, cx.negOff);
        c2A = hadd_relu(min1A, cx.negOff); c2B = hadd_relu(min1B, cx.negOff);
    } else {
        c1A = norm_eighths2<NK>(m2A, cx.norm_eighths); c1B = norm_eighths2<NK>(m2B, cx.norm_eighths);
        c2A = norm_eighths2<NK>(min1A, cx.norm_eighths); c2B = norm_eighths2<NK>(min1B, cx.norm_eighths);
    }
    // magnitude c2 + [|t| == min1] * (c1 - c2), sign sigma = (row parity ^ sign(t)) as +-1.0h: all exact on subnormal integers
    const u32 dcA = hsub(c1A, c2A), dcB = hsub(c1B, c2B);
    const u32 sgA = parA ^ kOne2, sgB = parB ^ kOne2;

    if constexpr (EXT) {   // only the sign of channel + new message is observable: ballot it into the hd vector
        const u32 aA = hfma((teA & kSignMask) ^ sgA, hfma(heq_one(habs(teA), min1A), dcA, c2A), ueA);
        const u32 aB = hfma((teB & kSignMask) ^ sgB, hfma(heq_one(habs(teB), min1B), dcB, c2B), ueB);
        const u32 b0 = __ballot_sync(0xffffffffu, (int)(aA << 16) < 0);   // lane i
        const u32 b1 = __ballot_sync(0xffffffffu, (int)aA < 0);           // lane i + W
        const u32 b2 = __ballot_sync(0xffffffffu, (int)(aB << 16) < 0);   // lane i + 2W
        const u32 b3 = __ballot_sync(0xffffffffu, (int)aB < 0);           // lane i + 3W
        if (cx.lane == 0) {
            u32 *h = hd_ext + cx.wis;
            const int wq = cx.wq;   // shift 0: the syndrome phase never reads the second copy of this vector
            h[0] = b0; h[wq] = b1; h[2 * wq] = b2; h[3 * wq] = b3;
        }
    }
    uint4 Yn = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
#pragma unroll
    for (int j = 0; j < DC; ++j) {
        if (EXACT || j < DC - 1 || j < nc) {
            u32 ta, tb, selW, ua, ub;
            char *a;
            if constexpr (MODE >= 2) {
                if ((j & 3) == 0) Tq = thr4[j >> 2];
                const EdgeEntry en = load_entry(erow + 32 * j, i, vcomp(Tq, j & 3));
                a = Li + en.off;
                selW = en.selW;
                if (j >= KU) {
                    if ((j & 3) == 0 || j == KU) Yq = lds128_volatile(ysrc + (j >> 2) * ystride);
                    const u32 Y = vcomp(Yq, j & 3);
                    const u32 yA = prmt(Y, 0u, 0x4140u), yB = prmt(Y, 0u, 0x4342u);
                    ua = hsub(prmt(xk[j], 0u, en.selA), yA);
                    ub = hsub(prmt(xk[j], 0u, en.selB), yB);
                } else {
                    ua = uA[j]; ub = uB[j];
                }
            } else {
                selW = sw[j]; a = ad[j]; ua = uA[j]; ub = uB[j];
            }
            if constexpr (MODE == 0) {
                ta = tA[j]; tb = tB[j];
            } else {
                ta = clip_msg(cx, ua);
                tb = clip_msg(cx, ub);
            }
            // |t| == min1 ? c1 : c2, then the edge's own sign (:73-75)
            const u32 bA = hfma((ta & kSignMask) ^ sgA, hfma(heq_one(habs(ta), min1A), dcA, c2A), k128);   // biased new message
            const u32 bB = hfma((tb & kSignMask) ^ sgB, hfma(heq_one(habs(tb), min1B), dcB, c2B), k128);
            const u32 lA = clip_belief(ua, bA);                       // clip(L - R_old + R_new) biased (:88-91)
            const u32 lB = clip_belief(ub, bB);
            vset(Yn, j & 3, prmt(bA, bB, 0x6420u));
            *reinterpret_cast<u32 *>(a) = prmt(lA, lB, selW);
        }
        if ((j & 3) == 3 || j == DC - 1) gdst[(j >> 2) * W] = Yn;
    }
}

template <int NK>
__device__ __forceinline__ void dispatch_row(int variant, const Cx &cx, char *Li, u32 erow, const int4 *thr4, int nc, int i,
                                             uint2 m1init, const uint4 *ysrc, int ystride, uint4 *gdst, int W,
                                             const unsigned char *extb, u32 *hd_ext)
{
#define QL_ROW(DCV, EXACTV, EXTV, BIGV) \
    process_row<NK, DCV, EXACTV, EXTV, (BIGV ? 3 : (DCV > kKeepT ? 1 : 0))>(cx, Li, erow, thr4, nc, i, m1init, ysrc, ystride, gdst, W, \
                                                                            extb, hd_ext)
    switch (variant) {
    case 0: QL_ROW(1, true, false, false); break;
    case 1: QL_ROW(1, true, true, false); break;
    case 2: QL_ROW(2, true, false, false); break;
    case 3: QL_ROW(2, true, true, false); break;
    case 4: QL_ROW(3, true, false, false); break;
    case 5: QL_ROW(3, true, true, false); break;
    case 6: QL_ROW(4, true, false, false); break;
    case 7: QL_ROW(4, true, true, false); break;
    case 8: QL_ROW(5, true, false, false); break;
    case 9: QL_ROW(5, true, true, false); break;
    case 10: QL_ROW(6, true, false, false); break;
    case 11: QL_ROW(6, true, true, false); break;
    case 12: QL_ROW(7, true, false, false); break;
    case 13: QL_ROW(7, true, true, false); break;
    case 14: QL_ROW(8, true, false, false); break;
    case 15: QL_ROW(8, true, true, false); break;
    case 16: QL_ROW(9, true, false, false); break;
    case 17: QL_ROW(9, true, true, false); break;
    case 18: QL_ROW(10, true, false, false); break;
    case 19: QL_ROW(10, true, true, false); break;
    case 20: QL_ROW(12, false, false, true); break;
    case 21: QL_ROW(12, false, true, true); break;
    case 22: QL_ROW(14, false, false, true); break;
    case 23: QL_ROW(14, false, true, true); break;
    case 24: QL_ROW(16, false, false, true); break;
    case 25: QL_ROW(16, false, true, true); break;
    case 26: QL_ROW(18, false, false, true); break;
    case 27: QL_ROW(18, false, true, true); break;
    case 28: QL_ROW(20, false, false, true); break;
    default: QL_ROW(20, false, true, true); break;
    }
#undef QL_ROW
}

// Biased int8 LLRs (L + 128) of four consecutive lanes from their key bits: +mag for a 0 bit, -mag for a 1 bit.
// nib: the four bits, first lane in bit 3 (MSB-first packing); mag: one magnitude byte (0..127) per lane, first lane lowest.
__device__ __forceinline__ u32 synth_biased4(u32 nib, u32 mag)
{
    const u32 m = ((nib * 0x08040201u) >> 3) & 0x01010101u;   // byte b = bit (3 - b) of nib
    const u32 s = m * 0xffu;
    return ((0x80808080u | mag) ^ s) + m;                      // per byte 0x80 + mag, or (0x7f - mag) + 1 = 0x80 - mag: no carries
}

// 4x4 byte transpose: in[k] holds lanes 4j..4j+3 of quarter k; out[m] = belief word of lane 4j+m
__device__ __forceinline__ void transpose4x4(const u32 (&in)[4], u32 (&out)[4])
{
    const u32 t0 = prmt(in[0], in[1], 0x5140u), t1 = prmt(in[2], in[3], 0x5140u);
    const u32 t2 = prmt(in[0], in[1], 0x7362u), t3 = prmt(in[2], in[3], 0x7362u);
    out[0] = prmt(t0, t1, 0x5410u);
    out[1] = prmt(t0, t1, 0x7632u);
    out[2] = prmt(t2, t3, 0x5410u);
    out[3] = prmt(t2, t3, 0x7632u);
}

// WT: W = Z/4 at compile time (0: read it from the parameters); with a constant W every stride of the ballot and
// syndrome code folds into instruction immediates.
//
// Synchronisation inside a frame group (W threads, one frame):
//   full[2]  transaction mbarriers: ONE elected thread stages the next block row -- its check-to-variable messages
//            (bulk copy from the L2-resident scratch) and the raw channel bytes of its extension column (bulk copy
//            from the frame's LLRs) -- into ring stage (trip+1)&1; every thread waits on full[trip&1] before it
//            reads the stage.  No per-thread cp.async, no address arithmetic in the other 95 threads.
//   bar.sync per group after every block row: the belief updates of a row are visible to the next one.
// The messages are written with plain stores and read back, one iteration later, by the bulk-copy engine (async
// proxy): every thread issues fence.proxy.async.global after its last store of an iteration, and the first row of
// the next iteration is staged only after all threads of the group have passed that fence.
template <int NK, int WT>
__global__ void __maxnreg__(kMaxRegs) layered_i8s_kernel(const LayeredI8sParams p)
{
    extern __shared__ __align__(16) char smem[];
    const int W = WT ? WT : p.W;
    const int Z = 4 * W, ZW32 = W >> 3, wq = W >> 5, wq4 = W >> 2;
    const int R = p.brows;
    // Thread -> (frame group g, lane-word i): group = threadIdx.y, so that the warps of a group land on different schedulers
    // (warp w runs on scheduler w % 4) and a block row is worked on by three schedulers at once.
    int g = threadIdx.y;
    const int i = threadIdx.x;
    // the group index is the same for all lanes of a warp (W is a multiple of 32): taking it through a warp reduction puts it
    // and everything derived from it (slot base, barrier id, mbarrier addresses) on the uniform datapath, out of the way of the
    // 120 vector registers, instead of being recomputed from %tid.y in every block row
    if (W % 32 == 0) g = (int)__reduce_max_sync(0xffffffffu, (unsigned)g);
    const int lane = i & 31, wis = i >> 5;
    const int bar_id = 1 + g;
    const int nthreads_cta = W * blockDim.y;

    char *slot = smem + p.tab_bytes + kLi8sSlotBase + g * p.slot_bytes;
    const u32 slot_saddr = (u32)__cvta_generic_to_shared(slot);
    const u32 mb_full = slot_saddr + p.off_mbar, mb_stg = mb_full + 24;   // full[0], full[1], next-frame word, stg
    {   // shared tables (all slots) + one block of biased zero messages; mbarriers of this group
        const int tid = g * W + i;
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tab);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int k = tid; k < (p.tab_bytes >> 4); k += nthreads_cta) dst[k] = src[k];
        if (tid == 0) dst[p.tab_bytes >> 4] = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
        if (p.bits != nullptr) {   // bit input: the per-position magnitudes, one table for all slots of the CTA
            const uint4 *ms = reinterpret_cast<const uint4 *>(p.mag);
            uint4 *md = reinterpret_cast<uint4 *>(smem + p.off_magtab);
            for (int k = tid; k < (p.N >> 4); k += nthreads_cta) md[k] = __ldg(ms + k);
        }
        if (i == 0) {
            mbar_init(mb_full, 1);
            mbar_init(mb_full + 8, 1);
            mbar_init(mb_stg, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    __syncthreads();
    const char *rowsc = smem + p.off_rows;               // Li8sRow[brows], read as int4 pairs
    const Li8sCol *pcols = reinterpret_cast<const Li8sCol *>(smem + p.off_pcols);
    const uint4 *zero_blk = reinterpret_cast<const uint4 *>(smem + p.tab_bytes);
    const u32 tab_saddr = (u32)__cvta_generic_to_shared(smem);

    u32 *Lw = reinterpret_cast<u32 *>(slot);
    char *Li = slot + 4 * i;
    char *ring_i = slot + p.off_ring + 16 * i;                        // 2 stages of stage_bytes
    unsigned char *extb_i = reinterpret_cast<unsigned char *>(slot + p.off_ext) + i;   // 2 buffers of Z bytes
    u32 *hd = reinterpret_cast<u32 *>(slot + p.off_hd);              // bcols vectors of 2*ZW32 words
    u32 *synl = reinterpret_cast<u32 *>(slot + p.off_syn);
    char *rg_slot = reinterpret_cast<char *>(p.rg + (size_t)(blockIdx.x * p.slots + g) * p.rg_u4);   // this slot's scratch
    char *rg_i = rg_slot + 16 * i;
    const bool has_syn = p.syn != nullptr;
    // bit input: the slot's extension-column LLRs live in its own N-byte scratch (same offsets as a frame of LLRs)
    const bool fused = p.bits != nullptr;
    int8_t *ext_slot = fused ? p.ext_scratch + (size_t)(blockIdx.x * p.slots + g) * p.N : nullptr;

    Cx cx;
    cx.cLo = p.h2_lo; cx.cHi = p.h2_hi; cx.cCap = p.h2_cap; cx.cSpan = p.h2_span;
    cx.negOff = p.h2_negoff;
    cx.norm_eighths = p.norm_eighths;
    cx.lane = lane; cx.wis = wis; cx.wq = wq; cx.ZW32 = ZW32;

    // fixed decompositions of the thread index
    const int lc = i / wq4, lj = i - lc * wq4;        // load phase: (column step, word quad)
    const int rs = i / ZW32, w = i - rs * ZW32;       // syndrome / output phases: (row step, 32-bit word); W / ZW32 == 8

    // Stage block row `rn` (messages only when `with_msgs`) for trip `trip`: thread 0 of the group arms the transaction
    // barrier with the byte count and issues the bulk copies; expect_tx may be posted before or after a copy completes.
    auto stage_row = [&](int rn, bool with_msgs, u32 trip, const int8_t *frame) {
        const int4 na = *reinterpret_cast<const int4 *>(rowsc + 32 * rn);        // e_off, thr_off, g_off, nc|nv|variant
        const int es = *reinterpret_cast<const int *>(rowsc + 32 * rn + 16);     // ext_src
        const u32 st = trip & 1u;
        const u32 mb = mb_full + 8 * st;
        const u32 msg_bytes = with_msgs ? (u32)(((na.w >> 8) & 0xff) * W * 16) : 0u;
        const u32 ext_bytes = es >= 0 ? (u32)Z : 0u;
        if (msg_bytes + ext_bytes) mbar_arrive_tx(mb, msg_bytes + ext_bytes);
        else mbar_arrive(mb);
        if (msg_bytes) bulk_g2s(slot_saddr + p.off_ring + st * p.stage_bytes, rg_slot + na.z, msg_bytes, mb);
        if (ext_bytes) bulk_g2s(slot_saddr + p.off_ext + st * Z, frame + es, ext_bytes, mb);
    };

    // Frame prefetch (when the slot has room for it, p.off_stg >= 0): the raw LLR bytes of the core columns of the slot's
    // NEXT frame (bit input: its N/8 bytes of packed bits) are bulk-copied into a staging buffer while the current frame
    // is decoded, so that the frame switch -- which every other group of the CTA waits for at the iteration barrier -- works
    // from shared memory instead of two rounds of global-load latency.
    const bool use_stg = p.off_stg >= 0;
    auto stage_frame = [&](int fr) {   // thread 0 of the group only
        if (fused) {
            mbar_arrive_tx(mb_stg, (u32)(p.N >> 3));
            bulk_g2s(slot_saddr + p.off_stg, p.bits + (size_t)fr * (p.N >> 5), (u32)(p.N >> 3), mb_stg);
            return;
        }
        const int8_t *fsrc = p.llr + (size_t)fr * p.N;
        mbar_arrive_tx(mb_stg, (u32)(p.n_pack * Z));
        for (int c = 0; c < p.n_pack; ++c) bulk_g2s(slot_saddr + p.off_stg + c * Z, fsrc + pcols[c].llr_off, (u32)Z, mb_stg);
    };
    u32 sp = 0;   // parity of the staging barrier

    int f = blockIdx.x * p.slots + g;
    const int fstride = gridDim.x * p.slots;
    // the slot's next frame: f + fstride, or (frame queue) drawn from the launch's counter by thread 0 while the current
    // frame is loaded and handed to the group through a shared-memory word at the next group barrier
    const bool dynq = p.frame_ctr != nullptr;
    volatile int *nf_slot = reinterpret_cast<volatile int *>(slot + p.off_mbar + 16);
    int fnext = f + fstride;
    if (use_stg && i == 0 && f < p.F) stage_frame(f);
    bool active = false, need_load = true, conv = false;
    int it = 0;
    u32 tt = 0;   // trips staged / waited on full[] (runs on across frames)
#pragma unroll 1
    for (;;) {
      if (need_load) {
        need_load = false;
        active = f < p.F;
        if (active) {
        const int8_t *src = fused ? ext_slot : p.llr + (size_t)f * p.N;
        if (!fused && i == 0) stage_row(0, false, tt, src);   // extension bytes of the first row
        int fdraw = 0;
        if (dynq && i == 0) fdraw = fstride + (int)atomicAdd(p.frame_ctr, 1u);
        if (fused) {
            // ---- bit input: int8 LLRs synthesised from the staged key bits and the per-position magnitudes
            mbar_wait(mb_stg, sp);
            sp ^= 1u;
            const u32 *bw = reinterpret_cast<const u32 *>(slot + p.off_stg);   // the frame's packed bits, MSB first
            const char *magtab = smem + p.off_magtab;
            const u32 nsh = 28u - 4u * (u32)(lj & 7);
#pragma unroll kLoadUnroll
            for (int c = lc; c < p.n_pack; c += 4) {   // core columns -> interleaved biased belief words
                const int off = pcols[c].llr_off;
                const u32 *mq = reinterpret_cast<const u32 *>(magtab + off) + lj;
                const u32 *bq = bw + (off >> 5) + (lj >> 3);
                u32 in[4], out[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) in[k] = synth_biased4((bq[k * wq] >> nsh) & 0xfu, mq[k * wq4]);
                transpose4x4(in, out);
                *reinterpret_cast<uint4 *>(Lw + c * W + 4 * lj) = make_uint4(out[0], out[1], out[2], out[3]);
            }
            // extension columns -> raw int8 LLRs in the slot's scratch (word i of a column = lanes 4i .. 4i+3), from where
            // the row staging bulk-copies them exactly as it does from a frame of LLRs
            const u32 esh = 28u - 4u * (u32)(i & 7);
#pragma unroll 4
            for (int r = 0; r < R; ++r) {
                const int es = *reinterpret_cast<const int *>(rowsc + 32 * r + 16);
                if (es < 0) continue;
                const u32 b = synth_biased4((bw[(es >> 5) + (i >> 3)] >> esh) & 0xfu, reinterpret_cast<const u32 *>(magtab + es)[i]);
                reinterpret_cast<u32 *>(ext_slot + es)[i] = b ^ 0x80808080u;
            }
            asm volatile("fence.proxy.async.global;" ::: "memory");   // the bulk copies of the rows read what was just stored
        } else if (use_stg) {
            // ---- load: int8 LLRs of the core columns -> interleaved biased belief words
            // 4 aligned 32-bit loads (one per quarter of the column) -> 4x4 byte transpose -> one 128-bit store
            mbar_wait(mb_stg, sp);
            sp ^= 1u;
            const char *stg = slot + p.off_stg;
#pragma unroll kLoadUnroll
            for (int c = lc; c < p.n_pack; c += 4) {
                const u32 *q = reinterpret_cast<const u32 *>(stg + c * Z) + lj;
                u32 in[4], out[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) in[k] = q[k * wq4] ^ 0x80808080u;
                transpose4x4(in, out);
                *reinterpret_cast<uint4 *>(Lw + c * W + 4 * lj) = make_uint4(out[0], out[1], out[2], out[3]);
            }
        } else {
#pragma unroll kLoadUnroll
            for (int c = lc; c < p.n_pack; c += 4) {
                const u32 *q = reinterpret_cast<const u32 *>(src + pcols[c].llr_off) + lj;
                u32 in[4], out[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) in[k] = __ldg(q + k * wq4) ^ 0x80808080u;
                transpose4x4(in, out);
                *reinterpret_cast<uint4 *>(Lw + c * W + 4 * lj) = make_uint4(out[0], out[1], out[2], out[3]);
            }
        }
        if (has_syn) {   // syndrome rows, Z-bit little-endian vectors (row r, lane l -> bit l)
            const u32 *sf = p.syn + (size_t)f * p.syn_words;
            for (int r = rs; r < R; r += 8) synl[r * ZW32 + w] = __brev(__ldg(sf + r * ZW32 + w));
        }
        if (dynq && i == 0) *nf_slot = fdraw;
        bar_sync(bar_id, W);
        fnext = dynq ? *nf_slot : f + fstride;
        if (fused && i == 0) stage_row(0, false, tt, src);
        if (use_stg && i == 0 && fnext < p.F) stage_frame(fnext);
        if (!fused && fnext < p.F) {   // pull the rest of the slot's next frame towards L2 while the current one is decoded
            const char *nx = reinterpret_cast<const char *>(p.llr + (size_t)fnext * p.N);
            for (int o = i * 128; o < p.N; o += W * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
        }
        it = 0;
        conv = false;
        }
      }
      // All groups of a CTA start every iteration together (one bar.red.or, which also carries the "any group still has a
      // frame" vote): free-running groups execute different row variants at the same time and lose a third of the issue
      // slots to instruction-cache misses (255 KB of unrolled row code per kernel).
      if (!bar_red_or(0, nthreads_cta, active)) break;
      if (active) {
            bool finished = false;
            const int8_t *frame = fused ? ext_slot : p.llr + (size_t)f * p.N;
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
                const int4 la = *reinterpret_cast<const int4 *>(rowsc + 32 * r);        // e_off, thr_off, g_off, nc|nv|variant
                const int4 lb = *reinterpret_cast<const int4 *>(rowsc + 32 * r + 16);   // ext_src, ext_hd, syn_off, deg
                const u32 stage = tt & 1u;
                if (i == 0 && r + 1 < R) stage_row(r + 1, it > 0, tt + 1, frame);
                uint2 m1init = make_uint2(kInf2, kInf2);   // initial (min1, sign) of the two half2 pairs
                if (has_syn) {
                    const u32 *sr = synl + r * ZW32 + wis;
                    const u32 s0 = (sr[0] >> lane) & 1u, s1 = (sr[wq] >> lane) & 1u;
                    const u32 s2 = (sr[2 * wq] >> lane) & 1u, s3 = (sr[3 * wq] >> lane) & 1u;
                    m1init.x ^= (s0 << 15) | (s1 << 31);
                    m1init.y ^= (s2 << 15) | (s3 << 31);
                }
                const uint4 *ysrc = reinterpret_cast<const uint4 *>(ring_i + stage * p.stage_bytes);
                mbar_wait(mb_full + 8 * stage, (tt >> 1) & 1u);   // this row's messages / extension bytes have landed
                ++tt;
                // first iteration of a frame: every old message is zero, read from one shared block (stride 0)
                dispatch_row<NK>(la.w >> 16, cx, Li, tab_saddr + la.x, reinterpret_cast<const int4 *>(smem + la.y), la.w & 0xff, i,
                                 m1init, it ? ysrc : zero_blk, it ? W : 0, reinterpret_cast<uint4 *>(rg_i + la.z), W,
                                 extb_i + stage * Z, reinterpret_cast<u32 *>(reinterpret_cast<char *>(hd) + lb.y));
                if (r == R - 1) asm volatile("fence.proxy.async.global;" ::: "memory");
                bar_sync(bar_id, W);   // every belief update of this row is visible to the next
            }
            ++it;
            const bool more = it < p.max_iter;
            if (i == 0 && more) stage_row(0, true, tt, frame);   // first row of the next iteration (dropped if the frame ends)
            // the syndrome / hard-decision phase runs after every iteration when early stop is on,
            // otherwise once after the last iteration
            if (p.early_stop || !more) {
                // hard decisions of the core columns (extension columns were balloted in their rows)
                u32 *hl = hd + wis;
#pragma unroll 2
                for (int c = 0; c < p.n_pack; ++c) {
                    const u32 X = Lw[c * W + i];                 // biased: bit 7 clear <=> L < 0
                    const u32 b0 = __ballot_sync(0xffffffffu, (X & 0x80u) == 0u);
                    const u32 b1 = __ballot_sync(0xffffffffu, (X & 0x8000u) == 0u);
                    const u32 b2 = __ballot_sync(0xffffffffu, (X & 0x800000u) == 0u);
                    const u32 b3 = __ballot_sync(0xffffffffu, (int)X >= 0);
                    if (lane == 0) {
                        u32 *h = hl + pcols[c].hd_off;
                        h[0] = b0; h[wq] = b1; h[2 * wq] = b2; h[3 * wq] = b3;
                        h[ZW32] = b0; h[ZW32 + wq] = b1; h[ZW32 + 2 * wq] = b2; h[ZW32 + 3 * wq] = b3;
                    }
                }
                bar_sync(bar_id, W);
                // syndrome words: thread -> (row rs + 8k, word w); the doubled vectors make every rotated
                // window contiguous: word (w + shift/32), funnel-shifted by shift%32
                const char *hdw = reinterpret_cast<const char *>(hd + w);
                auto row_syndrome = [&](int r) {
                    const int4 lb = *reinterpret_cast<const int4 *>(rowsc + 32 * r + 16);
                    u32 acc = has_syn ? synl[r * ZW32 + w] : 0u;
                    const u32 *se = reinterpret_cast<const u32 *>(smem + lb.z);
#pragma unroll 4
                    for (int e = 0; e < lb.w; ++e) {
                        const u32 en = se[e];
                        const char *a = hdw + (en >> 5);
                        acc ^= __funnelshift_r(*reinterpret_cast<const u32 *>(a), *reinterpret_cast<const u32 *>(a + 4), en);
                    }
                    return acc;
                };
                // first the block rows 0..7 only: a frame that has not converged is almost always caught by them
                u32 bad = rs < R ? row_syndrome(rs) : 0u;
                bool any_bad = bar_red_or(bar_id, W, bad != 0u);
                if (!any_bad && R > 8) {
                    bad = 0;
#pragma unroll 1
                    for (int r = rs + 8; r < R; r += 8) bad |= row_syndrome(r);
                    any_bad = bar_red_or(bar_id, W, bad != 0u);
                }
                conv = !any_bad;
                finished = conv || !more;
            }
        if (finished) {
        // ---- outputs: MSB-first packed hard decisions of the first out_cols block columns
        uint32_t *of = p.out + (size_t)f * p.out_words;
        for (int c = rs; c < p.out_cols; c += 8) of[c * ZW32 + w] = __brev(hd[c * 2 * ZW32 + w]);
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
        if (more) {   // the speculative stage of the next iteration's first row: let it land, then forget it
            mbar_wait(mb_full + 8 * (tt & 1u), (tt >> 1) & 1u);
            ++tt;
        }
        if (p.discard_scratch) {
            // the messages of the finished frame are dead: drop their (dirty) L2 lines instead of letting them be written back
            const int rg_bytes = p.rg_u4 * 16;
            for (int o = i * 128; o < rg_bytes; o += W * 128) asm volatile("discard.global.L2 [%0], 128;" ::"l"(rg_slot + o) : "memory");
        }
        bar_sync(bar_id, W);   // hd / beliefs / ring are reused by the next frame of this slot
        f = fnext;
        need_load = true;
        }
      }
    }
}

template <int NK, int WT>
int launch_nkw(const LayeredI8sParams &p, int grid, int smem_bytes, cudaStream_t st)
{
    QLDPC_CUDA(cudaFuncSetAttribute(layered_i8s_kernel<NK, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    layered_i8s_kernel<NK, WT><<<grid, dim3(p.W, p.slots), smem_bytes, st>>>(p);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

}  // namespace

int layered_i8s_max_threads() { return kMaxThreads; }

int launch_layered_i8s(const LayeredI8sParams &p, int grid, int smem_bytes, cudaStream_t st)
{
    if (p.rule == QLDPC_RULE_OMS) return launch_nkw<0, 0>(p, grid, smem_bytes, st);
    if (p.norm_eighths == 6) return p.W == 96 ? launch_nkw<6, 96>(p, grid, smem_bytes, st) : launch_nkw<6, 0>(p, grid, smem_bytes, st);
    return launch_nkw<-1, 0>(p, grid, smem_bytes, st);
}

}  // namespace qldpc

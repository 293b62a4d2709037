// Layered int8 min-sum decoder, "streamed" kernel (v5): quasi-cyclic codes with Z % 128 == 0
// (BG1/BG2 lifting sizes 128, 256, 384).  Same arithmetic as layered_i8.cu -- ML/BPSK_nrldpc_sim_FP.m:35-94
// generalised by syndrome input, early stop and the shift-normalised rule; bit-exact with
// oracle/qldpc_oracle.c:ora_decode_layered_fixed -- but a different placement of the state, chosen so that
// FIVE frames (BG1 Z=384) are in flight per SM, 15 warps at 120 registers (see QL_S_MAXNREG):
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
// The QL_S_* switches below are the measured alternatives (profiles/r1_experiments.md); the defaults are the best set.
#include <cuda_fp16.h>

#include "kernels.hpp"

#ifndef QL_S_MAXTHREADS
#define QL_S_MAXTHREADS 480          // 5 frames x 96 threads = 15 warps (18 warps: 5 on one sub-partition -> 96 registers, spills)
#endif
#ifndef QL_S_MAXNREG
// 120, not the 128 that 15 warps would allow: the warps sit 4/4/4/3 on the four SM sub-partitions, and at 128 registers the
// three full ones have none left, so no other CTA can start on the SM.  At 120 each keeps 1024 registers free = one
// 32-register warp, and the LLR synthesis of the next chunk (bitops.cu, 4 warps x 32 registers) runs beside the decoder
// in the host pipeline: device-resident 52.4 vs 52.6 Gbit/s, through the host call 48.0 vs 46.6.  0: cap from the thread count.
#define QL_S_MAXNREG 120
#endif
#ifndef QL_S_FMACLIP
#define QL_S_FMACLIP 1               // message clip as relu forms on the FMA pipe
#endif
#ifndef QL_S_CENTRY
#define QL_S_CENTRY 1                // table entry read in plain C++ (address select) instead of predicated loads
#endif
#ifndef QL_S_KEEPT
#define QL_S_KEEPT 8                 // rows with more core edges re-clip L - R_old in the second pass instead of keeping t
#endif
#ifndef QL_S_ALIGN_ROWS
#define QL_S_ALIGN_ROWS 0            // > 0: re-align the frame groups with a CTA-wide barrier every so many block rows
#endif
#ifndef QL_S_BELMBAR
#define QL_S_BELMBAR 0               // 1: belief hand-over between rows through an arrive/wait mbarrier, 0: bar.sync
#endif
#ifndef QL_S_BIGMODE
#define QL_S_BIGMODE 3   // per-edge register diet of the rows with more than 10 core edges
#endif
#ifndef QL_S_LOADUNROLL
#define QL_S_LOADUNROLL 4            // column steps of the frame load in flight per thread (BG1: 26 core columns = 7 steps)
#endif
#ifndef QL_S_FMANORM
#define QL_S_FMANORM 1               // k/8 normalisation with two FMA-pipe instructions per term instead of shift+mask+add
#endif
#ifndef QL_S_BIGKEEPU
#define QL_S_BIGKEEPU 10             // edges of a big row whose L - R_old stays in registers between the passes
#endif
#ifndef QL_S_FIRSTSPEC
#define QL_S_FIRSTSPEC 0             // separate code for the first iteration of a frame: -8 % (groups at different iterations stop sharing code)
#endif
#ifndef QL_S_SMSP_LOCAL
#define QL_S_SMSP_LOCAL 0            // warps of a frame group on one scheduler: 50.2 vs 51.6 Gbit/s spread over the schedulers (off)
#endif
#ifndef QL_S_ZEROFILL
#define QL_S_ZEROFILL 0              // first iteration reads zero messages from pre-filled ring stages instead of one shared block: 50.0 vs 50.8 (off)
#endif
#ifndef QL_S_SPLITSTAGE
#define QL_S_SPLITSTAGE 0            // staging duty split over two warps of the group: 49.8 vs 50.9 (off)
#endif
#ifndef QL_S_CTAMBAR
#define QL_S_CTAMBAR 0               // iteration alignment through an arrive (after the rows) / wait (before the next rows) mbarrier: 47.6 vs 51.4 (off)
#endif
#ifndef QL_S_FMASEL
#define QL_S_FMASEL 1                // new message = sigma * (c2 + [|t| == min1] * (c1 - c2)) + 128 as two FMAs and one LOP3 (was two LOP3 and an add)
#endif
#ifndef QL_S_M2FMA
#define QL_S_M2FMA 1                 // second minimum as min(m2, |m1| + |t| - |min1'|): one HMNMX2 becomes two adds (bit 0: lanes A, bit 1: lanes B)
#endif
#ifndef QL_S_UNIFORM_G
#define QL_S_UNIFORM_G 1
#endif
#ifndef QL_S_ROTSTAGE
#define QL_S_ROTSTAGE 0
#endif
#ifndef QL_S_HOTCHAIN
#define QL_S_HOTCHAIN 0
#endif
#ifndef QL_S_FMACLIP8
#define QL_S_FMACLIP8 0              // belief clip (upper bound) on the FMA pipe
#endif

namespace qldpc {

namespace {

typedef unsigned int u32;

constexpr u32 kSignMask = 0x80008000u;
constexpr u32 kInf2 = 0x03ff03ffu;     // 1023 ulp: larger than any message magnitude
constexpr u32 kOne2 = 0x3c003c00u;     // 1.0h, 1.0h
constexpr u32 kMinusOne2 = 0xbc00bc00u;
constexpr u32 k128 = 0x00800080u;
constexpr u32 k255 = 0x00ff00ffu;
constexpr int kLoadUnroll = QL_S_LOADUNROLL;

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
__device__ __forceinline__ u32 heq_mask(u32 a, u32 b) { return __heq2_mask(h2(a), h2(b)); }
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
__device__ __forceinline__ void mbar_arrive_drop(u32 mb)
{
    asm volatile("mbarrier.arrive_drop.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
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
// entry 1 (16 bytes further) when it does.  Two predicated loads, one of which executes: no SEL chain.
// The tables are written once before the first __syncthreads and never again, so the load may be scheduled freely.
struct EdgeEntry { u32 off, selA, selB, selW; };
// VOL: a second, volatile flavour for the re-read in the second pass of the big rows (an identical non-volatile
// asm would be merged with the first-pass one and its results kept live across the passes)
template <bool VOL>
__device__ __forceinline__ EdgeEntry load_entry(u32 saddr, int i, int thresh)
{
    EdgeEntry e;
#if QL_S_CENTRY
    {
        const uint4 *tp = reinterpret_cast<const uint4 *>(__cvta_shared_to_generic(saddr)) + (i >= thresh ? 1 : 0);
        const uint4 v = *tp;
        e.off = v.x; e.selA = v.y; e.selB = v.z; e.selW = v.w;
        return e;
    }
#endif
    if constexpr (VOL) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %4, %5;\n\t"
                     "@p ld.shared.v4.u32 {%0,%1,%2,%3}, [%6+16];\n\t"
                     "@!p ld.shared.v4.u32 {%0,%1,%2,%3}, [%6];\n\t}"
                     : "=r"(e.off), "=r"(e.selA), "=r"(e.selB), "=r"(e.selW)
                     : "r"(i), "r"(thresh), "r"(saddr));
        return e;
    }
    asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %4, %5;\n\t"
        "@p ld.shared.v4.u32 {%0,%1,%2,%3}, [%6+16];\n\t"
        "@!p ld.shared.v4.u32 {%0,%1,%2,%3}, [%6];\n\t}"
        : "=r"(e.off), "=r"(e.selA), "=r"(e.selB), "=r"(e.selW)
        : "r"(i), "r"(thresh), "r"(saddr));
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
#if QL_S_FMACLIP
    // hi - relu(span - relu(u + cap)): three FMA-pipe instructions, none on the ALU pipe
    return hsub(cx.cHi, hrsub_relu(hadd_relu(u, cx.cCap), cx.cSpan));
#else
    return hmin(hmax(u, cx.cLo), cx.cHi);
#endif
}
// clip(L - R_old + R_new) in biased form: relu(u + b) capped at 255
__device__ __forceinline__ u32 clip_belief(u32 u, u32 b)
{
    const u32 l = hadd_relu(u, b);
#if QL_S_FMACLIP8
    return hsub(k255, hrsub_relu(l, k255));
#else
    return hmin(l, k255);
#endif
}

// integer k/8 normalisation on two packed non-negative fields: sum of floor(x/2), floor(x/4), floor(x/8) terms
// (AFF3CT's integer normalize).  FMA form: floor(x / 2^s) = (x with its s low bits cleared) * 2^-s, exact in fp16
// for integers held as subnormals; one LOP3 on the ALU pipe per term, the multiply-adds run on the FMA pipe.
template <int NK>
__device__ __forceinline__ u32 norm_eighths2(u32 x, int k_rt)
{
    const int k = NK < 0 ? k_rt : NK;
    if (k >= 8) return x;
#if QL_S_FMANORM
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
#endif
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
//          and pack selector; 1: no clipped value (re-clip); 2: only L - R_old (re-clip, re-read the table entry);
//          3: only the belief word (re-read the table entry and the old message, redo the subtraction and the clip)
//   Li     this thread's belief base (slot beliefs + 4*i); erow: shared address of the row's first table entry
//   thr4   wrap thresholds of the row's edges, four per int4
//   ysrc   this thread's message blocks (ystride uint4 apart); gdst: where the new ones go (W uint4 apart)
//   FIRST  first iteration of a frame: every old message is zero, nothing is read from the ring
template <int NK, int DC, bool EXACT, bool EXT, int MODE, bool FIRST>
__device__ __forceinline__ void process_row(const Cx &cx, char *Li, u32 erow, const int4 *thr4, int nc, int i, uint2 m1init,
                                            const uint4 *ysrc, int ystride, uint4 *gdst, int W,
                                            const unsigned char *extb, u32 *hd_ext)
{
    // MODE 3 keeps L - R_old for its first QL_S_BIGKEEPU edges and only the belief word for the others
    constexpr int KU = MODE <= 2 ? DC : (QL_S_BIGKEEPU < DC ? QL_S_BIGKEEPU : DC);
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
                if constexpr (!FIRST) Yq = ysrc[(j >> 2) * ystride];
                Tq = thr4[j >> 2];
            }
            const EdgeEntry en = load_entry<false>(erow + 32 * j, i, vcomp(Tq, j & 3));
            char *a = Li + en.off;
            const u32 X = *reinterpret_cast<const u32 *>(a);
            const u32 xA = prmt(X, 0u, en.selA), xB = prmt(X, 0u, en.selB);
            u32 yA = k128, yB = k128;                                 // biased zero message
            if constexpr (!FIRST) {
                const u32 Y = vcomp(Yq, j & 3);
                yA = prmt(Y, 0u, 0x4140u); yB = prmt(Y, 0u, 0x4342u);
            }
            const u32 ua = hsub(xA, yA), ub = hsub(xB, yB);          // L - R_old  (:51)
            const u32 ta = clip_msg(cx, ua), tb = clip_msg(cx, ub);   // clip to the message range (:54-55)
            if (j < KU) { uA[j] = ua; uB[j] = ub; } else { xk[j] = X; }
            if constexpr (MODE == 0) { tA[j] = ta; tB[j] = tb; }
            if constexpr (MODE <= 1) { sw[j] = en.selW; ad[j] = a; }
            two_min<(QL_S_M2FMA & 1) != 0>(m1A, m2A, ta);               // first minimum and sign product (:60,:63), second minimum (:61)
            two_min<(QL_S_M2FMA & 2) != 0>(m1B, m2B, tb);
        }
    }
    u32 ueA = 0, ueB = 0, teA = 0, teB = 0;
    if constexpr (EXT) {   // the extension edge: belief - message == channel LLR
        const u32 b0 = extb[0], b1 = extb[W], b2 = extb[2 * W], b3 = extb[3 * W];
        ueA = hsub((b0 | (b1 << 16)) ^ k128, k128);
        ueB = hsub((b2 | (b3 << 16)) ^ k128, k128);
        teA = clip_msg(cx, ueA);
        teB = clip_msg(cx, ueB);
        two_min<(QL_S_M2FMA & 1) != 0>(m1A, m2A, teA);
        two_min<(QL_S_M2FMA & 2) != 0>(m1B, m2B, teB);
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
#if QL_S_FMASEL
    // magnitude c2 + [|t| == min1] * (c1 - c2), sign sigma = (row parity ^ sign(t)) as +-1.0h: all exact on subnormal integers
    const u32 dcA = hsub(c1A, c2A), dcB = hsub(c1B, c2B);
    const u32 sgA = parA ^ kOne2, sgB = parB ^ kOne2;
#else
    c1A ^= parA; c2A ^= parA; c1B ^= parB; c2B ^= parB;               // parity folded into both candidates
#endif

    if constexpr (EXT) {   // only the sign of channel + new message is observable: ballot it into the hd vector
#if QL_S_FMASEL
        const u32 aA = hfma((teA & kSignMask) ^ sgA, hfma(heq_one(habs(teA), min1A), dcA, c2A), ueA);
        const u32 aB = hfma((teB & kSignMask) ^ sgB, hfma(heq_one(habs(teB), min1B), dcB, c2B), ueB);
#else
        const u32 eA = heq_mask(habs(teA), min1A), eB = heq_mask(habs(teB), min1B);
        const u32 rA = ((eA & c1A) | (~eA & c2A)) ^ (teA & kSignMask);
        const u32 rB = ((eB & c1B) | (~eB & c2B)) ^ (teB & kSignMask);
        const u32 aA = hadd(ueA, rA), aB = hadd(ueB, rB);
#endif
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
                const EdgeEntry en = load_entry<true>(erow + 32 * j, i, vcomp(Tq, j & 3));
                a = Li + en.off;
                selW = en.selW;
                if (j >= KU) {
                    u32 yA = k128, yB = k128;
                    if constexpr (!FIRST) {
                        if ((j & 3) == 0 || j == KU) Yq = lds128_volatile(ysrc + (j >> 2) * ystride);
                        const u32 Y = vcomp(Yq, j & 3);
                        yA = prmt(Y, 0u, 0x4140u); yB = prmt(Y, 0u, 0x4342u);
                    }
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
#if QL_S_FMASEL
            const u32 bA = hfma((ta & kSignMask) ^ sgA, hfma(heq_one(habs(ta), min1A), dcA, c2A), k128);   // biased new message
            const u32 bB = hfma((tb & kSignMask) ^ sgB, hfma(heq_one(habs(tb), min1B), dcB, c2B), k128);
#else
            const u32 eA = heq_mask(habs(ta), min1A), eB = heq_mask(habs(tb), min1B);
            const u32 rA = ((eA & c1A) | (~eA & c2A)) ^ (ta & kSignMask);
            const u32 rB = ((eB & c1B) | (~eB & c2B)) ^ (tb & kSignMask);
            const u32 bA = hadd(rA, k128), bB = hadd(rB, k128);       // biased new message
#endif
            const u32 lA = clip_belief(ua, bA);                       // clip(L - R_old + R_new) biased (:88-91)
            const u32 lB = clip_belief(ub, bB);
            vset(Yn, j & 3, prmt(bA, bB, 0x6420u));
            *reinterpret_cast<u32 *>(a) = prmt(lA, lB, selW);
        }
        if ((j & 3) == 3 || j == DC - 1) gdst[(j >> 2) * W] = Yn;
    }
}

template <int NK, bool FIRST>
__device__ __forceinline__ void dispatch_row(int variant, const Cx &cx, char *Li, u32 erow, const int4 *thr4, int nc, int i,
                                             uint2 m1init, const uint4 *ysrc, int ystride, uint4 *gdst, int W,
                                             const unsigned char *extb, u32 *hd_ext)
{
#define QL_ROW(DCV, EXACTV, EXTV, BIGV) \
    process_row<NK, DCV, EXACTV, EXTV, (BIGV ? QL_S_BIGMODE : (DCV > QL_S_KEEPT ? 1 : 0)), FIRST>(cx, Li, erow, thr4, nc, i, m1init, ysrc, ystride, gdst, \
                                                                                W, extb, hd_ext)
#if QL_S_HOTCHAIN
    // the four shapes that make up 36 of BG1's 46 block rows (3..6 core edges + the extension edge) are tested with plain
    // compare-and-branch before the jump table (an indexed constant load + indirect branch per row); the copy of `variant`
    // is opaque so that the compiler does not fold the chain back into the switch
    {
        int v2 = variant;
        asm volatile("" : "+r"(v2));
        if (v2 == 7) { QL_ROW(4, true, true, false); return; }
        if (v2 == 9) { QL_ROW(5, true, true, false); return; }
        if (v2 == 5) { QL_ROW(3, true, true, false); return; }
        if (v2 == 11) { QL_ROW(6, true, true, false); return; }
    }
#endif
    switch (variant) {
#ifndef QL_S_NOSMALL
    case 0: QL_ROW(1, true, false, false); break;
    case 1: QL_ROW(1, true, true, false); break;
    case 2: QL_ROW(2, true, false, false); break;
    case 3: QL_ROW(2, true, true, false); break;
    case 4: QL_ROW(3, true, false, false); break;
#if !QL_S_HOTCHAIN
    case 5: QL_ROW(3, true, true, false); break;
#endif
    case 6: QL_ROW(4, true, false, false); break;
#if !QL_S_HOTCHAIN
    case 7: QL_ROW(4, true, true, false); break;
#endif
    case 8: QL_ROW(5, true, false, false); break;
#if !QL_S_HOTCHAIN
    case 9: QL_ROW(5, true, true, false); break;
#endif
    case 10: QL_ROW(6, true, false, false); break;
#if !QL_S_HOTCHAIN
    case 11: QL_ROW(6, true, true, false); break;
#endif
    case 12: QL_ROW(7, true, false, false); break;
    case 13: QL_ROW(7, true, true, false); break;
    case 14: QL_ROW(8, true, false, false); break;
    case 15: QL_ROW(8, true, true, false); break;
    case 16: QL_ROW(9, true, false, false); break;
    case 17: QL_ROW(9, true, true, false); break;
    case 18: QL_ROW(10, true, false, false); break;
    case 19: QL_ROW(10, true, true, false); break;
#endif
#ifndef QL_S_NOBIG
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
#else
    default: break;
#endif
    }
#undef QL_ROW
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
//   bel      arrival mbarrier (count W): a thread arrives when its belief updates of a row are written and waits
//            only right before it reads beliefs again, so the row header runs in the shadow of slower warps.
// The messages are written with plain stores and read back, one iteration later, by the bulk-copy engine (async
// proxy): every thread issues fence.proxy.async.global after its last store of an iteration, and the first row of
// the next iteration is staged only after all threads of the group have passed that fence.
template <int NK, int WT>
#if QL_S_MAXNREG > 0
__global__ void __maxnreg__(QL_S_MAXNREG) layered_i8s_kernel(const LayeredI8sParams p)
#else
__global__ void __launch_bounds__(QL_S_MAXTHREADS, 1) layered_i8s_kernel(const LayeredI8sParams p)
#endif
{
    extern __shared__ __align__(16) char smem[];
    const int W = WT ? WT : p.W;
    const int Z = 4 * W, ZW32 = W >> 3, wq = W >> 5, wq4 = W >> 2;
    const int R = p.brows;
    // Thread -> (frame group g, lane-word i).  The hardware assigns warp w to scheduler w % 4.  Default: group = threadIdx.y,
    // its warps land on different schedulers and run in parallel.  QL_S_SMSP_LOCAL puts the warps of a group on ONE
    // scheduler (they reach the row barrier together, but a row then takes longer): measured slower.
    int g, i;
    {
        const int tid_lin = threadIdx.y * W + threadIdx.x, wl = tid_lin >> 5, wpg = W >> 5, S = blockDim.y;
#if QL_S_SMSP_LOCAL
        const int quad = 4 * wpg;                           // warps of four groups
        if (wl < (S / 4) * quad) {
            const int b = wl / quad, r = wl - b * quad;
            g = 4 * b + (r & 3);
            i = (r >> 2) * 32 + (tid_lin & 31);
        } else {
            const int r = wl - (S / 4) * quad;
            g = 4 * (S / 4) + r / wpg;
            i = (r % wpg) * 32 + (tid_lin & 31);
        }
#else
        (void)wl; (void)wpg; (void)S;
        g = threadIdx.y; i = threadIdx.x;
#endif
    }
#if QL_S_UNIFORM_G
    // the group index is the same for all lanes of a warp (W is a multiple of 32): taking it through a warp reduction puts it
    // and everything derived from it (slot base, barrier id, mbarrier addresses) on the uniform datapath, out of the way of the
    // 120 vector registers, instead of being recomputed from %tid.y in every block row
    if (W % 32 == 0) g = (int)__reduce_max_sync(0xffffffffu, (unsigned)g);
#if QL_S_UNIFORM_G > 1
    // same for the warp's index inside its group: i = 32 * (uniform) + lane lets per-thread addresses split into a uniform base
    // and a lane offset
    if (W % 32 == 0) i = 32 * (int)__reduce_max_sync(0xffffffffu, (unsigned)(i >> 5)) + (i & 31);
#endif
#endif
    const int lane = i & 31, wis = i >> 5;
    const int bar_id = 1 + g;

    char *slot = smem + p.tab_bytes + kLi8sSlotBase + g * p.slot_bytes;
    const u32 slot_saddr = (u32)__cvta_generic_to_shared(slot);
    const u32 mb_full = slot_saddr + p.off_mbar, mb_bel = mb_full + 16, mb_stg = mb_full + 24;   // full[0], full[1], bel, stg
    {   // shared tables (all slots) + one block of biased zero messages; mbarriers of this group
        const int tid = g * W + i, nthreads = W * blockDim.y;
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tab);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int k = tid; k < (p.tab_bytes >> 4); k += nthreads) dst[k] = src[k];
        if (tid == 0) {
            dst[p.tab_bytes >> 4] = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
            mbar_init((u32)__cvta_generic_to_shared(smem + p.tab_bytes + 16), (u32)nthreads);   // iteration barrier of the CTA
        }
        if (i == 0) {
            mbar_init(mb_full, 1);
            mbar_init(mb_full + 8, 1);
#if QL_S_BELMBAR
            mbar_init(mb_bel, (u32)W);
#endif
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

    Cx cx;
    cx.cLo = p.h2_lo; cx.cHi = p.h2_hi; cx.cCap = p.h2_cap; cx.cSpan = p.h2_span;
    cx.negOff = p.h2_negoff;
    cx.norm_eighths = p.norm_eighths;
    cx.lane = lane; cx.wis = wis; cx.wq = wq; cx.ZW32 = ZW32;

    // fixed decompositions of the thread index
    const int lc = i / wq4, lj = i - lc * wq4;        // load phase: (column step, word quad)
    const int rs = i / ZW32, w = i - rs * ZW32;       // syndrome / output phases: (row step, 32-bit word); W / ZW32 == 8

#if QL_S_BELMBAR
#define QL_BEL_ARRIVE() mbar_arrive(mb_bel)
#define QL_BEL_WAIT() do { mbar_wait(mb_bel, bt & 1u); ++bt; } while (0)
#else
#define QL_BEL_ARRIVE() bar_sync(bar_id, W)
#define QL_BEL_WAIT() do { } while (0)
#endif
    // Stage block row `rn` (messages only when `with_msgs`) for trip `trip` -- elected thread only.
    // Two threads of different warps share the duty, so that no single warp of the group carries all of it into the row
    // barrier: thread 0 arms the barrier with the byte count and copies the messages, thread `ext_thread` (first lane of the
    // second warp when there is one) copies the extension bytes.  expect_tx may be posted before or after a copy completes.
    const int ext_thread = (QL_S_SPLITSTAGE && W > 32) ? 32 : 0;
    auto stage_row = [&](int rn, bool with_msgs, u32 trip, const int8_t *frame) {
        const int4 na = *reinterpret_cast<const int4 *>(rowsc + 32 * rn);        // e_off, thr_off, g_off, nc|nv|variant
        const int es = *reinterpret_cast<const int *>(rowsc + 32 * rn + 16);     // ext_src
        const u32 st = trip & 1u;
        const u32 mb = mb_full + 8 * st;
        const u32 msg_bytes = with_msgs ? (u32)(((na.w >> 8) & 0xff) * W * 16) : 0u;
        const u32 ext_bytes = es >= 0 ? (u32)Z : 0u;
        if (i == 0) {
            if (msg_bytes + ext_bytes) mbar_arrive_tx(mb, msg_bytes + ext_bytes);
            else mbar_arrive(mb);
            if (msg_bytes) bulk_g2s(slot_saddr + p.off_ring + st * p.stage_bytes, rg_slot + na.z, msg_bytes, mb);
        }
        if (i == ext_thread && ext_bytes) bulk_g2s(slot_saddr + p.off_ext + st * Z, frame + es, ext_bytes, mb);
    };

#if QL_S_ROTSTAGE
    auto stage_row_any = [&](int rn, bool with_msgs, u32 trip, const int8_t *frame) {   // the caller elected ONE thread
        const int4 na = *reinterpret_cast<const int4 *>(rowsc + 32 * rn);
        const int es = *reinterpret_cast<const int *>(rowsc + 32 * rn + 16);
        const u32 st = trip & 1u;
        const u32 mb = mb_full + 8 * st;
        const u32 msg_bytes = with_msgs ? (u32)(((na.w >> 8) & 0xff) * W * 16) : 0u;
        const u32 ext_bytes = es >= 0 ? (u32)Z : 0u;
        if (msg_bytes + ext_bytes) mbar_arrive_tx(mb, msg_bytes + ext_bytes);
        else mbar_arrive(mb);
        if (msg_bytes) bulk_g2s(slot_saddr + p.off_ring + st * p.stage_bytes, rg_slot + na.z, msg_bytes, mb);
        if (ext_bytes) bulk_g2s(slot_saddr + p.off_ext + st * Z, frame + es, ext_bytes, mb);
    };
    int rot = 0;
#endif

    // Frame prefetch (when the slot has room for it, p.off_stg >= 0): the raw LLR bytes of the core columns of the slot's
    // NEXT frame are bulk-copied into a staging buffer while the current frame is decoded, so that the frame switch --
    // which every other group of the CTA waits for at the iteration barrier -- is a shared-memory transpose instead of
    // two rounds of global-load latency.
    const bool use_stg = p.off_stg >= 0;
    auto stage_frame = [&](int fr) {   // elected thread only
        const int8_t *fsrc = p.llr + (size_t)fr * p.N;
        mbar_arrive_tx(mb_stg, (u32)(p.n_pack * Z));
        for (int c = 0; c < p.n_pack; ++c) bulk_g2s(slot_saddr + p.off_stg + c * Z, fsrc + pcols[c].llr_off, (u32)Z, mb_stg);
    };
    u32 sp = 0;   // parity of the staging barrier

    int f = blockIdx.x * p.slots + g;
    const int fstride = gridDim.x * p.slots;
    // the slot's next frame: f + fstride, or (frame queue) drawn from the launch's counter by thread 0 while the current
    // frame is loaded and handed to the group through the word of the (unused) belief mbarrier
    const bool dynq = !QL_S_BELMBAR && p.frame_ctr != nullptr;
    volatile int *nf_slot = reinterpret_cast<volatile int *>(slot + p.off_mbar + 16);
    int fnext = f + fstride;
    if (use_stg && i == 0 && f < p.F) stage_frame(f);
    [[maybe_unused]] const int nthreads_cta = W * blockDim.y;
    bool active = false, need_load = true, conv = false;
    int it = 0;
    u32 tt = 0;
    [[maybe_unused]] u32 bt = 0;   // trips staged / waited on full[], phases waited on bel (both run on across frames)
#pragma unroll 1
    // QL_S_CTAMBAR (measured, off): the iteration alignment as an mbarrier every thread ARRIVES on right after its block rows
    // and WAITS on right before the next ones, so that the phases in between (hard decisions, syndrome, output, frame switch)
    // overlap with the waiting of the other groups.  It removes the 8 % barrier stall and loses more than that, because the
    // groups then run those phases -- different code -- while others are in their rows (instruction-cache sharing again).
    // A group that runs out of frames leaves with arrive_drop, after the round it has already arrived for has completed.
    const u32 mb_cta = (u32)__cvta_generic_to_shared(smem + p.tab_bytes + 16);
    [[maybe_unused]] u32 cpar = 0;
    [[maybe_unused]] bool first_round = true;
    for (;;) {
      if (need_load) {
        need_load = false;
        active = f < p.F;
#if QL_S_CTAMBAR
        if (!active) {
            if (!first_round) mbar_wait(mb_cta, cpar);
            mbar_arrive_drop(mb_cta);
            break;
        }
#endif
        if (active) {
        const int8_t *src = p.llr + (size_t)f * p.N;
        if (i == 0 || i == ext_thread) stage_row(0, false, tt, src);   // extension bytes of the first row
        int fdraw = 0;
        if (dynq && i == 0) fdraw = fstride + (int)atomicAdd(p.frame_ctr, 1u);
        // ---- load: int8 LLRs of the core columns -> interleaved biased belief words
        // 4 aligned 32-bit loads (one per quarter of the column) -> 4x4 byte transpose -> one 128-bit store
        if (use_stg) {
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
#if QL_S_ZEROFILL
        {   // the first iteration of a frame reads zero messages: both ring stages are filled with biased zeros (nothing is
            // staged into them before the second iteration)
            uint4 *rz = reinterpret_cast<uint4 *>(ring_i);
            const int n16 = (2 * p.stage_bytes) >> 4;
            for (int k = 0; k < n16; k += W) rz[k] = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
        }
#endif
        if (dynq && i == 0) *nf_slot = fdraw;
        QL_BEL_ARRIVE();
        fnext = dynq ? *nf_slot : f + fstride;
        if (use_stg) {
#if QL_S_BELMBAR
            bar_sync(bar_id, W);   // every thread has read the staging buffer
#endif
            if (i == 0 && fnext < p.F) stage_frame(fnext);
        }
        if (fnext < p.F) {   // pull the rest of the slot's next frame towards L2 while the current one is decoded
            const char *nx = reinterpret_cast<const char *>(p.llr + (size_t)fnext * p.N);
            for (int o = i * 128; o < p.N; o += W * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
        }
        it = 0;
        conv = false;
        }
      }
#if QL_S_CTAMBAR
      if (!first_round) { mbar_wait(mb_cta, cpar); cpar ^= 1u; }   // every group has finished the rows of the previous round
      first_round = false;
#else
      if (!bar_red_or(0, nthreads_cta, active)) break;   // also the alignment barrier of the iteration
#endif
      if (active) {
            bool finished = false;
            const int8_t *frame = p.llr + (size_t)f * p.N;
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
                const int4 la = *reinterpret_cast<const int4 *>(rowsc + 32 * r);        // e_off, thr_off, g_off, nc|nv|variant
                const int4 lb = *reinterpret_cast<const int4 *>(rowsc + 32 * r + 16);   // ext_src, ext_hd, syn_off, deg
                const u32 stage = tt & 1u;
                QL_BEL_WAIT();   // every belief update of the previous row is visible
#if QL_S_ROTSTAGE
                // the staging duty (30 instructions with one active lane) rotates over the warps of the group, so that no warp
                // is the slow one at every row barrier
                if (lane == 0 && wis == rot && r + 1 < R) stage_row_any(r + 1, it > 0, tt + 1, frame);
                rot = (rot + 1 == W / 32) ? 0 : rot + 1;
#else
                if ((i == 0 || i == ext_thread) && r + 1 < R) stage_row(r + 1, it > 0, tt + 1, frame);
#endif
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
#if QL_S_FIRSTSPEC
                if (it == 0)
                    dispatch_row<NK, true>(la.w >> 16, cx, Li, tab_saddr + la.x, reinterpret_cast<const int4 *>(smem + la.y),
                                           la.w & 0xff, i, m1init, ysrc, W, reinterpret_cast<uint4 *>(rg_i + la.z), W,
                                           extb_i + stage * Z, reinterpret_cast<u32 *>(reinterpret_cast<char *>(hd) + lb.y));
                else
#endif
                dispatch_row<NK, false>(la.w >> 16, cx, Li, tab_saddr + la.x, reinterpret_cast<const int4 *>(smem + la.y),
                                        la.w & 0xff, i, m1init, (QL_S_FIRSTSPEC || QL_S_ZEROFILL || it) ? ysrc : zero_blk,
                                        (QL_S_FIRSTSPEC || QL_S_ZEROFILL || it) ? W : 0,
                                        reinterpret_cast<uint4 *>(rg_i + la.z), W,
                                        extb_i + stage * Z, reinterpret_cast<u32 *>(reinterpret_cast<char *>(hd) + lb.y));
                if (r == R - 1) asm volatile("fence.proxy.async.global;" ::: "memory");
                QL_BEL_ARRIVE();
            }
#if QL_S_CTAMBAR
            mbar_arrive(mb_cta);
#endif
            ++it;
            const bool more = it < p.max_iter;
            QL_BEL_WAIT();
            if ((i == 0 || i == ext_thread) && more) stage_row(0, true, tt, frame);   // first row of the next iteration (dropped if the frame ends)
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
                u32 bad = 0;
                if (rs < R) {
                    const int4 lb = *reinterpret_cast<const int4 *>(rowsc + 32 * rs + 16);
                    u32 acc = has_syn ? synl[rs * ZW32 + w] : 0u;
                    const u32 *se = reinterpret_cast<const u32 *>(smem + lb.z);
#pragma unroll 4
                    for (int e = 0; e < lb.w; ++e) {
                        const u32 en = se[e];
                        const char *a = hdw + (en >> 5);
                        acc ^= __funnelshift_r(*reinterpret_cast<const u32 *>(a), *reinterpret_cast<const u32 *>(a + 4), en);
                    }
                    bad = acc;
                }
                bool any_bad = bar_red_or(bar_id, W, bad != 0u);
                if (!any_bad && R > 8) {
                    bad = 0;
#pragma unroll 1
                    for (int r = rs + 8; r < R; r += 8) {
                        const int4 lb = *reinterpret_cast<const int4 *>(rowsc + 32 * r + 16);
                        u32 acc = has_syn ? synl[r * ZW32 + w] : 0u;
                        const u32 *se = reinterpret_cast<const u32 *>(smem + lb.z);
#pragma unroll 4
                        for (int e = 0; e < lb.w; ++e) {
                            const u32 en = se[e];
                            const char *a = hdw + (en >> 5);
                            acc ^= __funnelshift_r(*reinterpret_cast<const u32 *>(a), *reinterpret_cast<const u32 *>(a + 4), en);
                        }
                        bad |= acc;
                    }
                    any_bad = bar_red_or(bar_id, W, bad != 0u);
                }
                conv = !any_bad;
                finished = conv || !more;
            }
        if (!finished) {
#if QL_S_BELMBAR
            mbar_arrive(mb_bel);   // matched by the wait of the next iteration's first row
#endif
        } else {
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

int layered_i8s_max_threads() { return QL_S_MAXTHREADS; }

int launch_layered_i8s(const LayeredI8sParams &p, int grid, int smem_bytes, cudaStream_t st)
{
    if (p.rule == QLDPC_RULE_OMS) return launch_nkw<0, 0>(p, grid, smem_bytes, st);
    if (p.norm_eighths == 6) return p.W == 96 ? launch_nkw<6, 96>(p, grid, smem_bytes, st) : launch_nkw<6, 0>(p, grid, smem_bytes, st);
    return launch_nkw<-1, 0>(p, grid, smem_bytes, st);
}

}  // namespace qldpc

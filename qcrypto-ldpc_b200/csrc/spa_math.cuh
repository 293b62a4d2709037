// The transcendentals of the sum-product check update, shared by every float SPA kernel.
//
// Exact flavour (default): tanh(|x| / 2) and 2 atanh(r) evaluated in double and rounded ONCE to float -- the arithmetic
// specification of DESIGN.md section 2, what oracle/qldpc_oracle.c does with libm.  These kernels are bound by the FP64
// pipe (about 45 double-precision instructions per call at ~32 lanes per clock and SM).  Measured and dropped: hand-written
// double kernels (Cody-Waite exp + degree-13 polynomial, log through (m - 1) / (m + 1)) with a Ziv rounding test and libm
// as the fallback -- bit-identical to libm on 4e8 random arguments, but the same number of FP64 instructions: N=65536 SPA
// 33.2 ms vs 32.6 ms.
//
// fp32 flavour (QLDPC_FLAG_FAST_SPA): special-function-unit approximations with short odd series where the closed forms
// cancel.
#pragma once

namespace qldpc {

static __device__ __noinline__ float tanh_half_exact(float a) { return (float)tanh((double)(a * 0.5f)); }
static __device__ __noinline__ float two_atanh_exact(float r) { return 2.0f * (float)atanh((double)r); }

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// tanh(a / 2) for a >= 0: 1 - 2 / (e^a + 1) (exact to the last bit near 1, where atanh is ill-conditioned); below 1/16 the
// difference cancels and the odd series h - h^3 / 3 (next term 2 h^5 / 15 < 1.3e-7 h) takes over.  Five instructions, two
// of them on the special-function unit; relative error < 2e-5 everywhere.
__device__ __forceinline__ float tanh_half_fast(float a)
{
    const float h = 0.5f * a;
    const float series = h * fmaf(h * h, -0.33333334f, 1.0f);
    const float e = ex2_approx(a * 1.4426950408889634f);      // +inf for a > 88 -> t = 1
    const float t = fmaf(rcp_approx(e + 1.0f), -2.0f, 1.0f);
    return a < 0.0625f ? series : t;
}
// 2 * atanh(r) for 0 <= r < 1: ln(1 + r) - ln(1 - r) (1 - r is exact), series 2 r (1 + r^2 / 3) below 1/16
__device__ __forceinline__ float two_atanh_fast(float r)
{
    const float series = (r + r) * fmaf(r * r, 0.33333334f, 1.0f);
    const float lg = 0.6931471805599453f * (lg2_approx(1.0f + r) - lg2_approx(1.0f - r));
    return r < 0.0625f ? series : lg;
}

}  // namespace qldpc

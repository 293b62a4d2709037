// fp32 flavour of the SPA transcendentals (QLDPC_FLAG_FAST_SPA): special-function-unit approximations with short odd series
// where the closed forms cancel.  Shared by the flooding kernels; the exact flavour (tanh / atanh in double, rounded once)
// stays next to each kernel.
#pragma once

namespace qldpc {

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

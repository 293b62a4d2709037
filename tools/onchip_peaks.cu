// Micro-benchmark of the on-chip ceilings the layered decoder is measured against (SURVEY.md 8d asks for a measured
// shared-memory peak next to the HBM figure): shared-memory load/store bandwidth at the access widths the kernels use,
// L2 read bandwidth over a buffer the size of the message scratch, and the issue rate of the two pipes the edge update
// runs on (ALU: LOP3/PRMT/IADD3, FMA: HFMA2).  Prints one JSON line.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o onchip_peaks tools_onchip_peaks.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int kThreads = 1024;
constexpr int kSmemBytes = 64 * 1024;

template <int WIDTH>   // bytes per thread per access: 4, 8 or 16
__global__ void __launch_bounds__(kThreads) smem_read(int iters, uint32_t *sink)
{
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < kSmemBytes / 4; i += kThreads) ((uint32_t *)sm)[i] = i;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * WIDTH;
    constexpr uint32_t step = kThreads * WIDTH;          // one pass of the CTA
    constexpr uint32_t mask = kSmemBytes - 1;
    uint32_t acc = 0, off = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t a = base + ((off + u * step) & mask);
            if (WIDTH == 4) { uint32_t x; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(x) : "r"(a)); acc ^= x; }
            else if (WIDTH == 8) { uint32_t x, y; asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(a)); acc ^= x ^ y; }
            else { uint32_t x, y, z, w; asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a)); acc ^= x ^ y ^ z ^ w; }
        }
        off += 8 * step;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int WIDTH>
__global__ void __launch_bounds__(kThreads) smem_write(int iters, uint32_t *sink)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * WIDTH;
    constexpr uint32_t step = kThreads * WIDTH;
    constexpr uint32_t mask = kSmemBytes - 1;
    uint32_t off = 0, v = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t a = base + ((off + u * step) & mask);
            if (WIDTH == 4) asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v));
            else asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" :: "r"(a), "r"(v));
        }
        off += 8 * step;
        v += it;
    }
    __syncthreads();
    if (((uint32_t *)sm)[threadIdx.x] == 0x12345678u) sink[0] = 1;
}

// every thread reads 16 B at a stride of one grid pass; the buffer (scratch-sized) stays in the 126 MB L2
__global__ void __launch_bounds__(512) l2_read(const uint4 *buf, size_t n_vec, int passes, uint32_t *sink)
{
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) sink[0] = acc;
}

// 8 independent chains per thread.  One op code per chain slot: chains u = 0..7 run op OPS[u % NOPS].
enum Op { LOP3 = 0, PRMT, HFMA2, HMNMX2, IADD3, IMAD, HSET2, FFMA, HADD2, VIMNMX16, VIADDMNMX16, VIMNMX3_16, VIMNMX32 };
template <int OP>
__device__ __forceinline__ void one(uint32_t &r, uint32_t k1, uint32_t k2)
{
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(k1), "r"(k2));
    else if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r) : "r"(k1), "r"(k2 & 0x7777u));
    else if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r) : "r"(k1), "r"(k2));
    else if (OP == HMNMX2) asm volatile("min.xorsign.abs.f16x2 %0, %0, %1;" : "+r"(r) : "r"(k1));
    else if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(r) : "r"(k1));
    else if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(k1), "r"(k2));
    else if (OP == HSET2) asm volatile("set.eq.f16x2.f16x2 %0, %0, %1;" : "+r"(r) : "r"(k1));
    else if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&r) : "f"(__uint_as_float(k1)), "f"(__uint_as_float(k2)));
    else if (OP == HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r) : "r"(k1));
    // DPX: packed signed 16-bit min / add-min-relu / three-input min, and the 32-bit scalar form
    else if (OP == VIMNMX16) r = __vmins2(r, k1 ^ r);
    else if (OP == VIADDMNMX16) r = __viaddmin_s16x2_relu(r, k1, k2);
    else if (OP == VIMNMX3_16) r = __vimin3_s16x2(r, k1, k2 ^ r);
    else r = (uint32_t)__vimin3_s32((int)r, (int)k1, (int)(k2 ^ r));
}
template <int A, int B, int C>
__global__ void __launch_bounds__(kThreads) issue_rate(int iters, uint32_t *sink, uint32_t seed)
{
    uint32_t r[9];
#pragma unroll
    for (int u = 0; u < 9; ++u) r[u] = seed * (threadIdx.x + 1) + u;
    const uint32_t k1 = seed | 0x3c003c00u, k2 = seed ^ 0x5555aaaau;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int u = 0; u < 9; u += 3) { one<A>(r[u], k1, k2); one<B>(r[u + 1], k1, k2); one<C>(r[u + 2], k1, k2); }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int u = 0; u < 9; ++u) acc ^= r[u];
    if (acc == 0x12345678u) sink[0] = acc;
}

template <class F>
static float time_ms(F launch, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch();                                    // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    uint32_t *sink;
    CK(cudaMalloc(&sink, 64));

    const int iters = 4096;
    const double smem_acc = (double)sms * kThreads * 8.0 * iters;      // accesses per launch
    auto set_smem = [&](const void *f) { return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes); };
    CK(set_smem((const void *)smem_read<4>)); CK(set_smem((const void *)smem_read<8>)); CK(set_smem((const void *)smem_read<16>));
    CK(set_smem((const void *)smem_write<4>)); CK(set_smem((const void *)smem_write<16>));
    const float r4 = time_ms([&] { smem_read<4><<<sms, kThreads, kSmemBytes>>>(iters, sink); }, 5);
    const float r8 = time_ms([&] { smem_read<8><<<sms, kThreads, kSmemBytes>>>(iters, sink); }, 5);
    const float r16 = time_ms([&] { smem_read<16><<<sms, kThreads, kSmemBytes>>>(iters, sink); }, 5);
    const float w4 = time_ms([&] { smem_write<4><<<sms, kThreads, kSmemBytes>>>(iters, sink); }, 5);
    const float w16 = time_ms([&] { smem_write<16><<<sms, kThreads, kSmemBytes>>>(iters, sink); }, 5);
    CK(cudaGetLastError());

    const size_t l2_bytes = 64u << 20;
    uint4 *buf;
    CK(cudaMalloc(&buf, l2_bytes));
    CK(cudaMemset(buf, 1, l2_bytes));
    const int passes = 16;
    const float l2 = time_ms([&] { l2_read<<<sms * 4, 512>>>(buf, l2_bytes / 16, passes, sink); }, 5);
    CK(cudaGetLastError());

    const double inst = (double)sms * (kThreads / 32) * 36.0 * iters;   // warp instructions per launch
    struct Mix { const char *name; float ms; };
    Mix mixes[] = {
#define MIX(a, b, c) { #a "+" #b "+" #c, time_ms([&] { issue_rate<a, b, c><<<sms, kThreads>>>(iters, sink, 3); }, 3) }
        MIX(LOP3, LOP3, LOP3), MIX(PRMT, PRMT, PRMT), MIX(HMNMX2, HMNMX2, HMNMX2), MIX(HSET2, HSET2, HSET2),
        MIX(HFMA2, HFMA2, HFMA2), MIX(HADD2, HADD2, HADD2), MIX(IMAD, IMAD, IMAD), MIX(FFMA, FFMA, FFMA), MIX(IADD3, IADD3, IADD3),
        MIX(LOP3, HFMA2, LOP3), MIX(LOP3, HFMA2, HFMA2), MIX(LOP3, HFMA2, IADD3), MIX(LOP3, HFMA2, FFMA), MIX(LOP3, FFMA, FFMA),
        MIX(HFMA2, FFMA, FFMA), MIX(HFMA2, IMAD, HFMA2), MIX(HMNMX2, HFMA2, HMNMX2), MIX(HSET2, LOP3, HSET2), MIX(HSET2, HFMA2, HSET2),
        MIX(HMNMX2, LOP3, HMNMX2), MIX(LOP3, IADD3, IADD3), MIX(HFMA2, IADD3, IADD3), MIX(LOP3, IMAD, LOP3),
        MIX(VIMNMX16, VIMNMX16, VIMNMX16), MIX(VIADDMNMX16, VIADDMNMX16, VIADDMNMX16), MIX(VIMNMX3_16, VIMNMX3_16, VIMNMX3_16),
        MIX(VIMNMX32, VIMNMX32, VIMNMX32), MIX(VIADDMNMX16, HFMA2, VIADDMNMX16), MIX(VIADDMNMX16, LOP3, VIADDMNMX16),
#undef MIX
    };
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());

    auto gbs = [](double bytes, float ms) { return bytes / (ms * 1e-3) / 1e9; };
    auto ginst = [](double n, float ms) { return n / (ms * 1e-3) / 1e9; };
    const double ghz = clock_khz * 1e-6;
    std::printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_nominal\": %.0f, "
                "\"smem_read_GBps\": {\"b32\": %.0f, \"b64\": %.0f, \"b128\": %.0f}, "
                "\"smem_write_GBps\": {\"b32\": %.0f, \"b128\": %.0f}, "
                "\"smem_read_B_per_clk_per_sm_at_nominal\": {\"b32\": %.1f, \"b64\": %.1f, \"b128\": %.1f}, "
                "\"l2_read_GBps\": %.0f, \"l2_buffer_MB\": %zu, "
                "\"issue_warpinst_per_clk_per_sm_at_nominal\": {",
                prop.name, sms, clock_khz * 1e-3,
                gbs(smem_acc * 4, r4), gbs(smem_acc * 8, r8), gbs(smem_acc * 16, r16),
                gbs(smem_acc * 4, w4), gbs(smem_acc * 16, w16),
                gbs(smem_acc * 4, r4) / ghz / sms, gbs(smem_acc * 8, r8) / ghz / sms, gbs(smem_acc * 16, r16) / ghz / sms,
                gbs((double)l2_bytes * passes, l2), l2_bytes >> 20);
    for (size_t k = 0; k < sizeof(mixes) / sizeof(mixes[0]); ++k)
        std::printf("%s\"%s\": %.2f", k ? ", " : "", mixes[k].name, ginst(inst, mixes[k].ms) / ghz / sms);
    std::printf("}}\n");
    return 0;
}

// Post-reconciliation kernels: privacy amplification and a per-frame confirmation CRC.
//
// Privacy amplification = EC/subcomponents/priv_amp.c:213-218: final key bit i is the parity of
//     XOR_j ( key_word[j] & prng_word[i * numwords + j] ),
// the PRNG words being successive 32-step outputs of the 32-bit LFSR of EC/subcomponents/rnd.c:118-127
// (state <<= 1, new bit = parity(state & 0xe0000200), EC/subcomponents/rnd.h:46), seeded per block (:186-187);
// the bits of the last key word beyond workbits are cleared first (:189-191); bit i of the final key is
// 1 << (31 - i % 32) of word i / 32 (EC/subcomponents/helpers.h:66-68).
//
// GPU mapping: one thread per final key bit.  The LFSR is linear over GF(2), so
//   * 32 steps are one word operation: with P = S ^ S<<1 ^ S<<2 ^ S<<22 (the taps 31,30,29,9 seen from the 32 new bits)
//     and Q = P ^ P>>10 ^ P>>20 ^ P>>30 (the lag-10 self reference), the next word is Q ^ Q>>30 ^ Q>>31;
//   * thread i jumps to word i*numwords with the precomputed matrices (T^32)^(2^b) (binary exponent, constant memory).
// The key words of the block sit in shared memory (<= 2048 words for ecd2's 65 535-bit blocks).
#include "kernels.hpp"

namespace qldpc {

namespace {

__host__ __device__ __forceinline__ uint32_t lfsr_step32(uint32_t s)
{
    const uint32_t p = s ^ (s << 1) ^ (s << 2) ^ (s << 22);
    const uint32_t q = p ^ (p >> 10) ^ (p >> 20) ^ (p >> 30);
    return q ^ (q >> 30) ^ (q >> 31);
}

struct PaJump {
    uint32_t col[32][32];   // col[b][k] = column k of (T^32)^(2^b)
};
__constant__ PaJump c_jump;

__device__ __forceinline__ uint32_t matvec(const uint32_t (&col)[32], uint32_t x)
{
    uint32_t y = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) y ^= ((x >> k) & 1u) ? col[k] : 0u;
    return y;
}

constexpr int kPaThreads = 256;

__global__ void __launch_bounds__(kPaThreads) privacy_amplify_kernel(const uint32_t *__restrict__ key, const int32_t *__restrict__ workbits,
                                                                     const int32_t *__restrict__ final_bits,
                                                                     const uint32_t *__restrict__ seeds, int key_stride,
                                                                     uint32_t *__restrict__ out, int out_stride)
{
    extern __shared__ uint32_t skey[];
    const int blk = blockIdx.y;
    const int wb = workbits[blk], nf = final_bits[blk];
    const int nw = (wb + 31) >> 5;
    if ((int)(blockIdx.x * kPaThreads) >= nf) return;
    for (int j = threadIdx.x; j < nw; j += kPaThreads) {
        uint32_t v = key[(size_t)blk * key_stride + j];
        if (j == nw - 1 && (wb & 31)) v &= 0xffffffffu << (32 - (wb & 31));   // priv_amp.c:189-191
        skey[j] = v;
    }
    __syncthreads();
    const int i = blockIdx.x * kPaThreads + threadIdx.x;
    uint32_t m = 0;
    if (i < nf) {
        uint32_t s = seeds[blk];
        unsigned long long k0 = (unsigned long long)i * (unsigned long long)nw;   // PRNG words consumed before bit i
        for (int b = 0; k0; ++b, k0 >>= 1)
            if (k0 & 1ull) s = matvec(c_jump.col[b], s);
#pragma unroll 4
        for (int j = 0; j < nw; ++j) {
            s = lfsr_step32(s);
            m ^= skey[j] & s;
        }
    }
    const uint32_t word = __brev(__ballot_sync(0xffffffffu, (__popc(m) & 1) != 0));
    if ((threadIdx.x & 31) == 0 && i < nf) out[(size_t)blk * out_stride + (i >> 5)] = word;
}

// CRC-32 (IEEE 802.3, reflected, the zlib / PNG polynomial 0xEDB88320) of each frame, taken over the frame's bytes in
// transmission order: MSB-first words -> big-endian byte order.  One thread per frame, table in shared memory.
__global__ void crc32_frames_kernel(const uint32_t *__restrict__ bits, int n_frames, int words_per_frame, int stride_words,
                                    uint32_t *__restrict__ crc_out)
{
    __shared__ uint32_t tab[256];
    for (int n = threadIdx.x; n < 256; n += blockDim.x) {
        uint32_t c = (uint32_t)n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
        tab[n] = c;
    }
    __syncthreads();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const uint32_t *p = bits + (size_t)f * stride_words;
    uint32_t c = 0xffffffffu;
    for (int w = 0; w < words_per_frame; ++w) {
        const uint32_t v = __ldg(p + w);
#pragma unroll
        for (int k = 3; k >= 0; --k) c = tab[(c ^ (v >> (8 * k))) & 0xffu] ^ (c >> 8);
    }
    crc_out[f] = c ^ 0xffffffffu;
}

}  // namespace

int pa_upload_jump_tables()
{
    static PaJump h;
    for (int k = 0; k < 32; ++k) h.col[0][k] = lfsr_step32(1u << k);
    for (int b = 1; b < 32; ++b)
        for (int k = 0; k < 32; ++k) {
            uint32_t x = h.col[b - 1][k], y = 0;   // column k of M^2 = M * (column k of M)
            for (int t = 0; t < 32; ++t)
                if ((x >> t) & 1u) y ^= h.col[b - 1][t];
            h.col[b][k] = y;
        }
    QLDPC_CUDA(cudaMemcpyToSymbol(c_jump, &h, sizeof(h)));
    return QLDPC_OK;
}

int launch_privacy_amplify(const uint32_t *d_key, const int32_t *d_workbits, const int32_t *d_final_bits, const uint32_t *d_seeds,
                           int n_blocks, int key_stride, int max_workbits, int max_final_bits, uint32_t *d_out, int out_stride,
                           cudaStream_t st)
{
    if (n_blocks <= 0 || max_final_bits <= 0) return QLDPC_OK;
    const int smem = ((max_workbits + 31) / 32) * 4;
    if (smem > 48 * 1024)
        QLDPC_CUDA(cudaFuncSetAttribute(privacy_amplify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const dim3 grid((max_final_bits + kPaThreads - 1) / kPaThreads, n_blocks);
    privacy_amplify_kernel<<<grid, kPaThreads, smem, st>>>(d_key, d_workbits, d_final_bits, d_seeds, key_stride, d_out, out_stride);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

int launch_crc32_frames(const uint32_t *d_bits, int n_frames, int words_per_frame, int stride_words, uint32_t *d_crc, cudaStream_t st)
{
    if (n_frames <= 0) return QLDPC_OK;
    crc32_frames_kernel<<<(n_frames + 127) / 128, 128, 0, st>>>(d_bits, n_frames, words_per_frame, stride_words, d_crc);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

}  // namespace qldpc

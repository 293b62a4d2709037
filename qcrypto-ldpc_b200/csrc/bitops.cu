// Bit-packed helpers around the decoders: syndrome, info-bit gather, LLR synthesis, NR encoder.
// All bit vectors are MSB-first in 32-bit words (errorcorrection/subcomponents/helpers.h:65-68).
#include "kernels.hpp"

namespace qldpc {

namespace {

__device__ __forceinline__ unsigned get_bit(const uint32_t *w, int i) { return (w[i >> 5] >> (31 - (i & 31))) & 1u; }

// syndrome = H * bits (ML/check_cword.m:9-19).  One thread per output word: 32 checks, each the
// XOR of its variables' bits; popc parity of the assembled word is what early termination tests.
__global__ void syndrome_csr_kernel(const uint32_t *__restrict__ bits, uint32_t *__restrict__ syn, int F, int M,
                                    int cw_words, int syn_words, const int32_t *__restrict__ row_ptr,
                                    const int32_t *__restrict__ col_idx)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)F * syn_words) return;
    const int f = (int)(t / syn_words), w = (int)(t - (long long)f * syn_words);
    const uint32_t *b = bits + (size_t)f * cw_words;
    uint32_t out = 0;
    for (int k = 0; k < 32; ++k) {
        const int m = 32 * w + k;
        if (m >= M) break;
        unsigned s = 0;
        for (int e = row_ptr[m]; e < row_ptr[m + 1]; ++e) s ^= get_bit(b, col_idx[e]);
        out |= (s & 1u) << (31 - k);
    }
    syn[t] = out;
}

// The same for quasi-cyclic codes with Z % 32 == 0: a check word is the XOR of one rotated 32-bit window per edge
// (two loads + funnel shift) instead of 32 x degree single-bit gathers.  One thread per output word.
__global__ void syndrome_qc_kernel(const uint32_t *__restrict__ bits, uint32_t *__restrict__ syn, int F, int brows, int ZW32,
                                   int cw_words, int syn_words, const QcLayer *__restrict__ layers,
                                   const QcEdgeAux *__restrict__ aux)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)F * syn_words) return;
    const int f = (int)(t / syn_words), ow = (int)(t - (long long)f * syn_words);
    const int r = ow / ZW32, w = ow - r * ZW32;
    if (r >= brows) return;
    const uint32_t *b = bits + (size_t)f * cw_words;
    const QcLayer ly = layers[r];
    uint32_t acc = 0;
    for (int e = 0; e < ly.degree; ++e) {
        const QcEdgeAux a = aux[ly.edge_begin + e];
        // check lane 32w + j reads variable lane (32w + j + shift) mod Z (ML/mul_sh.m:9); MSB-first words: a left funnel shift
        int w0 = w + (a.shift >> 5);
        if (w0 >= ZW32) w0 -= ZW32;
        const int w1 = w0 + 1 == ZW32 ? 0 : w0 + 1;
        acc ^= __funnelshift_l(__ldg(b + a.hdw + w1), __ldg(b + a.hdw + w0), a.shift & 31);
    }
    syn[t] = acc;
}

__global__ void gather_bits_kernel(const uint32_t *__restrict__ allbits, uint32_t *__restrict__ out, int F, int cw_words,
                                   int out_words, int K, const int32_t *__restrict__ info_pos)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)F * out_words) return;
    const int f = (int)(t / out_words), w = (int)(t - (long long)f * out_words);
    const uint32_t *b = allbits + (size_t)f * cw_words;
    uint32_t v = 0;
    for (int k = 0; k < 32; ++k) {
        const int i = 32 * w + k;
        if (i >= K) break;
        v |= get_bit(b, info_pos[i]) << (31 - k);
    }
    out[t] = v;
}

// Modem_OOK_BSC::demodulate + confirmed-parity / puncture override (BOOT/src/main.cpp:348-363):
// received 0 -> +mag, received 1 -> -mag.  One thread per 4 consecutive bits (one 32-bit store
// for int8 output).
template <typename T>
__global__ void make_llr_kernel(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ known,
                                const uint32_t *__restrict__ punct, T noisy, T known_mag, int F, int N, int cw_words,
                                T *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int groups = (N + 3) / 4;
    if (t >= (long long)F * groups) return;
    const int f = (int)(t / groups), g = (int)(t - (long long)f * groups);
    const uint32_t *b = bits + (size_t)f * cw_words;
    T *o = out + (size_t)f * N;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = 4 * g + k;
        if (i >= N) break;
        T mag = noisy;
        if (known && get_bit(known, i)) mag = known_mag;
        if (punct && get_bit(punct, i)) mag = (T)0;
        o[i] = get_bit(b, i) ? (T)(-mag) : mag;
    }
}

// int8 fast path of the same synthesis: one thread per 16 consecutive positions, one 128-bit store.
__device__ __forceinline__ uint32_t spread_nibble(uint32_t nib)   // bit 3..0 of nib -> 0x00/0xFF in bytes 0..3
{
    return (((nib * 0x08040201u) >> 3) & 0x01010101u) * 0xffu;
}
// 16 positions per item -> one 16-byte store.  In the host pipeline the kernel runs BESIDE a decode CTA of the other lane,
// which leaves 1024 registers on each of the four SM sub-partitions (layered_i8s.cu, kMaxRegs): one CTA of four warps
// at 32 registers per SM.  With so few warps the latency is hidden by loads in flight, not by occupancy: a CTA lives for
// kLlrRep rounds of kLlrUnroll independent loads per thread.  (A one-warp CTA with more registers does not get placed
// beside a decoder that uses 128 registers either: measured.)
constexpr int kLlrBlock = 128, kLlrUnroll = 4, kLlrRep = 4;
__global__ void __launch_bounds__(kLlrBlock, 16)
make_llr_i8x16_kernel(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ known, const uint32_t *__restrict__ punct,
                      int noisy, int known_mag, int F, int groups, int cw_words, uint4 *__restrict__ out)
{
    const long long total = (long long)F * groups;
    const uint32_t posN = (uint32_t)(noisy & 0xff) * 0x01010101u, negN = (uint32_t)(-noisy & 0xff) * 0x01010101u;
    const uint32_t posK = (uint32_t)(known_mag & 0xff) * 0x01010101u, negK = (uint32_t)(-known_mag & 0xff) * 0x01010101u;
    long long t0 = (long long)blockIdx.x * (kLlrBlock * kLlrUnroll * kLlrRep) + threadIdx.x;
#pragma unroll 1
    for (int rep = 0; rep < kLlrRep; ++rep, t0 += kLlrBlock * kLlrUnroll) {
        uint32_t b[kLlrUnroll];
        int g[kLlrUnroll];
#pragma unroll
        for (int u = 0; u < kLlrUnroll; ++u) {
            const long long t = t0 + kLlrBlock * u;
            b[u] = 0;
            g[u] = 0;
            if (t < total) {
                const int f = (int)(t / groups);
                g[u] = (int)(t - (long long)f * groups);
                b[u] = __ldg(bits + (size_t)f * cw_words + (g[u] >> 1));
            }
        }
#pragma unroll
        for (int u = 0; u < kLlrUnroll; ++u) {
            const long long t = t0 + kLlrBlock * u;
            if (t >= total) break;
            const int sh = (g[u] & 1) ? 0 : 16;                 // MSB-first words: positions 32w .. 32w+15 are bits 31 .. 16
            const uint32_t bb = (b[u] >> sh) & 0xffffu;
            const uint32_t k = known ? (__ldg(known + (g[u] >> 1)) >> sh) & 0xffffu : 0u;
            const uint32_t p = punct ? (__ldg(punct + (g[u] >> 1)) >> sh) & 0xffffu : 0u;
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t S = spread_nibble((bb >> (12 - 4 * q)) & 0xfu);
                const uint32_t K = spread_nibble((k >> (12 - 4 * q)) & 0xfu), P = spread_nibble((p >> (12 - 4 * q)) & 0xfu);
                const uint32_t rN = (posN & ~S) | (negN & S), rK = (posK & ~S) | (negK & S);
                w[q] = ((rN & ~K) | (rK & K)) & ~P;
            }
            out[t] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// 5G-NR double-diagonal encoder (ML/nrldpc_encode.m:12-40), one CTA per frame, bits unpacked in
// shared memory as bytes: cword = [msg | p1 | p2 p3 p4 | extension parities].
__global__ void encode_nr_kernel(const uint32_t *__restrict__ msg, uint32_t *__restrict__ cword, int F, int Z, int brows,
                                 int bcols, const int32_t *__restrict__ base, int msg_words, int cw_words)
{
    extern __shared__ unsigned char cw[];   // bcols*Z bytes + Z bytes temp
    const int kb = bcols - brows, N = bcols * Z, K = kb * Z;
    unsigned char *temp = cw + N;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int f = blockIdx.x; f < F; f += gridDim.x) {
        const uint32_t *mf = msg + (size_t)f * msg_words;
        for (int i = tid; i < N; i += nt) cw[i] = i < K ? (unsigned char)get_bit(mf, i) : 0;
        __syncthreads();
        // :18-23 temp = sum over rows 0..3 and message columns of mul_sh(msg_j, B(i,j))
        for (int l = tid; l < Z; l += nt) {
            unsigned s = 0;
            for (int r = 0; r < 4; ++r)
                for (int j = 0; j < kb; ++j) {
                    const int sh = base[r * bcols + j];
                    if (sh >= 0) s ^= cw[j * Z + (l + sh) % Z];
                }
            temp[l] = (unsigned char)(s & 1u);
        }
        __syncthreads();
        // :24-29 p1 = mul_sh(temp, z - p1_sh)
        const int p1_sh = base[1 * bcols + kb] == -1 ? base[2 * bcols + kb] : base[1 * bcols + kb];
        for (int l = tid; l < Z; l += nt) cw[K + l] = temp[(l + Z - p1_sh) % Z];
        __syncthreads();
        // :30-37 p2..p4 (each needs the previous one)
        for (int r = 0; r < 3; ++r) {
            for (int l = tid; l < Z; l += nt) {
                unsigned s = 0;
                for (int j = 0; j < kb + r + 1; ++j) {
                    const int sh = base[r * bcols + j];
                    if (sh >= 0) s ^= cw[j * Z + (l + sh) % Z];
                }
                cw[(kb + r + 1) * Z + l] = (unsigned char)(s & 1u);
            }
            __syncthreads();
        }
        // :38-45 extension parities, independent of each other
        for (int t = tid; t < (brows - 4) * Z; t += nt) {
            const int r = 4 + t / Z, l = t % Z;
            unsigned s = 0;
            for (int j = 0; j < kb + 4; ++j) {
                const int sh = base[r * bcols + j];
                if (sh >= 0) s ^= cw[j * Z + (l + sh) % Z];
            }
            cw[(kb + r) * Z + l] = (unsigned char)(s & 1u);
        }
        __syncthreads();
        uint32_t *of = cword + (size_t)f * cw_words;
        for (int w = tid; w < cw_words; w += nt) {
            uint32_t v = 0;
            for (int b = 0; b < 32; ++b) {
                const int i = 32 * w + b;
                if (i < N) v |= (uint32_t)cw[i] << (31 - b);
            }
            of[w] = v;
        }
        __syncthreads();
    }
}

// The same encoder on PACKED words for lifting sizes that are a multiple of 32 (every size the packed decoders take): a block
// column is Z / 32 MSB-first words, mul_sh (ML/mul_sh.m:9) is a word rotation plus one funnel shift, and a thread owns one
// word of every vector.  blockDim = (Z / 32, frames per CTA); the message and the four core parity columns of a frame sit in
// shared memory, the extension parities go straight to the output.  Same steps and the same order of dependencies as
// encode_nr_kernel above (ML/nrldpc_encode.m:18-45); the byte-per-bit version walks all kb + 4 columns of the base matrix
// for every parity BIT (18.8 ms for 65 536 BG1 Z=384 frames, slower than decoding them), this one walks the row's edges
// for every parity WORD.
__global__ void encode_nr_packed_kernel(const uint32_t *__restrict__ msg, uint32_t *__restrict__ cword, int F, int ZW, int brows,
                                        int bcols, int p1_rot, int msg_words, int cw_words, const QcLayer *__restrict__ layers,
                                        const QcEdgeAux *__restrict__ aux)
{
    extern __shared__ uint32_t esm[];
    const int kb = bcols - brows, ncore = kb + 4, w = threadIdx.x, fs = threadIdx.y, FP = blockDim.y;
    uint32_t *cw = esm + (size_t)fs * (ncore + 1) * ZW, *temp = cw + ncore * ZW;
    // word w of vector v (ZW words) rotated by `shift` lanes: result lane l = v lane (l + shift) mod Z
    auto rot = [&](const uint32_t *v, int shift) {
        int w0 = w + (shift >> 5);
        if (w0 >= ZW) w0 -= ZW;
        const int w1 = w0 + 1 == ZW ? 0 : w0 + 1;
        return __funnelshift_l(v[w1], v[w0], shift & 31);
    };
    for (int f0 = blockIdx.x * FP; f0 < F; f0 += gridDim.x * FP) {
        const int f = f0 + fs;
        const bool act = f < F;
        if (act)
            for (int j = 0; j < kb; ++j) cw[j * ZW + w] = __ldg(msg + (size_t)f * msg_words + j * ZW + w);
        __syncthreads();
        uint32_t lam[4] = {0u, 0u, 0u, 0u};           // :18-23 message part of the four core rows
        if (act) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const QcLayer ly = layers[r];
                for (int e = 0; e < ly.degree; ++e) {
                    const QcEdgeAux a = aux[ly.edge_begin + e];
                    if (a.col < kb) lam[r] ^= rot(cw + a.col * ZW, a.shift);
                }
            }
            temp[w] = lam[0] ^ lam[1] ^ lam[2] ^ lam[3];
        }
        __syncthreads();
        if (act) cw[kb * ZW + w] = rot(temp, p1_rot);   // :24-29 p1 = mul_sh(temp, z - p1_sh)
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 3; ++r) {                 // :30-37 p2..p4, each needs the previous one
            if (act) {
                uint32_t acc = lam[r];
                const QcLayer ly = layers[r];
                for (int e = 0; e < ly.degree; ++e) {
                    const QcEdgeAux a = aux[ly.edge_begin + e];
                    if (a.col >= kb && a.col < kb + r + 1) acc ^= rot(cw + a.col * ZW, a.shift);
                }
                cw[(kb + r + 1) * ZW + w] = acc;
            }
            __syncthreads();
        }
        if (act) {
            uint32_t *of = cword + (size_t)f * cw_words;
            for (int j = 0; j < ncore; ++j) of[j * ZW + w] = cw[j * ZW + w];
            for (int r = 4; r < brows; ++r) {         // :38-45 extension parities, independent of each other
                uint32_t acc = 0;
                const QcLayer ly = layers[r];
                for (int e = 0; e < ly.degree; ++e) {
                    const QcEdgeAux a = aux[ly.edge_begin + e];
                    if (a.col < ncore) acc ^= rot(cw + a.col * ZW, a.shift);
                }
                of[(kb + r) * ZW + w] = acc;
            }
        }
        __syncthreads();
    }
}

// Per-position LLR magnitudes (one byte each, shared by all frames) for the decoders that synthesise their LLRs from key
// bits themselves (layered_i8s.cu, bit input): noisy / known / punctured as in make_llr_kernel.
__global__ void make_mag_i8_kernel(const uint32_t *__restrict__ known, const uint32_t *__restrict__ punct, int noisy,
                                   int known_mag, int N, uint8_t *__restrict__ mag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int m = noisy;
    if (known && get_bit(known, i)) m = known_mag;
    if (punct && get_bit(punct, i)) m = 0;
    mag[i] = (uint8_t)m;
}

inline int grid_for(long long items, int block) { return (int)((items + block - 1) / block); }

}  // namespace

int launch_make_mag_i8(const uint32_t *known, const uint32_t *punct, int noisy, int known_mag, int N, uint8_t *mag, cudaStream_t st)
{
    make_mag_i8_kernel<<<grid_for(N, 256), 256, 0, st>>>(known, punct, noisy, known_mag, N, mag);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

int launch_syndrome_qc(const uint32_t *bits, uint32_t *syn, int F, int brows, int Z, int cw_words, int syn_words,
                       const QcLayer *layers, const QcEdgeAux *aux, cudaStream_t st)
{
    if (F <= 0) return QLDPC_OK;
    syndrome_qc_kernel<<<grid_for((long long)F * syn_words, 128), 128, 0, st>>>(bits, syn, F, brows, Z / 32, cw_words, syn_words,
                                                                                 layers, aux);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

int launch_syndrome_csr(const uint32_t *bits, uint32_t *syn, int F, int N, int M, int cw_words, int syn_words,
                        const int32_t *row_ptr, const int32_t *col_idx, cudaStream_t st)
{
    (void)N;
    if (F <= 0) return QLDPC_OK;
    syndrome_csr_kernel<<<grid_for((long long)F * syn_words, 128), 128, 0, st>>>(bits, syn, F, M, cw_words, syn_words,
                                                                                  row_ptr, col_idx);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

int launch_gather_bits(const uint32_t *allbits, uint32_t *out, int F, int cw_words, int out_words, int K,
                       const int32_t *info_pos, cudaStream_t st)
{
    if (F <= 0) return QLDPC_OK;
    gather_bits_kernel<<<grid_for((long long)F * out_words, 128), 128, 0, st>>>(allbits, out, F, cw_words, out_words, K,
                                                                                info_pos);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

int launch_make_llr(const uint32_t *bits, const uint32_t *known, const uint32_t *punct, float noisy, float known_mag,
                    int F, int N, int cw_words, int dtype, void *llr_out, cudaStream_t st)
{
    if (F <= 0) return QLDPC_OK;
    const int grid = grid_for((long long)F * ((N + 3) / 4), 256);
    switch (dtype) {
    case QLDPC_DTYPE_F32:
        make_llr_kernel<float><<<grid, 256, 0, st>>>(bits, known, punct, noisy, known_mag, F, N, cw_words, (float *)llr_out);
        break;
    case QLDPC_DTYPE_I16:
        make_llr_kernel<int16_t><<<grid, 256, 0, st>>>(bits, known, punct, (int16_t)lrintf(noisy), (int16_t)lrintf(known_mag),
                                                       F, N, cw_words, (int16_t *)llr_out);
        break;
    case QLDPC_DTYPE_I8:
        if (N % 16 == 0 && (reinterpret_cast<uintptr_t>(llr_out) & 15) == 0) {
            const int groups = N / 16;
            make_llr_i8x16_kernel<<<grid_for((long long)F * groups, kLlrBlock * kLlrUnroll * kLlrRep), kLlrBlock, 0, st>>>(
                bits, known, punct, (int)(int8_t)lrintf(noisy), (int)(int8_t)lrintf(known_mag), F, groups, cw_words,
                (uint4 *)llr_out);
            break;
        }
        make_llr_kernel<int8_t><<<grid, 256, 0, st>>>(bits, known, punct, (int8_t)lrintf(noisy), (int8_t)lrintf(known_mag), F,
                                                      N, cw_words, (int8_t *)llr_out);
        break;
    default: return QLDPC_ERR_UNSUPPORTED;
    }
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

int launch_encode_nr(const uint32_t *msg, uint32_t *cword, int F, int Z, int brows, int bcols, const int32_t *base,
                     int msg_words, int cw_words, cudaStream_t st)
{
    if (F <= 0) return QLDPC_OK;
    const int smem = bcols * Z + Z;
    QLDPC_CUDA(cudaFuncSetAttribute(encode_nr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int block = ((Z + 31) / 32) * 32;
    if (block > 1024) block = 1024;
    if (block < 128) block = 128;
    const int grid = F < 148 * 8 ? F : 148 * 8;
    encode_nr_kernel<<<grid, block, smem, st>>>(msg, cword, F, Z, brows, bcols, base, msg_words, cw_words);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

// Z % 32 == 0: packed words (p1_rot = (Z - p1_sh) % Z, the rotation that turns the sum of the core rows into p1)
int launch_encode_nr_packed(const uint32_t *msg, uint32_t *cword, int F, int Z, int brows, int bcols, int p1_rot, int msg_words,
                            int cw_words, const QcLayer *layers, const QcEdgeAux *aux, cudaStream_t st)
{
    if (F <= 0) return QLDPC_OK;
    const int ZW = Z / 32, ncore = bcols - brows + 4;
    const int fp = std::max(1, std::min(32, 256 / ZW));
    const int smem = fp * (ncore + 1) * ZW * 4;
    if (smem > 48 * 1024) QLDPC_CUDA(cudaFuncSetAttribute(encode_nr_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int groups = (F + fp - 1) / fp;
    encode_nr_packed_kernel<<<std::min(groups, 148 * 8), dim3(ZW, fp), smem, st>>>(msg, cword, F, ZW, brows, bcols, p1_rot, msg_words,
                                                                                     cw_words, layers, aux);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

}  // namespace qldpc

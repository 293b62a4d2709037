// Post-reconciliation kernels: privacy amplification and a per-frame confirmation CRC.
//
// Privacy amplification = EC/subcomponents/priv_amp.c:213-218: final key bit i is the parity of
//     XOR_j ( key_word[j] & prng_word[i * numwords + j] ),
// the PRNG words being successive 32-step outputs of the 32-bit LFSR of EC/subcomponents/rnd.c:118-127
// (state <<= 1, new bit = parity(state & 0xe0000200), EC/subcomponents/rnd.h:46), seeded per block (:186-187);
// the bits of the last key word beyond workbits are cleared first (:189-191); bit i of the final key is
// 1 << (31 - i % 32) of word i / 32 (EC/subcomponents/helpers.h:66-68).
//
// GPU mapping.  The LFSR is linear over GF(2): with T the 32 x 32 matrix of 32 steps (one PRNG word), the word used for key
// word j of final bit i is T^(i nw + j + 1) s0, and
//     bit i = sum_j < key_j , T^(i nw + j + 1) s0 >  =  < w , U^i s0 >,   w = sum_j (T^t)^(j+1) key_j,   U = T^nw.
// So a block costs ONE pass over its key (w: nw products with the transposed step matrix, Horner order, split over the threads
// and recombined with the jump tables) plus a 32-bit matrix-vector product per final bit -- about 3e6 instructions per
// 65 535-bit block instead of the 1e9 of the bit-serial definition (40 000 final bits x 2 048 PRNG words), which the first
// version of this kernel executed literally (one thread per final bit, 28 ms for 512 blocks; now the call is its copies).
// Tables: T^(2^b) and its transpose in constant memory (binary jumps), U^(2^b) per block in shared memory.
#include <mutex>

#include "kernels.hpp"

namespace qldpc {

namespace {

__host__ __device__ __forceinline__ uint32_t lfsr_step32(uint32_t s)
{
    // 32 steps as one word operation: P = S ^ S<<1 ^ S<<2 ^ S<<22 (the taps 31,30,29,9 seen from the 32 new bits),
    // Q = P ^ P>>10 ^ P>>20 ^ P>>30 (the lag-10 self reference), next word = Q ^ Q>>30 ^ Q>>31
    const uint32_t p = s ^ (s << 1) ^ (s << 2) ^ (s << 22);
    const uint32_t q = p ^ (p >> 10) ^ (p >> 20) ^ (p >> 30);
    return q ^ (q >> 30) ^ (q >> 31);
}

struct PaJump {
    uint32_t col[32][32];    // col[b][k]  = column k of T^(2^b)
    uint32_t colT[32][32];   // colT[b][k] = column k of its transpose
};
__constant__ PaJump c_jump;

__device__ __forceinline__ uint32_t matvec(const uint32_t *col, uint32_t x)
{
    uint32_t y = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) y ^= ((x >> k) & 1u) ? col[k] : 0u;
    return y;
}

constexpr int kPaThreads = 256;
constexpr int kPaPow = 22;   // U^(2^b) for b < 22: final bits per block below 2^22

__global__ void __launch_bounds__(kPaThreads) privacy_amplify_kernel(const uint32_t *__restrict__ key, const int32_t *__restrict__ workbits,
                                                                     const int32_t *__restrict__ final_bits,
                                                                     const uint32_t *__restrict__ seeds, int key_stride,
                                                                     uint32_t *__restrict__ out, int out_stride)
{
    __shared__ uint32_t upow[kPaPow][32];
    __shared__ uint32_t wred[kPaThreads / 32];
    const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int wb = workbits[blk], nf = final_bits[blk];
    const int nw = (wb + 31) >> 5;
    // ---- w = sum_j (T^t)^(j+1) key_j: thread t takes the words [a, b), Horner from the last, then the jump by a words
    const int per = (nw + kPaThreads - 1) / kPaThreads;
    const int a = tid * per, b = min(nw, a + per);
    uint32_t acc = 0;
    for (int j = b - 1; j >= a; --j) {
        uint32_t v = key[(size_t)blk * key_stride + j];
        if (j == nw - 1 && (wb & 31)) v &= 0xffffffffu << (32 - (wb & 31));   // priv_amp.c:189-191
        acc = matvec(c_jump.colT[0], acc ^ v);
    }
    if (a < b)
        for (int bit = 0, e = a; e; ++bit, e >>= 1)
            if (e & 1) acc = matvec(c_jump.colT[bit], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) wred[tid >> 5] = acc;
    // ---- U = T^nw (column `lane`: the commuting factors T^(2^b) applied to the unit vector), then its powers by squaring
    if (tid < 32) {
        uint32_t c = 1u << lane;
        for (int bit = 0, e = nw; e; ++bit, e >>= 1)
            if (e & 1) c = matvec(c_jump.col[bit], c);
        upow[0][lane] = c;
        __syncwarp();
        for (int p = 1; p < kPaPow; ++p) {
            c = matvec(upow[p - 1], upow[p - 1][lane]);
            upow[p][lane] = c;
            __syncwarp();
        }
    }
    __syncthreads();
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < kPaThreads / 32; ++k) w ^= wred[k];
    // ---- final bits: one thread per output word (32 consecutive bits)
    const uint32_t s0 = seeds[blk];
    for (int ow = tid; 32 * ow < nf; ow += kPaThreads) {
        uint32_t u = s0;
        for (int bit = 0, e = 32 * ow; e; ++bit, e >>= 1)
            if (e & 1) u = matvec(upow[bit], u);
        uint32_t word = 0;
        for (int r = 0; r < 32 && 32 * ow + r < nf; ++r) {
            word |= (uint32_t)(__popc(w & u) & 1) << (31 - r);     // helpers.h:66-68
            u = matvec(upow[0], u);
        }
        out[(size_t)blk * out_stride + ow] = word;
    }
}

// CRC-32 (IEEE 802.3, reflected, the zlib / PNG polynomial 0xEDB88320) of each frame, taken over the frame's bytes in
// transmission order: MSB-first words -> big-endian byte order.
__device__ __forceinline__ uint32_t crc_word(const uint32_t *tab, uint32_t c, uint32_t v)
{
#pragma unroll
    for (int k = 3; k >= 0; --k) c = tab[(c ^ (v >> (8 * k))) & 0xffu] ^ (c >> 8);
    return c;
}

// One thread per frame, byte-serial (any frame length).
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
    for (int w = 0; w < words_per_frame; ++w) c = crc_word(tab, c, __ldg(p + w));
    crc_out[f] = c ^ 0xffffffffu;
}

// One WARP per frame.  The CRC register is linear in (state, data): with Z_n the map "feed n zero bytes",
//     crc(A || B, init) = Z_|B|( crc(A, init) ) ^ crc(B, 0),
// so every lane runs the byte loop over its own contiguous chunk from state 0, advances the result over the bytes that follow
// its chunk with the binary jump tables Z_(2^b) (constant memory), and the warp XORs the 32 contributions together with the
// image of the initial state, Z_L(0xffffffff) (the same for every frame: a kernel argument).  The frame is read once, coalesced,
// through shared memory.  (The thread-per-frame kernel above walks 1 056 dependent table look-ups per frame with a stride of a
// whole frame between the lanes of a warp: 17 GB/s on 65 536 frames.)
struct CrcJump { uint32_t col[24][32]; };     // col[b][k] = column k of Z_(2^b bytes)
__constant__ CrcJump c_crc;
constexpr int kCrcWarps = 4, kCrcMaxWords = 2048;

__global__ void __launch_bounds__(32 * kCrcWarps) crc32_frames_warp_kernel(const uint32_t *__restrict__ bits, int n_frames,
                                                                          int words_per_frame, int stride_words, uint32_t init_image,
                                                                          uint32_t *__restrict__ crc_out)
{
    __shared__ uint32_t tab[256];
    extern __shared__ uint32_t fr[];              // kCrcWarps frames of words_per_frame words
    for (int n = threadIdx.x; n < 256; n += blockDim.x) {
        uint32_t c = (uint32_t)n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
        tab[n] = c;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t *my = fr + wid * words_per_frame;
    const int per = (words_per_frame + 31) / 32;
    const int a = min(words_per_frame, lane * per), b = min(words_per_frame, a + per);
    for (int f = blockIdx.x * kCrcWarps + wid; f < n_frames; f += gridDim.x * kCrcWarps) {
        const uint32_t *p = bits + (size_t)f * stride_words;
        for (int w = lane; w < words_per_frame; w += 32) my[w] = __ldg(p + w);
        __syncwarp();
        uint32_t c = 0;
        for (int w = a; w < b; ++w) c = crc_word(tab, c, my[w]);
        for (int bit = 0, e = 4 * (words_per_frame - b); e; ++bit, e >>= 1)      // bytes after this lane's chunk
            if (e & 1) c = matvec(c_crc.col[bit], c);
#pragma unroll
        for (int o = 16; o; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) crc_out[f] = c ^ init_image ^ 0xffffffffu;
        __syncwarp();
    }
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
    for (int b = 0; b < 32; ++b)
        for (int k = 0; k < 32; ++k) {             // column k of the transpose = row k: bit k of every column
            uint32_t r = 0;
            for (int m = 0; m < 32; ++m) r |= ((h.col[b][m] >> k) & 1u) << m;
            h.colT[b][k] = r;
        }
    QLDPC_CUDA(cudaMemcpyToSymbol(c_jump, &h, sizeof(h)));
    return QLDPC_OK;
}

int launch_privacy_amplify(const uint32_t *d_key, const int32_t *d_workbits, const int32_t *d_final_bits, const uint32_t *d_seeds,
                           int n_blocks, int key_stride, int max_workbits, int max_final_bits, uint32_t *d_out, int out_stride,
                           cudaStream_t st)
{
    (void)max_workbits;
    if (n_blocks <= 0 || max_final_bits <= 0) return QLDPC_OK;
    if (max_final_bits >= (1 << kPaPow)) return QLDPC_ERR_UNSUPPORTED;
    privacy_amplify_kernel<<<n_blocks, kPaThreads, 0, st>>>(d_key, d_workbits, d_final_bits, d_seeds, key_stride, d_out, out_stride);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

static uint32_t crc_zero_byte(uint32_t c)
{
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;    // one zero byte through the reflected register
    return c;
}

int launch_crc32_frames(const uint32_t *d_bits, int n_frames, int words_per_frame, int stride_words, uint32_t *d_crc, cudaStream_t st)
{
    if (n_frames <= 0) return QLDPC_OK;
    if (words_per_frame < 32 || words_per_frame > kCrcMaxWords) {
        crc32_frames_kernel<<<(n_frames + 127) / 128, 128, 0, st>>>(d_bits, n_frames, words_per_frame, stride_words, d_crc);
        QLDPC_CUDA(cudaGetLastError());
        return QLDPC_OK;
    }
    // jump tables Z_(2^b bytes), once per device; the image of the initial state under Z_L, per call
    static CrcJump h;
    static std::once_flag built;
    static thread_local int tables_on = -1;
    std::call_once(built, [] {
        for (int k = 0; k < 32; ++k) h.col[0][k] = crc_zero_byte(1u << k);
        for (int b = 1; b < 24; ++b)
            for (int k = 0; k < 32; ++k) {
                uint32_t x = h.col[b - 1][k], y = 0;
                for (int t = 0; t < 32; ++t)
                    if ((x >> t) & 1u) y ^= h.col[b - 1][t];
                h.col[b][k] = y;
            }
    });
    int dev = 0;
    QLDPC_CUDA(cudaGetDevice(&dev));
    if (tables_on != dev) {
        QLDPC_CUDA(cudaMemcpyToSymbol(c_crc, &h, sizeof(h)));
        tables_on = dev;
    }
    uint32_t init_image = 0xffffffffu;
    for (int bit = 0, e = 4 * words_per_frame; e; ++bit, e >>= 1)
        if (e & 1) {
            uint32_t y = 0;
            for (int t = 0; t < 32; ++t)
                if ((init_image >> t) & 1u) y ^= h.col[bit][t];
            init_image = y;
        }
    const int smem = kCrcWarps * words_per_frame * 4;
    if (smem > 40 * 1024)
        QLDPC_CUDA(cudaFuncSetAttribute(crc32_frames_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int groups = (n_frames + kCrcWarps - 1) / kCrcWarps;
    crc32_frames_warp_kernel<<<std::min(groups, 148 * 16), 32 * kCrcWarps, smem, st>>>(d_bits, n_frames, words_per_frame, stride_words,
                                                                                      init_image, d_crc);
    QLDPC_CUDA(cudaGetLastError());
    return QLDPC_OK;
}

}  // namespace qldpc

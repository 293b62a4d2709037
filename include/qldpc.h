/*
 * qldpc.h -- C ABI of the B200 LDPC reconciliation engine (libqldpc_b200.so).
 *
 * This is the drop-in boundary for the LDPC decode path of JarryChou/qcrypto-ldpc.
 * Plain C: opaque handles, pointers and sizes; every call returns an int
 * (0 = ok, >0 = error code, see qldpc_strerror) -- the convention of the ecd2 packet
 * handlers (errorcorrection/definitions/algorithms/packet_manager.h:33, error table
 * errorcorrection/ecd2.h:251-337).  No exceptions, no globals, callable from C.
 *
 * Reference interfaces each entry point replaces (paths relative to the reference root;
 * BOOT = errorcorrection/ldpc_examples/my_project_with_aff3ct/examples/bootstrap,
 * VAR  = BOOT/src/variants (copy out as main.cpp to use),
 * ML   = errorcorrection/ldpc_examples/matlab_code_Base_matrices/matlab_code & Base_matrices):
 *
 *   qldpc_code_from_alist_file   tools::LDPC_matrix_handler::read      VAR/main.cpp (alist):340
 *   qldpc_code_from_qc_file      tools::LDPC_matrix_handler::read      VAR/main.cpp (5g-qc):389
 *   qldpc_code_from_qc           `load base_matrices/NR_1_1_24.txt`    ML/BPSK_nrldpc_sim_FP.m:8-11
 *   qldpc_code_set_info_bits_pos info_bits_pos ctor argument           BOOT/src/main.cpp:193
 *   qldpc_decoder_create         Decoder_LDPC_BP_flooding<..> ctor     BOOT/src/main.cpp:193
 *                                Decoder_LDPC_BP_horizontal_layered    VAR/main.cpp (5g-qc):256-270
 *                                decoder constants                     ML/BPSK_nrldpc_sim_FP.m:1-6
 *   qldpc_decode[_device]        decoder->decode_siho(LLRs, dec_bits)  BOOT/src/main.cpp:365
 *                                layered fixed-point decode loop       ML/BPSK_nrldpc_sim_FP.m:35-94
 *   qldpc_syndrome[_device]      check_cword(B,z,c)                    ML/check_cword.m:9-19
 *   qldpc_make_llr[_device]      Modem_OOK_BSC::demodulate + parity/puncture override
 *                                                                      BOOT/src/main.cpp:348-363
 *   qldpc_decode_bits[_device]   the two steps above fused for callers that hold key BITS: the ecd2
 *                                handlers work on pb->mainBufPtr       errorcorrection/definitions/processblock.h:105
 *   qldpc_encode_nr[_device]     nrldpc_encode(B,z,msg)                ML/nrldpc_encode.m:12-40
 *   qldpc_get_stats              Monitor_BFER / FER bookkeeping        BOOT/src/main.cpp:366-388
 *   (decoder.reset(), BOOT/src/main.cpp:389, has no counterpart: every decode call starts
 *    from zero check-to-variable messages.)
 *
 * Bit vectors are packed MSB-first in 32-bit words, each frame padded to whole words --
 * the ecd2 convention (errorcorrection/subcomponents/helpers.h:65-68):
 *   bit i of a frame lives in word i/32 under mask 1u << (31 - i%32).
 *
 * LLR sign convention: LLR >= 0 <=> bit 0 (Modem_OOK_BSC, BOOT/src/main.cpp:353).
 *
 * There is no CPU fallback: every compute entry point needs a CUDA device of compute
 * capability 10.x and returns QLDPC_ERR_NO_DEVICE / QLDPC_ERR_CUDA otherwise.
 */
#ifndef QLDPC_H
#define QLDPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QLDPC_VERSION 200

/* error codes */
enum {
    QLDPC_OK = 0,
    QLDPC_ERR_ARG = 1,         /* bad argument / null pointer / out of range           */
    QLDPC_ERR_IO = 2,          /* file could not be opened                              */
    QLDPC_ERR_FORMAT = 3,      /* matrix file malformed                                 */
    QLDPC_ERR_NOMEM = 4,       /* host or device allocation failed                      */
    QLDPC_ERR_CUDA = 5,        /* CUDA runtime error (see qldpc_last_cuda_error)        */
    QLDPC_ERR_UNSUPPORTED = 6, /* combination of schedule / rule / dtype not available  */
    QLDPC_ERR_NO_DEVICE = 7    /* no sm_100 device visible                              */
};

/* schedules, update rules, message types (names follow the AFF3CT driver's decoderTypeNames,
 * VAR/main.cpp (5g-qc):117-148) */
enum { QLDPC_SCHED_FLOODING = 0, QLDPC_SCHED_LAYERED = 1 };
enum { QLDPC_RULE_SPA = 0, QLDPC_RULE_NMS = 1, QLDPC_RULE_OMS = 2 };
enum { QLDPC_DTYPE_F32 = 0, QLDPC_DTYPE_I16 = 1, QLDPC_DTYPE_I8 = 2 };
enum { QLDPC_OUT_INFO = 0, QLDPC_OUT_ALL = 1 };

#define QLDPC_MAX_DEVICES 8
/* decoder flags */
#define QLDPC_FLAG_L2_PERSIST     1u  /* int8 layered decoder for Z % 128 == 0: keep the check-to-variable message scratch
                                         resident in L2.  OPT-IN because it has side effects outside the decoder: creation raises
                                         the device-wide cudaLimitPersistingL2CacheSize and the decode calls set
                                         cudaStreamAttributeAccessPolicyWindow on the stream they run on.                */
#define QLDPC_FLAG_LI8_RESIDENT   2u  /* diagnostics: int8 layered decoding on the previous-generation kernel (layered_i8.cu)  */
#define QLDPC_FLAG_LI8_STREAM     4u  /*   with its messages resident in shared memory / streamed through an L2 scratch   */
#define QLDPC_FLAG_NO_FUSED_BITS  8u  /* diagnostics: qldpc_decode_bits runs LLR synthesis and decoding as two kernels      */
#define QLDPC_FLAG_FAST_SPA      32u  /* float SPA (flooding and layered kernels): tanh / atanh in fp32 on the
                                         special-function units instead of double.  Decoded bits are unchanged on all test
                                         batches, iteration counts on all but about one frame in 3e4; posteriors stay within
                                         1e-3 except for saturated messages (|m| > 14), which may move by ln 2 (one last-bit
                                         difference in 1 - r)                                                                 */
#define QLDPC_FLAG_NO_ZERO_COPY  16u  /* diagnostics: qldpc_decode_bits stages pinned host buffers through device copies too */
#define QLDPC_FLAG_DISCARD_SCRATCH 64u /* int8 layered decoder for Z % 128 == 0: when a frame ends, drop the (dirty, dead) L2 lines
                                         of its message scratch with discard.global.L2 instead of letting them be written back.
                                         Measured on B200, BG1 Z=384 at QBER 3 %: DRAM traffic 194 -> 80 KB per frame (writes
                                         115 -> 22 KB), throughput -1.4 % (the kernel is not HBM-bound): opt-in, for a decoder
                                         that shares the device's memory bandwidth with other work.  Results are unchanged.   */

typedef struct qldpc_code qldpc_code;
typedef struct qldpc_decoder qldpc_decoder;

typedef struct qldpc_code_info {
    int32_t n;               /* codeword length (variables)            */
    int32_t m;               /* parity checks                          */
    int32_t k;               /* information bits (n - m, or as set)    */
    int32_t edges;           /* ones in H                              */
    int32_t z;               /* lifting size, 0 for non-QC codes       */
    int32_t base_rows;       /* QC only                                */
    int32_t base_cols;       /* QC only                                */
    int32_t max_chk_degree;
    int32_t max_var_degree;
} qldpc_code_info;

/* ---- codes ---------------------------------------------------------------------------- */
int  qldpc_code_from_alist_file(const char *path, qldpc_code **out);
int  qldpc_code_from_qc_file(const char *path, qldpc_code **out);
/* base: rows*cols shifts row-major, -1 = zero block, shifts reduced mod z on load */
int  qldpc_code_from_qc(const int32_t *base, int32_t rows, int32_t cols, int32_t z, qldpc_code **out);
/* H in CSR by check: row_ptr[m+1], col_idx[row_ptr[m]] */
int  qldpc_code_from_csr(int32_t n, int32_t m, const int32_t *row_ptr, const int32_t *col_idx, qldpc_code **out);
/* positions of the k information bits inside the codeword; default: [0,k) for QC codes,
 * [n-k,n) for alist / CSR codes (the layout of BOOT/matrices/G/PEGReg504x1008.alist) */
int  qldpc_code_set_info_bits_pos(qldpc_code *code, const int32_t *pos, int32_t k);
int  qldpc_code_get_info(const qldpc_code *code, qldpc_code_info *info);
void qldpc_code_free(qldpc_code *code);

/* ---- decoders ------------------------------------------------------------------------- */
typedef struct qldpc_decoder_config {
    int32_t schedule;        /* QLDPC_SCHED_*                                                  */
    int32_t rule;            /* QLDPC_RULE_*                                                   */
    int32_t dtype;           /* QLDPC_DTYPE_*: type of the LLR input and of the messages       */
    int32_t max_iter;        /* n_ite (BOOT/src/main.cpp:100) / MaxItrs (BPSK_nrldpc_sim_FP.m:2) */
    int32_t early_stop;      /* enable_syndrome (BOOT/src/main.cpp:101)                        */
    int32_t syndrome_depth;  /* syndrome_depth (BOOT/src/main.cpp:102); flooding only          */
    float   norm_factor;     /* NMS factor (BOOT/src/main.cpp:98); integer dtypes: k/8, k=1..8 */
    float   offset;          /* OMS offset (BOOT/src/main.cpp:99; BPSK_nrldpc_sim_FP.m:6)      */
    int32_t msg_max;         /* integer dtypes: messages clipped to [-(msg_max+1), msg_max];
                                0 = default (31 for i8 = maxqr, BPSK_nrldpc_sim_FP.m:4; 511 for i16); at most 32766 */
    int32_t app_max;         /* integer dtypes, layered: beliefs clipped to [-(app_max+1), app_max];
                                0 = default (127 for i8 = maxqL, BPSK_nrldpc_sim_FP.m:5; 8191 for i16); at most 32767 */
    int32_t out_mode;        /* QLDPC_OUT_INFO: k info bits per frame; QLDPC_OUT_ALL: n bits   */
    int32_t device;          /* CUDA device ordinal (used when n_devices <= 1)                 */
    /* Frames are independent (AFF3CT n_frames semantics; ML/BPSK_nrldpc_sim_RM_FP.m:27 `parfor`): with n_devices > 1 the
     * HOST-pointer entry points (qldpc_decode, qldpc_decode_bits, qldpc_syndrome, qldpc_make_llr, qldpc_encode_nr) cut the
     * batch into contiguous frame ranges, one per device, each driven by its own host thread and stream pair; statistics
     * are summed on the host (qldpc_get_stats).  No collective, no peer access.  The *_device entry points take pointers
     * of ONE device and return QLDPC_ERR_UNSUPPORTED on such a decoder. */
    int32_t n_devices;       /* 0 or 1: `device` only; 2..QLDPC_MAX_DEVICES: devices[0..n_devices)             */
    int32_t devices[QLDPC_MAX_DEVICES];
    uint32_t flags;          /* QLDPC_FLAG_* */
} qldpc_decoder_config;

void qldpc_decoder_config_default(qldpc_decoder_config *cfg);
int  qldpc_decoder_create(const qldpc_code *code, const qldpc_decoder_config *cfg, qldpc_decoder **out);
void qldpc_decoder_free(qldpc_decoder *dec);

/* words per frame of the packed vectors this decoder reads / writes */
int32_t qldpc_out_words(const qldpc_decoder *dec);       /* ceil(k/32) or ceil(n/32)  */
int32_t qldpc_syndrome_words(const qldpc_decoder *dec);  /* ceil(m/32)                */
int32_t qldpc_codeword_words(const qldpc_decoder *dec);  /* ceil(n/32)                */

/*
 * Decode n_frames independent frames.
 *   llr        n_frames * n values of the decoder's dtype (float / int16_t / int8_t)
 *   syndrome   n_frames * qldpc_syndrome_words() packed words, or NULL (zero syndrome: the
 *              reference's send-parity formulation, BOOT/src/main.cpp:351-354)
 *   out_bits   n_frames * qldpc_out_words() packed hard decisions
 *   ok         n_frames bytes, 1 if H*hard == syndrome when the decoder stopped   (may be NULL)
 *   iters      n_frames check-node sweeps executed                                (may be NULL)
 *   posterior  n_frames * n a-posteriori values (float for f32, int32_t for i16/i8) (may be NULL)
 * qldpc_decode takes HOST pointers (copies are inside the call); qldpc_decode_device takes
 * DEVICE pointers, enqueues on `cuda_stream` (a cudaStream_t, NULL = default stream) and
 * returns without synchronising.
 * Concurrency: a decoder owns one set of device scratch buffers, so at most one qldpc_decode_device call per decoder may be
 * in flight at a time (use one decoder per stream); the host-pointer calls are self-contained and synchronise before
 * returning.  The library changes no device-wide state unless QLDPC_FLAG_L2_PERSIST is set in the configuration.
 */
int qldpc_decode(qldpc_decoder *dec, const void *llr, const uint32_t *syndrome, int32_t n_frames,
                 uint32_t *out_bits, uint8_t *ok, uint16_t *iters, void *posterior);
int qldpc_decode_device(qldpc_decoder *dec, const void *d_llr, const uint32_t *d_syndrome, int32_t n_frames,
                        uint32_t *d_out_bits, uint8_t *d_ok, uint16_t *d_iters, void *d_posterior,
                        void *cuda_stream);

/* syndrome = H * bits; bits: n_frames * qldpc_codeword_words(); syndrome: n_frames * qldpc_syndrome_words() */
int qldpc_syndrome(qldpc_decoder *dec, const uint32_t *bits, int32_t n_frames, uint32_t *syndrome);
int qldpc_syndrome_device(qldpc_decoder *dec, const uint32_t *d_bits, int32_t n_frames, uint32_t *d_syndrome,
                          void *cuda_stream);

/*
 * LLR synthesis from sifted-key bits (BSC):  llr[i] = (bit ? -1 : +1) * mag, with
 *   mag = llr_known  where known_mask has a 1 (parity bits Alice sent / shortened bits),
 *   mag = 0          where punct_mask has a 1 (punctured),
 *   mag = llr_noisy  elsewhere (ln((1-q)/q) for QBER q, quantised by the caller for integer dtypes).
 * bits: n_frames * qldpc_codeword_words(); the masks are ONE frame long (shared by all frames) or NULL.
 */
int qldpc_make_llr(qldpc_decoder *dec, const uint32_t *bits, const uint32_t *known_mask,
                   const uint32_t *punct_mask, float llr_noisy, float llr_known, int32_t n_frames, void *llr_out);
int qldpc_make_llr_device(qldpc_decoder *dec, const uint32_t *d_bits, const uint32_t *d_known_mask,
                          const uint32_t *d_punct_mask, float llr_noisy, float llr_known, int32_t n_frames,
                          void *d_llr_out, void *cuda_stream);

/*
 * Fused LLR synthesis + decode for callers that hold sifted-key bits (what an ecd2 LDPC handler has in
 * pb->mainBufPtr, MSB-first words): only n/8 bytes per frame cross the bus instead of n LLR values.
 * Arguments as in qldpc_make_llr followed by qldpc_decode; results are identical to calling the two.
 * On the int8 layered decoder for Z % 128 == 0 the synthesis happens INSIDE the decoder kernel (the frame's bits are
 * bulk-copied to shared memory, +-magnitude bytes are formed there): no LLR array exists in device memory.  If, in
 * addition, bits / syndrome / out_bits / ok / iters are PINNED host memory (cudaHostAlloc, cudaHostRegister), the kernel
 * reads and writes them in place over PCIe (zero copy): one launch per call, no staging buffers.  Pageable buffers
 * are staged through a two-stream chunked copy pipeline.
 * Integer dtypes: llr_noisy / llr_known are rounded to the nearest integer and must lie in [0, 127] (int8) or
 * [0, 32767] (int16), else QLDPC_ERR_ARG.
 */
int qldpc_decode_bits(qldpc_decoder *dec, const uint32_t *bits, const uint32_t *known_mask, const uint32_t *punct_mask,
                      float llr_noisy, float llr_known, const uint32_t *syndrome, int32_t n_frames,
                      uint32_t *out_bits, uint8_t *ok, uint16_t *iters);
int qldpc_decode_bits_device(qldpc_decoder *dec, const uint32_t *d_bits, const uint32_t *d_known_mask,
                             const uint32_t *d_punct_mask, float llr_noisy, float llr_known,
                             const uint32_t *d_syndrome, int32_t n_frames, uint32_t *d_out_bits, uint8_t *d_ok,
                             uint16_t *d_iters, void *cuda_stream);

/* 5G-NR systematic encoder (QC codes with the NR double-diagonal core only):
 * msg: n_frames * ceil(k/32) packed words; cword: n_frames * qldpc_codeword_words() */
int qldpc_encode_nr(qldpc_decoder *dec, const uint32_t *msg, int32_t n_frames, uint32_t *cword);
int qldpc_encode_nr_device(qldpc_decoder *dec, const uint32_t *d_msg, int32_t n_frames, uint32_t *d_cword,
                           void *cuda_stream);

/* ---- encoders for codes without the NR structure (Alice's side of the send-parity formulation) --------------
 * qldpc_encoder_from_h            module::Encoder_LDPC_from_H<B>(K, N, H, "IDENTITY" | "LU_DEC", ...)   VAR/main.cpp (alist-v1.0.1):144
 * qldpc_encoder_from_g_alist_file module::Encoder_LDPC<B>(K, N, G, n_frames), G read from the .alist files of BOOT/matrices/G   VAR/main.cpp (alist):143,333
 * qldpc_encode[_device]           m.encoder->encode(ref_bits, enc_bits)                               VAR/main.cpp (alist):417
 * from_H: Gauss-Jordan on the host; the code's info_bits_pos stay information positions wherever the rank of H allows it
 * (always for a full-rank parity part); qldpc_encoder_info_bits_pos returns the positions actually used, k = n - rank(H).
 * msg: n_frames * ceil(k/32) words, message bit i of a frame = information position pos[i]; cword: n_frames * ceil(n/32). */
typedef struct qldpc_encoder qldpc_encoder;
int  qldpc_encoder_from_h(const qldpc_code *code, int32_t device, qldpc_encoder **out);
int  qldpc_encoder_from_g_alist_file(const char *path, int32_t device, qldpc_encoder **out);
int  qldpc_encoder_get_info(const qldpc_encoder *enc, int32_t *k, int32_t *n);
int  qldpc_encoder_info_bits_pos(const qldpc_encoder *enc, int32_t *pos);
int  qldpc_encode(qldpc_encoder *enc, const uint32_t *msg, int32_t n_frames, uint32_t *cword);
int  qldpc_encode_device(qldpc_encoder *enc, const uint32_t *d_msg, int32_t n_frames, uint32_t *d_cword, void *cuda_stream);
void qldpc_encoder_free(qldpc_encoder *enc);

/* ---- statistics (host-side reduction across GPUs is the caller's job) ------------------ */
#define QLDPC_ITER_HIST_BINS 64
typedef struct qldpc_stats {
    uint64_t frames;                            /* frames decoded since create / reset          */
    uint64_t failures;                          /* frames whose final syndrome check failed      */
    uint64_t iter_sum;                          /* sum of iterations                             */
    uint64_t iter_hist[QLDPC_ITER_HIST_BINS];   /* histogram of iterations (last bin = overflow) */
    uint64_t kernel_launches;                   /* kernels this decoder launched                 */
} qldpc_stats;
int qldpc_get_stats(qldpc_decoder *dec, qldpc_stats *out);   /* synchronises the decoder's device */
int qldpc_reset_stats(qldpc_decoder *dec);

/* ---- after reconciliation: confirmation CRC and privacy amplification ---------------
 *
 * qldpc_privacy_amplify replaces the loop of errorcorrection/subcomponents/priv_amp.c:213-218 (privAmp_doPrivAmp) and
 * its PRNG, errorcorrection/subcomponents/rnd.c:118-127 (rnd_getPrngValue2_32, feedback 0xe0000200, rnd.h:46), bit for
 * bit: final key bit i = parity( XOR_j key[j] & prng_word[i*numwords + j] ), bits beyond workbits of the last key word
 * cleared first (priv_amp.c:189-191), final bits MSB-first (helpers.h:66-68).  A call serves n_blocks process blocks:
 *   key          n_blocks * key_stride_words words (pb->mainBufPtr of each block, MSB-first)
 *   workbits     n_blocks values (pb->workbits), each <= 32 * key_stride_words
 *   final_bits   n_blocks values (pb->finalKeyBits, computed by the caller from leakage / sneakloss as priv_amp.c:166 does)
 *   seeds        n_blocks PRNG seeds (the seed of EC packet subtype 8, priv_amp.c:45,54,76)
 *   final_key    n_blocks * out_stride_words words, out_stride_words >= ceil(max final_bits / 32); words past a block's
 *                ceil(final_bits/32) are left untouched
 * HOST pointers; `device` is the CUDA device ordinal.  There is no CPU fallback (QLDPC_ERR_NO_DEVICE).
 *
 * qldpc_crc32_frames: CRC-32 (IEEE 802.3 / zlib) of every frame of packed bits, over the frame's bytes in transmission
 * order (MSB-first words = big-endian bytes) -- the confirmation step the reference planned but never wrote
 * (errorcorrection/README_LDPC.md:784-788, README_AFF3CT.md:94): both sides compare the CRCs of their corrected frames.
 *   bits  n_frames * stride_words words; the first words_per_frame words of each frame are hashed.
 */
int qldpc_privacy_amplify(int32_t device, const uint32_t *key, int32_t key_stride_words, const int32_t *workbits,
                          const int32_t *final_bits, const uint32_t *seeds, int32_t n_blocks, uint32_t *final_key,
                          int32_t out_stride_words);
int qldpc_crc32_frames(int32_t device, const uint32_t *bits, int32_t n_frames, int32_t words_per_frame, int32_t stride_words,
                       uint32_t *crc_out);

/* name of the kernel family the decoder dispatches to (for logs / tests) */
const char *qldpc_decoder_kernel_name(const qldpc_decoder *dec);
const char *qldpc_strerror(int code);
const char *qldpc_last_cuda_error(void);
int qldpc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QLDPC_H */

// Kernel parameter blocks and launcher prototypes (host <-> device contract inside the library).
#pragma once

#include "qldpc_internal.hpp"

namespace qldpc {

// ---- layered int8, QC, Z % 4 == 0 ("zpack4" family) ------------------------------------------
struct LayeredI8Params {
    const int8_t *llr;        // F * N
    const uint32_t *syn;      // F * syn_words (MSB-first) or null
    uint32_t *out;            // F * out_words (MSB-first)
    uint8_t *ok;              // F or null
    uint16_t *iters;          // F or null
    DevStats *stats;
    const Li8Edge *edges;     // nnz
    const Li8Layer *layers;   // brows
    const QcEdgeAux *aux;     // nnz
    const uint16_t *pack_cols;  // n_pack block columns whose hard decisions are packed from the beliefs
    int F;
    int Z, W, ZW32;           // lanes, belief words per column (Z/4), 32-bit words per Z-bit vector
    int brows, bcols, nnz, N;
    int n_store;              // edges with a stored message (shared memory)
    int n_pack;
    int regdc;                // 0, or unrolled degree of the four register-resident rows
    int out_cols;             // block columns written to `out`
    int out_words, syn_words; // per frame
    int max_iter, early_stop;
    int rule, offset, norm_eighths, msg_max;
    uint32_t h2_lo, h2_hi, h2_cap, h2_negoff;   // half2 bit patterns: -(msg_max+1), msg_max, msg_max+1, -offset (units of 2^-24)
    int slots, tpg;           // frames in flight per CTA, threads per frame group
    int tab_bytes;            // shared tables at the start of dynamic smem
    int slot_bytes;           // bytes per frame slot
    int off_R, off_hd, off_syn;  // byte offsets inside a slot (beliefs start at 0); off_R = message store or ring
    // streamed mode (messages in an L2-resident scratch, staged through a 2-deep cp.async ring at off_R)
    int stream;
    uint32_t *rg;             // grid * slots * rg_words
    int rg_words;             // words per frame slot
    int stage_words;          // words per ring stage
};
int launch_layered_i8(const LayeredI8Params &p, int grid, int smem_bytes, cudaStream_t st);
int layered_i8_max_threads();
int layered_i8_reg_rows();
int layered_i8_max_threads_stream();

// ---- layered int8, QC, Z % 128 == 0, streamed kernel (layered_i8s.cu) ----------------------------
struct LayeredI8sParams {
    const int8_t *llr;        // F * N, 16-byte aligned (null with bit input)
    // bit input (qldpc_decode_bits): packed sifted-key bits instead of LLRs; the kernel synthesises +-mag[position] itself
    const uint32_t *bits;     // F * N/32 words, MSB-first, 16-byte aligned; null: LLR input
    const uint8_t *mag;       // N magnitudes 0..127 (shared by all frames), 16-byte aligned; copied to shared memory (off_magtab)
    int8_t *ext_scratch;      // grid * slots * N bytes: the slots' synthesised extension-column LLRs
    const uint32_t *syn;      // F * syn_words (MSB-first) or null
    uint32_t *out;            // F * out_words (MSB-first)
    uint8_t *ok;
    uint16_t *iters;
    DevStats *stats;
    const uint8_t *tab;       // table blob (tab_bytes, multiple of 16)
    uint4 *rg;                // message scratch: grid * slots * rg_u4
    int F;
    int Z, W, ZW32;
    int brows, bcols, N;
    int n_pack;               // core block columns (beliefs in shared memory)
    int out_cols, out_words, syn_words;
    int max_iter, early_stop;
    int rule, norm_eighths;
    uint32_t h2_lo, h2_hi, h2_cap, h2_span, h2_negoff;   // half2 patterns: -(m+1), m, m+1, 2m+1, -offset (m = msg_max)
    int slots;
    int tab_bytes, off_rows, off_pcols;
    int slot_bytes, off_ring, stage_bytes, off_ext, off_hd, off_syn, off_mbar;
    int off_stg;              // staging buffer of the next frame's core LLRs (n_pack * Z bytes; bit input: N/8 bytes), -1: none
    int off_magtab;           // bit input: byte offset of the CTA's magnitude table (N bytes) in dynamic shared memory
    int rg_u4;                // uint4 per frame slot in the scratch
    unsigned int *frame_ctr;  // zeroed before the launch: frames beyond the first grid * slots are handed out through it
                              // (null: frame f of a slot is followed by f + grid * slots)
    int discard_scratch;      // QLDPC_FLAG_DISCARD_SCRATCH: discard.global.L2 of the slot's message scratch when its frame ends
};
constexpr int kLi8sSlotBase = 32;   // after the tables: 16 bytes of zero messages, the CTA's iteration mbarrier (8 bytes, padded)
int launch_layered_i8s(const LayeredI8sParams &p, int grid, int smem_bytes, cudaStream_t st);
int layered_i8s_max_threads();

// ---- generic layered (QC, any Z; f32 / i16 / i8) ---------------------------------------------
struct LayeredGenParams {
    const void *llr;
    const uint32_t *syn;
    uint32_t *allbits;        // F * cw_words, MSB-first, all N hard decisions
    uint8_t *ok;
    uint16_t *iters;
    void *posterior;          // F * N (float or int32) or null
    DevStats *stats;
    const QcEdgeAux *aux;
    const QcLayer *layers;
    void *msg;                // grid * (nnz * Z messages (float / int16) | compressed check states), L2-resident global scratch
    void *app;                // null: beliefs in shared memory; else grid * N beliefs (float / int16) in a global scratch
    int F;
    int Z, brows, bcols, nnz, N, M;
    int cw_words, syn_words;
    int max_iter, early_stop, syndrome_depth;
    int rule, dtype;
    float norm, offset;
    int offset_int, norm_eighths, msg_max, app_max;
    int compressed;           // min-sum rules, every row within the compiled degrees: {c1, c2, index, signs} per check lane
    int fast_spa;             // QLDPC_FLAG_FAST_SPA
};
int layered_generic_max_compiled_degree();
size_t layered_generic_msg_scratch_bytes(int dtype, int compressed, int brows, int nnz, int Z);
int layered_generic_belief_bytes(int dtype);
int layered_generic_msg_bytes(int dtype);
int layered_generic_smem_bytes(int brows, int nnz, int N, int dtype);       // N = 0: tables only (beliefs in global memory)
int layered_generic_blocks_per_sm(int dtype, int Z, int smem_bytes, int beliefs_global);
int launch_layered_generic(const LayeredGenParams &p, int grid, cudaStream_t st);

// ---- layered, any H (CSR), float: one thread per frame (layered_csr.cu) ------------------------------------------
struct LayeredCsrParams {
    const float *llr;
    const uint32_t *syn;
    uint32_t *allbits;
    uint8_t *ok;
    uint16_t *iters;
    float *posterior;
    DevStats *stats;
    const int32_t *row_ptr, *col_idx;
    float *var, *branch;       // frame-minor state: N x threads and E x threads floats
    int F, N, M, E;
    int cw_words, syn_words;
    int max_iter, early_stop, syndrome_depth;
    int rule;
    float norm, offset;
    int fast_spa;             // QLDPC_FLAG_FAST_SPA
};
int layered_csr_max_degree();
int launch_layered_csr(const LayeredCsrParams &p, int threads, cudaStream_t st);   // threads: multiple of 128

// ---- flooding (any H; f32 / i16 / i8) -----------------------------------------------------------
struct FloodParams {
    const void *llr;
    const uint32_t *syn;
    uint32_t *allbits;
    uint8_t *ok;
    uint16_t *iters;
    void *posterior;
    DevStats *stats;
    const int32_t *row_ptr, *col_idx, *var_ptr, *var_edge;
    void *c2v;                // grid * E messages (global scratch) when they do not fit in smem
    void *post;               // grid * N
    int F, N, M, E;
    int cw_words, syn_words;
    int max_iter, early_stop, syndrome_depth;
    int rule, dtype;
    float norm, offset;
    int offset_int, norm_eighths, vmax;
    int use_smem;             // messages + posteriors in shared memory
    int tanh_cache;           // !use_smem, float SPA: 8 * block floats of shared memory cache tanh between the two passes
    int fast_spa;             // SPA transcendentals in fp32 on the SFUs instead of double (QLDPC_FLAG_FAST_SPA)
};
int launch_flooding(const FloodParams &p, int grid, int block, int smem_bytes, cudaStream_t st);

// ---- flooding, quasi-cyclic codes, float (flooding_qc.cu) -----------------------------------------
struct FloodQcParams {
    const float *llr;
    const uint32_t *syn;
    uint32_t *allbits;
    uint8_t *ok;
    uint16_t *iters;
    float *posterior;
    DevStats *stats;
    const QcEdgeAux *aux;      // nnz entries, row-major, columns ascending
    const QcLayer *layers;     // brows
    const int32_t *col_ptr;    // bcols + 1
    const int2 *col_edges;     // per block column: (edge id, shift) in ascending block-row order
    float *c2v;                // grid * nnz * Z (global scratch) when the messages do not fit in smem
    float *post;               // grid * N
    int F, Z, nnz, N, M;
    int cw_words, syn_words;
    int max_iter, early_stop, syndrome_depth;
    int rule;
    float norm, offset;
    int use_smem;
};
int layered_flood_qc_threads();
int launch_flooding_qc(const FloodQcParams &p, int grid, int block, int smem_bytes, cudaStream_t st);

// ---- flooding, large quasi-cyclic codes, one frame per thread-block cluster (flooding_qcx.cu) -----------------------
struct FloodQcxParams {
    const void *llr;           // F * N values of the decoder's dtype
    const uint32_t *syn;
    uint32_t *allbits;
    uint8_t *ok;
    uint16_t *iters;
    void *posterior;           // float (f32) or int32 (i16 / i8)
    DevStats *stats;
    const QcEdgeAux *aux;
    const QcLayer *layers;
    const int32_t *col_ptr;
    const int2 *col_edges;
    void *c2v;                 // n_clusters * nnz * (Z + 4) messages (float / int16 / int8), L2-resident scratch
    void *post;                // n_clusters * bcols * (Z + 4) posteriors (float / int32 / int16)
    int F, Z, nnz, N, M, brows, bcols;
    int cw_words, syn_words;
    int max_iter, early_stop, syndrome_depth;
    int rule, dtype;
    float norm, offset;
    int offset_int, norm_eighths, vmax;
    int lanes;                 // consecutive circulant lanes per thread: 4 (heaviest row <= 8 edges) or 2
    int fast_spa;              // SPA transcendentals in fp32 on the SFUs instead of double (QLDPC_FLAG_FAST_SPA)
};
int flooding_qcx_table_bytes(int brows, int bcols, int nnz);
int flooding_qcx_smem_bytes(int brows, int bcols, int nnz, int dtype);
int flooding_qcx_msg_bytes(int dtype);
int flooding_qcx_post_bytes(int dtype);
int flooding_qcx_run_lanes(int Z);                    // stride of a run of Z lanes in the scratch (Z + copy of the first lanes)
int flooding_qcx_lanes_per_thread(int max_row_degree);
int flooding_qcx_max_clusters(int dtype, int lanes, int cl, int smem_bytes);   // co-resident clusters of `cl` blocks on the current device
int launch_flooding_qcx(const FloodQcxParams &p, int n_clusters, int cl, int smem_bytes, cudaStream_t st);

// ---- bit-level helpers ---------------------------------------------------------------------------
// syndrome of packed frames: QC codes with Z % 32 == 0 use word-wise rotate + XOR (launch_syndrome_qc), others a CSR gather
int launch_syndrome_qc(const uint32_t *bits, uint32_t *syn, int F, int brows, int Z, int cw_words, int syn_words,
                       const QcLayer *layers, const QcEdgeAux *aux, cudaStream_t st);
int launch_make_mag_i8(const uint32_t *known, const uint32_t *punct, int noisy, int known_mag, int N, uint8_t *mag, cudaStream_t st);
int launch_syndrome_csr(const uint32_t *bits, uint32_t *syn, int F, int N, int M, int cw_words, int syn_words,
                        const int32_t *row_ptr, const int32_t *col_idx, cudaStream_t st);
int launch_gather_bits(const uint32_t *allbits, uint32_t *out, int F, int cw_words, int out_words, int K,
                       const int32_t *info_pos, cudaStream_t st);
int launch_make_llr(const uint32_t *bits, const uint32_t *known, const uint32_t *punct, float noisy, float known_mag,
                    int F, int N, int cw_words, int dtype, void *llr_out, cudaStream_t st);
int launch_encode_nr(const uint32_t *msg, uint32_t *cword, int F, int Z, int brows, int bcols, const int32_t *base,
                     int msg_words, int cw_words, cudaStream_t st);
int launch_encode_nr_packed(const uint32_t *msg, uint32_t *cword, int F, int Z, int brows, int bcols, int p1_rot, int msg_words,
                            int cw_words, const QcLayer *layers, const QcEdgeAux *aux, cudaStream_t st);


// ---- post-reconciliation (postproc.cu) -------------------------------------------------------------
int pa_upload_jump_tables();
int launch_privacy_amplify(const uint32_t *d_key, const int32_t *d_workbits, const int32_t *d_final_bits, const uint32_t *d_seeds,
                           int n_blocks, int key_stride, int max_workbits, int max_final_bits, uint32_t *d_out, int out_stride,
                           cudaStream_t st);
int launch_crc32_frames(const uint32_t *d_bits, int n_frames, int words_per_frame, int stride_words, uint32_t *d_crc, cudaStream_t st);

}  // namespace qldpc

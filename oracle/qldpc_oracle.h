/*
 * qldpc_oracle.h -- CPU restatement of the reference's LDPC reconciliation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (qcrypto-ldpc_b200/, include/)
 * links, imports or executes this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, and only as the
 * checker / reported baseline.
 *
 * PARITY PIN STATUS (see DESIGN.md "Oracle"; tests/test_oracle.py, tests/test_oracle_pins.py, tests/test_gpu_pins.py):
 *   - float flooding SPA decoded bits: PINNED exactly by the reference's one known-answer
 *     vector (PEGReg504x1008, "main.cpp (alist)":443-462) -> tests/golden/kat_pegreg504x1008.json,
 *     and statistically by the (5g-qc) driver's recorded sweep on NR_1_0_2.qc
 *     (README_LDPC.md:941-974): all 11 QBER steps inside their 95 % intervals, BER of the sweep within
 *     a few per cent (min-sum rules give 2-6 x that BER, so the rule is pinned too).
 *   - alist / qc parsing, circulant convention, NR encoder: PINNED by structural
 *     golden facts of the reference's own matrices (H*encode(msg)=0, NR_1_1_384 mod 192
 *     == NR_1_1_192.qc, edge counts, degree profiles).
 *   - fixed-point layered offset min-sum (int8 tier): restates ML/BPSK_nrldpc_sim_FP.m /
 *     _RM_FP.m line by line and is PINNED STATISTICALLY by the table the reference recorded with it
 *     (ML/sim_results.m:2-5,9-12, NR_1_1_24 and NR_2_6_52, 100..20000 frames per point): all eight
 *     points inside their 95 % intervals -- with the noise variance taken from the rate k/n; the script
 *     as committed uses k/(n-2z), which is 0.19 / 0.41 dB better than the table it is filed with
 *     (tests/refpins.py explains, tests/test_oracle_pins.py keeps both facts true).  MATLAB/Octave are
 *     absent, so there is no bit-exact pin: iteration counts with early stop, the syndrome extension and
 *     the shift-normalised rule remain defined by this file.
 *   - normalised min-sum, int16, flooding min-sum, float posteriors:
 *     the arithmetic lives in AFF3CT v2.3.5 (commit 1ceddfc), which is NOT in
 *     /root/reference (fetched by git clone in ci/build-linux-macos.sh:50).  These
 *     restate AFF3CT's published algorithm from memory: "parity unpinned".
 *
 * ML/  = errorcorrection/ldpc_examples/matlab_code_Base_matrices/matlab_code & Base_matrices/
 * BOOT/= errorcorrection/ldpc_examples/my_project_with_aff3ct/examples/bootstrap/
 */
#ifndef QLDPC_ORACLE_H
#define QLDPC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* update rules named by the drivers (BOOT/src/main.cpp:193, "main.cpp (5g-qc)":236-251) */
enum { ORA_RULE_SPA = 0, ORA_RULE_NMS = 1, ORA_RULE_OMS = 2 };

typedef struct ora_code {
    int N, M, E;          /* variables, checks, edges                               */
    int *row_ptr;         /* M+1 : CSR by check                                      */
    int *col_idx;         /* E   : variable index of every edge, ascending per check */
    int *col_ptr;         /* N+1 : CSC by variable                                   */
    int *row_edge;        /* E   : for variable-major slot k, the CSR edge id        */
    int is_qc, Z, brows, bcols;
    int *base;            /* brows*bcols shifts (already mod Z), -1 = zero block     */
} ora_code;

/* matrix readers: AFF3CT LDPC_matrix_handler::read ("main.cpp (alist)":333,340; "(5g-qc)":389) */
ora_code *ora_code_from_alist_file(const char *path);
ora_code *ora_code_from_qc_file(const char *path);      /* .qc with header, or NR_*.txt with z>0 */
ora_code *ora_code_from_nr_txt(const char *path, int Z);
ora_code *ora_code_from_base(const int *base, int brows, int bcols, int Z);
ora_code *ora_code_from_csr(int N, int M, const int *row_ptr, const int *col_idx);
void      ora_code_free(ora_code *c);
int       ora_code_N(const ora_code *c);
int       ora_code_M(const ora_code *c);
int       ora_code_E(const ora_code *c);
int       ora_code_Z(const ora_code *c);
int       ora_code_brows(const ora_code *c);
int       ora_code_bcols(const ora_code *c);
const int *ora_code_row_ptr(const ora_code *c);
const int *ora_code_col_idx(const ora_code *c);
const int *ora_code_base(const ora_code *c);

/* ML/check_cword.m:9-19 -- syndrome H*c (one byte per bit) */
void ora_syndrome(const ora_code *c, const uint8_t *bits, uint8_t *syn);
/* ML/nrldpc_encode.m:12-40 -- 5G-NR double-diagonal encoder. returns 0 ok */
int  ora_nr_encode(const ora_code *c, const uint8_t *msg, uint8_t *cword);

/* AFF3CT Decoder_LDPC_BP_flooding<B,Q,Update_rule_{SPA,NMS,OMS}> restated.
 * llr[N] float; syn[M] bytes or NULL (zero syndrome = the reference's send-parity case);
 * post[N] (may be NULL) posterior LLR; hard[N] bytes; *iters = check-node sweeps executed.
 * returns 1 if H*hard == syn at exit, else 0. */
int ora_decode_flooding_f32(const ora_code *c, const float *llr, const uint8_t *syn,
                            int rule, int n_ite, int early_stop, int syndrome_depth,
                            float norm, float offset,
                            float *post, uint8_t *hard, int *iters);

/* AFF3CT Decoder_LDPC_BP_horizontal_layered restated; row-serial on the expanded H. */
int ora_decode_layered_f32(const ora_code *c, const float *llr, const uint8_t *syn,
                           int rule, int n_ite, int early_stop, int syndrome_depth,
                           float norm, float offset,
                           float *post, uint8_t *hard, int *iters);

/* ML/BPSK_nrldpc_sim_FP.m:35-94 restated (QC codes only), generalised by
 *   - optional syndrome (parity sign of check c multiplied by (-1)^syn[c]),
 *   - optional early stop after each full iteration (H*hard == syn),
 *   - rule OMS(offset, integer) as in the .m file, or NMS(norm_eighths/8 via shifts).
 * llr[N] integers (already quantised), msg_max=31 / app_max=127 reproduce the .m constants.
 * app[N] (may be NULL) final total beliefs. */
int ora_decode_layered_fixed(const ora_code *c, const int *llr, const uint8_t *syn,
                             int rule, int n_ite, int early_stop,
                             int offset, int norm_eighths, int msg_max, int app_max,
                             int *app, uint8_t *hard, int *iters);

/* fixed-point flooding min-sum (int8/int16 tiers): AFF3CT flooding with integer Q,
 * saturating at +-vmax on every stored message. */
int ora_decode_flooding_fixed(const ora_code *c, const int *llr, const uint8_t *syn,
                              int rule, int n_ite, int early_stop,
                              int offset, int norm_eighths, int vmax,
                              int *post, uint8_t *hard, int *iters);

/* batched helpers used by tests / cpu_baseline (frames split over pthreads).
 * llr: F*N int8; syn: F*M bytes or NULL; hard: F*N bytes; iters: F ints; ok: F bytes.
 * returns number of threads used. */
int ora_batch_layered_fixed_i8(const ora_code *c, const int8_t *llr, const uint8_t *syn, int F,
                               int rule, int n_ite, int early_stop,
                               int offset, int norm_eighths, int msg_max, int app_max,
                               uint8_t *hard, int *iters, uint8_t *ok, int n_threads);
int ora_batch_flooding_f32(const ora_code *c, const float *llr, const uint8_t *syn, int F,
                           int rule, int n_ite, int early_stop, float norm, float offset,
                           float *post, uint8_t *hard, int *iters, uint8_t *ok, int n_threads);

/* Privacy amplification, restating errorcorrection/subcomponents/priv_amp.c:186-218 with the PRNG of
 * errorcorrection/subcomponents/rnd.c:118-127 (rnd_getPrngValue2_32: 32 steps of state <<= 1, state += parity(state &
 * 0xe0000200), rnd.h:46).  key: ceil(workbits/32) words MSB-first (the last word is masked as priv_amp.c:189-191 does, on a
 * copy); out: ceil(final_bits/32) words, zeroed first (priv_amp.c:207).  `prng32` = NULL uses the restatement below; tests
 * pass the reference's own rnd_getPrngValue2_32 from oracle/_ref/librnd_ref.so to pin it. */
typedef unsigned int (*ora_prng32_fn)(unsigned int *state);
unsigned int ora_prng32(unsigned int *state);
void ora_privacy_amplify(const uint32_t *key, int workbits, int final_bits, uint32_t seed, uint32_t *out, ora_prng32_fn prng32);
/* CRC-32 (IEEE 802.3 / zlib), bit-serial, over words taken MSB-first (big-endian bytes) */
uint32_t ora_crc32_words(const uint32_t *words, int n_words);

/* integer normalisation by k/8 with shifts (AFF3CT Update_rule_NMS for integer Q) */
int ora_normalize_eighths(int v, int eighths);

#ifdef __cplusplus
}
#endif
#endif

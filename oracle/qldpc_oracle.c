/*
 * qldpc_oracle.c -- scalar CPU restatement of the reference's LDPC decode path.
 * TEST INFRASTRUCTURE ONLY (see qldpc_oracle.h header comment for the pin status).
 *
 * Every function cites the reference file:line it follows.  Paths are relative to
 * /root/reference/; ML/ and BOOT/ are expanded in qldpc_oracle.h.
 */
#include "qldpc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------ utilities */

static char *read_line(FILE *f, char **buf, size_t *cap)
{
    size_t len = 0;
    int ch;
    if (!*buf) { *cap = 1 << 16; *buf = (char *)malloc(*cap); }
    while ((ch = fgetc(f)) != EOF) {
        if (len + 2 > *cap) { *cap *= 2; *buf = (char *)realloc(*buf, *cap); }
        if (ch == '\n') break;
        (*buf)[len++] = (char)ch;
    }
    if (ch == EOF && len == 0) return NULL;
    (*buf)[len] = 0;
    return *buf;
}

/* parse all integers on a line; returns count (<= max) */
static int parse_ints(const char *s, int *out, int max)
{
    int n = 0;
    char *end;
    while (*s) {
        long v = strtol(s, &end, 10);
        if (end == s) { ++s; continue; }
        if (n < max) out[n] = (int)v;
        ++n;
        s = end;
    }
    return n;
}

static char *next_nonblank(FILE *f, char **buf, size_t *cap)
{
    char *l;
    while ((l = read_line(f, buf, cap))) {
        const char *p = l;
        while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
        if (*p) return l;
    }
    return NULL;
}

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

/* build CSC (variable-major) view from the CSR (check-major) one */
static void build_csc(ora_code *c)
{
    int *cnt = (int *)calloc((size_t)c->N + 1, sizeof(int));
    c->col_ptr = (int *)malloc(((size_t)c->N + 1) * sizeof(int));
    c->row_edge = (int *)malloc((size_t)c->E * sizeof(int));
    for (int e = 0; e < c->E; ++e) cnt[c->col_idx[e] + 1]++;
    c->col_ptr[0] = 0;
    for (int v = 0; v < c->N; ++v) c->col_ptr[v + 1] = c->col_ptr[v] + cnt[v + 1];
    memset(cnt, 0, ((size_t)c->N + 1) * sizeof(int));
    /* checks visited in ascending order => each variable's slots are in ascending check order */
    for (int m = 0; m < c->M; ++m)
        for (int e = c->row_ptr[m]; e < c->row_ptr[m + 1]; ++e) {
            int v = c->col_idx[e];
            c->row_edge[c->col_ptr[v] + cnt[v]++] = e;
        }
    free(cnt);
}

ora_code *ora_code_from_csr(int N, int M, const int *row_ptr, const int *col_idx)
{
    ora_code *c = (ora_code *)calloc(1, sizeof(*c));
    c->N = N; c->M = M; c->E = row_ptr[M];
    c->row_ptr = (int *)malloc(((size_t)M + 1) * sizeof(int));
    c->col_idx = (int *)malloc((size_t)c->E * sizeof(int));
    memcpy(c->row_ptr, row_ptr, ((size_t)M + 1) * sizeof(int));
    memcpy(c->col_idx, col_idx, (size_t)c->E * sizeof(int));
    for (int m = 0; m < M; ++m)   /* edge order inside a check: ascending variable index */
        qsort(c->col_idx + c->row_ptr[m], (size_t)(c->row_ptr[m + 1] - c->row_ptr[m]), sizeof(int), cmp_int);
    build_csc(c);
    return c;
}

/* .alist reader -- format as used by BOOT/matrices/H/PEGReg504x1008.alist:1-4 and
 * read by AFF3CT's LDPC_matrix_handler::read at "main.cpp (alist)":340.
 * line1 "N M", line2 "max_col_deg max_row_deg", line3 N col degrees, line4 M row degrees,
 * N lines of 1-based check indices (0 padded), M lines of 1-based variable indices. */
ora_code *ora_code_from_alist_file(const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) return NULL;
    char *buf = NULL; size_t cap = 0; char *l;
    int hdr[2], N, M, dvmax, dcmax;
    ora_code *c = NULL;
    int *tmp = NULL, *vdeg = NULL, *cdeg = NULL, *rp = NULL, *ci = NULL, *fill = NULL;

    if (!(l = next_nonblank(f, &buf, &cap)) || parse_ints(l, hdr, 2) < 2) goto done;
    N = hdr[0]; M = hdr[1];
    if (!(l = next_nonblank(f, &buf, &cap)) || parse_ints(l, hdr, 2) < 2) goto done;
    dvmax = hdr[0]; dcmax = hdr[1];
    if (N <= 0 || M <= 0 || dvmax <= 0 || dcmax <= 0) goto done;
    vdeg = (int *)malloc((size_t)N * sizeof(int));
    cdeg = (int *)malloc((size_t)M * sizeof(int));
    if (!(l = next_nonblank(f, &buf, &cap)) || parse_ints(l, vdeg, N) != N) goto done;
    if (!(l = next_nonblank(f, &buf, &cap)) || parse_ints(l, cdeg, M) != M) goto done;
    rp = (int *)calloc((size_t)M + 1, sizeof(int));
    for (int m = 0; m < M; ++m) rp[m + 1] = rp[m] + cdeg[m];
    ci = (int *)malloc((size_t)rp[M] * sizeof(int));
    fill = (int *)calloc((size_t)M, sizeof(int));
    tmp = (int *)malloc(((size_t)(N > M ? N : M) + 8) * sizeof(int));
    /* variable lines define the edges */
    for (int v = 0; v < N; ++v) {
        if (!(l = next_nonblank(f, &buf, &cap))) goto done;
        int n = parse_ints(l, tmp, N > M ? N : M), d = 0;
        for (int j = 0; j < n; ++j) {
            if (tmp[j] == 0) continue;
            int m = tmp[j] - 1;
            if (m < 0 || m >= M || fill[m] >= cdeg[m]) goto done;
            ci[rp[m] + fill[m]++] = v;
            ++d;
        }
        if (d != vdeg[v]) goto done;
    }
    /* check lines must agree with what the variable lines said */
    for (int m = 0; m < M; ++m) {
        if (fill[m] != cdeg[m]) goto done;
        if (!(l = next_nonblank(f, &buf, &cap))) goto done;
        int n = parse_ints(l, tmp, N > M ? N : M), d = 0;
        for (int j = 0; j < n; ++j) {
            if (tmp[j] == 0) continue;
            int v = tmp[j] - 1, found = 0;
            for (int e = rp[m]; e < rp[m + 1]; ++e) found |= (ci[e] == v);
            if (!found) goto done;
            ++d;
        }
        if (d != cdeg[m]) goto done;
    }
    c = ora_code_from_csr(N, M, rp, ci);
done:
    free(buf); free(tmp); free(vdeg); free(cdeg); free(rp); free(ci); free(fill);
    fclose(f);
    return c;
}

/* QC expansion.  Convention from ML/mul_sh.m:9 and ML/check_cword.m:12:
 * y = mul_sh(x,k) has y[i] = x[(i+k) mod z], so check lane i of block-row r touches
 * variable lane (i + B(r,c)) mod z of block-column c. */
ora_code *ora_code_from_base(const int *base, int brows, int bcols, int Z)
{
    ora_code *c = (ora_code *)calloc(1, sizeof(*c));
    int nnz = 0;
    for (int i = 0; i < brows * bcols; ++i) nnz += (base[i] >= 0);
    c->N = bcols * Z; c->M = brows * Z; c->E = nnz * Z;
    c->is_qc = 1; c->Z = Z; c->brows = brows; c->bcols = bcols;
    c->base = (int *)malloc((size_t)brows * bcols * sizeof(int));
    for (int i = 0; i < brows * bcols; ++i) c->base[i] = base[i] < 0 ? -1 : base[i] % Z;
    c->row_ptr = (int *)malloc(((size_t)c->M + 1) * sizeof(int));
    c->col_idx = (int *)malloc((size_t)c->E * sizeof(int));
    int e = 0;
    for (int r = 0; r < brows; ++r)
        for (int i = 0; i < Z; ++i) {
            c->row_ptr[r * Z + i] = e;
            for (int b = 0; b < bcols; ++b) {
                int s = c->base[r * bcols + b];
                if (s >= 0) c->col_idx[e++] = b * Z + (i + s) % Z;
            }
        }
    c->row_ptr[c->M] = e;
    build_csc(c);
    return c;
}

/* .qc: "cols rows Z", blank, rows lines of cols ints (BOOT/matrices/H/NR_1_1_192.qc:1-3);
 * shifts >= Z occur in BOOT/matrices/H/test2.qc:3 and are reduced mod Z. */
ora_code *ora_code_from_qc_file(const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) return NULL;
    char *buf = NULL; size_t cap = 0; char *l;
    int hdr[3];
    ora_code *c = NULL;
    int *base = NULL;
    if (!(l = next_nonblank(f, &buf, &cap)) || parse_ints(l, hdr, 3) != 3) goto done;
    int bcols = hdr[0], brows = hdr[1], Z = hdr[2];
    if (bcols <= 0 || brows <= 0 || Z <= 0) goto done;
    base = (int *)malloc((size_t)brows * bcols * sizeof(int));
    for (int r = 0; r < brows; ++r) {
        if (!(l = next_nonblank(f, &buf, &cap))) goto done;
        if (parse_ints(l, base + r * bcols, bcols) != bcols) goto done;
    }
    c = ora_code_from_base(base, brows, bcols, Z);
done:
    free(base); free(buf); fclose(f);
    return c;
}

/* NR_*.txt: the same table without the header (errorcorrection/ldpc_examples/README.md:7);
 * loaded by ML/BPSK_nrldpc_sim_FP.m:8-11 with z given separately. */
ora_code *ora_code_from_nr_txt(const char *path, int Z)
{
    FILE *f = fopen(path, "r");
    if (!f) return NULL;
    char *buf = NULL; size_t cap = 0; char *l;
    int *base = NULL, brows = 0, bcols = 0, capr = 0;
    int tmp[512];
    ora_code *c = NULL;
    while ((l = next_nonblank(f, &buf, &cap))) {
        int n = parse_ints(l, tmp, 512);
        if (n == 0) continue;
        if (bcols == 0) bcols = n;
        if (n != bcols || n > 512) goto done;
        if (brows == capr) { capr = capr ? capr * 2 : 64; base = (int *)realloc(base, (size_t)capr * bcols * sizeof(int)); }
        memcpy(base + brows * bcols, tmp, (size_t)bcols * sizeof(int));
        ++brows;
    }
    if (brows > 0 && Z > 0) c = ora_code_from_base(base, brows, bcols, Z);
done:
    free(base); free(buf); fclose(f);
    return c;
}

void ora_code_free(ora_code *c)
{
    if (!c) return;
    free(c->row_ptr); free(c->col_idx); free(c->col_ptr); free(c->row_edge); free(c->base);
    free(c);
}
int ora_code_N(const ora_code *c) { return c->N; }
int ora_code_M(const ora_code *c) { return c->M; }
int ora_code_E(const ora_code *c) { return c->E; }
int ora_code_Z(const ora_code *c) { return c->is_qc ? c->Z : 0; }
int ora_code_brows(const ora_code *c) { return c->brows; }
int ora_code_bcols(const ora_code *c) { return c->bcols; }
const int *ora_code_row_ptr(const ora_code *c) { return c->row_ptr; }
const int *ora_code_col_idx(const ora_code *c) { return c->col_idx; }
const int *ora_code_base(const ora_code *c) { return c->base; }

/* ML/check_cword.m:9-19 (syn = H c^T over GF(2)) */
void ora_syndrome(const ora_code *c, const uint8_t *bits, uint8_t *syn)
{
    for (int m = 0; m < c->M; ++m) {
        unsigned s = 0;
        for (int e = c->row_ptr[m]; e < c->row_ptr[m + 1]; ++e) s ^= bits[c->col_idx[e]];
        syn[m] = (uint8_t)(s & 1u);
    }
}

static int syndrome_matches(const ora_code *c, const uint8_t *hard, const uint8_t *syn)
{
    for (int m = 0; m < c->M; ++m) {
        unsigned s = syn ? syn[m] : 0u;
        for (int e = c->row_ptr[m]; e < c->row_ptr[m + 1]; ++e) s ^= hard[c->col_idx[e]];
        if (s & 1u) return 0;
    }
    return 1;
}

/* y = mul_sh(x,k): ML/mul_sh.m:9 */
static void mul_sh_xor(uint8_t *acc, const uint8_t *x, int k, int z)
{
    if (k < 0) return;
    for (int i = 0; i < z; ++i) acc[i] ^= x[(i + k) % z];
}

/* ML/nrldpc_encode.m:12-40.  cword[0:k) = msg, then p1 (double-diagonal core),
 * p2..p4, then the degree-1 extension parities. */
int ora_nr_encode(const ora_code *c, const uint8_t *msg, uint8_t *cword)
{
    if (!c->is_qc || c->brows < 4) return -1;
    const int z = c->Z, m = c->brows, n = c->bcols, kb = n - m;
    const int *B = c->base;
    uint8_t *temp = (uint8_t *)calloc((size_t)z, 1);
    memset(cword, 0, (size_t)n * z);
    memcpy(cword, msg, (size_t)kb * z);
    /* :18-23  temp = sum over rows 1..4 and message columns */
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < kb; ++j) mul_sh_xor(temp, msg + j * z, B[i * n + j], z);
    /* :24-29  p1 = mul_sh(temp, z - p1_sh) */
    int p1_sh = (B[1 * n + kb] == -1) ? B[2 * n + kb] : B[1 * n + kb];
    for (int i = 0; i < z; ++i) cword[kb * z + i] = temp[(i + (z - p1_sh)) % z];
    /* :30-37  p2,p3,p4 */
    for (int i = 0; i < 3; ++i) {
        memset(temp, 0, (size_t)z);
        for (int j = 0; j < kb + i + 1; ++j) mul_sh_xor(temp, cword + j * z, B[i * n + j], z);
        memcpy(cword + (kb + i + 1) * z, temp, (size_t)z);
    }
    /* :38-45  remaining parities */
    for (int i = 4; i < m; ++i) {
        memset(temp, 0, (size_t)z);
        for (int j = 0; j < kb + 4; ++j) mul_sh_xor(temp, cword + j * z, B[i * n + j], z);
        memcpy(cword + (kb + i) * z, temp, (size_t)z);
    }
    free(temp);
    return 0;
}

/* ------------------------------------------------- AFF3CT update rules (float) */
/* [RECALL] AFF3CT v2.3.5 Update_rule_{SPA,MS,OMS,NMS}; SURVEY.md section 3.4.
 * sign is kept as 0 / 1, var_sign = signbit(x). */

typedef struct {
    int rule;
    float norm, offset;
    int sign;
    float product;     /* SPA */
    float min1, min2;  /* MS family */
    float cst1, cst2;
    float *values;     /* SPA: tanh(|x|/2) per edge */
} rule_f32;

static inline void rule_begin(rule_f32 *u, int syn_bit)
{
    u->sign = syn_bit;             /* syndrome bit folded into the running sign */
    u->product = 1.0f;
    u->min1 = FLT_MAX; u->min2 = FLT_MAX;
}

static inline void rule_in(rule_f32 *u, int j, float x)
{
    const float a = fabsf(x);
    u->sign ^= signbit(x) ? 1 : 0;
    if (u->rule == ORA_RULE_SPA) {
        const float t = (float)tanh((double)(a * 0.5f));   /* correctly rounded float tanh, see DESIGN.md */
        const float r = (t != 0.0f) ? t : 1e-12f;
        u->product *= r;
        u->values[j] = r;
    } else {
        u->min2 = fminf(u->min2, fmaxf(a, u->min1));
        u->min1 = fminf(u->min1, a);
    }
}

static inline void rule_end_in(rule_f32 *u)
{
    if (u->rule == ORA_RULE_NMS) {
        u->cst1 = u->min2 * u->norm;
        u->cst2 = u->min1 * u->norm;
    } else if (u->rule == ORA_RULE_OMS) {
        u->cst1 = fmaxf(0.0f, u->min2 - u->offset);
        u->cst2 = fmaxf(0.0f, u->min1 - u->offset);
    }
}

static inline float rule_out(const rule_f32 *u, int j, float x)
{
    const int s = u->sign ^ (signbit(x) ? 1 : 0);
    float mag;
    if (u->rule == ORA_RULE_SPA) {
        float r = u->product / u->values[j];
        r = (r < 1.0f) ? r : 1.0f - FLT_EPSILON;
        mag = 2.0f * (float)atanh((double)r);
    } else {
        mag = (fabsf(x) == u->min1) ? u->cst1 : u->cst2;
    }
    return s ? -mag : mag;
}

static int max_chk_degree(const ora_code *c)
{
    int d = 0;
    for (int m = 0; m < c->M; ++m) if (c->row_ptr[m + 1] - c->row_ptr[m] > d) d = c->row_ptr[m + 1] - c->row_ptr[m];
    return d;
}

/* AFF3CT Decoder_LDPC_BP_flooding::_decode / _initialize_var_to_chk / _decode_single_ite /
 * _compute_post restated ([RECALL]); call site BOOT/src/main.cpp:193,365;
 * early stop = enable_syndrome, syndrome_depth (BOOT/src/main.cpp:101-102). */
int ora_decode_flooding_f32(const ora_code *c, const float *llr, const uint8_t *syn,
                            int rule, int n_ite, int early_stop, int syndrome_depth,
                            float norm, float offset,
                            float *post_out, uint8_t *hard, int *iters)
{
    const int N = c->N, M = c->M, E = c->E;
    float *c2v = (float *)calloc((size_t)E, sizeof(float));   /* decoder.reset(): BOOT/src/main.cpp:389 */
    float *v2c = (float *)malloc((size_t)E * sizeof(float));
    float *post = (float *)malloc((size_t)N * sizeof(float));
    rule_f32 u; memset(&u, 0, sizeof(u));
    u.rule = rule; u.norm = norm; u.offset = offset;
    u.values = (float *)malloc((size_t)(max_chk_degree(c) + 1) * sizeof(float));
    int ite = 0, depth = 0, executed = 0;
    if (syndrome_depth < 1) syndrome_depth = 1;

    for (; ite < n_ite; ++ite) {
        for (int v = 0; v < N; ++v) {                         /* _initialize_var_to_chk */
            float sum = 0.0f;
            for (int k = c->col_ptr[v]; k < c->col_ptr[v + 1]; ++k) sum += c2v[c->row_edge[k]];
            const float tmp = llr[v] + sum;
            for (int k = c->col_ptr[v]; k < c->col_ptr[v + 1]; ++k) v2c[c->row_edge[k]] = tmp - c2v[c->row_edge[k]];
        }
        for (int m = 0; m < M; ++m) {                         /* _decode_single_ite */
            const int e0 = c->row_ptr[m], d = c->row_ptr[m + 1] - e0;
            rule_begin(&u, syn ? (syn[m] & 1) : 0);
            for (int j = 0; j < d; ++j) rule_in(&u, j, v2c[e0 + j]);
            rule_end_in(&u);
            for (int j = 0; j < d; ++j) c2v[e0 + j] = rule_out(&u, j, v2c[e0 + j]);
        }
        ++executed;
        if (early_stop && ite != n_ite - 1) {                 /* check_syndrome_soft */
            for (int v = 0; v < N; ++v) {
                float sum = 0.0f;
                for (int k = c->col_ptr[v]; k < c->col_ptr[v + 1]; ++k) sum += c2v[c->row_edge[k]];
                post[v] = llr[v] + sum;
                hard[v] = (uint8_t)(post[v] < 0.0f);
            }
            if (syndrome_matches(c, hard, syn)) { if (++depth == syndrome_depth) break; }
            else depth = 0;
        }
    }
    for (int v = 0; v < N; ++v) {                             /* _compute_post + hard decision */
        float sum = 0.0f;
        for (int k = c->col_ptr[v]; k < c->col_ptr[v + 1]; ++k) sum += c2v[c->row_edge[k]];
        post[v] = llr[v] + sum;
        hard[v] = (uint8_t)(post[v] < 0.0f);
    }
    if (post_out) memcpy(post_out, post, (size_t)N * sizeof(float));
    if (iters) *iters = executed;
    const int ok = syndrome_matches(c, hard, syn);
    free(c2v); free(v2c); free(post); free(u.values);
    return ok;
}

/* AFF3CT Decoder_LDPC_BP_horizontal_layered restated ([RECALL]);
 * call sites block-commented at "main.cpp (5g-qc)":256-270. */
int ora_decode_layered_f32(const ora_code *c, const float *llr, const uint8_t *syn,
                           int rule, int n_ite, int early_stop, int syndrome_depth,
                           float norm, float offset,
                           float *post_out, uint8_t *hard, int *iters)
{
    const int N = c->N, M = c->M, E = c->E;
    const int dmax = max_chk_degree(c);
    float *branches = (float *)calloc((size_t)E, sizeof(float));
    float *var = (float *)malloc((size_t)N * sizeof(float));
    float *contrib = (float *)malloc((size_t)(dmax + 1) * sizeof(float));
    rule_f32 u; memset(&u, 0, sizeof(u));
    u.rule = rule; u.norm = norm; u.offset = offset;
    u.values = (float *)malloc((size_t)(dmax + 1) * sizeof(float));
    int ite = 0, depth = 0, executed = 0;
    if (syndrome_depth < 1) syndrome_depth = 1;
    memcpy(var, llr, (size_t)N * sizeof(float));

    for (; ite < n_ite; ++ite) {
        for (int m = 0; m < M; ++m) {
            const int e0 = c->row_ptr[m], d = c->row_ptr[m + 1] - e0;
            rule_begin(&u, syn ? (syn[m] & 1) : 0);
            for (int j = 0; j < d; ++j) {
                contrib[j] = var[c->col_idx[e0 + j]] - branches[e0 + j];
                rule_in(&u, j, contrib[j]);
            }
            rule_end_in(&u);
            for (int j = 0; j < d; ++j) {
                branches[e0 + j] = rule_out(&u, j, contrib[j]);
                var[c->col_idx[e0 + j]] = contrib[j] + branches[e0 + j];
            }
        }
        ++executed;
        if (early_stop && ite != n_ite - 1) {
            for (int v = 0; v < N; ++v) hard[v] = (uint8_t)(var[v] < 0.0f);
            if (syndrome_matches(c, hard, syn)) { if (++depth == syndrome_depth) break; }
            else depth = 0;
        }
    }
    for (int v = 0; v < N; ++v) hard[v] = (uint8_t)(var[v] < 0.0f);
    if (post_out) memcpy(post_out, var, (size_t)N * sizeof(float));
    if (iters) *iters = executed;
    const int ok = syndrome_matches(c, hard, syn);
    free(branches); free(var); free(contrib); free(u.values);
    return ok;
}

/* integer normalisation by k/8 with shifts -- [RECALL] AFF3CT Update_rule_NMS,
 * integer specialisation of normalize<>(): only factors in {1/8..8/8}. v >= 0. */
int ora_normalize_eighths(int v, int eighths)
{
    switch (eighths) {
    case 1: return v >> 3;
    case 2: return v >> 2;
    case 3: return (v >> 2) + (v >> 3);
    case 4: return v >> 1;
    case 5: return (v >> 1) + (v >> 3);
    case 6: return (v >> 1) + (v >> 2);
    case 7: return (v >> 1) + (v >> 2) + (v >> 3);
    default: return v;
    }
}

static inline int clipi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ML/BPSK_nrldpc_sim_FP.m:35-94 restated.  Names follow the .m file:
 *   L    total belief (:40)          R    row-processing storage, Slen x z (:42)
 *   treg per-layer register, ti x z  maxqr = msg_max (:4), maxqL = app_max (:5), offset (:6)
 * Extensions (not in the .m): syndrome sign, early stop, NMS, degree-1 rows (min2 := msg_max+1). */
int ora_decode_layered_fixed(const ora_code *c, const int *llr, const uint8_t *syn,
                             int rule, int n_ite, int early_stop,
                             int offset, int norm_eighths, int msg_max, int app_max,
                             int *app, uint8_t *hard, int *iters)
{
    if (!c->is_qc) return -1;
    const int z = c->Z, mb = c->brows, nb = c->bcols, N = c->N;
    const int *B = c->base;
    int Slen = 0, dmax = 0;
    for (int r = 0; r < mb; ++r) {
        int d = 0;
        for (int b = 0; b < nb; ++b) d += (B[r * nb + b] != -1);
        Slen += d; if (d > dmax) dmax = d;
    }
    int *L = (int *)malloc((size_t)N * sizeof(int));
    int *R = (int *)calloc((size_t)Slen * z, sizeof(int));            /* :42 */
    int *treg = (int *)malloc((size_t)dmax * z * sizeof(int));        /* :14 */
    int *S = (int *)malloc((size_t)dmax * sizeof(int));
    int itr = 0, ok = 0;
    memcpy(L, llr, (size_t)N * sizeof(int));                           /* :40  L = rq */

    while (itr < n_ite) {                                              /* :43 */
        int Ri = 0;
        for (int lyr = 0; lyr < mb; ++lyr) {                           /* :45 */
            int ti = 0;
            for (int col = 0; col < nb; ++col) {                       /* :47 ascending columns */
                const int sh = B[lyr * nb + col];
                if (sh == -1) continue;
                int *Lc = L + col * z;
                for (int i = 0; i < z; ++i) Lc[i] -= R[(Ri) * z + i];  /* :51 subtraction */
                for (int i = 0; i < z; ++i)                            /* :53-56 row alignment + clip */
                    treg[ti * z + i] = clipi(Lc[(i + sh) % z], -(msg_max + 1), msg_max);
                ++ti; ++Ri;
            }
            for (int i1 = 0; i1 < z; ++i1) {                           /* :59 min-sum per lane */
                int min1 = 1 << 30, pos = 0, min2 = 1 << 30, parity = 1;
                for (int j = 0; j < ti; ++j) {                         /* :60 first minimum (first index on ties) */
                    const int a = abs(treg[j * z + i1]);
                    if (a < min1) { min1 = a; pos = j; }
                }
                for (int j = 0; j < ti; ++j) {                         /* :61 second minimum */
                    if (j == pos) continue;
                    const int a = abs(treg[j * z + i1]);
                    if (a < min2) min2 = a;
                }
                if (ti == 1) min2 = msg_max + 1;                       /* extension: the .m assumes ti >= 2 */
                for (int j = 0; j < ti; ++j) {                         /* :62-63 zero counts as + */
                    S[j] = (treg[j * z + i1] >= 0) ? 1 : -1;
                    parity *= S[j];
                }
                if (syn && syn[lyr * z + i1]) parity = -parity;       /* extension: syndrome decoding */
                if (rule == ORA_RULE_OMS) {                            /* :65-72 offset */
                    min1 -= offset; if (min1 < 0) min1 = 0;
                    min2 -= offset; if (min2 < 0) min2 = 0;
                } else if (rule == ORA_RULE_NMS) {
                    min1 = ora_normalize_eighths(min1, norm_eighths);
                    min2 = ora_normalize_eighths(min2, norm_eighths);
                }
                for (int j = 0; j < ti; ++j)                           /* :73-75 */
                    treg[j * z + i1] = parity * S[j] * ((j == pos) ? min2 : min1);
            }
            Ri -= ti; ti = 0;                                          /* :79-80 */
            for (int col = 0; col < nb; ++col) {                       /* :81 */
                const int sh = B[lyr * nb + col];
                if (sh == -1) continue;
                int *Lc = L + col * z;
                for (int i = 0; i < z; ++i)                            /* :86 column alignment mul_sh(.., z-sh) */
                    R[Ri * z + i] = treg[ti * z + (i + (z - sh)) % z];
                for (int i = 0; i < z; ++i)                            /* :88-91 addition + clip */
                    Lc[i] = clipi(Lc[i] + R[Ri * z + i], -(app_max + 1), app_max);
                ++ti; ++Ri;
            }
        }
        for (int v = 0; v < N; ++v) hard[v] = (uint8_t)(L[v] < 0);     /* :92 decision */
        ++itr;                                                         /* :93 */
        if (early_stop && (ok = syndrome_matches(c, hard, syn))) break;
    }
    if (n_ite <= 0) for (int v = 0; v < N; ++v) hard[v] = (uint8_t)(L[v] < 0);
    if (!ok) ok = syndrome_matches(c, hard, syn);
    if (app) memcpy(app, L, (size_t)N * sizeof(int));
    if (iters) *iters = itr;
    free(L); free(R); free(treg); free(S);
    return ok;
}

/* Fixed-point flooding min-sum: the flooding schedule of ora_decode_flooding_f32 with
 * integer messages saturated to [-vmax, vmax]; OMS / NMS(k/8) / plain MS (NMS 8/8). */
int ora_decode_flooding_fixed(const ora_code *c, const int *llr, const uint8_t *syn,
                              int rule, int n_ite, int early_stop,
                              int offset, int norm_eighths, int vmax,
                              int *post_out, uint8_t *hard, int *iters)
{
    const int N = c->N, M = c->M, E = c->E;
    int *c2v = (int *)calloc((size_t)E, sizeof(int));
    int *v2c = (int *)malloc((size_t)E * sizeof(int));
    int *post = (int *)malloc((size_t)N * sizeof(int));
    int ite = 0, executed = 0;
    if (rule == ORA_RULE_SPA) { free(c2v); free(v2c); free(post); return -1; }
    for (; ite < n_ite; ++ite) {
        for (int v = 0; v < N; ++v) {
            int sum = 0;
            for (int k = c->col_ptr[v]; k < c->col_ptr[v + 1]; ++k) sum += c2v[c->row_edge[k]];
            const int tmp = llr[v] + sum;
            for (int k = c->col_ptr[v]; k < c->col_ptr[v + 1]; ++k)
                v2c[c->row_edge[k]] = clipi(tmp - c2v[c->row_edge[k]], -vmax, vmax);
        }
        for (int m = 0; m < M; ++m) {
            const int e0 = c->row_ptr[m], d = c->row_ptr[m + 1] - e0;
            int sign = syn ? (syn[m] & 1) : 0, min1 = vmax, min2 = vmax;
            for (int j = 0; j < d; ++j) {
                const int x = v2c[e0 + j], a = abs(x);
                sign ^= (x < 0);
                const int t = a > min1 ? a : min1;
                if (t < min2) min2 = t;
                if (a < min1) min1 = a;
            }
            int cst1, cst2;
            if (rule == ORA_RULE_OMS) {
                cst1 = min2 - offset; if (cst1 < 0) cst1 = 0;
                cst2 = min1 - offset; if (cst2 < 0) cst2 = 0;
            } else {
                cst1 = ora_normalize_eighths(min2, norm_eighths);
                cst2 = ora_normalize_eighths(min1, norm_eighths);
            }
            for (int j = 0; j < d; ++j) {
                const int x = v2c[e0 + j];
                const int mag = (abs(x) == min1) ? cst1 : cst2;
                c2v[e0 + j] = (sign ^ (x < 0)) ? -mag : mag;
            }
        }
        ++executed;
        if (early_stop && ite != n_ite - 1) {
            for (int v = 0; v < N; ++v) {
                int sum = 0;
                for (int k = c->col_ptr[v]; k < c->col_ptr[v + 1]; ++k) sum += c2v[c->row_edge[k]];
                hard[v] = (uint8_t)((llr[v] + sum) < 0);
            }
            if (syndrome_matches(c, hard, syn)) break;
        }
    }
    for (int v = 0; v < N; ++v) {
        int sum = 0;
        for (int k = c->col_ptr[v]; k < c->col_ptr[v + 1]; ++k) sum += c2v[c->row_edge[k]];
        post[v] = llr[v] + sum;
        hard[v] = (uint8_t)(post[v] < 0);
    }
    if (post_out) memcpy(post_out, post, (size_t)N * sizeof(int));
    if (iters) *iters = executed;
    const int ok = syndrome_matches(c, hard, syn);
    free(c2v); free(v2c); free(post);
    return ok;
}

/* ----------------------------------------------------------- batched helpers */
/* frames are independent (ML/BPSK_nrldpc_sim_RM_FP.m:27 parfor): split them over pthreads */

typedef struct {
    const ora_code *c;
    int kind;                 /* 0 = layered fixed i8, 1 = flooding f32 */
    const int8_t *llr8; const float *llrf; const uint8_t *syn;
    int F, rule, n_ite, early_stop, offset, norm_eighths, msg_max, app_max;
    float norm, foffset;
    float *post; uint8_t *hard; int *iters; uint8_t *ok;
    int tid, nt;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const ora_code *c = j->c;
    const int N = c->N, M = c->M;
    int *l = (int *)malloc((size_t)N * sizeof(int));
    for (int f = j->tid; f < j->F; f += j->nt) {
        int it = 0, r;
        const uint8_t *s = j->syn ? j->syn + (size_t)f * M : NULL;
        if (j->kind == 0) {
            for (int v = 0; v < N; ++v) l[v] = j->llr8[(size_t)f * N + v];
            r = ora_decode_layered_fixed(c, l, s, j->rule, j->n_ite, j->early_stop, j->offset, j->norm_eighths,
                                         j->msg_max, j->app_max, NULL, j->hard + (size_t)f * N, &it);
        } else {
            r = ora_decode_flooding_f32(c, j->llrf + (size_t)f * N, s, j->rule, j->n_ite, j->early_stop, 1,
                                        j->norm, j->foffset, j->post ? j->post + (size_t)f * N : NULL,
                                        j->hard + (size_t)f * N, &it);
        }
        if (j->iters) j->iters[f] = it;
        if (j->ok) j->ok[f] = (uint8_t)(r == 1);
    }
    free(l);
    return NULL;
}

static int run_batch(batch_job *proto, int n_threads)
{
    int nt = n_threads;
    if (nt <= 0) { long n = sysconf(_SC_NPROCESSORS_ONLN); nt = n > 0 ? (int)n : 1; }
    if (nt > proto->F) nt = proto->F > 0 ? proto->F : 1;
    pthread_t *th = (pthread_t *)malloc((size_t)nt * sizeof(pthread_t));
    batch_job *jobs = (batch_job *)malloc((size_t)nt * sizeof(batch_job));
    for (int t = 0; t < nt; ++t) {
        jobs[t] = *proto; jobs[t].tid = t; jobs[t].nt = nt;
        if (t > 0) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    batch_worker(&jobs[0]);
    for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return nt;
}

int ora_batch_layered_fixed_i8(const ora_code *c, const int8_t *llr, const uint8_t *syn, int F,
                               int rule, int n_ite, int early_stop,
                               int offset, int norm_eighths, int msg_max, int app_max,
                               uint8_t *hard, int *iters, uint8_t *ok, int n_threads)
{
    batch_job j; memset(&j, 0, sizeof(j));
    j.c = c; j.kind = 0; j.llr8 = llr; j.syn = syn; j.F = F; j.rule = rule; j.n_ite = n_ite;
    j.early_stop = early_stop; j.offset = offset; j.norm_eighths = norm_eighths;
    j.msg_max = msg_max; j.app_max = app_max; j.hard = hard; j.iters = iters; j.ok = ok;
    return run_batch(&j, n_threads);
}

int ora_batch_flooding_f32(const ora_code *c, const float *llr, const uint8_t *syn, int F,
                           int rule, int n_ite, int early_stop, float norm, float offset,
                           float *post, uint8_t *hard, int *iters, uint8_t *ok, int n_threads)
{
    batch_job j; memset(&j, 0, sizeof(j));
    j.c = c; j.kind = 1; j.llrf = llr; j.syn = syn; j.F = F; j.rule = rule; j.n_ite = n_ite;
    j.early_stop = early_stop; j.norm = norm; j.foffset = offset;
    j.post = post; j.hard = hard; j.iters = iters; j.ok = ok;
    return run_batch(&j, n_threads);
}

/* ---- privacy amplification (errorcorrection/subcomponents/priv_amp.c:186-218, rnd.c:118-127) ------------------- */
unsigned int ora_prng32(unsigned int *state)
{
    for (int k0 = 32; k0; k0--) {                                   /* rnd.c:120-124 */
        const int b = __builtin_parity(*state & 0xe0000200u);       /* PRNG_FEEDBACK, rnd.h:46; calcParity, rnd.h:49 */
        *state <<= 1;
        *state += (unsigned int)b;
    }
    return *state;
}

void ora_privacy_amplify(const uint32_t *key, int workbits, int final_bits, uint32_t seed, uint32_t *out, ora_prng32_fn prng32)
{
    const int numwords = (workbits + 31) / 32;                      /* wordCount, helpers.h:67 */
    uint32_t *k = (uint32_t *)malloc((size_t)(numwords > 0 ? numwords : 1) * sizeof(uint32_t));
    unsigned int state = seed;                                      /* priv_amp.c:187 */
    if (!prng32) prng32 = ora_prng32;
    for (int j = 0; j < numwords; ++j) k[j] = key[j];
    if ((workbits & 31) != 0) k[numwords - 1] &= 0xffffffffu << (32 - (workbits & 31));   /* :189-191 */
    for (int j = 0; j < (final_bits + 31) / 32; ++j) out[j] = 0;    /* :207 */
    for (int i = 0; i < final_bits; i++) {                          /* :213-218 */
        uint32_t m = 0;
        for (int j = 0; j < numwords; j++) m ^= k[j] & prng32(&state);
        if (__builtin_parity(m)) out[i / 32] |= 1u << (31 - (i & 31));   /* uint32AllZeroExceptAtN, helpers.h:69 */
    }
    free(k);
}

uint32_t ora_crc32_words(const uint32_t *words, int n_words)
{
    uint32_t c = 0xffffffffu;
    for (int w = 0; w < n_words; ++w)
        for (int byte = 3; byte >= 0; --byte) {
            c ^= (words[w] >> (8 * byte)) & 0xffu;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
        }
    return c ^ 0xffffffffu;
}

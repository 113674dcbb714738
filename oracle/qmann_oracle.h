/*
 * qmann_oracle.h -- CPU restatement of Q-MANN's quantized MemN2N inference forward.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under q-mann_b200/ may include, link or call this.
 * Allowed users: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference leg.
 *
 * The reference (seongsikpark/Q-MANN) has no live CPU forward (lib/layer.c:254-437 and
 * :1854-1929 are commented out and print "NOT YET FIX CPU MODE"); the normative implementation
 * is the CUDA code in lib/layer_cuda.cu.  This file restates THAT arithmetic on the CPU,
 * following the macros literally (float/double types as the macro expansion produces them):
 *
 *   quantiser / FIXED_MUL / FIXED_ADD      lib/layer_cuda.h:207-259
 *   _cuda_mat_vec_product                  lib/layer_cuda.cu:49-83
 *   _cuda_mat_mat_trans_product            lib/layer_cuda.cu:105-172
 *   _cuda_hamming_similarity               lib/layer_cuda.cu:218-326
 *   _cuda_approximate_attention            lib/layer_cuda.cu:355-541
 *   _cuda_mat_trans_mat_product            lib/layer_cuda.cu:547-635
 *   _cuda_vec_vec_sum                      lib/layer_cuda.cu:1535-1542
 *   _cuda_max / _cuda_max_i                lib/layer_cuda.cu:1895-1939
 *   _cuda_softmax_fwd                      lib/layer_cuda.cu:1969-2060
 *   _cuda_cross_entropy_cost               lib/layer_cuda.cu:2191-2218
 *   layer wiring / per-hop formats         MemN2N/MemN2N.c:714-775, 826-912, 2410-2548, 2626-2697
 *
 * Parity pin: tests/golden/ holds outputs of the UNMODIFIED reference CUDA kernels
 * (oracle/_ref, built from /root/reference by oracle/Makefile) run on a B200; this oracle is
 * checked against them in tests/test_oracle_golden.py.  Known, documented deviation: the
 * reference uses the GPU's __expf (MUFU.EX2); this file uses libm expf, so softmax values agree
 * to ~1e-6 relative, not bit-for-bit.  qmo_forward() reports, per story, how close any
 * attention weight came to a truncation boundary so a test can attribute a downstream difference.
 */
#ifndef QMANN_ORACLE_H
#define QMANN_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMO_MAX_HOP 8

/* fixed-point format: iwl integer bits, frac fractional bits, 1 sign bit (sign-magnitude) */
typedef struct { uint32_t iwl, frac; } qmo_fmt;

typedef struct {
    uint32_t V;          /* dim_input  = dictionary + time-encoding columns          */
    uint32_t d;          /* dim_emb                                                  */
    uint32_t H;          /* NUM_HOP                                                  */
    uint32_t mode;       /* attention mode: 1 float dot, 2 fixed dot, 3 approximate  */
    uint32_t lin_map;    /* EN_LINEAR_MAPPING                                        */
    uint32_t f_fixed;    /* EN_FIXED_POINT                                           */
    int32_t  const_scale;/* ATTENTION_CONST_SCALE (mode 3), define.h:67  (= -3)      */
    qmo_fmt  fmt[QMO_MAX_HOP];      /* iwl[h],frac[h]       (read + update)           */
    qmo_fmt  fmt_w[QMO_MAX_HOP];    /* iwl_w[h],frac_w[h]   (weight-bearing layers)   */
    qmo_fmt  fmt_att[QMO_MAX_HOP];  /* iwl_att[h],frac_att[h]                         */
    qmo_fmt  fmt_bin;               /* iwl_bin,frac_bin                               */
    const float *B;                 /* emb_q      [d][V]                              */
    const float *A[QMO_MAX_HOP];    /* emb_m[h]   [d][V]                              */
    const float *C[QMO_MAX_HOP];    /* emb_c[h]   [d][V]                              */
    const float *Hm[QMO_MAX_HOP];   /* lin_map[h] [d][d]                              */
    const float *W;                 /* ds_ans     [V][d]                              */
} qmo_model;

/* Optional per-story dumps.  Any pointer may be NULL.  Ragged tensors are packed by the story's
 * sentence offset (sum of n_sen of earlier stories), exactly like the m arena. */
typedef struct {
    float    *u0;      /* [N][d]                  emb_q output                         */
    float    *M;       /* [H][sum_sen][d]         emb_m[h] outputs                     */
    float    *C;       /* [H][sum_sen][d]         emb_c[h] outputs                     */
    float    *s;       /* [H][sum_sen]            attention scores                     */
    float    *p;       /* [H][sum_sen]            attention weights (fp32 softmax)     */
    float    *o;       /* [H][N][d]               weighted read                        */
    float    *g;       /* [H][N][d]               lin_map output                       */
    float    *u;       /* [H][N][d]               hop update output                    */
    float    *z;       /* [N][V]                  answer logits                        */
    float    *h;       /* [N][V]                  answer probabilities                 */
    uint32_t *pred;    /* [N]                     argmax_last(h)                       */
    float    *h_true;  /* [N]                     h[y] (cost contribution is -h[y])    */
    float    *risk;    /* [N]  min over hops/slots of the distance of p*2^frac to an integer,
                               in units of 2^-frac (small => a 1-ulp change of p can flip Q(p)) */
    float    *risk_ans;/* [N]  (h_top - h_second)/h_top of the answer softmax           */
} qmo_dump;

/* scalar helpers (exposed for unit tests) --------------------------------------------------- */
/* CUDA_FLOAT2FIXED: sign-magnitude code word (bit 31 = sign).  lib/layer_cuda.h:233,246 */
uint32_t qmo_float2fixed(double x, uint32_t iwl, uint32_t frac);
/* CUDA_FIXED2FLOAT.  lib/layer_cuda.h:247 */
float    qmo_fixed2float(uint32_t code, uint32_t iwl, uint32_t frac);
/* CUDA_FLOAT_QUANT (incl. the binary iwl+frac==0 branch).  lib/layer_cuda.h:253 */
double   qmo_quant(double x, uint32_t iwl, uint32_t frac);
/* CUDA_FIXED_MUL / CUDA_FIXED_ADD.  lib/layer_cuda.h:257-258 */
float    qmo_fixed_mul(float a, float b, uint32_t iwl_a, uint32_t frac_a, uint32_t iwl_b, uint32_t frac_b);
float    qmo_fixed_add(float a, float b, uint32_t iwl_a, uint32_t frac_a, uint32_t iwl_b, uint32_t frac_b);
/* integer closed forms used by the CUDA path; unit tests prove them equal
 * to the float-literal forms above on all 8-bit operands */
int32_t  qmo_int_quant(double x, uint32_t iwl, uint32_t frac);          /* signed integer code */
int32_t  qmo_int_requant(int32_t n, uint32_t frac_from, uint32_t iwl_to, uint32_t frac_to);
int32_t  qmo_int_mul(int32_t a, int32_t b, uint32_t iwl_a, uint32_t frac_a, uint32_t frac_b);
/* one element of _cuda_approximate_attention before the const scale: +-(sum of 2^-i over matching
 * bits i=1..num_bit-1).  lib/layer_cuda.cu:384-428, 218-326 */
float    qmo_appx_element(float m, float v, uint32_t iwl, uint32_t num_bit);

/* layer-level functions (one call == one reference kernel launch) --------------------------- */
void qmo_mat_vec_product(const float *mat, const float *vec, float *out, uint32_t dim_out, uint32_t dim_in,
                         int f_fixed, qmo_fmt fm, qmo_fmt fv);
void qmo_mat_mat_trans_product(const float *a, const float *b, float *out, uint32_t rows, uint32_t cols,
                               uint32_t dim_in, int f_fixed, qmo_fmt fm, qmo_fmt fv, qmo_fmt fout);
void qmo_mat_trans_mat_product(const float *p, const float *Cm, float *out, uint32_t S, uint32_t d,
                               int f_fixed, qmo_fmt f);
void qmo_approximate_attention(const float *M, const float *u, float *out, uint32_t S, uint32_t d,
                               uint32_t iwl, uint32_t num_bit, int32_t const_scale);
void qmo_softmax(const float *in, float *out, uint32_t dim);
void qmo_expf_shifted(const float *in, float mx, float *out, uint32_t dim);
void qmo_vec_vec_sum(const float *a, const float *b, float *out, uint32_t dim, int f_fixed, qmo_fmt f);
uint32_t qmo_argmax_last(const float *in, uint32_t dim);

/* whole forward over N stories ------------------------------------------------------------- */
/* m: packed [sum n_sen][V] dense BoW, q: [N][V], a: [N][V] one-hot or NULL, n_sen: [N].
 * Returns the match count (0 when a == NULL); *cost accumulates -h[y] in story order (fp32). */
uint32_t qmo_forward(const qmo_model *mdl, const float *m, const float *q, const float *a,
                     const uint32_t *n_sen, uint32_t N, qmo_dump *dump, float *cost, int n_threads);

/* number of OpenMP threads the library will use for n_threads <= 0 */
int qmo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

/*
 * qmann_abi.h -- C ABI of libqmann_b200.so, the B200 (sm_100a) implementation of Q-MANN's
 * quantized MemN2N inference forward.
 *
 * Part 1 re-declares the reference's `extern "C" void cuda_*` surface (lib/layer_cuda.cu:2296-4943).
 *   The reference ships NO header for it: lib/layer.c and MemN2N/MemN2N.c call these through
 *   implicit declarations (MemN2N/Makefile:16 `-w`).  libqmann_b200.so exports every symbol
 *   with exactly the reference's parameter list, so the unmodified layer.o / MemN2N.o link
 *   against it instead of layer_cuda.o.  Forward / lifetime / data entry
 *   points are real; training entry points (`*_bwd`, `*_w_up`, ...) exist only to satisfy the
 *   linker and abort with a message (training is out of scope).
 *   Error convention is the reference's: void return, CUDA errors print
 *   "[*E] CUDA : <fn> : <msg>" to stderr and exit(code)          (lib/layer_cuda.h:13-22).
 *
 * Part 2 is the batched entry the reference does not have: one call runs the whole forward of
 *   N stories (MemN2N/MemN2N.c:2377-2703 is one story per 31 launches).  It takes the same
 *   device tensors the reference's layer structs own (fp32 weights, dense fp32 BoW arenas).
 *   These return 0 on success or a negative QMANN_E_* code; qmann_last_error() has the text.
 *
 * All pointers named dev_* are CUDA device pointers; everything else is host memory.
 * No C++ or torch types appear in any signature.
 */
#ifndef QMANN_ABI_H
#define QMANN_ABI_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ============================================================================================
 * Part 1 -- the reference's cuda_* surface.  "ref:" = definition line in lib/layer_cuda.cu.
 * ========================================================================================== */

/* ---- dot_mat_vec (addressing + weighted read) ------------------------------------------- */
void cuda_dot_mat_vec_constructor(float **dev_out_vec, float **dev_grad_out_vec, float **dev_grad_out_mat, float **dev_f_overflow, float **dev_cliff_marker, unsigned int dim_mat_r, unsigned int dim_mat_c, bool f_trans);                                           /* ref: 2297 */
void cuda_dot_mat_vec_init(float *dev_out_vec, float *dev_grad_out_vec, float *dev_grad_out_mat, float *dev_f_overflow, float *dev_cliff_marker, unsigned int dim_mat_r, unsigned int dim_mat_c, bool f_trans);                                                       /* ref: 2365 */
void cuda_dot_mat_vec_fwd(float *dev_in_mat, float *dev_in_vec, float *dev_out_vec, float *dev_f_overflow, unsigned int dim_mat_r, unsigned int dim_mat_c, bool f_trans, bool f_fixed, unsigned int iwl_m, unsigned int frac_m, unsigned int iwl_v, unsigned int frac_v, unsigned int f_mode, bool verbose);   /* ref: 2406 */
void cuda_dot_mat_vec_fwd_appx(float *dev_in_mat, float *dev_in_vec, float *dev_out_vec, float *dev_f_overflow, float *dev_cliff_marker, unsigned int dim_mat_r, unsigned int dim_mat_c, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode, unsigned int num_bit_attention, bool f_trans, bool verbose);   /* ref: 2491 */
void cuda_dot_mat_vec_bwd(float *dev_in_mat, float *dev_in_vec, float *dev_grad_in, float *dev_grad_out_mat, float *dev_grad_out_vec, float *dev_f_overflow, unsigned int dim_mat_r, unsigned int dim_mat_c, bool f_trans, bool f_fixed, unsigned int iwl_m, unsigned int frac_m, unsigned int iwl_v, unsigned int frac_v, unsigned int f_mode, bool verbose);   /* ref: 2561, training */
void cuda_dot_mat_vec_bwd_appx(float *dev_in_mat, float *dev_in_vec, float *dev_grad_in, float *dev_grad_out_mat, float *dev_grad_out_vec, float *dev_f_overflow, float *dev_cliff_marker, unsigned int dim_mat_r, unsigned int dim_mat_c, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode, unsigned int num_bit_attention, bool f_trans, bool verbose, unsigned int hop);   /* ref: 2667, training */
void cuda_dot_mat_vec_destructor(float *dev_out_vec, float *dev_grad_out_vec, float *dev_grad_out_mat, float *dev_f_overflow, float *dev_cliff_marker);   /* ref: 2770 */

/* ---- softmax ---------------------------------------------------------------------------- */
void cuda_softmax_constructor(float **dev_out_vec, float **dev_grad_out, float **dev_max, unsigned int dim);   /* ref: 2792 */
void cuda_softmax_init(float *dev_out_vec, float *dev_grad_out, float *dev_max, unsigned int dim);              /* ref: 2820 */
void cuda_softmax_fwd(float *dev_out_vec, float *dev_in_vec, float *out_vec, float *in_vec, float *dev_max, unsigned int dim, bool f_shift_based, bool verbose);   /* ref: 2845 */
void cuda_softmax_bwd(float *dev_grad_in, float *dev_out_vec, float *dev_grad_out, float *dev_in_vec, unsigned int dim, bool f_shift_based, bool verbose);         /* ref: 2886, training */
void cuda_softmax_destructor(float *dev_out_vec, float *dev_grad_out, float *dev_max);                          /* ref: 2924 */

/* ---- sum_vec (hop update) --------------------------------------------------------------- */
void cuda_sum_vec_constructor(float **dev_out_vec, float **dev_grad_out, unsigned int dim);   /* ref: 2942 */
void cuda_sum_vec_init(float *dev_out_vec, float *dev_grad_out, unsigned int dim);             /* ref: 2969 */
void cuda_sum_vec_fwd(float *dev_in_vec_a, float *dev_in_vec_b, float *dev_out_vec, unsigned int dim, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode, bool verbose);   /* ref: 2992 */
void cuda_sum_vec_bwd(float *dev_grad_out, float *dev_grad_in, float *grad_in, float *grad_out, unsigned int dim);   /* ref: 3032, training */
void cuda_sum_vec_destructor(float *dev_out_vec, float *dev_grad_out);                         /* ref: 3064 */

/* ---- dense (question embedding B, linear mapping H, answer projection W) ---------------- */
void cuda_dense_constructor(float **dev_w_mat, float **dev_w_mat_del, float **dev_w_mat_best, float **dev_bias, float **dev_bias_del, float **dev_out_vec, float **dev_grad_out, float **dev_grad_l2_norm, float **dev_grad_bias_l2_norm, float **dev_f_overflow, unsigned int dim_in, unsigned int dim_out);   /* ref: 3080 */
void cuda_dense_init(float *dev_out_vec, float *dev_grad_out, float *dev_w_mat_del, float *dev_w_mat, float *dev_bias, float *dev_bias_del, float *w_mat, float *bias, float *dev_f_overflow, unsigned int dim_in, unsigned int dim_out);   /* ref: 3126 */
void cuda_dense_fwd(float *dev_w_mat, float *dev_bias, float *dev_in_vec, float *dev_out_vec, float *dev_f_overflow, unsigned int dim_in, unsigned int dim_out, char *activation, bool f_fixed, unsigned int iwl_in, unsigned int frac_in, unsigned int iwl_w, unsigned int frac_w, unsigned int f_mode, bool verbose);   /* ref: 3163 */
void cuda_dense_bwd(float *dev_w_mat, float *dev_w_mat_del, float *dev_bias, float *dev_bias_del, float *dev_in_vec, float *dev_out_vec, float *dev_grad_in, float *dev_grad_out, float *dev_f_overflow, unsigned int dim_in, unsigned int dim_out, char *activation, bool f_fixed, unsigned int iwl_in, unsigned int frac_in, unsigned int iwl_w, unsigned int frac_w, unsigned int f_mode, bool verbose);   /* ref: 3233, training */
void cuda_dense_w_up(float *dev_w_mat, float *dev_w_mat_del, float *dev_bias, float *dev_bias_del, float *dev_grad_l2_norm, float *dev_grad_bias_l2_norm, unsigned int dim_in, unsigned int dim_out, unsigned int batch_size, float *lr, float *lambda, float *max_grad_l2_norm, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode, bool verbose);   /* ref: 3318, training */
void cuda_dense_destructor(float *dev_w_mat, float *dev_w_mat_del, float *dev_w_mat_best, float *dev_out_vec, float *dev_grad_out, float *dev_grad_l2_norm, float *dev_grad_bias_l2_norm, float *dev_f_overflow);   /* ref: 3366 */
void cuda_dense_test_dtoh(float *dev_in_vec, float *in_vec, unsigned int dim_in, unsigned int dim_out);   /* ref: 3390 */
void cuda_dense_test_htod(float *dev_in_vec, float *in_vec, unsigned int dim_in, unsigned int dim_out);   /* ref: 3402 */

/* ---- dense_mat (memory embeddings A / C incl. temporal-encoding columns) ----------------- */
void cuda_dense_mat_constructor(float **dev_w_mat, float **dev_w_mat_del, float **dev_w_mat_best, float **dev_bias, float **dev_bias_del, float **dev_out_mat, float **dev_grad_out, float **dev_grad_l2_norm, float **dev_grad_bias_l2_norm, float **dev_f_overflow, unsigned int dim_in, unsigned int dim_out, unsigned int dim_len);   /* ref: 3418 */
void cuda_dense_mat_init(float *dev_out_mat, float *dev_grad_out, float *dev_w_mat, float *dev_w_mat_del, float *dev_bias, float *dev_bias_del, float *w_mat, float *bias, float *dev_f_overflow, unsigned int dim_in, unsigned int dim_out, unsigned int dim_len);   /* ref: 3469 */
void cuda_dense_mat_fwd(float *dev_w_mat, float *dev_bias, float *dev_in_mat, float *dev_out_mat, float *dev_f_overflow, unsigned int dim_in, unsigned int dim_out, unsigned int dim_len, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode, bool verbose);   /* ref: 3512 */
void cuda_dense_mat_bwd(float *dev_in_mat, float *dev_w_mat, float *dev_w_mat_del, float *dev_bias, float *dev_bias_del, float *dev_grad_in, float *dev_grad_out, float *dev_f_overflow, unsigned int dim_in, unsigned int dim_out, unsigned int dim_len, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode, bool verbose);   /* ref: 3571, training */
void cuda_dense_mat_w_up(float *dev_w_mat, float *dev_w_mat_del, float *dev_bias, float *dev_bias_del, float *dev_grad_l2_norm, float *dev_grad_bias_l2_norm, float *w_mat, float *w_mat_del, unsigned int dim_in, unsigned int dim_out, unsigned int batch_size, float *lr, float *lambda, float *max_grad_l2_norm, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode, bool verbose);   /* ref: 3613, training */
void cuda_dense_mat_destructor(float *dev_w_mat, float *dev_w_mat_del, float *dev_w_mat_best, float *dev_out_mat, float *dev_grad_out, float *dev_grad_l2_norm, float *dev_f_overflow);   /* ref: 3654 */

/* ---- cross_entropy (prediction, cost, match count) -------------------------------------- */
void cuda_cross_entropy_constructor(float **dev_cost_train, float **dev_cost_valid, float **dev_cost_test, unsigned int **dev_m_cnt_train, unsigned int **dev_m_cnt_valid, unsigned int **dev_m_cnt_test, unsigned int **dev_pred_i, float **dev_grad_out, unsigned int dim);   /* ref: 3680 */
void cuda_cross_entropy_init(float *dev_cost_train, float *dev_cost_valid, float *dev_cost_test, unsigned int *dev_m_cnt_train, unsigned int *dev_m_cnt_valid, unsigned int *dev_m_cnt_test, float *dev_grad_out, unsigned int dim);   /* ref: 3717 */
void cuda_cross_entropy_run(float *dev_cost_train, float *dev_cost_valid, float *dev_cost_test, unsigned int *dev_m_cnt_train, unsigned int *dev_m_cnt_valid, unsigned int *dev_m_cnt_test, unsigned int *dev_pred_i, float *cost, float *dev_h, float *dev_y, float *h, float *y, float *dev_grad_out, float *grad_out, unsigned int dim, unsigned int mode);   /* ref: 3750 */
void cuda_cross_entropy_cost_load(float *dev_cost_train, float *dev_cost_valid, float *dev_cost_test, float *cost_train, float *cost_valid, float *cost_test);   /* ref: 3813 */
void cuda_cross_entropy_m_cnt_load(unsigned int *dev_m_cnt_train, unsigned int *dev_m_cnt_valid, unsigned int *dev_m_cnt_test, unsigned int *m_cnt_train, unsigned int *m_cnt_valid, unsigned int *m_cnt_test);   /* ref: 3833 */
void cuda_cross_entropy_destructor(float *dev_cost_train, float *dev_cost_valid, float *dev_cost_test, float *dev_m_cnt_train, float *dev_m_cnt_valid, float *dev_m_cnt_test, float *dev_pred_i, float *dev_grad_out);   /* ref: 3854 */

/* ---- residual-gradient buffers (allocated unconditionally by MemN2N.c:977-984) ---------- */
void cuda_dup_grad_constructor(float **dev_dup_grad, unsigned int num_hop, unsigned int dim);   /* ref: 3885 */
void cuda_dup_grad_bwd(float *dev_dup_grad, float *dotmv_dev_grad_out_vec, float *sv_dev_grad_out_vec, float *dup_grad, unsigned int dim, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode);   /* ref: 3909, training */
void cuda_dup_grad_destructor(float *dev_dup_grad);                                             /* ref: 3949 */

/* ---- data arenas (one H2D of a whole split, MemN2N.c:2294-2350) -------------------------- */
void cuda_data_constructor(float **dev_m, float **dev_q, float **dev_a, unsigned int dim_len, unsigned int dim_in, unsigned int num_sample);   /* ref: 3960 */
void cuda_data_in(float *dev_m, float *dev_q, float *dev_a, float *m, float *q, float *a, unsigned int dim_len, unsigned int dim_in, unsigned int num_sample);   /* ref: 3991 */
void cuda_data_destructor(float *dev_m, float *dev_q, float *dev_a);                           /* ref: 4023 */

/* ---- matrix utilities -------------------------------------------------------------------- */
void cuda_copy_mat(float *dev_src, float *dev_dest, unsigned int dim_col, unsigned int dim_row, bool f_trans);    /* ref: 4117 */
void cuda_accum_mat(float *dev_src, float *dev_dest, unsigned int dim_col, unsigned int dim_row, bool f_trans);   /* ref: 4153, training */
void cuda_set_value(float *dest, float value, unsigned int dim, unsigned int start_idx, unsigned int stride);     /* ref: 4646 */
void cuda_memcpy_dev_to_host(float *host, float *dev, unsigned int size);                                         /* ref: 4436 */
void cuda_copy_dev2host(float *host, float *dev, unsigned int size);                                              /* ref: 4932 */
void cuda_binarization(float *dev_in_vec, unsigned int size);                                                     /* ref: 4454 */
void cuda_quantization(float *dev_in_vec, unsigned int size, unsigned int iwl, unsigned int frac, unsigned int f_mode);   /* ref: 4476 */

/* ---- element-wise product layers (never instantiated by MemN2N.c; link-only) ------------ */
void cuda_mult_e_vec_constructor(float **dev_out_vec, float **dev_grad_out_a, float **dev_grad_out_b, unsigned int dim);   /* ref: 4176 */
void cuda_mult_e_vec_init(float *dev_out_vec, float *dev_grad_out_a, float *dev_grad_out_b, unsigned int dim);             /* ref: 4205 */
void cuda_mult_e_vec_fwd(float *dev_in_vec_a, float *dev_in_vec_b, float *dev_out_vec, float *in_vec_a, float *in_vec_b, float *out_vec, unsigned int dim);   /* ref: 4234 */
void cuda_mult_e_vec_bwd(float *dev_in_vec_a, float *dev_in_vec_b, float *dev_grad_out_a, float *dev_grad_out_b, float *dev_grad_in, float *grad_in, float *grad_out_a, float *grad_out_b, unsigned int dim);   /* ref: 4266 */
void cuda_mult_e_vec_destructor(void);                                                                                    /* ref: 4302 */
void cuda_mult_e_mat_constructor(float **dev_out_mat, float **dev_grad_out_a, float **dev_grad_out_b, unsigned int dim_row, unsigned int dim_col);   /* ref: 4313 */
void cuda_mult_e_mat_init(float *dev_out_mat, float *dev_grad_out_a, float *dev_grad_out_b, unsigned int dim_row, unsigned int dim_col);             /* ref: 4343 */
void cuda_mult_e_mat_fwd(float *dev_in_mat_a, float *dev_in_mat_b, float *dev_out_mat, float *in_mat_a, float *in_mat_b, float *out_mat, unsigned int dim_row, unsigned int dim_col);   /* ref: 4373 */
void cuda_mult_e_mat_bwd(float *dev_in_mat_a, float *dev_in_mat_b, float *dev_grad_out_a, float *dev_grad_out_b, float *dev_grad_in, float *grad_in, float *grad_out_a, float *grad_out_b, unsigned int dim_row, unsigned int dim_col);   /* ref: 4397 */
void cuda_mult_e_mat_destructor(void);                                                                                    /* ref: 4428 */

/* ---- optional forward layers, default off (EN_NON_LINEARITY, EN_SC_ATT) ------------------ */
void cuda_activation_constructor(float **dev_out, float **dev_grad_out, unsigned int dim);   /* ref: 4504 */
void cuda_activation_init(float *dev_out, float *dev_grad_out, unsigned int dim);             /* ref: 4528 */
void cuda_activation_fwd(float *dev_in, float *dev_out, char *type_act, unsigned int dim, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode);   /* ref: 4548 */
void cuda_activation_bwd(float *dev_out, float *dev_grad_in, float *dev_grad_out, char *type_act, unsigned int dim, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode);   /* ref: 4587, training */
void cuda_activation_destructor(float *dev_out, float *dev_grad_out);                         /* ref: 4617 */
void cuda_scale_constructor(float **dev_w, float **dev_w_del, float **dev_w_best, float **dev_out, float **dev_grad_out, unsigned int dim);   /* ref: 4748 */
void cuda_scale_init(float *dev_w, float *dev_w_del, float *dev_out, float *dev_grad_out, float *w, unsigned int dim);                       /* ref: 4778 */
void cuda_scale_fwd(float *dev_in, float *dev_w, float *dev_out, unsigned int dim, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode, bool verbose);   /* ref: 4805 */
void cuda_scale_bwd(float *dev_in, float *dev_grad_in, float *dev_w, float *dev_w_del, float *dev_grad_out, unsigned int dim, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode, bool verbose);   /* ref: 4830, training */
void cuda_scale_w_up(float *dev_w, float *dev_w_del, unsigned int dim, unsigned int batch_size, float *lr, float *lambda, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode, bool verbose);   /* ref: 4861, training */
void cuda_scale_destructor(float *dev_w, float *dev_w_del, float *dev_w_best, float *dev_out, float *dev_grad_out);   /* ref: 4911 */

/* ============================================================================================
 * Part 2 -- batched forward (new).  Replaces the host loop MemN2N/MemN2N.c:2377-2703.
 * ========================================================================================== */

#define QMANN_MAX_HOP 8

enum {
    QMANN_OK = 0,
    QMANN_E_ARG = -1,        /* bad argument / unsupported configuration            */
    QMANN_E_CUDA = -2,       /* a CUDA runtime call failed                          */
    QMANN_E_NOMEM = -3,      /* configuration does not fit the SM's shared memory   */
};

/* Compile-time configuration of the reference (MemN2N/define.h) + argv-derived formats
 * (MemN2N/MemN2N.c:714-775), made run-time. */
typedef struct {
    uint32_t V;                       /* dim_input  (dictionary + time columns)      MemN2N.c:566-582 */
    uint32_t d;                       /* dim_emb                                     define.h:159     */
    uint32_t S_max;                   /* max_line   (memory slots)                   define.h:154     */
    uint32_t H;                       /* NUM_HOP                                     define.h:254     */
    uint32_t mode;                    /* ATTENTION_MODE 2 (fixed dot) | 3 (Hamming)  define.h:15      */
    uint32_t lin_map;                 /* EN_LINEAR_MAPPING                           define.h:291     */
    int32_t  const_scale;             /* ATTENTION_CONST_SCALE (mode 3)              define.h:67      */
    uint32_t iwl[QMANN_MAX_HOP], frac[QMANN_MAX_HOP];           /* read + update     MemN2N.c:714-716 */
    uint32_t iwl_w[QMANN_MAX_HOP], frac_w[QMANN_MAX_HOP];       /* weight layers     MemN2N.c:718-754 */
    uint32_t iwl_att[QMANN_MAX_HOP], frac_att[QMANN_MAX_HOP];   /* addressing        MemN2N.c:721-722 */
    uint32_t iwl_bin, frac_bin;                                 /* u operand         MemN2N.c:767-773 */
    /* optional layers of the reference's graph, default off (zero): */
    uint32_t en_sc_att;               /* EN_SC_ATT (define.h:58): scale layer between scorer and softmax, s' = s * w_h in fp32
                                         (scale_fwd, lib/layer.c:4450; MemN2N.c:852, 2446, 2647)                             */
    float    sc_att_w[QMANN_MAX_HOP]; /* its one weight per hop (scale.w)                                                      */
    uint32_t en_non_lin;              /* EN_NON_LINEARITY (define.h:294): RELU activation layer after the hop update,
                                         u' = Q_(iwl[h],frac[h])(max(u', 0)) (activation_fwd, MemN2N.c:894, 2670)              */
} qmann_config;

/* Device pointers to the fp32 weights as they live in the reference's layer structs
 * (dense.dev_w_mat / dense_mat.dev_w_mat), row-major [dim_out][dim_in]. */
typedef struct {
    const float *dev_B;                       /* emb_q      [d][V]   MemN2N.c:826 */
    const float *dev_A[QMANN_MAX_HOP];        /* emb_m[h]   [d][V]   MemN2N.c:835 */
    const float *dev_C[QMANN_MAX_HOP];        /* emb_c[h]   [d][V]   MemN2N.c:838 */
    const float *dev_Hm[QMANN_MAX_HOP];       /* lin_map[h] [d][d]   MemN2N.c:873 (NULL if !lin_map) */
    const float *dev_W;                       /* ds_ans     [V][d]   MemN2N.c:906 */
} qmann_weights;

/* Optional per-story intermediates for parity checks (any member may be NULL).  fp32, same
 * values the reference's layers would hold; ragged tensors are packed by sentence offset. */
typedef struct {
    float *dev_u0;     /* [N][d]            */
    float *dev_M;      /* [H][sum_sen][d]   */
    float *dev_C;      /* [H][sum_sen][d]   */
    float *dev_s;      /* [H][sum_sen]      */
    float *dev_p;      /* [H][sum_sen]      */
    float *dev_o;      /* [H][N][d]         */
    float *dev_g;      /* [H][N][d]         */
    float *dev_u;      /* [H][N][d]         */
    float *dev_z;      /* [N][V]            */
    float *dev_h;      /* [N][V]            */
    /* -- what the PRODUCTION kernels computed (round 2) -- */
    uint8_t *dev_pcode;  /* [H][sum_sen]  code of Q_f(p) per slot; the slots with a non-zero code are the selected slots of
                                          the weighted read (lib/layer_cuda.cu:561)                                     */
    uint8_t *dev_path;   /* [N]           kernel tier that produced the prediction: 1 packed, 2 unpacked, 3 general    */
    uint8_t *dev_cand;   /* [N][V]        answer rows whose fp32 logit was computed: 1 = prefilter candidate, 2 = every
                                          row was computed (near-tie, h[y] requested, general kernel), 0 = row skipped  */
    uint32_t production; /* 0: every story through the instrumented general kernel (all members above are filled);
                            1: the production tiers (k_story packed / unpacked, then the general kernel for what they
                               decline) with dumps switched on: dev_u0, dev_s, dev_pcode, dev_o, dev_g, dev_u, dev_z (exact
                               rows; -inf where skipped), dev_cand, dev_path.  dev_M/dev_C/dev_p/dev_h are written only for
                               the stories the general kernel finished (dev_path == 3).                                 */
} qmann_debug;

typedef struct qmann_model qmann_model;   /* quantised weight images resident in HBM */
typedef struct qmann_batch qmann_batch;   /* per-story sentence counts / offsets on the device */

/* Quantise the fp32 weights into per-hop int8 images (device) and size the kernels.
 * The model is bound to the CUDA device that is current at the time of the call. */
int  qmann_model_create(qmann_model **out, const qmann_config *cfg, const qmann_weights *w);
void qmann_model_destroy(qmann_model *m);

/* Weight files in the layout of the reference driver's own raw dumps (MemN2N/MemN2N.c:2553-2618 load, :2853-2978 dump; the loops
 * are commented out upstream): per file, per hop, for each input column j, for each output row i, one little-endian fp32
 * w_mat[i][j] -- w_emb_a_float.bin (A, [d][V] x H), w_emb_c_float.bin (C), w_emb_q_float.bin (B), w_float.bin (W, [V][d]) and, in
 * the same convention, w_lin_map_float.bin ([d][d] x H; required when cfg->lin_map).
 * qmann_model_load() replaces those read loops plus the cuda_dense_init / cuda_dense_mat_init uploads (lib/layer.c:1790, :2600);
 * qmann_weights_dump() writes the files from the fp32 device tensors of the layer structs. */
int  qmann_model_load(qmann_model **out, const qmann_config *cfg, const char *dir);
int  qmann_weights_dump(const qmann_config *cfg, const qmann_weights *w, const char *dir);
const char *qmann_weights_last_error(void);

/* Describe a packed batch: n_sen[i] sentences for story i (host array, like n_sen_test_arr,
 * MemN2N.c:2294-2333).  Uploads the offsets once. */
int  qmann_batch_create(qmann_batch **out, const uint32_t *n_sen, uint32_t N);
void qmann_batch_destroy(qmann_batch *b);

/* Forward of all N stories on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 * dev_m [sum_sen][V], dev_q [N][V]: the arenas cuda_data_in() fills.  dev_a [N][V] one-hot or NULL.
 * Outputs (device; any may be NULL): dev_pred[N] predicted answer index (argmax with the
 * reference's last-index tie-break), dev_h_true[N] = h[y] (needs dev_a), *dev_match += #(pred==y).
 * Asynchronous; returns after the launches are enqueued. */
int  qmann_forward_batch(qmann_model *m, const qmann_batch *b, const float *dev_m, const float *dev_q,
                         const float *dev_a, uint32_t *dev_pred, float *dev_h_true, uint32_t *dev_match,
                         const qmann_debug *dbg, void *stream);

/* End-to-end convenience over HOST arenas (pinned or pageable): chunked H2D copies overlapped
 * with the forward, predictions copied back.  Returns the match count through *match (if a). */
int  qmann_infer_host(qmann_model *m, const float *m_host, const float *q_host, const float *a_host,
                      const uint32_t *n_sen, uint32_t N, uint32_t *pred_host, uint32_t *match, float *cost);

/* Word-id input: the lists MemN2N/sample.c:413-575 (sample_vectorization) scatters into
 * the dense arenas, before the scatter.  Rows are story-major: for story i the question row, then its n_sen[i]
 * sentence rows (N + sum_sen rows in all).  row_off[r] .. row_off[r+1] delimit row r inside ids[]; every
 * occurrence of an id adds 1.0 to that column of the row (sample.c:547-568), a sentence's time column
 * V_dict + n_sen-1-j (sample.c:474) is one more id of its row.  ans[i] is the answer column of story i (or NULL).
 * Same outputs and the same arithmetic as qmann_forward_batch on the equivalent dense arenas; ids must be < V. */
int  qmann_forward_ids(qmann_model *m, const qmann_batch *b, const uint16_t *dev_ids, const uint32_t *dev_row_off,
                       const uint32_t *dev_ans, uint32_t *dev_pred, float *dev_h_true, uint32_t *dev_match,
                       const qmann_debug *dbg, void *stream);
int  qmann_infer_ids_host(qmann_model *m, const uint16_t *ids_host, const uint32_t *row_off_host, const uint32_t *ans_host,
                          const uint32_t *n_sen, uint32_t N, uint32_t *pred_host, uint32_t *match, float *cost);

/* Contiguous story range [first, first+count) for `rank` of `world`, balanced by sentence count
 * (batch sharding across GPUs: stories are independent, no collective). */
int  qmann_shard_plan(const uint32_t *n_sen, uint32_t N, uint32_t world, uint32_t rank,
                      uint32_t *first, uint32_t *count);

/* Device-side error flag of the model.  The asynchronous entries (qmann_forward_batch, qmann_forward_ids) cannot
 * return what a kernel finds: a story that overflows the compaction heap or names an id >= V gets the prediction
 * 0xFFFFFFFF (QMANN_PRED_NONE) in dev_pred and sets the flag.  qmann_check_errors() synchronises `stream`, stores the
 * flag in *flags (0 = clean), clears it and returns QMANN_OK, or QMANN_E_CUDA if the stream itself failed.  The host
 * entries (qmann_infer_host, qmann_infer_ids_host) do this themselves: they return a QMANN_E_* code for the call in
 * which it happened and the next call starts clean.
 *
 * Threading: a qmann_model owns ONE set of scratch buffers (compact records, heap, work lists, staging arenas), so it
 * supports one forward in flight at a time.  Calls on the same model must be issued from one thread and on one stream
 * (or be externally ordered); use one model per stream for concurrency (the weight images are a few hundred KB). */
#define QMANN_PRED_NONE 0xFFFFFFFFu
int  qmann_check_errors(qmann_model *m, void *stream, uint32_t *flags);

/* Stories that entered each tier of the production path since the last call (then reset): tiers[0] packed kernel,
 * tiers[1] unpacked kernel, tiers[2] general kernel.  Synchronises `stream`. */
int  qmann_path_counts(qmann_model *m, void *stream, uint64_t tiers[3]);

/* Optional per-kernel timing: while enabled, qmann_forward_batch/qmann_infer_host bracket each of
 * their two kernels with CUDA events on the launching stream.  qmann_profile_read() synchronises
 * on them, returns the summed durations (ms) and the number of (compact, forward) launch pairs,
 * and resets the accumulation. */
int  qmann_profile_enable(qmann_model *m, int enable);
int  qmann_profile_read(qmann_model *m, float *ms_compact, float *ms_forward, uint32_t *n_pairs);

/* ============================================================================================
 * Part 3 -- one very large pre-embedded memory, slot-sharded across GPUs (BASELINE config 5).
 * The reference cannot launch this shape (every dimension is capped at 1024 by its <<<d, S>>> launches,
 * lib/layer_cuda.cu:547-559); the arithmetic per slot is the reference's (scorer lib/layer_cuda.cu:105-141 /
 * :355-541, softmax :1969-2060, weighted read :547-579, linear map :49-68, update :1535-1542).
 *
 * Each rank (one process per GPU) owns S_local contiguous slots [slot0, slot0+S_local) of S_total.  A hop is
 * three calls with one collective after each of the first two (issued by the caller, e.g. torch.distributed
 * all_reduce over NCCL, on the same stream):
 *     qmann_bigmem_hop_scores(h, dev_hist)              local scores + local per-query score histogram
 *         all-reduce SUM of dev_hist   (uint32 [Q][qmann_bigmem_num_bins()])
 *     qmann_bigmem_hop_read(h, dev_hist, dev_partial)   global max/total from the histogram (identical on every
 *                                                       rank, fixed summation order), Q_f(p), local partial read
 *         all-reduce SUM of dev_partial (int32 [Q][d])
 *     qmann_bigmem_hop_update(h, dev_partial)           final clamp, linear map, u_{h+1} = u_h (+) o_h
 * With one rank the collectives are simply skipped.  The result does not depend on the number of shards.
 * ========================================================================================== */
typedef struct qmann_bigmem qmann_bigmem;

/* dev_M[h], dev_C[h]: int8 codes [S_local][d] in hop h's weight format (iwl_w[h], frac_w[h]) -- what
 * emb_m[h] / emb_c[h] output (MemN2N.c:835-838); not copied, must outlive the object AND stay unchanged while it
 * lives (create() derives per-row maxima and, where Q_att(M) != M, a re-quantised int8 copy of M from it).  w: only dev_Hm[h]
 * (fp32 [d][d], if cfg.lin_map) and dev_W (fp32 [V][d], may be NULL with cfg.V == 0) are used.
 * d must be a multiple of 16; cfg.S_max is ignored. */
int  qmann_bigmem_create(qmann_bigmem **out, const qmann_config *cfg, const qmann_weights *w,
                         const int8_t *const *dev_M, const int8_t *const *dev_C,
                         uint64_t S_total, uint64_t slot0, uint64_t S_local, uint32_t Q_max);
void qmann_bigmem_destroy(qmann_bigmem *b);
uint32_t qmann_bigmem_num_bins(const qmann_bigmem *b);
/* Start a batch of Q queries: dev_u0 int8 [Q][d], codes in the hop-0 weight format (the output of emb_q). */
int  qmann_bigmem_begin(qmann_bigmem *b, const int8_t *dev_u0, uint32_t Q, void *stream);
int  qmann_bigmem_hop_scores(qmann_bigmem *b, uint32_t h, uint32_t *dev_hist, void *stream);
/* dev_pbin: optional fp32 [Q][num_bins] attention weight per score bin (parity checks), else NULL */
int  qmann_bigmem_hop_read(qmann_bigmem *b, uint32_t h, const uint32_t *dev_hist, int32_t *dev_partial,
                           float *dev_pbin, void *stream);
/* dev_o, dev_g: optional int8 [Q][d] dumps of the read and of the linear-map output, else NULL */
int  qmann_bigmem_hop_update(qmann_bigmem *b, uint32_t h, const int32_t *dev_partial, int8_t *dev_o, int8_t *dev_g,
                             void *stream);
/* Current controller state u: int8 codes [Q][d] and their fractional bits. */
int  qmann_bigmem_state(qmann_bigmem *b, int8_t *dev_u, int32_t *frac_bits, void *stream);
/* Answer projection (fp32, index order), softmax, argmax_last.  dev_z/dev_h: optional fp32 [Q][V]. */
int  qmann_bigmem_finish(qmann_bigmem *b, uint32_t *dev_pred, float *dev_z, float *dev_h, void *stream);
/* The whole forward of a query batch in ONE call (SURVEY 8b "qmann_forward_sharded"): begin, then per hop scores -> all-reduce(SUM, u32
 * histograms) -> read -> all-reduce(SUM, i32 partial reads) -> update, then the answer projection.  nccl_comm: the caller's ncclComm_t
 * spanning the shards (one rank per GPU, created by the host with ncclCommInitRank), or NULL for a single shard; ncclAllReduce is resolved
 * at run time from the libnccl.so.2 of the process (the library does not link NCCL).  On a non-default stream the sequence is captured once
 * per (u0, Q, pred, comm) into a CUDA graph and replayed (QMANN_BIGMEM_GRAPH=0 disables that).  dev_pred may be NULL (no answer layer);
 * the controller state is available through qmann_bigmem_state() afterwards.  Collective: every rank must make the same calls. */
int  qmann_bigmem_forward_sharded(qmann_bigmem *b, void *nccl_comm, const int8_t *dev_u0, uint32_t Q, uint32_t *dev_pred, void *stream);
/* Optional CUDA-event timing of the dominant kernel (k_big_scores, one launch per hop).  _read folds the pending
 * event pairs (at most one forward's worth: call it after every forward while enabled), synchronising on them. */
int  qmann_bigmem_profile_enable(qmann_bigmem *b, int enable);
int  qmann_bigmem_profile_read(qmann_bigmem *b, float *ms_scores, uint32_t *n_launches, int reset);
const char *qmann_bigmem_last_error(void);

/* Number of kernel launches issued by this library since load (for benchmarks' accounting). */
uint64_t qmann_launch_count(void);
const char *qmann_last_error(void);
const char *qmann_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QMANN_ABI_H */

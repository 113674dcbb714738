/*
 * ref_harness.c -- drives the reference's OWN layer API (lib/layer.h, compiled unmodified from
 * /root/reference by oracle/Makefile) through one inference pass, wired exactly like the test
 * phase of MemN2N/MemN2N.c:2410-2548 (pointer wiring) and :2626-2697 (forward order), on inputs
 * read from a case file, and dumps every intermediate tensor.
 *
 * TEST INFRASTRUCTURE ONLY.  The same source is linked twice:
 *   oracle/_ref/ref_harness_refcuda : layer.o + common.o + the reference's layer_cuda.o
 *                                     -> generator of tests/golden/ (the parity pin)
 *   oracle/_ref/ref_harness_b200    : layer.o + common.o + libqmann_b200.so (our cuda_* shim)
 *                                     -> proves the shim is a link-level drop-in
 * This file contains no reference code; it only calls the reference's public functions.
 *
 * usage: ref_harness <case.bin> <dump.bin> [time_reps]
 *   time_reps > 0: additionally run the forward over all stories time_reps times without
 *   dumping and print "TIME_S <seconds per pass>" (wall clock around cuda_copy_dev2host sync).
 */
#include "layer.h"
#include <stdint.h>
#include <time.h>

bool en_gpu_model = true;   /* MemN2N/MemN2N.h:8-10 define these in the reference driver */
bool en_cpu = false;

#define MAXH 8

/* the cuda_* surface has no header in the reference (implicit declarations); declare what we use */
void cuda_data_constructor(float **dev_m, float **dev_q, float **dev_a, unsigned int dim_len, unsigned int dim_in, unsigned int num_sample);
void cuda_data_in(float *dev_m, float *dev_q, float *dev_a, float *m, float *q, float *a, unsigned int dim_len, unsigned int dim_in, unsigned int num_sample);
void cuda_data_destructor(float *dev_m, float *dev_q, float *dev_a);
void cuda_copy_dev2host(float *host, float *dev, unsigned int size);
void cuda_dense_init(float *dev_out_vec, float *dev_grad_out, float *dev_w_mat_del, float *dev_w_mat, float *dev_bias, float *dev_bias_del, float *w_mat, float *bias, float *dev_f_overflow, unsigned int dim_in, unsigned int dim_out);
void cuda_scale_init(float *dev_w, float *dev_w_del, float *dev_out, float *dev_grad_out, float *w, unsigned int dim);
void cuda_dense_mat_init(float *dev_out_mat, float *dev_grad_out, float *dev_w_mat, float *dev_w_mat_del, float *dev_bias, float *dev_bias_del, float *w_mat, float *bias, float *dev_f_overflow, unsigned int dim_in, unsigned int dim_out, unsigned int dim_len);

static void die(const char *msg) { fprintf(stderr, "ref_harness: %s\n", msg); exit(2); }
static void rd(void *p, size_t sz, size_t n, FILE *f) { if (fread(p, sz, n, f) != n) die("short read"); }
static void wr(const void *p, size_t sz, size_t n, FILE *f) { if (fwrite(p, sz, n, f) != n) die("short write"); }
static float *falloc(size_t n) { float *p = (float *)calloc(n ? n : 1, sizeof(float)); if (!p) die("oom"); return p; }
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int main(int argc, char **argv)
{
    if (argc < 3) die("usage: ref_harness <case.bin> <dump.bin> [time_reps]");
    const int time_reps = argc > 3 ? atoi(argv[3]) : 0;
    FILE *fc = fopen(argv[1], "rb");
    if (!fc) die("cannot open case file");

    char magic[8];
    rd(magic, 1, 8, fc);
    if (memcmp(magic, "QMNCASE1", 8)) die("bad case magic");
    uint32_t V, d, S_max, H, N, mode, lin_map, f_fixed_u; int32_t const_scale;
    rd(&V, 4, 1, fc); rd(&d, 4, 1, fc); rd(&S_max, 4, 1, fc); rd(&H, 4, 1, fc); rd(&N, 4, 1, fc);
    rd(&mode, 4, 1, fc); rd(&lin_map, 4, 1, fc); rd(&f_fixed_u, 4, 1, fc); rd(&const_scale, 4, 1, fc);
    if (H > MAXH) die("too many hops");
    if (const_scale != ATTENTION_CONST_SCALE) die("case const_scale differs from the reference's compile-time ATTENTION_CONST_SCALE");
    uint32_t iwl[MAXH], frac[MAXH], iwl_w[MAXH], frac_w[MAXH], iwl_att[MAXH], frac_att[MAXH], iwl_bin, frac_bin, sum_sen;
    rd(iwl, 4, H, fc); rd(frac, 4, H, fc); rd(iwl_w, 4, H, fc); rd(frac_w, 4, H, fc);
    rd(iwl_att, 4, H, fc); rd(frac_att, 4, H, fc); rd(&iwl_bin, 4, 1, fc); rd(&frac_bin, 4, 1, fc);
    rd(&sum_sen, 4, 1, fc);
    const bool f_fixed = f_fixed_u != 0;
    const unsigned int f_mode = 3;      /* QUANT_MODE, define.h:36-47 */

    float *B = falloc((size_t)d * V), *W = falloc((size_t)V * d);
    float *A[MAXH], *C[MAXH], *Hm[MAXH];
    rd(B, 4, (size_t)d * V, fc);
    for (uint32_t h = 0; h < H; h++) { A[h] = falloc((size_t)d * V); rd(A[h], 4, (size_t)d * V, fc); }
    for (uint32_t h = 0; h < H; h++) { C[h] = falloc((size_t)d * V); rd(C[h], 4, (size_t)d * V, fc); }
    for (uint32_t h = 0; h < H; h++) { Hm[h] = falloc((size_t)d * d); rd(Hm[h], 4, (size_t)d * d, fc); }
    rd(W, 4, (size_t)V * d, fc);
    uint32_t *n_sen = (uint32_t *)calloc(N ? N : 1, 4);
    rd(n_sen, 4, N, fc);
    float *m = falloc((size_t)sum_sen * V), *q = falloc((size_t)N * V), *a = falloc((size_t)N * V);
    rd(m, 4, (size_t)sum_sen * V, fc); rd(q, 4, (size_t)N * V, fc); rd(a, 4, (size_t)N * V, fc);
    /* optional extension: the default-off layers of the reference's graph (EN_SC_ATT, EN_NON_LINEARITY; MemN2N.c:852, 894) */
    uint32_t en_sc_att = 0, en_non_lin = 0;
    float sc_w[MAXH] = {0};
    {
        char ext[8];
        if (fread(ext, 1, 8, fc) == 8) {
            if (memcmp(ext, "QMNEXT01", 8)) die("bad extension magic");
            rd(&en_sc_att, 4, 1, fc); rd(sc_w, 4, H, fc); rd(&en_non_lin, 4, 1, fc);
        }
    }
    fclose(fc);

    FILE *fp_log = fopen("/dev/null", "w");

    /* ---- layer construction, as MemN2N.c:826-912 ---- */
    dense emb_q, lin[MAXH], ds_ans;
    dense_mat emb_m[MAXH], emb_c[MAXH];
    dot_mat_vec dotmv[MAXH], w_sum[MAXH];
    softmax sf_in[MAXH], sf_out;
    sum_vec sv[MAXH];
    scale sc_sf_in[MAXH];
    activation non_lin[MAXH];
    cross_entropy ce;

    dense_constructor(&emb_q, V, d, true, 40.0f, "NULL", f_fixed, iwl_w[0], frac_w[0], iwl_w[0], frac_w[0], f_mode, fp_log);
    for (uint32_t h = 0; h < H; h++) {
        dense_mat_constructor(&emb_m[h], S_max, V, d, true, 40.0f, f_fixed, iwl_w[h], frac_w[h], f_mode, fp_log);
        dense_mat_constructor(&emb_c[h], S_max, V, d, true, 40.0f, f_fixed, iwl_w[h], frac_w[h], f_mode, fp_log);
        if (mode == 2)
            dot_mat_vec_constructor(&dotmv[h], S_max, d, d, false, f_fixed, iwl_att[h], frac_att[h], iwl_bin, frac_bin, f_mode, mode, fp_log);
        else
            dot_mat_vec_constructor(&dotmv[h], S_max, d, d, false, f_fixed, iwl_att[h], frac_att[h], iwl_att[h], frac_att[h], f_mode, mode, fp_log);
        if (en_sc_att) scale_constructor(&sc_sf_in[h], S_max, f_fixed, iwl_att[h], frac_att[h], f_mode, fp_log);      /* MemN2N.c:852-854 */
        softmax_constructor(&sf_in[h], S_max, false, false, fp_log);
        dot_mat_vec_constructor(&w_sum[h], S_max, d, S_max, true, f_fixed, iwl[h], frac[h], iwl[h], frac[h], f_mode, mode, fp_log);
        if (lin_map)
            dense_constructor(&lin[h], d, d, true, 20.0f, "NULL", f_fixed, iwl_bin, frac_bin, iwl_w[h], frac_w[h], f_mode, fp_log);
        sum_vec_constructor(&sv[h], d, f_fixed, iwl[h], frac[h], f_mode, fp_log);
        if (en_non_lin) activation_constructor(&non_lin[h], d, "RELU", f_fixed, iwl[h], frac[h], f_mode, fp_log);    /* MemN2N.c:894-896 */
    }
    dense_constructor(&ds_ans, d, V, true, 40.0f, "NULL", false, 8, 7, 8, 7, f_mode, fp_log);
    softmax_constructor(&sf_out, V, false, false, fp_log);
    cross_entropy_constructor(&ce, V, fp_log);

    /* ---- init (random weights), then overwrite with the case's weights and re-upload ---- */
    dense_init(&emb_q);
    memcpy(emb_q.w_mat[0], B, sizeof(float) * d * V);
    cuda_dense_init(emb_q.dev_out_vec, emb_q.dev_grad_out, emb_q.dev_w_mat_del, emb_q.dev_w_mat, emb_q.dev_bias, emb_q.dev_bias_del, emb_q.w_mat[0], emb_q.bias, emb_q.dev_f_overflow, emb_q.dim_in, emb_q.dim_out);
    for (uint32_t h = 0; h < H; h++) {
        dense_mat_init(&emb_m[h]);
        memcpy(emb_m[h].w_mat[0], A[h], sizeof(float) * d * V);
        cuda_dense_mat_init(emb_m[h].dev_out_mat, emb_m[h].dev_grad_out, emb_m[h].dev_w_mat, emb_m[h].dev_w_mat_del, emb_m[h].dev_bias, emb_m[h].dev_bias_del, emb_m[h].w_mat[0], emb_m[h].bias, emb_m[h].dev_f_overflow, emb_m[h].dim_in, emb_m[h].dim_out, emb_m[h].dim_len_max);
        dense_mat_init(&emb_c[h]);
        memcpy(emb_c[h].w_mat[0], C[h], sizeof(float) * d * V);
        cuda_dense_mat_init(emb_c[h].dev_out_mat, emb_c[h].dev_grad_out, emb_c[h].dev_w_mat, emb_c[h].dev_w_mat_del, emb_c[h].dev_bias, emb_c[h].dev_bias_del, emb_c[h].w_mat[0], emb_c[h].bias, emb_c[h].dev_f_overflow, emb_c[h].dim_in, emb_c[h].dim_out, emb_c[h].dim_len_max);
        /* the reference's cuda_dot_mat_vec_init writes out of bounds when d > max_line
         * (layer_cuda.cu:2393-2394); outputs are fully overwritten by fwd, skip it then */
        if (d <= S_max) dot_mat_vec_init(&dotmv[h]);
        softmax_init(&sf_in[h]);
        if (d <= S_max) dot_mat_vec_init(&w_sum[h]);
        if (lin_map) {
            dense_init(&lin[h]);
            memcpy(lin[h].w_mat[0], Hm[h], sizeof(float) * d * d);
            cuda_dense_init(lin[h].dev_out_vec, lin[h].dev_grad_out, lin[h].dev_w_mat_del, lin[h].dev_w_mat, lin[h].dev_bias, lin[h].dev_bias_del, lin[h].w_mat[0], lin[h].bias, lin[h].dev_f_overflow, lin[h].dim_in, lin[h].dim_out);
        }
        sum_vec_init(&sv[h]);
        if (en_sc_att) {
            scale_init(&sc_sf_in[h]);                       /* random weight; replace it by the case's and upload again */
            *(sc_sf_in[h].w) = sc_w[h];
            cuda_scale_init(sc_sf_in[h].dev_w, sc_sf_in[h].dev_w_del, sc_sf_in[h].dev_out, sc_sf_in[h].dev_grad_out, sc_sf_in[h].w, sc_sf_in[h].dim_max);
        }
        if (en_non_lin) activation_init(&non_lin[h]);
    }
    dense_init(&ds_ans);
    memcpy(ds_ans.w_mat[0], W, sizeof(float) * V * d);
    cuda_dense_init(ds_ans.dev_out_vec, ds_ans.dev_grad_out, ds_ans.dev_w_mat_del, ds_ans.dev_w_mat, ds_ans.dev_bias, ds_ans.dev_bias_del, ds_ans.w_mat[0], ds_ans.bias, ds_ans.dev_f_overflow, ds_ans.dim_in, ds_ans.dim_out);
    softmax_init(&sf_out);
    cross_entropy_init(&ce);

    /* ---- data arenas: one H2D of the packed split, MemN2N.c:2294-2350 ---- */
    float *dev_m, *dev_q, *dev_a;
    cuda_data_constructor(&dev_m, &dev_q, &dev_a, sum_sen ? sum_sen : 1, V, N ? N : 1);
    cuda_data_in(dev_m, dev_q, dev_a, m, q, a, sum_sen, V, N);

    /* ---- dump buffers ---- */
    float *o_u0 = falloc((size_t)N * d), *o_M = falloc((size_t)H * sum_sen * d), *o_C = falloc((size_t)H * sum_sen * d);
    float *o_s = falloc((size_t)H * sum_sen), *o_p = falloc((size_t)H * sum_sen);
    float *o_o = falloc((size_t)H * N * d), *o_g = falloc((size_t)H * N * d), *o_u = falloc((size_t)H * N * d);
    float *o_z = falloc((size_t)N * V), *o_h = falloc((size_t)N * V);
    uint32_t *o_pred = (uint32_t *)calloc(N ? N : 1, 4);

    /* The GPU path never reads the host-side tensors, but softmax_fwd loads in_vec[0] before
     * dispatching (lib/layer.c:1163); hand every *_in() a valid dummy instead of NULL. */
    const size_t dummy_n = (size_t)(V > d ? V : d) > S_max ? (size_t)(V > d ? V : d) : S_max;
    float *hv = falloc(dummy_n);
    float **hm = (float **)calloc(S_max ? S_max : 1, sizeof(float *));
    for (uint32_t r = 0; r < S_max; r++) hm[r] = hv;

    const int passes = 1 + (time_reps > 0 ? time_reps : 0);
    double t_acc = 0.0;
    for (int pass = 0; pass < passes; pass++) {
        const bool dump = (pass == 0);
        if (pass >= 1) cross_entropy_init(&ce);
        const double t0 = now_s();
        size_t addr_m = 0;
        for (uint32_t i = 0; i < N; i++) {
            const uint32_t ns = n_sen[i];
            /* wiring: MemN2N.c:2410-2548 (host tensors are unused by the GPU path: dummies) */
            dense_in(&emb_q, hv, hv, dev_q + (size_t)i * V, NULL);
            float *dev_u = emb_q.dev_out_vec;
            for (uint32_t h = 0; h < H; h++) {
                dense_mat_in(&emb_m[h], ns, hm, hm, dev_m + addr_m * V, NULL);
                dense_mat_in(&emb_c[h], ns, hm, hm, dev_m + addr_m * V, NULL);
                dot_mat_vec_in(&dotmv[h], ns, hm, hv, hv, emb_m[h].dev_out_mat, dev_u, NULL);
                if (en_sc_att) {                                /* MemN2N.c:2446-2448 */
                    scale_in(&sc_sf_in[h], ns, hv, hv, dotmv[h].dev_out_vec, NULL);
                    softmax_in(&sf_in[h], ns, hv, hv, sc_sf_in[h].dev_out, NULL);
                } else
                softmax_in(&sf_in[h], ns, hv, hv, dotmv[h].dev_out_vec, NULL);
                dot_mat_vec_in(&w_sum[h], ns, hm, hv, hv, emb_c[h].dev_out_mat, sf_in[h].dev_out_vec, NULL);
                float *dev_a_in = dev_u;
                if (lin_map) {
                    dense_in(&lin[h], hv, hv, dev_u, NULL);
                    dev_a_in = lin[h].dev_out_vec;
                }
                sum_vec_in(&sv[h], hv, hv, hv, dev_a_in, w_sum[h].dev_out_vec, NULL);
                dev_u = sv[h].dev_out_vec;
                if (en_non_lin) {                               /* MemN2N.c:2423-2431: the next hop reads non_lin[h].out */
                    activation_in(&non_lin[h], hv, hv, sv[h].dev_out_vec, NULL);
                    dev_u = non_lin[h].dev_out;
                }
            }
            dense_in(&ds_ans, hv, hv, dev_u, NULL);
            softmax_in(&sf_out, V, hv, hv, ds_ans.dev_out_vec, NULL);
            cross_entropy_in(&ce, hv, hv, sf_out.dev_out_vec, dev_a + (size_t)i * V);

            /* forward: MemN2N.c:2626-2697 */
            dense_fwd(&emb_q, false);
            if (dump) cuda_copy_dev2host(o_u0 + (size_t)i * d, emb_q.dev_out_vec, d);
            for (uint32_t h = 0; h < H; h++) {
                dense_mat_fwd(&emb_m[h], false);
                dense_mat_fwd(&emb_c[h], false);
                dot_mat_vec_fwd(&dotmv[h], false);
                if (en_sc_att) scale_fwd(&sc_sf_in[h], false);              /* MemN2N.c:2647 */
                softmax_fwd(&sf_in[h], false);
                dot_mat_vec_fwd(&w_sum[h], false);
                if (lin_map) dense_fwd(&lin[h], false);
                sum_vec_fwd(&sv[h], false);
                if (en_non_lin) activation_fwd(&non_lin[h], false);         /* MemN2N.c:2670 */
                if (dump) {
                    const size_t so = (size_t)h * sum_sen + addr_m, vo = ((size_t)h * N + i) * d;
                    if (ns) {
                        cuda_copy_dev2host(o_M + so * d, emb_m[h].dev_out_mat, ns * d);
                        cuda_copy_dev2host(o_C + so * d, emb_c[h].dev_out_mat, ns * d);
                        cuda_copy_dev2host(o_s + so, dotmv[h].dev_out_vec, ns);
                        cuda_copy_dev2host(o_p + so, sf_in[h].dev_out_vec, ns);
                    }
                    cuda_copy_dev2host(o_o + vo, w_sum[h].dev_out_vec, d);
                    if (lin_map) cuda_copy_dev2host(o_g + vo, lin[h].dev_out_vec, d);
                    cuda_copy_dev2host(o_u + vo, en_non_lin ? non_lin[h].dev_out : sv[h].dev_out_vec, d);
                }
            }
            dense_fwd(&ds_ans, false);
            softmax_fwd(&sf_out, false);
            cross_entropy_run(&ce, 3);
            if (dump) {
                cuda_copy_dev2host(o_z + (size_t)i * V, ds_ans.dev_out_vec, V);
                cuda_copy_dev2host(o_h + (size_t)i * V, sf_out.dev_out_vec, V);
                cuda_copy_dev2host((float *)(o_pred + i), (float *)ce.dev_pred_i, 1);
            }
            addr_m += ns;
        }
        float c_tr, c_va, c_te; unsigned int m_tr, m_va, m_te;
        cross_entropy_cost_load(&ce, &c_tr, &c_va, &c_te);       /* D2H => sync, MemN2N.c:2701 */
        cross_entropy_m_cnt_load(&ce, &m_tr, &m_va, &m_te);
        const double t1 = now_s();
        if (pass == 0) {
            FILE *fo = fopen(argv[2], "wb");
            if (!fo) die("cannot open dump file");
            wr("QMNDUMP1", 1, 8, fo);
            wr(&N, 4, 1, fo); wr(&H, 4, 1, fo); wr(&d, 4, 1, fo); wr(&V, 4, 1, fo); wr(&sum_sen, 4, 1, fo);
            wr(o_u0, 4, (size_t)N * d, fo); wr(o_M, 4, (size_t)H * sum_sen * d, fo); wr(o_C, 4, (size_t)H * sum_sen * d, fo);
            wr(o_s, 4, (size_t)H * sum_sen, fo); wr(o_p, 4, (size_t)H * sum_sen, fo);
            wr(o_o, 4, (size_t)H * N * d, fo); wr(o_g, 4, (size_t)H * N * d, fo); wr(o_u, 4, (size_t)H * N * d, fo);
            wr(o_z, 4, (size_t)N * V, fo); wr(o_h, 4, (size_t)N * V, fo); wr(o_pred, 4, N, fo);
            wr(&c_te, 4, 1, fo); wr(&m_te, 4, 1, fo);
            fclose(fo);
            printf("DUMPED N=%u match=%u cost=%.9g\n", N, m_te, (double)c_te);
        } else {
            t_acc += t1 - t0;
        }
    }
    if (time_reps > 0) printf("TIME_S %.9f\n", t_acc / time_reps);

    /* ---- teardown through the reference's destructors ---- */
    cuda_data_destructor(dev_m, dev_q, dev_a);
    dense_destructor(&emb_q);
    for (uint32_t h = 0; h < H; h++) {
        dense_mat_destructor(&emb_m[h]); dense_mat_destructor(&emb_c[h]);
        dot_mat_vec_destructor(&dotmv[h]); softmax_destructor(&sf_in[h]); dot_mat_vec_destructor(&w_sum[h]);
        if (lin_map) dense_destructor(&lin[h]);
        sum_vec_destructor(&sv[h]);
        if (en_sc_att) scale_destructor(&sc_sf_in[h]);
        if (en_non_lin) activation_destructor(&non_lin[h]);
    }
    dense_destructor(&ds_ans); softmax_destructor(&sf_out); cross_entropy_destructor(&ce);
    return 0;
}

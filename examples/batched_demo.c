/* batched_demo.c -- C host code (the reference's language) driving the batched entries of include/qmann_abi.h
 * through plain pointers: the same synthetic stories once as dense fp32 arenas (qmann_infer_host, the reference's
 * boundary format, MemN2N.c:2294-2350) and once as word-id lists (qmann_infer_ids_host, the lists
 * sample_vectorization scatters, MemN2N/sample.c:413-575); both must give the same predictions.
 *
 *   gcc -O2 -I include examples/batched_demo.c -o examples/batched_demo \
 *       -L q-mann_b200 -lqmann_b200 -Wl,-rpath,$PWD/q-mann_b200 -L/usr/local/cuda/lib64 -lcudart -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qmann_abi.h"

/* the four CUDA runtime calls a C driver needs to own device weights (no CUDA headers required) */
extern int cudaMalloc(void **p, size_t n);
extern int cudaMemcpy(void *dst, const void *src, size_t n, int kind);   /* 1 = host to device */
extern int cudaFree(void *p);

static uint64_t rng_state = 0x5EED5EEDull;
static uint32_t rnd(void) { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng_state >> 33); }
static float gauss(void)
{
    const double u1 = (rnd() + 1.0) / 2147483649.0, u2 = rnd() / 2147483648.0;
    return (float)(sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
}
static float *dev_gauss(size_t n, float sigma, size_t zero_col, size_t row_len)
{
    float *h = (float *)malloc(n * sizeof(float)), *d = NULL;
    for (size_t i = 0; i < n; i++) h[i] = sigma * gauss();
    if (row_len) for (size_t i = zero_col; i < n; i += row_len) h[i] = 0.0f;      /* NULL word column, MemN2N.c:1821-1851 */
    if (cudaMalloc((void **)&d, n * sizeof(float)) || cudaMemcpy(d, h, n * sizeof(float), 1)) { fprintf(stderr, "cuda alloc/copy failed\n"); exit(2); }
    free(h);
    return d;
}

int main(void)
{
    enum { V_DICT = 20, S_MAX = 50, V = V_DICT + S_MAX, D = 20, H = 3, N = 500 };
    qmann_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.V = V; cfg.d = D; cfg.S_max = S_MAX; cfg.H = H; cfg.mode = 2; cfg.lin_map = 1; cfg.const_scale = -3;
    for (int h = 0; h < H; h++) {                                 /* run.sh: iwl 5 -> base (5,2); EN_MQ weight formats, MemN2N.c:714-775 */
        cfg.iwl[h] = 5; cfg.frac[h] = 2; cfg.iwl_att[h] = 5; cfg.frac_att[h] = 2;
        cfg.iwl_w[h] = (uint32_t)(6 - h); cfg.frac_w[h] = (uint32_t)(1 + h);
    }
    cfg.iwl_bin = 5; cfg.frac_bin = 2;

    qmann_weights w;
    memset(&w, 0, sizeof(w));
    w.dev_B = dev_gauss((size_t)D * V, 0.5f, 0, V);
    float *A = dev_gauss((size_t)D * V, 0.5f, 0, V), *C = dev_gauss((size_t)D * V, 0.5f, 0, V), *Hm = dev_gauss((size_t)D * D, 0.5f, 0, 0);
    for (int h = 0; h < H; h++) { w.dev_A[h] = A; w.dev_C[h] = C; w.dev_Hm[h] = Hm; }        /* layer-wise tying, define.h:287 */
    w.dev_W = dev_gauss((size_t)V * D, 0.5f, 0, 0);

    qmann_model *model = NULL;
    if (qmann_model_create(&model, &cfg, &w)) { fprintf(stderr, "%s\n", qmann_last_error()); return 1; }

    /* stories: 1..S_MAX sentences of 2..6 word ids + the time id, a 3-word question, an answer id */
    uint32_t *n_sen = (uint32_t *)malloc(N * sizeof(uint32_t)), *ans = (uint32_t *)malloc(N * sizeof(uint32_t));
    size_t sum_sen = 0;
    for (int i = 0; i < N; i++) { n_sen[i] = 1 + rnd() % S_MAX; sum_sen += n_sen[i]; ans[i] = 1 + rnd() % (V_DICT - 1); }
    uint16_t *ids = (uint16_t *)malloc((sum_sen + N) * 8 * sizeof(uint16_t));
    uint32_t *row_off = (uint32_t *)malloc((sum_sen + N + 1) * sizeof(uint32_t));
    float *m = (float *)calloc(sum_sen * V, sizeof(float)), *q = (float *)calloc((size_t)N * V, sizeof(float)), *a = (float *)calloc((size_t)N * V, sizeof(float));
    size_t r = 0, k = 0, srow = 0;
    for (int i = 0; i < N; i++) {
        row_off[r++] = (uint32_t)k;                                /* question row first */
        for (int j = 0; j < 3; j++) { const uint16_t id = (uint16_t)(1 + rnd() % (V_DICT - 1)); ids[k++] = id; q[(size_t)i * V + id] += 1.0f; }
        for (uint32_t s = 0; s < n_sen[i]; s++, srow++) {
            row_off[r++] = (uint32_t)k;
            const int nw = 2 + (int)(rnd() % 5);
            for (int j = 0; j < nw; j++) { const uint16_t id = (uint16_t)(1 + rnd() % (V_DICT - 1)); ids[k++] = id; m[srow * V + id] += 1.0f; }
            const uint16_t te = (uint16_t)(V_DICT + n_sen[i] - s - 1);                     /* sample.c:474 */
            ids[k++] = te; m[srow * V + te] = 1.0f;
        }
        a[(size_t)i * V + ans[i]] = 1.0f;
    }
    row_off[r] = (uint32_t)k;

    uint32_t *pred_dense = (uint32_t *)malloc(N * sizeof(uint32_t)), *pred_ids = (uint32_t *)malloc(N * sizeof(uint32_t));
    uint32_t match_dense = 0, match_ids = 0;
    if (qmann_infer_host(model, m, q, a, n_sen, N, pred_dense, &match_dense, NULL)) { fprintf(stderr, "%s\n", qmann_last_error()); return 1; }
    if (qmann_infer_ids_host(model, ids, row_off, ans, n_sen, N, pred_ids, &match_ids, NULL)) { fprintf(stderr, "%s\n", qmann_last_error()); return 1; }
    int same = (match_dense == match_ids);
    for (int i = 0; i < N; i++) same &= (pred_dense[i] == pred_ids[i]) && pred_dense[i] < V;
    printf("%s: %d stories, %u ids, match %u / %u, launches %llu\n", same ? "DEMO_OK" : "DEMO_MISMATCH", N, (unsigned)k, match_dense, match_ids,
           (unsigned long long)qmann_launch_count());
    qmann_model_destroy(model);
    return same ? 0 : 3;
}

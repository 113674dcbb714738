// qmann_weights.cu -- weight files in the layout of the reference driver's own (commented-out) raw dumps, SURVEY.md 8f-2.
//
// MemN2N/MemN2N.c:2553-2618 (load) and :2853-2978 (dump) read / write, per file, for each hop, for each INPUT column j, for each
// OUTPUT row i, one little-endian fp32 = w_mat[i][j]:
//     w_emb_a_float.bin   emb_m[h].w_mat  [d][V] x H        w_emb_c_float.bin   emb_c[h].w_mat  [d][V] x H
//     w_emb_q_float.bin   emb_q.w_mat     [d][V]            w_float.bin         ds_ans.w_mat    [V][d]
// The reference never dumped lin_map[h].w_mat [d][d]; it is kept in the same convention in w_lin_map_float.bin.
// qmann_model_load() is what a C host calls instead of those loops + the cuda_*_init uploads; qmann_weights_dump() writes the
// files from the fp32 device tensors of the layer structs (dense.dev_w_mat / dense_mat.dev_w_mat).
#include "../../include/qmann_abi.h"

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace {
thread_local std::string g_werr;
int wfail(int code, const std::string &msg) { g_werr = msg; return code; }

// file -> n matrices [dim_out][dim_in] (row-major, as the layer structs hold them), back to back in `dst`
int read_mats(const std::string &path, unsigned n, unsigned dim_out, unsigned dim_in, std::vector<float> &dst)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return wfail(QMANN_E_ARG, path + ": cannot open");
    const size_t per = (size_t)dim_out * dim_in;
    std::vector<float> raw(per);
    dst.assign(per * n, 0.0f);
    for (unsigned k = 0; k < n; k++) {
        if (fread(raw.data(), sizeof(float), per, f) != per) { fclose(f); return wfail(QMANN_E_ARG, path + ": shorter than " + std::to_string(n) + " x [" + std::to_string(dim_out) + "][" + std::to_string(dim_in) + "] fp32"); }
        for (unsigned j = 0; j < dim_in; j++)
            for (unsigned i = 0; i < dim_out; i++) dst[k * per + (size_t)i * dim_in + j] = raw[(size_t)j * dim_out + i];     // file order: for j, for i
    }
    const bool extra = (fgetc(f) != EOF);
    fclose(f);
    if (extra) return wfail(QMANN_E_ARG, path + ": longer than expected");
    return QMANN_OK;
}
int write_mats(const std::string &path, const float *const *dev, unsigned n, unsigned dim_out, unsigned dim_in)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return wfail(QMANN_E_ARG, path + ": cannot create");
    const size_t per = (size_t)dim_out * dim_in;
    std::vector<float> host(per), raw(per);
    for (unsigned k = 0; k < n; k++) {
        if (cudaMemcpy(host.data(), dev[k], per * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) { fclose(f); return wfail(QMANN_E_CUDA, "cudaMemcpy of a weight tensor failed"); }
        for (unsigned j = 0; j < dim_in; j++)
            for (unsigned i = 0; i < dim_out; i++) raw[(size_t)j * dim_out + i] = host[(size_t)i * dim_in + j];
        if (fwrite(raw.data(), sizeof(float), per, f) != per) { fclose(f); return wfail(QMANN_E_ARG, path + ": write failed"); }
    }
    fclose(f);
    return QMANN_OK;
}
}  // namespace

extern "C" {

const char *qmann_weights_last_error(void) { return g_werr.c_str(); }

int qmann_weights_dump(const qmann_config *cfg, const qmann_weights *w, const char *dir)
{
    if (!cfg || !w || !dir) return wfail(QMANN_E_ARG, "null argument");
    const std::string d = std::string(dir) + "/";
    int rc;
    if ((rc = write_mats(d + "w_emb_a_float.bin", w->dev_A, cfg->H, cfg->d, cfg->V))) return rc;
    if ((rc = write_mats(d + "w_emb_c_float.bin", w->dev_C, cfg->H, cfg->d, cfg->V))) return rc;
    if ((rc = write_mats(d + "w_emb_q_float.bin", &w->dev_B, 1, cfg->d, cfg->V))) return rc;
    if ((rc = write_mats(d + "w_float.bin", &w->dev_W, 1, cfg->V, cfg->d))) return rc;
    if (cfg->lin_map && (rc = write_mats(d + "w_lin_map_float.bin", w->dev_Hm, cfg->H, cfg->d, cfg->d))) return rc;
    return QMANN_OK;
}

int qmann_model_load(qmann_model **out, const qmann_config *cfg, const char *dir)
{
    if (!out || !cfg || !dir) return wfail(QMANN_E_ARG, "null argument");
    *out = nullptr;
    if (cfg->H == 0 || cfg->H > QMANN_MAX_HOP) return wfail(QMANN_E_ARG, "H must be in 1..8");
    const std::string d = std::string(dir) + "/";
    std::vector<float> A, C, B, W, Hm;
    int rc;
    if ((rc = read_mats(d + "w_emb_a_float.bin", cfg->H, cfg->d, cfg->V, A))) return rc;
    if ((rc = read_mats(d + "w_emb_c_float.bin", cfg->H, cfg->d, cfg->V, C))) return rc;
    if ((rc = read_mats(d + "w_emb_q_float.bin", 1, cfg->d, cfg->V, B))) return rc;
    if ((rc = read_mats(d + "w_float.bin", 1, cfg->V, cfg->d, W))) return rc;
    if (cfg->lin_map && (rc = read_mats(d + "w_lin_map_float.bin", cfg->H, cfg->d, cfg->d, Hm))) return rc;
    // one device arena for all tensors; the model quantises them into its own images, so it is released afterwards
    const size_t total = A.size() + C.size() + B.size() + W.size() + Hm.size();
    float *dev = nullptr;
    if (cudaMalloc((void **)&dev, total * sizeof(float)) != cudaSuccess) return wfail(QMANN_E_NOMEM, "cudaMalloc of the weight arena failed");
    qmann_weights w = {};
    size_t off = 0;
    bool ok = true;
    auto put = [&](const std::vector<float> &v) { const float *p = dev + off; ok = ok && (v.empty() || cudaMemcpy(dev + off, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess); off += v.size(); return p; };
    const float *pA = put(A), *pC = put(C);
    w.dev_B = put(B);
    w.dev_W = put(W);
    const float *pH = put(Hm);
    for (unsigned h = 0; h < cfg->H; h++) {
        w.dev_A[h] = pA + (size_t)h * cfg->d * cfg->V;
        w.dev_C[h] = pC + (size_t)h * cfg->d * cfg->V;
        w.dev_Hm[h] = cfg->lin_map ? pH + (size_t)h * cfg->d * cfg->d : nullptr;
    }
    if (!ok) { cudaFree(dev); return wfail(QMANN_E_CUDA, "upload of the weights failed"); }
    rc = qmann_model_create(out, cfg, &w);
    if (rc != QMANN_OK) g_werr = qmann_last_error();
    cudaFree(dev);
    return rc;
}

}  // extern "C"

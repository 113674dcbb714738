// qmann_layers.cu -- the reference's per-layer `cuda_*` call surface (one story per call),
// re-implemented for sm_100a behind the exact C ABI of lib/layer_cuda.cu so that the unmodified
// lib/layer.c / MemN2N.c objects link against libqmann_b200.so (include/qmann_abi.h, part 1).
//
// These entry points operate on the reference's own fp32 device tensors, so they use the literal
// quantiser (lit_* in qmann_fixed.cuh).  They are the drop-in/correctness surface; throughput
// comes from the batched path in qmann_forward.cu.  Differences from the reference kernels are
// structural only: one warp per output with shuffle reductions instead of <<<out, in>>> blocks
// whose thread 0 sums serially (lib/layer_cuda.cu:58-66), grid-stride fills instead of the
// mis-sized zero fills, no 1024 limit on any dimension.
//
// Exactness notes.  In fixed-point mode every summand is a multiple of 2^-frac bounded by the
// format maximum, so fp32 sums are exact and order-independent and a tree
// reduction reproduces the reference bit for bit.  In floating-point mode (f_fixed == false: the
// answer projection, attention mode 1) the reference adds in ascending index order; those paths
// run one thread per output, sequentially, with __fmul_rn/__fadd_rn (no FMA contraction).
#include "qmann_fixed.cuh"
#include "qmann_common.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace qmann;

namespace {

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_fill(float *p, float v, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_fill_u32(unsigned *p, unsigned v, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out[r*cols + c] = Q_out( sum_t Q_m( Q_m(a[r][t]) * Q_v(b[c][t]) ) )      one warp per output.
// Covers _cuda_mat_vec_product (a = W rows, b = the single input vector, cols = 1, out fmt = m),
// _cuda_mat_mat_trans_product for dense_mat (a = BoW rows, b = W rows) and for the scorer
// (a = memory rows, b = u, cols = 1).            reference lib/layer_cuda.cu:49-83, 105-172
__global__ void k_rowdot_fixed(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out,
                               unsigned rows, unsigned cols, unsigned dim_in, Fmt fm, Fmt fv, Fmt fo)
{
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows * cols) return;
    const unsigned r = warp / cols, c = warp % cols;
    const float *ar = a + (size_t)r * dim_in, *bc = b + (size_t)c * dim_in;
    float sum = 0.0f;
    for (unsigned t = lane; t < dim_in; t += 32)
        sum += lit_fixed_mul(ar[t], bc[t], fm.iwl, fm.frac, fv.iwl, fv.frac);
    sum = warp_sum(sum);
    if (lane == 0) out[warp] = (float)lit_quant((double)sum, fo.iwl, fo.frac);
}

// floating-point variant: sequential ascending sum of fp32 products, one thread per output
__global__ void k_rowdot_float(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out,
                               unsigned rows, unsigned cols, unsigned dim_in)
{
    const unsigned o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= rows * cols) return;
    const unsigned r = o / cols, c = o % cols;
    const float *ar = a + (size_t)r * dim_in, *bc = b + (size_t)c * dim_in;
    float sum = 0.0f;
    for (unsigned t = 0; t < dim_in; t++) sum = __fadd_rn(sum, __fmul_rn(ar[t], bc[t]));
    out[o] = sum;
}

// weighted read: out[c] = Q( sum_t Q( Q(p[t]) * Q(C[t][c]) ) )   one thread per column c,
// coalesced over c (the reference strides by d inside a block).   lib/layer_cuda.cu:547-635
__global__ void k_weighted_read(const float *__restrict__ p, const float *__restrict__ Cm, float *__restrict__ out,
                                unsigned S, unsigned d, bool f_fixed, Fmt f)
{
    const unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    float sum = 0.0f;
    if (f_fixed) {
        for (unsigned t = 0; t < S; t++) sum += lit_fixed_mul(p[t], Cm[(size_t)t * d + c], f.iwl, f.frac, f.iwl, f.frac);
        out[c] = (float)lit_quant((double)sum, f.iwl, f.frac);
    } else {
        for (unsigned t = 0; t < S; t++) sum = __fadd_rn(sum, __fmul_rn(p[t], Cm[(size_t)t * d + c]));
        out[c] = sum;
    }
}

// approximate (Hamming) attention, one warp per memory slot.     lib/layer_cuda.cu:355-541
__global__ void k_appx_attention(const float *__restrict__ M, const float *__restrict__ u, float *__restrict__ out,
                                 unsigned S, unsigned d, int iwl, int frac, unsigned num_bit, float scale)
{
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= S) return;
    float sum = 0.0f;
    for (unsigned t = lane; t < d; t += 32) {
        unsigned fm = lit_float2fixed((double)M[(size_t)warp * d + t], iwl, frac);
        unsigned fv = lit_float2fixed((double)u[t], iwl, frac);
        const unsigned sm = fm & 0x80000000u, sv = fv & 0x80000000u;
        const unsigned am = fm & 0x7FFFFFFFu, av = fv & 0x7FFFFFFFu;
        const unsigned amin = min(am, av);
        if (sm == sv) { fm = sm | (am - amin); fv = sv | (av - amin); }
        else if (am >= av) { fm = sm + am + amin; fv = sv; }    // three-input add, see appx_element_x128
        else { fm = sm; fv = sv + av + amin; }
        // weighted bit match over bits 30 .. 32-num_bit, weight 2^-i            :261-296
        float sim = 0.0f;
        for (unsigned i = 1; i < num_bit; i++)
            if (((fm ^ fv) & (0x80000000u >> i)) == 0u) sim += __int_as_float((127 - (int)i) << 23);   // 2^-i, exact
        if ((fm ^ fv) & 0x80000000u) sim = -sim;
        float tmp = sim * scale;
        tmp = (float)lit_quant((double)tmp, iwl, frac);
        sum += tmp;
    }
    sum = warp_sum(sum);
    if (lane == 0) out[warp] = (float)lit_quant((double)sum, iwl, frac);
}

// softmax over one vector (single block): max, __expf, double total in ascending order, divide.
//                                                             lib/layer_cuda.cu:1895-1916, 1969-2060
__global__ void k_softmax(const float *__restrict__ in, float *__restrict__ out, float *__restrict__ dev_max,
                          unsigned dim, bool f_shift_based)
{
    __shared__ float red[32];
    __shared__ double total;
    float mx = -INFINITY;
    for (unsigned i = threadIdx.x; i < dim; i += blockDim.x) mx = fmaxf(mx, in[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (unsigned w = 1; w < (blockDim.x + 31) / 32; w++) mx = fmaxf(mx, red[w]);
    for (unsigned i = threadIdx.x; i < dim; i += blockDim.x) out[i] = __expf(in[i] - mx);
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (unsigned i = 0; i < dim; i++) t += out[i];
        total = t;
        if (dev_max) *dev_max = mx;
    }
    __syncthreads();
    const double t = total;
    for (unsigned i = threadIdx.x; i < dim; i += blockDim.x) {
        if (f_shift_based) out[i] = out[i] / llrintf(log2f(t));          // :2038
        else               out[i] = out[i] / t;                          // float / double -> double -> float
    }
}

__global__ void k_vec_sum(const float *a, const float *b, float *out, unsigned dim, bool f_fixed, Fmt f)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    out[i] = f_fixed ? lit_fixed_add(a[i], b[i], f.iwl, f.frac, f.iwl, f.frac) : a[i] + b[i];
}

// argmax with the reference tree's tie-break (left operand wins only on strict '>': the highest
// index among equal maxima), then cost / match / gradient.     lib/layer_cuda.cu:1918-1939, 2191-2250
__global__ void k_cross_entropy(const float *__restrict__ h, const float *__restrict__ y, float *cost, unsigned *m_cnt,
                                unsigned *pred, float *grad_out, unsigned dim)
{
    __shared__ float bv[32];
    __shared__ unsigned bi[32];
    __shared__ unsigned s_pred;
    // per thread: ascending strided scan, a later index replaces the incumbent unless the
    // incumbent is strictly greater; then merge by (value, index)
    float v = -INFINITY;
    unsigned idx = 0xFFFFFFFFu;
    for (unsigned i = threadIdx.x; i < dim; i += blockDim.x) {
        const float x = h[i];
        if (idx == 0xFFFFFFFFu || !(v > x)) { v = x; idx = i; }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const unsigned oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (oi != 0xFFFFFFFFu && (idx == 0xFFFFFFFFu || ov > v || (ov == v && oi > idx))) { v = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = v; bi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float best = -INFINITY; unsigned b = 0xFFFFFFFFu;
        for (unsigned w = 0; w < (blockDim.x + 31) / 32; w++) {
            if (bi[w] == 0xFFFFFFFFu) continue;
            if (b == 0xFFFFFFFFu || bv[w] > best || (bv[w] == best && bi[w] > b)) { best = bv[w]; b = bi[w]; }
        }
        s_pred = b;
        *pred = b;
    }
    __syncthreads();
    const unsigned p = s_pred;
    for (unsigned i = threadIdx.x; i < dim; i += blockDim.x) {
        const bool hot = (y[i] == 1.0f);
        if (hot && cost) {
            *cost = (float)((double)*cost + -1.0 * (double)h[i]);       // one thread has y == 1
            if (i == p) *m_cnt += 1;
        }
        if (grad_out) grad_out[i] = hot ? (float)(1.0 - (double)h[i]) : -h[i];
    }
}

__global__ void k_copy_mat(const float *src, float *dest, unsigned col, unsigned row, bool f_trans)
{
    const size_t n = (size_t)col * row;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned x = (unsigned)(i % col), yy = (unsigned)(i / col);
        dest[f_trans ? ((size_t)x * row + yy) : i] = src[i];
    }
}
__global__ void k_accum_mat(const float *src, float *dest, unsigned col, unsigned row, bool f_trans)
{
    const size_t n = (size_t)col * row;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned x = (unsigned)(i % col), yy = (unsigned)(i / col);
        dest[f_trans ? ((size_t)x * row + yy) : i] += src[i];
    }
}
__global__ void k_set_value(float *dest, float value, unsigned dim, unsigned start_idx, unsigned stride)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < dim; i += (size_t)gridDim.x * blockDim.x)
        if (i % stride == start_idx) dest[i] = value;
}
__global__ void k_quantize_inplace(float *p, size_t n, int iwl, int frac)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = (float)lit_quant((double)p[i], iwl, frac);
}
__global__ void k_binarize_inplace(float *p, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = (p[i] >= 0.0f) ? 1.0f : -1.0f;
}
// activation forward: bypass / sigmoid / relu then quantise     lib/layer_cuda.cu:1664-1703
__global__ void k_activation(const float *in, float *out, unsigned dim, int kind, bool f_fixed, Fmt f)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    double v;
    if (kind == 1) v = 1.0 / (1.0 + (double)expf(-in[i]));
    else if (kind == 2) v = (in[i] > 0.0f) ? in[i] : 0.0f;
    else v = in[i];
    if (kind == 2) v = (float)v;
    out[i] = f_fixed ? (float)lit_quant(v, f.iwl, f.frac) : (float)v;
}
__global__ void k_scale(const float *in, const float *w, float *out, unsigned dim)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dim) out[i] = in[i] * (*w);
}
__global__ void k_mult_e(const float *a, const float *b, float *out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = a[i] * b[i];
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
inline unsigned blocks_for(size_t n, unsigned per_block) { size_t b = (n + per_block - 1) / per_block; return (unsigned)(b ? (b > 65535u * 16u ? 65535u * 16u : b) : 1); }

void fill(float *p, float v, size_t n, const char *fn)
{
    if (!p || !n) return;
    k_fill<<<blocks_for(n, 256), 256>>>(p, v, n);
    count_launch();
    check_cuda(fn, cudaPeekAtLastError());
}
void dmalloc(const char *fn, float **p, size_t n_floats)
{
    check_cuda(fn, cudaMalloc((void **)p, (n_floats ? n_floats : 1) * sizeof(float)));
}
void dfree(const char *fn, void *p) { if (p) check_cuda(fn, cudaFree(p)); }

[[noreturn]] void training_only(const char *fn)
{
    fprintf(stderr, "[*E] qmann_b200 : %s : training (backward / weight update) is outside this library's scope; "
                    "it implements the inference forward of Q-MANN only\n", fn);
    exit(3);
}
void no_binary(const char *fn, unsigned iwl, unsigned frac)
{
    if (iwl + frac == 0) {
        fprintf(stderr, "[*E] qmann_b200 : %s : binary weights (iwl+frac==0) are not supported: the reference's XNOR-style "
                        "output scale accumulates into an uninitialised buffer (lib/layer_cuda.cu:3189-3195)\n", fn);
        exit(3);
    }
}
}  // namespace

// =============================================================================================
// extern "C" surface
// =============================================================================================
extern "C" {

// ---- dot_mat_vec ------------------------------------------------------------------------------
void cuda_dot_mat_vec_constructor(float **dev_out_vec, float **dev_grad_out_vec, float **dev_grad_out_mat, float **dev_f_overflow,
                                  float **dev_cliff_marker, unsigned int dim_mat_r, unsigned int dim_mat_c, bool f_trans)
{
    const char *fn = "cuda_dot_mat_vec_constructor";
    const size_t r = dim_mat_r, c = dim_mat_c;
    dmalloc(fn, dev_out_vec, f_trans ? c : r);
    dmalloc(fn, dev_grad_out_vec, f_trans ? r : c);
    dmalloc(fn, dev_f_overflow, f_trans ? c : r);
    dmalloc(fn, dev_grad_out_mat, r * c);
    dmalloc(fn, dev_cliff_marker, r * c);
    fill(*dev_grad_out_mat, 0.0f, r * c, fn);
}
void cuda_dot_mat_vec_init(float *dev_out_vec, float *dev_grad_out_vec, float *dev_grad_out_mat, float *dev_f_overflow,
                           float *dev_cliff_marker, unsigned int dim_mat_r, unsigned int dim_mat_c, bool f_trans)
{
    const char *fn = "cuda_dot_mat_vec_init";
    const size_t r = dim_mat_r, c = dim_mat_c;
    // sized like the ALLOCATIONS (the reference swaps the two lengths in the non-trans branch,
    // lib/layer_cuda.cu:2393-2394, and writes out of bounds when d > max_line)
    fill(dev_out_vec, 0.0f, f_trans ? c : r, fn);
    fill(dev_grad_out_vec, 0.0f, f_trans ? r : c, fn);
    fill(dev_f_overflow, 0.0f, f_trans ? c : r, fn);
    fill(dev_grad_out_mat, 0.0f, r * c, fn);
    fill(dev_cliff_marker, 0.0f, r * c, fn);
}
void cuda_dot_mat_vec_fwd(float *dev_in_mat, float *dev_in_vec, float *dev_out_vec, float *dev_f_overflow, unsigned int dim_mat_r,
                          unsigned int dim_mat_c, bool f_trans, bool f_fixed, unsigned int iwl_m, unsigned int frac_m,
                          unsigned int iwl_v, unsigned int frac_v, unsigned int f_mode, bool verbose)
{
    const char *fn = "cuda_dot_mat_vec_fwd";
    (void)dev_f_overflow; (void)f_mode; (void)verbose;
    if (dim_mat_r == 0 || dim_mat_c == 0) return;
    const Fmt fm{(int)iwl_m, (int)frac_m}, fv{(int)iwl_v, (int)frac_v};
    if (f_trans) {
        // out[c] = sum_t in_vec[t] * in_mat[t][c]; both operands quantised in the matrix format  :2430
        k_weighted_read<<<blocks_for(dim_mat_c, 128), 128>>>(dev_in_vec, dev_in_mat, dev_out_vec, dim_mat_r, dim_mat_c, f_fixed, fm);
    } else if (f_fixed) {
        k_rowdot_fixed<<<blocks_for((size_t)dim_mat_r * 32, 128), 128>>>(dev_in_mat, dev_in_vec, dev_out_vec, dim_mat_r, 1, dim_mat_c, fm, fv, fm);
    } else {
        k_rowdot_float<<<blocks_for(dim_mat_r, 128), 128>>>(dev_in_mat, dev_in_vec, dev_out_vec, dim_mat_r, 1, dim_mat_c);
    }
    count_launch();
    check_cuda(fn, cudaPeekAtLastError());
}
void cuda_dot_mat_vec_fwd_appx(float *dev_in_mat, float *dev_in_vec, float *dev_out_vec, float *dev_f_overflow, float *dev_cliff_marker,
                               unsigned int dim_mat_r, unsigned int dim_mat_c, bool f_fixed, unsigned int iwl, unsigned int frac,
                               unsigned int f_mode, unsigned int num_bit_attention, bool f_trans, bool verbose)
{
    const char *fn = "cuda_dot_mat_vec_fwd_appx";
    (void)dev_f_overflow; (void)dev_cliff_marker; (void)f_mode; (void)verbose;
    if (dim_mat_r == 0 || dim_mat_c == 0) return;
    if (f_trans) {
        const Fmt f{(int)iwl, (int)frac};
        k_weighted_read<<<blocks_for(dim_mat_c, 128), 128>>>(dev_in_vec, dev_in_mat, dev_out_vec, dim_mat_r, dim_mat_c, f_fixed, f);
    } else {
        // the scorer ignores `frac` and encodes at 32-1-iwl fractional bits            :2515
        k_appx_attention<<<blocks_for((size_t)dim_mat_r * 32, 128), 128>>>(dev_in_mat, dev_in_vec, dev_out_vec, dim_mat_r, dim_mat_c,
                                                                          (int)iwl, 32 - 1 - (int)iwl, num_bit_attention,
                                                                          ldexpf(1.0f, QMANN_ATTENTION_CONST_SCALE));
    }
    count_launch();
    check_cuda(fn, cudaPeekAtLastError());
}
void cuda_dot_mat_vec_bwd(float *, float *, float *, float *, float *, float *, unsigned int, unsigned int, bool, bool, unsigned int,
                          unsigned int, unsigned int, unsigned int, unsigned int, bool) { training_only("cuda_dot_mat_vec_bwd"); }
void cuda_dot_mat_vec_bwd_appx(float *, float *, float *, float *, float *, float *, float *, unsigned int, unsigned int, bool,
                               unsigned int, unsigned int, unsigned int, unsigned int, bool, bool, unsigned int) { training_only("cuda_dot_mat_vec_bwd_appx"); }
void cuda_dot_mat_vec_destructor(float *dev_out_vec, float *dev_grad_out_vec, float *dev_grad_out_mat, float *dev_f_overflow, float *dev_cliff_marker)
{
    const char *fn = "cuda_dot_mat_vec_destructor";
    dfree(fn, dev_out_vec); dfree(fn, dev_grad_out_vec); dfree(fn, dev_grad_out_mat); dfree(fn, dev_f_overflow); dfree(fn, dev_cliff_marker);
}

// ---- softmax ----------------------------------------------------------------------------------
void cuda_softmax_constructor(float **dev_out_vec, float **dev_grad_out, float **dev_max, unsigned int dim)
{
    const char *fn = "cuda_softmax_constructor";
    dmalloc(fn, dev_out_vec, dim); dmalloc(fn, dev_grad_out, dim); dmalloc(fn, dev_max, 1);
}
void cuda_softmax_init(float *dev_out_vec, float *dev_grad_out, float *dev_max, unsigned int dim)
{
    const char *fn = "cuda_softmax_init";
    fill(dev_out_vec, 0.0f, dim, fn); fill(dev_grad_out, 0.0f, dim, fn); fill(dev_max, 0.0f, 1, fn);
}
void cuda_softmax_fwd(float *dev_out_vec, float *dev_in_vec, float *out_vec, float *in_vec, float *dev_max, unsigned int dim,
                      bool f_shift_based, bool verbose)
{
    const char *fn = "cuda_softmax_fwd";
    (void)out_vec; (void)in_vec; (void)verbose;
    if (dim == 0) return;
    const unsigned threads = dim >= 1024 ? 1024 : ((dim + 31) / 32) * 32;
    k_softmax<<<1, threads>>>(dev_in_vec, dev_out_vec, dev_max, dim, f_shift_based);
    count_launch();
    check_cuda(fn, cudaPeekAtLastError());
}
void cuda_softmax_bwd(float *, float *, float *, float *, unsigned int, bool, bool) { training_only("cuda_softmax_bwd"); }
void cuda_softmax_destructor(float *dev_out_vec, float *dev_grad_out, float *dev_max)
{
    const char *fn = "cuda_softmax_destructor";
    dfree(fn, dev_out_vec); dfree(fn, dev_grad_out); dfree(fn, dev_max);
}

// ---- sum_vec ------------------------------------------------------------------------------------
void cuda_sum_vec_constructor(float **dev_out_vec, float **dev_grad_out, unsigned int dim)
{
    dmalloc("cuda_sum_vec_constructor", dev_out_vec, dim); dmalloc("cuda_sum_vec_constructor", dev_grad_out, dim);
}
void cuda_sum_vec_init(float *dev_out_vec, float *dev_grad_out, unsigned int dim)
{
    fill(dev_out_vec, 0.0f, dim, "cuda_sum_vec_init"); fill(dev_grad_out, 0.0f, dim, "cuda_sum_vec_init");
}
void cuda_sum_vec_fwd(float *dev_in_vec_a, float *dev_in_vec_b, float *dev_out_vec, unsigned int dim, bool f_fixed, unsigned int iwl,
                      unsigned int frac, unsigned int f_mode, bool verbose)
{
    (void)f_mode; (void)verbose;
    if (dim == 0) return;
    k_vec_sum<<<blocks_for(dim, 128), 128>>>(dev_in_vec_a, dev_in_vec_b, dev_out_vec, dim, f_fixed, Fmt{(int)iwl, (int)frac});
    count_launch();
    check_cuda("cuda_sum_vec_fwd", cudaPeekAtLastError());
}
void cuda_sum_vec_bwd(float *, float *, float *, float *, unsigned int) { training_only("cuda_sum_vec_bwd"); }
void cuda_sum_vec_destructor(float *dev_out_vec, float *dev_grad_out)
{
    dfree("cuda_sum_vec_destructor", dev_out_vec); dfree("cuda_sum_vec_destructor", dev_grad_out);
}

// ---- dense ----------------------------------------------------------------------------------------
void cuda_dense_constructor(float **dev_w_mat, float **dev_w_mat_del, float **dev_w_mat_best, float **dev_bias, float **dev_bias_del,
                            float **dev_out_vec, float **dev_grad_out, float **dev_grad_l2_norm, float **dev_grad_bias_l2_norm,
                            float **dev_f_overflow, unsigned int dim_in, unsigned int dim_out)
{
    const char *fn = "cuda_dense_constructor";
    const size_t n = (size_t)dim_in * dim_out;
    dmalloc(fn, dev_w_mat, n); dmalloc(fn, dev_w_mat_del, n); dmalloc(fn, dev_w_mat_best, n);
    dmalloc(fn, dev_bias, dim_out); dmalloc(fn, dev_bias_del, dim_out);
    dmalloc(fn, dev_out_vec, dim_out); dmalloc(fn, dev_grad_out, dim_in);
    dmalloc(fn, dev_grad_l2_norm, 1); dmalloc(fn, dev_grad_bias_l2_norm, 1);
    dmalloc(fn, dev_f_overflow, dim_out);
}
void cuda_dense_init(float *dev_out_vec, float *dev_grad_out, float *dev_w_mat_del, float *dev_w_mat, float *dev_bias, float *dev_bias_del,
                     float *w_mat, float *bias, float *dev_f_overflow, unsigned int dim_in, unsigned int dim_out)
{
    const char *fn = "cuda_dense_init";
    const size_t n = (size_t)dim_in * dim_out;
    fill(dev_out_vec, 0.0f, dim_out, fn); fill(dev_grad_out, 0.0f, dim_in, fn); fill(dev_w_mat_del, 0.0f, n, fn);
    fill(dev_bias_del, 0.0f, dim_out, fn); fill(dev_f_overflow, 0.0f, dim_out, fn);
    check_cuda(fn, cudaMemcpy(dev_w_mat, w_mat, n * sizeof(float), cudaMemcpyHostToDevice));
    if (bias) check_cuda(fn, cudaMemcpy(dev_bias, bias, dim_out * sizeof(float), cudaMemcpyHostToDevice));
}
void cuda_dense_fwd(float *dev_w_mat, float *dev_bias, float *dev_in_vec, float *dev_out_vec, float *dev_f_overflow, unsigned int dim_in,
                    unsigned int dim_out, char *activation, bool f_fixed, unsigned int iwl_in, unsigned int frac_in, unsigned int iwl_w,
                    unsigned int frac_w, unsigned int f_mode, bool verbose)
{
    const char *fn = "cuda_dense_fwd";
    (void)dev_bias; (void)dev_f_overflow; (void)f_mode; (void)verbose;     // bias is never used in forward (:3184)
    if (dim_in == 0 || dim_out == 0) return;
    const Fmt fw{(int)iwl_w, (int)frac_w}, fi{(int)iwl_in, (int)frac_in};
    if (f_fixed) {
        no_binary(fn, iwl_w, frac_w);
        k_rowdot_fixed<<<blocks_for((size_t)dim_out * 32, 128), 128>>>(dev_w_mat, dev_in_vec, dev_out_vec, dim_out, 1, dim_in, fw, fi, fw);
    } else {
        k_rowdot_float<<<blocks_for(dim_out, 128), 128>>>(dev_w_mat, dev_in_vec, dev_out_vec, dim_out, 1, dim_in);
    }
    count_launch();
    const int kind = !strcmp(activation, "SIGMOID") ? 1 : (!strcmp(activation, "RELU") ? 2 : 0);
    if (kind) {
        k_activation<<<blocks_for(dim_out, 128), 128>>>(dev_out_vec, dev_out_vec, dim_out, kind, f_fixed, fw);
        count_launch();
    }
    check_cuda(fn, cudaPeekAtLastError());
}
void cuda_dense_bwd(float *, float *, float *, float *, float *, float *, float *, float *, float *, unsigned int, unsigned int, char *, bool,
                    unsigned int, unsigned int, unsigned int, unsigned int, unsigned int, bool) { training_only("cuda_dense_bwd"); }
void cuda_dense_w_up(float *, float *, float *, float *, float *, float *, unsigned int, unsigned int, unsigned int, float *, float *, float *,
                     bool, unsigned int, unsigned int, unsigned int, bool) { training_only("cuda_dense_w_up"); }
void cuda_dense_destructor(float *dev_w_mat, float *dev_w_mat_del, float *dev_w_mat_best, float *dev_out_vec, float *dev_grad_out,
                           float *dev_grad_l2_norm, float *dev_grad_bias_l2_norm, float *dev_f_overflow)
{
    const char *fn = "cuda_dense_destructor";
    dfree(fn, dev_w_mat); dfree(fn, dev_w_mat_del); dfree(fn, dev_w_mat_best); dfree(fn, dev_out_vec); dfree(fn, dev_grad_out);
    dfree(fn, dev_grad_l2_norm); dfree(fn, dev_grad_bias_l2_norm); dfree(fn, dev_f_overflow);
}
void cuda_dense_test_dtoh(float *dev_in_vec, float *in_vec, unsigned int dim_in, unsigned int dim_out)
{
    (void)dim_out;
    check_cuda("cuda_dense_test_dtoh", cudaMemcpy(in_vec, dev_in_vec, (size_t)dim_in * sizeof(float), cudaMemcpyDeviceToHost));
}
void cuda_dense_test_htod(float *dev_in_vec, float *in_vec, unsigned int dim_in, unsigned int dim_out)
{
    (void)dim_out;
    check_cuda("cuda_dense_test_htod", cudaMemcpy(dev_in_vec, in_vec, (size_t)dim_in * sizeof(float), cudaMemcpyHostToDevice));
}

// ---- dense_mat ------------------------------------------------------------------------------------
void cuda_dense_mat_constructor(float **dev_w_mat, float **dev_w_mat_del, float **dev_w_mat_best, float **dev_bias, float **dev_bias_del,
                                float **dev_out_mat, float **dev_grad_out, float **dev_grad_l2_norm, float **dev_grad_bias_l2_norm,
                                float **dev_f_overflow, unsigned int dim_in, unsigned int dim_out, unsigned int dim_len)
{
    const char *fn = "cuda_dense_mat_constructor";
    const size_t n = (size_t)dim_in * dim_out;
    dmalloc(fn, dev_w_mat, n); dmalloc(fn, dev_w_mat_del, n); dmalloc(fn, dev_w_mat_best, n);
    dmalloc(fn, dev_bias, dim_out); dmalloc(fn, dev_bias_del, dim_out);
    dmalloc(fn, dev_out_mat, (size_t)dim_len * dim_out); dmalloc(fn, dev_grad_out, (size_t)dim_len * dim_in);
    dmalloc(fn, dev_grad_l2_norm, 1); dmalloc(fn, dev_grad_bias_l2_norm, 1);
    dmalloc(fn, dev_f_overflow, (size_t)dim_len * dim_out);
}
void cuda_dense_mat_init(float *dev_out_mat, float *dev_grad_out, float *dev_w_mat, float *dev_w_mat_del, float *dev_bias, float *dev_bias_del,
                         float *w_mat, float *bias, float *dev_f_overflow, unsigned int dim_in, unsigned int dim_out, unsigned int dim_len)
{
    const char *fn = "cuda_dense_mat_init";
    const size_t n = (size_t)dim_in * dim_out;
    fill(dev_out_mat, 0.0f, (size_t)dim_len * dim_out, fn); fill(dev_grad_out, 0.0f, (size_t)dim_len * dim_in, fn);
    fill(dev_w_mat_del, 0.0f, n, fn);
    fill(dev_bias_del, 0.0f, dim_out, fn);          // the reference fills dim_out*dim_in floats here (:3500), out of bounds
    fill(dev_f_overflow, 0.0f, (size_t)dim_len * dim_out, fn);
    check_cuda(fn, cudaMemcpy(dev_w_mat, w_mat, n * sizeof(float), cudaMemcpyHostToDevice));
    if (bias) check_cuda(fn, cudaMemcpy(dev_bias, bias, dim_out * sizeof(float), cudaMemcpyHostToDevice));
}
void cuda_dense_mat_fwd(float *dev_w_mat, float *dev_bias, float *dev_in_mat, float *dev_out_mat, float *dev_f_overflow, unsigned int dim_in,
                        unsigned int dim_out, unsigned int dim_len, bool f_fixed, unsigned int iwl, unsigned int frac, unsigned int f_mode,
                        bool verbose)
{
    const char *fn = "cuda_dense_mat_fwd";
    (void)dev_bias; (void)dev_f_overflow; (void)f_mode; (void)verbose;
    if (dim_in == 0 || dim_out == 0 || dim_len == 0) return;
    const Fmt f{(int)iwl, (int)frac};
    const size_t outs = (size_t)dim_len * dim_out;
    if (f_fixed) {
        no_binary(fn, iwl, frac);
        k_rowdot_fixed<<<blocks_for(outs * 32, 128), 128>>>(dev_in_mat, dev_w_mat, dev_out_mat, dim_len, dim_out, dim_in, f, f, f);
    } else {
        k_rowdot_float<<<blocks_for(outs, 128), 128>>>(dev_in_mat, dev_w_mat, dev_out_mat, dim_len, dim_out, dim_in);
    }
    count_launch();
    check_cuda(fn, cudaPeekAtLastError());
}
void cuda_dense_mat_bwd(float *, float *, float *, float *, float *, float *, float *, float *, unsigned int, unsigned int, unsigned int, bool,
                        unsigned int, unsigned int, unsigned int, bool) { training_only("cuda_dense_mat_bwd"); }
void cuda_dense_mat_w_up(float *, float *, float *, float *, float *, float *, float *, float *, unsigned int, unsigned int, unsigned int,
                         float *, float *, float *, bool, unsigned int, unsigned int, unsigned int, bool) { training_only("cuda_dense_mat_w_up"); }
void cuda_dense_mat_destructor(float *dev_w_mat, float *dev_w_mat_del, float *dev_w_mat_best, float *dev_out_mat, float *dev_grad_out,
                               float *dev_grad_l2_norm, float *dev_f_overflow)
{
    const char *fn = "cuda_dense_mat_destructor";
    dfree(fn, dev_w_mat); dfree(fn, dev_w_mat_del); dfree(fn, dev_w_mat_best); dfree(fn, dev_out_mat); dfree(fn, dev_grad_out);
    dfree(fn, dev_grad_l2_norm); dfree(fn, dev_f_overflow);
}

// ---- cross_entropy --------------------------------------------------------------------------------
void cuda_cross_entropy_constructor(float **dev_cost_train, float **dev_cost_valid, float **dev_cost_test, unsigned int **dev_m_cnt_train,
                                    unsigned int **dev_m_cnt_valid, unsigned int **dev_m_cnt_test, unsigned int **dev_pred_i,
                                    float **dev_grad_out, unsigned int dim)
{
    const char *fn = "cuda_cross_entropy_constructor";
    dmalloc(fn, dev_cost_train, 1); dmalloc(fn, dev_cost_valid, 1); dmalloc(fn, dev_cost_test, 1);
    dmalloc(fn, (float **)dev_m_cnt_train, 1); dmalloc(fn, (float **)dev_m_cnt_valid, 1); dmalloc(fn, (float **)dev_m_cnt_test, 1);
    dmalloc(fn, (float **)dev_pred_i, 1); dmalloc(fn, dev_grad_out, dim);
}
void cuda_cross_entropy_init(float *dev_cost_train, float *dev_cost_valid, float *dev_cost_test, unsigned int *dev_m_cnt_train,
                             unsigned int *dev_m_cnt_valid, unsigned int *dev_m_cnt_test, float *dev_grad_out, unsigned int dim)
{
    const char *fn = "cuda_cross_entropy_init";
    fill(dev_cost_train, 0.0f, 1, fn); fill(dev_cost_valid, 0.0f, 1, fn); fill(dev_cost_test, 0.0f, 1, fn);
    fill((float *)dev_m_cnt_train, 0.0f, 1, fn); fill((float *)dev_m_cnt_valid, 0.0f, 1, fn); fill((float *)dev_m_cnt_test, 0.0f, 1, fn);
    fill(dev_grad_out, 0.0f, dim, fn);
}
void cuda_cross_entropy_run(float *dev_cost_train, float *dev_cost_valid, float *dev_cost_test, unsigned int *dev_m_cnt_train,
                            unsigned int *dev_m_cnt_valid, unsigned int *dev_m_cnt_test, unsigned int *dev_pred_i, float *cost, float *dev_h,
                            float *dev_y, float *h, float *y, float *dev_grad_out, float *grad_out, unsigned int dim, unsigned int mode)
{
    const char *fn = "cuda_cross_entropy_run";
    (void)cost; (void)h; (void)y; (void)grad_out;
    if (dim == 0) return;
    float *c = mode == 1 ? dev_cost_train : mode == 2 ? dev_cost_valid : mode == 3 ? dev_cost_test : nullptr;
    unsigned *m = mode == 1 ? dev_m_cnt_train : mode == 2 ? dev_m_cnt_valid : mode == 3 ? dev_m_cnt_test : nullptr;
    const unsigned threads = dim >= 1024 ? 1024 : ((dim + 31) / 32) * 32;
    k_cross_entropy<<<1, threads>>>(dev_h, dev_y, c, m, dev_pred_i, dev_grad_out, dim);
    count_launch();
    check_cuda(fn, cudaPeekAtLastError());
}
void cuda_cross_entropy_cost_load(float *dev_cost_train, float *dev_cost_valid, float *dev_cost_test, float *cost_train, float *cost_valid,
                                  float *cost_test)
{
    const char *fn = "cuda_cross_entropy_cost_load";
    check_cuda(fn, cudaMemcpy(cost_train, dev_cost_train, sizeof(float), cudaMemcpyDeviceToHost));
    check_cuda(fn, cudaMemcpy(cost_valid, dev_cost_valid, sizeof(float), cudaMemcpyDeviceToHost));
    check_cuda(fn, cudaMemcpy(cost_test, dev_cost_test, sizeof(float), cudaMemcpyDeviceToHost));
}
void cuda_cross_entropy_m_cnt_load(unsigned int *dev_m_cnt_train, unsigned int *dev_m_cnt_valid, unsigned int *dev_m_cnt_test,
                                   unsigned int *m_cnt_train, unsigned int *m_cnt_valid, unsigned int *m_cnt_test)
{
    const char *fn = "cuda_cross_entropy_m_cnt_load";
    check_cuda(fn, cudaMemcpy(m_cnt_train, dev_m_cnt_train, sizeof(unsigned), cudaMemcpyDeviceToHost));
    check_cuda(fn, cudaMemcpy(m_cnt_valid, dev_m_cnt_valid, sizeof(unsigned), cudaMemcpyDeviceToHost));
    check_cuda(fn, cudaMemcpy(m_cnt_test, dev_m_cnt_test, sizeof(unsigned), cudaMemcpyDeviceToHost));
}
void cuda_cross_entropy_destructor(float *dev_cost_train, float *dev_cost_valid, float *dev_cost_test, float *dev_m_cnt_train,
                                   float *dev_m_cnt_valid, float *dev_m_cnt_test, float *dev_pred_i, float *dev_grad_out)
{
    const char *fn = "cuda_cross_entropy_destructor";
    dfree(fn, dev_cost_train); dfree(fn, dev_cost_valid); dfree(fn, dev_cost_test); dfree(fn, dev_m_cnt_train);
    dfree(fn, dev_m_cnt_valid); dfree(fn, dev_m_cnt_test); dfree(fn, dev_pred_i); dfree(fn, dev_grad_out);
}

// ---- dup_grad ---------------------------------------------------------------------------------------
void cuda_dup_grad_constructor(float **dev_dup_grad, unsigned int num_hop, unsigned int dim)
{
    dmalloc("cuda_dup_grad_constructor", dev_dup_grad, (size_t)num_hop * dim);
}
void cuda_dup_grad_bwd(float *, float *, float *, float *, unsigned int, bool, unsigned int, unsigned int, unsigned int) { training_only("cuda_dup_grad_bwd"); }
void cuda_dup_grad_destructor(float *dev_dup_grad) { dfree("cuda_dup_grad_destructor", dev_dup_grad); }

// ---- data arenas --------------------------------------------------------------------------------------
void cuda_data_constructor(float **dev_m, float **dev_q, float **dev_a, unsigned int dim_len, unsigned int dim_in, unsigned int num_sample)
{
    const char *fn = "cuda_data_constructor";
    dmalloc(fn, dev_m, (size_t)dim_len * dim_in); dmalloc(fn, dev_q, (size_t)num_sample * dim_in); dmalloc(fn, dev_a, (size_t)num_sample * dim_in);
}
void cuda_data_in(float *dev_m, float *dev_q, float *dev_a, float *m, float *q, float *a, unsigned int dim_len, unsigned int dim_in,
                  unsigned int num_sample)
{
    const char *fn = "cuda_data_in";
    check_cuda(fn, cudaMemcpy(dev_m, m, (size_t)dim_len * dim_in * sizeof(float), cudaMemcpyHostToDevice));
    check_cuda(fn, cudaMemcpy(dev_q, q, (size_t)num_sample * dim_in * sizeof(float), cudaMemcpyHostToDevice));
    check_cuda(fn, cudaMemcpy(dev_a, a, (size_t)num_sample * dim_in * sizeof(float), cudaMemcpyHostToDevice));
}
void cuda_data_destructor(float *dev_m, float *dev_q, float *dev_a)
{
    const char *fn = "cuda_data_destructor";
    dfree(fn, dev_m); dfree(fn, dev_q); dfree(fn, dev_a);
}

// ---- matrix utilities -----------------------------------------------------------------------------------
void cuda_copy_mat(float *dev_src, float *dev_dest, unsigned int dim_col, unsigned int dim_row, bool f_trans)
{
    const size_t n = (size_t)dim_col * dim_row;
    if (!n) return;
    k_copy_mat<<<blocks_for(n, 256), 256>>>(dev_src, dev_dest, dim_col, dim_row, f_trans);
    count_launch();
    check_cuda("cuda_copy_mat", cudaPeekAtLastError());
}
void cuda_accum_mat(float *dev_src, float *dev_dest, unsigned int dim_col, unsigned int dim_row, bool f_trans)
{
    const size_t n = (size_t)dim_col * dim_row;
    if (!n) return;
    k_accum_mat<<<blocks_for(n, 256), 256>>>(dev_src, dev_dest, dim_col, dim_row, f_trans);
    count_launch();
    check_cuda("cuda_accum_mat", cudaPeekAtLastError());
}
void cuda_set_value(float *dest, float value, unsigned int dim, unsigned int start_idx, unsigned int stride)
{
    if (!dim || !stride) return;
    k_set_value<<<blocks_for(dim, 256), 256>>>(dest, value, dim, start_idx, stride);   // bounded by dim (the reference over-runs, :4658)
    count_launch();
    check_cuda("cuda_set_value", cudaPeekAtLastError());
}
void cuda_memcpy_dev_to_host(float *host, float *dev, unsigned int size)
{
    check_cuda("cuda_memcpy_dev_to_host", cudaMemcpy(host, dev, (size_t)size * sizeof(float), cudaMemcpyDeviceToHost));
}
void cuda_copy_dev2host(float *host, float *dev, unsigned int size)
{
    check_cuda("cuda_copy_dev2host", cudaMemcpy(host, dev, (size_t)size * sizeof(float), cudaMemcpyDeviceToHost));
}
void cuda_binarization(float *dev_in_vec, unsigned int size)
{
    if (!size) return;
    k_binarize_inplace<<<blocks_for(size, 256), 256>>>(dev_in_vec, size);
    count_launch();
    check_cuda("cuda_binarization", cudaPeekAtLastError());
}
void cuda_quantization(float *dev_in_vec, unsigned int size, unsigned int iwl, unsigned int frac, unsigned int f_mode)
{
    (void)f_mode;
    if (!size) return;
    k_quantize_inplace<<<blocks_for(size, 256), 256>>>(dev_in_vec, size, (int)iwl, (int)frac);
    count_launch();
    check_cuda("cuda_quantization", cudaPeekAtLastError());
}

// ---- element-wise product layers (never instantiated by the driver) -----------------------------------------
void cuda_mult_e_vec_constructor(float **dev_out_vec, float **dev_grad_out_a, float **dev_grad_out_b, unsigned int dim)
{
    const char *fn = "cuda_mult_e_vec_constructor";
    dmalloc(fn, dev_out_vec, dim); dmalloc(fn, dev_grad_out_a, dim); dmalloc(fn, dev_grad_out_b, dim);
}
void cuda_mult_e_vec_init(float *dev_out_vec, float *dev_grad_out_a, float *dev_grad_out_b, unsigned int dim)
{
    const char *fn = "cuda_mult_e_vec_init";
    fill(dev_out_vec, 0.0f, dim, fn); fill(dev_grad_out_a, 0.0f, dim, fn); fill(dev_grad_out_b, 0.0f, dim, fn);
}
void cuda_mult_e_vec_fwd(float *dev_in_vec_a, float *dev_in_vec_b, float *dev_out_vec, float *, float *, float *, unsigned int dim)
{
    if (!dim) return;
    k_mult_e<<<blocks_for(dim, 256), 256>>>(dev_in_vec_a, dev_in_vec_b, dev_out_vec, dim);
    count_launch();
    check_cuda("cuda_mult_e_vec_fwd", cudaPeekAtLastError());
}
void cuda_mult_e_vec_bwd(float *, float *, float *, float *, float *, float *, float *, float *, unsigned int) { training_only("cuda_mult_e_vec_bwd"); }
void cuda_mult_e_vec_destructor(void) {}
void cuda_mult_e_mat_constructor(float **dev_out_mat, float **dev_grad_out_a, float **dev_grad_out_b, unsigned int dim_row, unsigned int dim_col)
{
    const char *fn = "cuda_mult_e_mat_constructor";
    const size_t n = (size_t)dim_row * dim_col;
    dmalloc(fn, dev_out_mat, n); dmalloc(fn, dev_grad_out_a, n); dmalloc(fn, dev_grad_out_b, n);
}
void cuda_mult_e_mat_init(float *dev_out_mat, float *dev_grad_out_a, float *dev_grad_out_b, unsigned int dim_row, unsigned int dim_col)
{
    const char *fn = "cuda_mult_e_mat_init";
    const size_t n = (size_t)dim_row * dim_col;
    fill(dev_out_mat, 0.0f, n, fn); fill(dev_grad_out_a, 0.0f, n, fn); fill(dev_grad_out_b, 0.0f, n, fn);
}
void cuda_mult_e_mat_fwd(float *dev_in_mat_a, float *dev_in_mat_b, float *dev_out_mat, float *, float *, float *, unsigned int dim_row,
                         unsigned int dim_col)
{
    const size_t n = (size_t)dim_row * dim_col;
    if (!n) return;
    k_mult_e<<<blocks_for(n, 256), 256>>>(dev_in_mat_a, dev_in_mat_b, dev_out_mat, n);
    count_launch();
    check_cuda("cuda_mult_e_mat_fwd", cudaPeekAtLastError());
}
void cuda_mult_e_mat_bwd(float *, float *, float *, float *, float *, float *, float *, float *, unsigned int, unsigned int) { training_only("cuda_mult_e_mat_bwd"); }
void cuda_mult_e_mat_destructor(void) {}

// ---- optional forward layers ------------------------------------------------------------------------------------
void cuda_activation_constructor(float **dev_out, float **dev_grad_out, unsigned int dim)
{
    dmalloc("cuda_activation_constructor", dev_out, dim); dmalloc("cuda_activation_constructor", dev_grad_out, dim);
}
void cuda_activation_init(float *dev_out, float *dev_grad_out, unsigned int dim)
{
    fill(dev_out, 0.0f, dim, "cuda_activation_init"); fill(dev_grad_out, 0.0f, dim, "cuda_activation_init");
}
void cuda_activation_fwd(float *dev_in, float *dev_out, char *type_act, unsigned int dim, bool f_fixed, unsigned int iwl, unsigned int frac,
                         unsigned int f_mode)
{
    (void)f_mode;
    if (!dim) return;
    int kind;
    if (!strcmp(type_act, "NULL")) kind = 0;
    else if (!strcmp(type_act, "SIGMOID")) kind = 1;
    else if (!strcmp(type_act, "RELU")) kind = 2;
    else return;                                           // the reference launches nothing for other names
    k_activation<<<blocks_for(dim, 128), 128>>>(dev_in, dev_out, dim, kind, f_fixed, Fmt{(int)iwl, (int)frac});
    count_launch();
    check_cuda("cuda_activation_fwd", cudaPeekAtLastError());
}
void cuda_activation_bwd(float *, float *, float *, char *, unsigned int, bool, unsigned int, unsigned int, unsigned int) { training_only("cuda_activation_bwd"); }
void cuda_activation_destructor(float *dev_out, float *dev_grad_out)
{
    dfree("cuda_activation_destructor", dev_out); dfree("cuda_activation_destructor", dev_grad_out);
}
void cuda_scale_constructor(float **dev_w, float **dev_w_del, float **dev_w_best, float **dev_out, float **dev_grad_out, unsigned int dim)
{
    const char *fn = "cuda_scale_constructor";
    dmalloc(fn, dev_w, 1); dmalloc(fn, dev_w_del, 1); dmalloc(fn, dev_w_best, 1); dmalloc(fn, dev_out, dim); dmalloc(fn, dev_grad_out, dim);
}
void cuda_scale_init(float *dev_w, float *dev_w_del, float *dev_out, float *dev_grad_out, float *w, unsigned int dim)
{
    const char *fn = "cuda_scale_init";
    fill(dev_w_del, 0.0f, 1, fn); fill(dev_out, 0.0f, dim, fn); fill(dev_grad_out, 0.0f, dim, fn);
    check_cuda(fn, cudaMemcpy(dev_w, w, sizeof(float), cudaMemcpyHostToDevice));
}
void cuda_scale_fwd(float *dev_in, float *dev_w, float *dev_out, unsigned int dim, bool f_fixed, unsigned int iwl, unsigned int frac,
                    unsigned int f_mode, bool verbose)
{
    (void)f_fixed; (void)iwl; (void)frac; (void)f_mode; (void)verbose;     // the reference does not quantise here (:4822)
    if (!dim) return;
    k_scale<<<blocks_for(dim, 128), 128>>>(dev_in, dev_w, dev_out, dim);
    count_launch();
    check_cuda("cuda_scale_fwd", cudaPeekAtLastError());
}
void cuda_scale_bwd(float *, float *, float *, float *, float *, unsigned int, bool, unsigned int, unsigned int, unsigned int, bool) { training_only("cuda_scale_bwd"); }
void cuda_scale_w_up(float *, float *, unsigned int, unsigned int, float *, float *, bool, unsigned int, unsigned int, unsigned int, bool) { training_only("cuda_scale_w_up"); }
void cuda_scale_destructor(float *dev_w, float *dev_w_del, float *dev_w_best, float *dev_out, float *dev_grad_out)
{
    const char *fn = "cuda_scale_destructor";
    dfree(fn, dev_w); dfree(fn, dev_w_del); dfree(fn, dev_w_best); dfree(fn, dev_out); dfree(fn, dev_grad_out);
}

}  // extern "C"

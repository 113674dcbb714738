// qmann_forward.cu -- batched quantized MemN2N inference forward for sm_100a (include/qmann_abi.h
// part 2).  Replaces the reference's one-story-at-a-time host loop (MemN2N/MemN2N.c:2377-2703,
// 31 kernel launches per story) by two kernels per chunk of stories:
//
//   k_compact  : streams the dense fp32 bag-of-words arenas (the reference's boundary format,
//                MemN2N.c:2294-2350) once, HBM-bound, and writes per story an ordered list of its
//                non-zero (column, value) entries with per-row end offsets.
//   k_forward  : one warp per story, persistent CTAs, all quantised weight tables resident in
//                shared memory as int8 codes.  Per hop: bag-of-words embedding as a gather-and-sum
//                of int8 table rows (128-bit shared loads, dp4a accumulation), scorer (fixed-point
//                dot product or Hamming/approximate similarity), fp32 softmax with the reference's
//                sequential double-precision denominator, quantised weights, weighted read over the
//                slots whose quantised weight is non-zero, linear mapping, saturating update; then
//                the fp32 answer projection in the reference's summation order and the argmax.
//
// Arithmetic follows SURVEY.md Appendix A (integer forms proven against the reference by
// tests/golden/kat_*.npz); every formula cites the reference line it reproduces.
#include "qmann_fixed.cuh"
#include "qmann_common.h"
#include "../../include/qmann_abi.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace qmann;

// =============================================================================================
// shared host helpers
// =============================================================================================
namespace qmann {
static std::atomic<unsigned long long> g_launches{0};
static thread_local std::string g_err;

void check_cuda(const char *fn, cudaError_t code)
{
    if (code != cudaSuccess) {
        fprintf(stderr, "[*E] CUDA : %s : %s\n", fn, cudaGetErrorString(code));
        exit((int)code);
    }
}
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace qmann

namespace {

int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define QCUDA(expr)                                                                          \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) return fail(QMANN_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

constexpr int MAXH = QMANN_MAX_HOP;
constexpr unsigned REC_HDR_BYTES = 16;          // n_ent, flags, ans_idx, heap_off
constexpr unsigned FLAG_HEAP = 1u, FLAG_ERROR = 2u;
constexpr unsigned MAX_EXC = 16;                // non-unit BoW values a warp keeps in shared memory per story
constexpr unsigned ANS_NONE = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// kernel parameter blocks
// ---------------------------------------------------------------------------------------------
struct CompactParams {
    const float *m, *q, *a;            // dense fp32 arenas
    const unsigned long long *sen_off; // [N+1] sentence prefix sums
    unsigned V, S_max;
    unsigned story0, n_stories;        // this chunk
    unsigned char *rec;                // per-story records of this chunk
    unsigned rec_stride, off_rend, off_ent, lcap;
    uint2 *heap;                       // overflow entries
    unsigned long long heap_cap;
    unsigned long long *heap_used;
    int vec4;                          // 1: V % 4 == 0 and arenas 16-byte aligned
};

struct FwdParams {
    // quantised images (global) and their layout, copied into shared memory per CTA
    const unsigned char *img;
    unsigned img_bytes;
    unsigned offB, offA[MAXH], offC[MAXH], offH[MAXH], offW;
    unsigned V, d, S_max, H, lin_map;
    int const_scale;
    unsigned DP, HS, WS;
    // formats: fractional bits and code limits
    int fw[MAXH], lw[MAXH], iw[MAXH];      // weight layers
    int fa[MAXH], la[MAXH], ia[MAXH];      // addressing
    int ff[MAXH], lf[MAXH], iff[MAXH];     // read + update
    int fb, lb;                            // u operand of scorer / linear map
    // compact records
    const unsigned char *rec;
    unsigned rec_stride, off_rend, off_ent;
    const uint2 *heap;
    const unsigned long long *sen_off;
    unsigned story0, n_stories, n_total;
    unsigned long long sum_sen;
    // per-warp shared-memory scratch layout
    unsigned warp_bytes, LW, S_pad, o_rend, o_sc, o_ex, o_pq, o_uvec, o_ubvec, o_ovec, o_ufl, o_exc, tables_bytes;
    // outputs
    unsigned *pred;
    float *h_true;
    unsigned *match;
    unsigned *counter;
    unsigned *err_flag;
    int want_h;
    qmann_debug dbg;
};

// =============================================================================================
// weight preparation: fp32 [dim_out][dim_in] -> int8 codes, transposed to [dim_in][row_stride]
// =============================================================================================
// emb tables (B, A_h, C_h): img[v*DP + c] = code(w[c][v]);   CUDA_FLOAT_QUANT of the weight inside
// FIXED_MUL, reference lib/layer_cuda.cu:120 with formats from MemN2N.c:826-838.
__global__ void k_prep_emb(const float *__restrict__ w, signed char *__restrict__ img, unsigned V, unsigned d, unsigned DP, int iwl, int frac)
{
    const size_t n = (size_t)V * DP;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned v = (unsigned)(i / DP), c = (unsigned)(i % DP);
        img[i] = (c < d) ? (signed char)qi_encode(w[(size_t)c * V + v], iwl, frac) : (signed char)0;
    }
}
// linear map Hm_h: img[i*HS + j] = code(Hm[i][j])                           MemN2N.c:873
__global__ void k_prep_lin(const float *__restrict__ w, signed char *__restrict__ img, unsigned d, unsigned HS, int iwl, int frac)
{
    const size_t n = (size_t)d * HS;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned r = (unsigned)(i / HS), c = (unsigned)(i % HS);
        img[i] = (c < d) ? (signed char)qi_encode(w[(size_t)r * d + c], iwl, frac) : (signed char)0;
    }
}
// answer projection W stays fp32 (f_fixed = false, MemN2N.c:902-906), rows padded to WS floats
__global__ void k_prep_ans(const float *__restrict__ w, float *__restrict__ img, unsigned V, unsigned d, unsigned WS)
{
    const size_t n = (size_t)V * WS;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned r = (unsigned)(i / WS), c = (unsigned)(i % WS);
        img[i] = (c < d) ? w[(size_t)r * d + c] : 0.0f;
    }
}

// =============================================================================================
// k_compact: dense fp32 BoW -> ordered (column, value) lists.  One warp per story.
// Record layout (rec_stride bytes per story):
//   +0   u32 n_ent      total entries (question row + all sentence rows)
//   +4   u32 flags      FLAG_HEAP: entries live in the overflow heap at heap_off; FLAG_ERROR: heap exhausted
//   +8   u32 ans_idx    index of the 1.0 in the answer row (last one), ANS_NONE without answers
//   +12  u32 heap_off
//   +off_rend  u16 rend[S_max+2]   rend[k] = end (exclusive) of row k; row 0 is the question, rows 1..S the sentences
//   +off_ent   uint2 ent[lcap]     {column, fp32 bits of the value}
// =============================================================================================
__device__ __forceinline__ float4 ldg_stream4(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float *p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ bool nonzero(float x) { return (__float_as_uint(x) << 1) != 0u; }

// Scans one row of V floats; appends its non-zero entries in ascending column order at `base`
// (warp-uniform running count).  Stores are dropped beyond `cap`.  Returns the new count.
template <bool VEC4>
__device__ __forceinline__ unsigned scan_row(const float *__restrict__ row, unsigned V, uint2 *__restrict__ dst, unsigned cap, unsigned base, unsigned lane)
{
    const unsigned lt = (1u << lane) - 1u;
    if (VEC4) {
        const float4 *r4 = reinterpret_cast<const float4 *>(row);
        const unsigned V4 = V >> 2;
        for (unsigned c0 = 0; c0 < V4; c0 += 64) {
            // two independent 128-bit loads in flight per lane
            const unsigned ca = c0 + lane, cb = c0 + 32 + lane;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (ca < V4) va = ldg_stream4(r4 + ca);
            if (cb < V4) vb = ldg_stream4(r4 + cb);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const float4 v = half ? vb : va;
                const unsigned c4 = half ? cb : ca;
                const unsigned m = (nonzero(v.x) ? 1u : 0u) | (nonzero(v.y) ? 2u : 0u) | (nonzero(v.z) ? 4u : 0u) | (nonzero(v.w) ? 8u : 0u);
                const unsigned any = __ballot_sync(0xffffffffu, m != 0u);
                if (any == 0u) continue;
                const unsigned cnt = __popc(m);
                unsigned pos = base + __popc(any & lt);
                unsigned tot = __popc(any);
                const unsigned b2 = __ballot_sync(0xffffffffu, cnt >= 2u);
                if (b2) {       // rare: two or more non-zeros inside one float4
                    const unsigned b3 = __ballot_sync(0xffffffffu, cnt >= 3u), b4 = __ballot_sync(0xffffffffu, cnt >= 4u);
                    pos += __popc(b2 & lt) + __popc(b3 & lt) + __popc(b4 & lt);
                    tot += __popc(b2) + __popc(b3) + __popc(b4);
                }
                if (m) {
                    const unsigned col = c4 << 2;
                    if (m & 1u) { if (pos < cap) dst[pos] = make_uint2(col, __float_as_uint(v.x)); pos++; }
                    if (m & 2u) { if (pos < cap) dst[pos] = make_uint2(col + 1, __float_as_uint(v.y)); pos++; }
                    if (m & 4u) { if (pos < cap) dst[pos] = make_uint2(col + 2, __float_as_uint(v.z)); pos++; }
                    if (m & 8u) { if (pos < cap) dst[pos] = make_uint2(col + 3, __float_as_uint(v.w)); pos++; }
                }
                base += tot;
            }
        }
    } else {
        for (unsigned c0 = 0; c0 < V; c0 += 64) {
            const unsigned ca = c0 + lane, cb = c0 + 32 + lane;
            float xa = 0.f, xb = 0.f;
            if (ca < V) xa = ldg_stream1(row + ca);
            if (cb < V) xb = ldg_stream1(row + cb);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const float x = half ? xb : xa;
                const unsigned c = half ? cb : ca;
                const bool nz = nonzero(x);
                const unsigned any = __ballot_sync(0xffffffffu, nz);
                if (any == 0u) continue;
                if (nz) {
                    const unsigned pos = base + __popc(any & lt);
                    if (pos < cap) dst[pos] = make_uint2(c, __float_as_uint(x));
                }
                base += __popc(any);
            }
        }
    }
    return base;
}

template <bool VEC4>
__device__ __forceinline__ unsigned scan_story(const CompactParams &p, unsigned story, unsigned S, unsigned long long soff,
                                               uint2 *dst, unsigned cap, unsigned short *rend, unsigned lane)
{
    unsigned cnt = scan_row<VEC4>(p.q + (size_t)story * p.V, p.V, dst, cap, 0u, lane);
    if (rend && lane == 0) rend[0] = (unsigned short)min(cnt, 0xFFFFu);
    const float *mrow = p.m + (size_t)soff * p.V;
    for (unsigned r = 0; r < S; r++) {
        cnt = scan_row<VEC4>(mrow + (size_t)r * p.V, p.V, dst, cap, cnt, lane);
        if (rend && lane == 0) rend[r + 1] = (unsigned short)min(cnt, 0xFFFFu);
    }
    return cnt;
}

template <bool VEC4>
__global__ void __launch_bounds__(256) k_compact(const CompactParams p)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < p.n_stories; w += warps) {
        const unsigned story = p.story0 + w;
        const unsigned long long soff = p.sen_off[story];
        const unsigned S = (unsigned)(p.sen_off[story + 1] - soff);
        unsigned char *rec = p.rec + (size_t)w * p.rec_stride;
        unsigned *hdr = reinterpret_cast<unsigned *>(rec);
        unsigned short *rend = reinterpret_cast<unsigned short *>(rec + p.off_rend);
        uint2 *ent = reinterpret_cast<uint2 *>(rec + p.off_ent);

        unsigned n = scan_story<VEC4>(p, story, S, soff, ent, p.lcap, rend, lane);
        unsigned flags = 0, heap_off = 0;
        if (n > p.lcap || n > 0xFFFFu) {
            // rare: a story denser than the fixed slot.  Reserve exactly n entries in the heap and rescan.
            unsigned long long off = 0;
            if (lane == 0) off = atomicAdd(p.heap_used, (unsigned long long)n);
            off = __shfl_sync(0xffffffffu, off, 0);
            if (n > 0xFFFFu || off + n > p.heap_cap || off + n > 0xFFFFFFFFull) flags = FLAG_ERROR;
            else {
                flags = FLAG_HEAP;
                heap_off = (unsigned)off;
                scan_story<VEC4>(p, story, S, soff, p.heap + off, n, nullptr, lane);
            }
        }
        // answer: index of the (last) 1.0 in the one-hot row; the reference tests y == 1.0 per
        // class (lib/layer_cuda.cu:2196)
        unsigned ans = ANS_NONE;
        if (p.a) {
            const float *arow = p.a + (size_t)story * p.V;
            for (unsigned c0 = 0; c0 < p.V; c0 += 32) {
                const unsigned c = c0 + lane;
                const bool hot = (c < p.V) && (ldg_stream1(arow + c) == 1.0f);
                const unsigned b = __ballot_sync(0xffffffffu, hot);
                if (b) ans = c0 + 31 - __clz(b);
            }
        }
        if (lane == 0) { hdr[0] = n; hdr[1] = flags; hdr[2] = ans; hdr[3] = heap_off; }
    }
}

// =============================================================================================
// k_forward
// =============================================================================================
__device__ __forceinline__ int sbyte(unsigned w, int b) { return (int)(signed char)((w >> (8 * b)) & 0xFFu); }

struct WarpCtx {
    unsigned lane;
    unsigned char *ws;             // this warp's scratch
    const unsigned char *tab;      // tables in shared memory
    // entries of the current story
    const unsigned *ent_s;         // shared copy (col | (exc+1)<<16), or
    const uint2 *ent_g;            // global entries when the story does not fit
    bool ent_in_smem;
    const unsigned short *rend;
    const float *exc;
};

// Gather-and-sum embedding of up to 32/LPR rows at once (one row per LPR-lane group; lane q of a
// group owns dims 16q..16q+15).  acc[j] = sum over the row's entries of Q_w(Q_w(x) * Q_w(T[c][id]))
// -- the per-product quantise + clamp of the reference (lib/layer_cuda.cu:120) is kept, which is why
// this is not a dp4a dot product over ids.  For x == 1.0 the term is the table code itself
// (Q_w(1.0) = 2^frac_w when iwl_w >= 1) and dp4a with a one-hot selector does the sign-extending
// byte accumulate.  `row` is the record row index (0 = question) or -1 for an idle group.
template <int LPR>
__device__ __forceinline__ void embed_rows(const WarpCtx &c, const unsigned char *table, unsigned DP, int row, int iwl_w, int frac_w, int lim_w, int acc[16])
{
    const unsigned q = c.lane % LPR;
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0;
    unsigned beg = 0, len = 0;
    if (row >= 0) {
        beg = row ? c.rend[row - 1] : 0u;
        len = c.rend[row] - beg;
    }
    const unsigned maxlen = __reduce_max_sync(0xffffffffu, len);
    const bool unit_ok = (iwl_w >= 1);
    for (unsigned k = 0; k < maxlen; k++) {
        if (k < len) {
            unsigned col;
            float x = 1.0f;
            bool unit;
            if (c.ent_in_smem) {
                const unsigned e = c.ent_s[beg + k];
                col = e & 0xFFFFu;
                const unsigned xi = e >> 16;
                unit = (xi == 0u);
                if (!unit) x = c.exc[xi - 1];
            } else {
                const uint2 e = c.ent_g[beg + k];
                col = e.x;
                x = __uint_as_float(e.y);
                unit = (x == 1.0f);
            }
            const uint4 t = *reinterpret_cast<const uint4 *>(table + (size_t)col * DP + 16u * q);
            const unsigned tw[4] = {t.x, t.y, t.z, t.w};
            if (unit && unit_ok) {
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    acc[4 * w + 0] = __dp4a((int)tw[w], 0x00000001, acc[4 * w + 0]);
                    acc[4 * w + 1] = __dp4a((int)tw[w], 0x00000100, acc[4 * w + 1]);
                    acc[4 * w + 2] = __dp4a((int)tw[w], 0x00010000, acc[4 * w + 2]);
                    acc[4 * w + 3] = __dp4a((int)tw[w], 0x01000000, acc[4 * w + 3]);
                }
            } else {
                const int xq = qi_encode(x, iwl_w, frac_w);
#pragma unroll
                for (int w = 0; w < 4; w++)
#pragma unroll
                    for (int b = 0; b < 4; b++) acc[4 * w + b] += qi_mul(xq, sbyte(tw[w], b), lim_w, frac_w);
            }
        }
    }
}

template <int LPR>
__device__ __forceinline__ int group_sum(int v)
{
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int LPR, int MODE, bool DEBUG>
__global__ void __launch_bounds__(512, 1) k_forward(const __grid_constant__ FwdParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int G = 32 / LPR;                 // rows embedded concurrently by one warp
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned g = lane / LPR, q = lane % LPR;

    // ---- stage the quantised tables into shared memory (once per CTA) ----
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.img);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (unsigned i = threadIdx.x; i < p.img_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    WarpCtx c;
    c.lane = lane;
    c.tab = smem;
    c.ws = smem + p.tables_bytes + (size_t)wid * p.warp_bytes;
    unsigned *ent_s = reinterpret_cast<unsigned *>(c.ws);
    unsigned short *rend_s = reinterpret_cast<unsigned short *>(c.ws + p.o_rend);
    int *sc = reinterpret_cast<int *>(c.ws + p.o_sc);
    float *ex = reinterpret_cast<float *>(c.ws + p.o_ex);
    unsigned char *pq = c.ws + p.o_pq;
    signed char *uvec = reinterpret_cast<signed char *>(c.ws + p.o_uvec);
    signed char *ubvec = reinterpret_cast<signed char *>(c.ws + p.o_ubvec);
    signed char *ovec = reinterpret_cast<signed char *>(c.ws + p.o_ovec);
    float *ufl = reinterpret_cast<float *>(c.ws + p.o_ufl);
    float *exc = reinterpret_cast<float *>(c.ws + p.o_exc);
    float *zbuf = reinterpret_cast<float *>(c.ws);      // aliases the entry list (dead by the answer phase)
    c.ent_s = ent_s;
    c.rend = rend_s;
    c.exc = exc;
    const unsigned d = p.d, DP = p.DP, V = p.V;

    for (;;) {
        unsigned w = 0;
        if (lane == 0) w = atomicAdd(p.counter, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= p.n_stories) break;
        const unsigned story = p.story0 + w;
        const unsigned long long soff = p.sen_off[story];
        const unsigned S = (unsigned)(p.sen_off[story + 1] - soff);

        // ---- load this story's compact record ----
        const unsigned char *rec = p.rec + (size_t)w * p.rec_stride;
        const unsigned *hdr = reinterpret_cast<const unsigned *>(rec);
        const unsigned n_ent = hdr[0], flags = hdr[1], ans_idx = hdr[2], heap_off = hdr[3];
        if (flags & FLAG_ERROR) {
            if (lane == 0) { atomicExch(p.err_flag, 1u); if (p.pred) p.pred[story] = ANS_NONE; }
            continue;
        }
        const uint2 *ent_g = (flags & FLAG_HEAP) ? (p.heap + heap_off) : reinterpret_cast<const uint2 *>(rec + p.off_ent);
        {
            const unsigned short *rend_g = reinterpret_cast<const unsigned short *>(rec + p.off_rend);
            for (unsigned r = lane; r < S + 1; r += 32) rend_s[r] = rend_g[r];
        }
        bool in_smem = (n_ent <= p.LW);
        if (in_smem) {
            // compress {col, fp32} to col | (exception index + 1) << 16; 1.0 is the common value
            unsigned n_exc = 0;
            for (unsigned k0 = 0; k0 < n_ent; k0 += 32) {
                const unsigned k = k0 + lane;
                uint2 e = make_uint2(0u, 0x3F800000u);
                if (k < n_ent) e = ent_g[k];
                const bool special = (e.y != 0x3F800000u) || (e.x > 0xFFFFu);
                const unsigned b = __ballot_sync(0xffffffffu, special);
                unsigned code = e.x;
                if (special) {
                    const unsigned xi = n_exc + __popc(b & ((1u << lane) - 1u));
                    if (xi < MAX_EXC) exc[xi] = __uint_as_float(e.y);
                    code = (e.x & 0xFFFFu) | ((xi + 1u) << 16);
                }
                if (k < n_ent) ent_s[k] = code;
                n_exc += __popc(b);
            }
            if (n_exc > MAX_EXC) in_smem = false;      // warp-uniform
        }
        c.ent_in_smem = in_smem;
        c.ent_g = ent_g;
        __syncwarp();

        int acc[16];
        // ---- question embedding: u0 = Q_w0( sum_j Q_w0(Q_w0(B[i][j]) * Q_w0(q[j])) )   MemN2N.c:826, layer_cuda.cu:49 ----
        embed_rows<LPR>(c, c.tab + p.offB, DP, (g == 0) ? 0 : -1, p.iw[0], p.fw[0], p.lw[0], acc);
        if (g == 0) {
            unsigned packed[4];
#pragma unroll
            for (int w4 = 0; w4 < 4; w4++) {
                unsigned v = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) v |= ((unsigned)(qi_clamp(acc[4 * w4 + b], p.lw[0]) & 0xFF)) << (8 * b);
                packed[w4] = v;
            }
            *reinterpret_cast<uint4 *>(uvec + 16 * q) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        __syncwarp();
        int fu = p.fw[0];                         // fractional bits of the codes currently in uvec
        if (DEBUG && p.dbg.dev_u0)
            for (unsigned j = lane; j < d; j += 32) p.dbg.dev_u0[(size_t)story * d + j] = (float)uvec[j] / (float)(1 << fu);

        for (unsigned h = 0; h < p.H; h++) {
            const int fw = p.fw[h], lw = p.lw[h], iw = p.iw[h];
            const int fa = p.fa[h], la = p.la[h];
            const int ff = p.ff[h], lf = p.lf[h];
            const int fb = p.fb, lb = p.lb;

            // u operand: Q_bin(u) for the scorer (mode 2) and the linear map          MemN2N.c:847,873
            for (unsigned j = lane; j < DP; j += 32) ubvec[j] = (j < d) ? (signed char)qi_requant((int)uvec[j], fu, lb, fb) : (signed char)0;
            __syncwarp();
            int ub[16];
            unsigned au[16];
            unsigned su_bits = 0;
            {
                const uint4 t = *reinterpret_cast<const uint4 *>((MODE == 3 ? uvec : ubvec) + 16 * q);
                const unsigned tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int v = sbyte(tw[j >> 2], j & 3);
                    if (MODE == 3) {
                        // Hamming scorer quantises BOTH operands with the addressing format at
                        // 31-iwl fractional bits (layer.c:215-233, layer_cuda.cu:2515)
                        unsigned s_, m_;
                        appx_encode(v, fu, p.ia[h], s_, m_);
                        au[j] = m_;
                        su_bits |= (s_ >> 31) << j;
                        ub[j] = 0;
                    } else {
                        ub[j] = v;
                        au[j] = 0;
                    }
                }
            }

            // ---- memory embedding + addressing, G rows per pass ----
            for (unsigned r0 = 0; r0 < S; r0 += G) {
                const unsigned r = r0 + g;
                embed_rows<LPR>(c, c.tab + p.offA[h], DP, (r < S) ? (int)(r + 1) : -1, iw, fw, lw, acc);
                int part = 0;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int m_w = qi_clamp(acc[j], lw);                       // M_h[r][t], weight format
                    if (MODE == 3) {
                        unsigned sm, am;
                        appx_encode(m_w, fw, p.ia[h], sm, am);
                        const unsigned sv = ((su_bits >> j) & 1u) << 31;
                        // dims >= d hold zero codes on both sides: all 7 bits match, e = +127; masked below
                        part += (16u * q + j < d) ? appx_element_x128(sm, am, sv, au[j]) : 0;
                    } else {
                        // s[r] = Q_att( sum_t Q_att( Q_att(M[r][t]) * Q_bin(u[t]) ) )      layer_cuda.cu:105-141
                        const int m_att = qi_requant(m_w, fw, la, fa);
                        part += qi_mul(m_att, ub[j], la, fb);
                    }
                }
                if (DEBUG && p.dbg.dev_M && r < S) {
                    float *dstM = p.dbg.dev_M + ((size_t)h * p.sum_sen + soff + r) * d;
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (16u * q + j < d) dstM[16 * q + j] = (float)qi_clamp(acc[j], lw) / (float)(1 << fw);
                }
                const int tot = group_sum<LPR>(part);
                if (q == 0 && r < S) sc[r] = (MODE == 3) ? tot : qi_clamp(tot, la);
            }
            __syncwarp();

            // ---- attention normalisation: fp32 __expf softmax, double total in ascending slot order
            //      (layer_cuda.cu:1895-1916, 1969-2060) ----
            float mx = -INFINITY;
            for (unsigned r = lane; r < S; r += 32) {
                float sv;
                if (MODE == 3) {
                    // Q_(iwl,31-iwl) of the sum of e*2^const_scale: exact unless |sum| >= 2^iwl;
                    // +-2^iwl saturate, exactly -2^iwl encodes to magnitude 0
                    const int sh = 7 - p.const_scale;
                    const float v = (float)sc[r] / (float)(1 << sh);
                    const float lim = (float)(1 << p.ia[h]);
                    sv = (v >= lim) ? lim : (v < -lim ? -lim : (v == -lim ? 0.0f : v));
                } else {
                    sv = (float)sc[r] / (float)(1 << fa);
                }
                ex[r] = sv;
                mx = fmaxf(mx, sv);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (DEBUG && p.dbg.dev_s)
                for (unsigned r = lane; r < S; r += 32) p.dbg.dev_s[(size_t)h * p.sum_sen + soff + r] = ex[r];
            for (unsigned r = lane; r < S; r += 32) ex[r] = __expf(ex[r] - mx);
            __syncwarp();
            double total = 0.0;
            for (unsigned r = 0; r < S; r++) total += (double)ex[r];
            unsigned nnz = 0;
            for (unsigned r0 = 0; r0 < S; r0 += 32) {
                const unsigned r = r0 + lane;
                unsigned code = 0;
                if (r < S) {
                    const float pr = (float)((double)ex[r] / total);
                    // Q_f(p) inside the weighted read                               layer_cuda.cu:561
                    code = (unsigned)qi_encode(pr, p.iff[h], ff);
                    if (DEBUG && p.dbg.dev_p) p.dbg.dev_p[(size_t)h * p.sum_sen + soff + r] = pr;
                }
                // compact the slots whose quantised weight is non-zero: the rest contribute
                // Q(0 * c) = 0 to every output dimension
                const unsigned b = __ballot_sync(0xffffffffu, code != 0u);
                if (code) {
                    const unsigned k = nnz + __popc(b & ((1u << lane) - 1u));
                    sc[k] = (int)r;
                    pq[k] = (unsigned char)code;
                }
                nnz += __popc(b);
            }
            __syncwarp();

            if (DEBUG && p.dbg.dev_C) {
                for (unsigned r0 = 0; r0 < S; r0 += G) {
                    const unsigned r = r0 + g;
                    embed_rows<LPR>(c, c.tab + p.offC[h], DP, (r < S) ? (int)(r + 1) : -1, iw, fw, lw, acc);
                    if (r < S) {
                        float *dstC = p.dbg.dev_C + ((size_t)h * p.sum_sen + soff + r) * d;
#pragma unroll
                        for (int j = 0; j < 16; j++)
                            if (16u * q + j < d) dstC[16 * q + j] = (float)qi_clamp(acc[j], lw) / (float)(1 << fw);
                    }
                }
            }

            // ---- weighted read: o[c] = Q_f( sum_r Q_f( Q_f(p[r]) * Q_f(C_h[r][c]) ) )   layer_cuda.cu:547-579 ----
            int oacc[16];
#pragma unroll
            for (int j = 0; j < 16; j++) oacc[j] = 0;
            for (unsigned k0 = 0; k0 < nnz; k0 += G) {
                const unsigned k = k0 + g;
                const int r = (k < nnz) ? sc[k] : -1;
                const int pc = (k < nnz) ? (int)pq[k] : 0;
                embed_rows<LPR>(c, c.tab + p.offC[h], DP, (r >= 0) ? r + 1 : -1, iw, fw, lw, acc);
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int c_f = qi_requant(qi_clamp(acc[j], lw), fw, lf, ff);
                    oacc[j] += qi_mul(pc, c_f, lf, ff);
                }
            }
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
                for (int j = 0; j < 16; j++) oacc[j] += __shfl_xor_sync(0xffffffffu, oacc[j], o);
            if (g == 0) {
                unsigned packed[4];
#pragma unroll
                for (int w4 = 0; w4 < 4; w4++) {
                    unsigned v = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) v |= ((unsigned)(qi_clamp(oacc[4 * w4 + b], lf) & 0xFF)) << (8 * b);
                    packed[w4] = v;
                }
                *reinterpret_cast<uint4 *>(ovec + 16 * q) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
            __syncwarp();

            // ---- linear map g = Q_w( sum_j Q_w( Q_w(Hm[i][j]) * Q_bin(u[j]) ) )  (MemN2N.c:873, layer_cuda.cu:49-68)
            //      and update u' = Q_f( Q_f(g) + Q_f(o) )                            (MemN2N.c:889, layer_cuda.cu:1535) ----
            for (unsigned i0 = 0; i0 < d; i0 += 32) {
                const unsigned i = i0 + lane;
                int a_f = 0, g_w = 0;
                if (i < d) {
                    if (p.lin_map) {
                        const unsigned *hrow = reinterpret_cast<const unsigned *>(c.tab + p.offH[h] + (size_t)i * p.HS);
                        const unsigned *ubw = reinterpret_cast<const unsigned *>(ubvec);
                        int s_ = 0;
                        for (unsigned j4 = 0; j4 < (d + 3) / 4; j4++) {
                            const unsigned hw = hrow[j4], uw = ubw[j4];
#pragma unroll
                            for (int b = 0; b < 4; b++) s_ += qi_mul(sbyte(hw, b), sbyte(uw, b), lw, fb);
                        }
                        g_w = qi_clamp(s_, lw);
                        a_f = qi_requant(g_w, fw, lf, ff);
                    } else {
                        g_w = (int)uvec[i];
                        a_f = qi_requant(g_w, fu, lf, ff);
                    }
                }
                __syncwarp();
                if (i < d) {
                    const int un = qi_clamp(a_f + (int)ovec[i], lf);
                    if (DEBUG) {
                        const size_t vo = ((size_t)h * p.n_total + story) * d + i;
                        if (p.dbg.dev_o) p.dbg.dev_o[vo] = (float)ovec[i] / (float)(1 << ff);
                        if (p.dbg.dev_g) p.dbg.dev_g[vo] = (float)g_w / (float)(1 << (p.lin_map ? fw : fu));
                        if (p.dbg.dev_u) p.dbg.dev_u[vo] = (float)un / (float)(1 << ff);
                    }
                    uvec[i] = (signed char)un;
                }
            }
            fu = ff;
            __syncwarp();
        }

        // ---- answer projection z[i] = sum_j fl(W[i][j]*u[j]), sequential fp32 (MemN2N.c:902-906,
        //      layer_cuda.cu:69-82), softmax and argmax on the probabilities (layer_cuda.cu:1918-1939) ----
        for (unsigned j = lane; j < DP; j += 32) ufl[j] = (j < d) ? (float)uvec[j] / (float)(1 << fu) : 0.0f;
        __syncwarp();
        const float *Wt = reinterpret_cast<const float *>(c.tab + p.offW);
        const unsigned d4 = (d + 3) / 4;
        float zmax = -INFINITY;
        for (unsigned i0 = 0; i0 < V; i0 += 128) {
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            for (unsigned j4 = 0; j4 < d4; j4++) {
                const float4 uu = *reinterpret_cast<const float4 *>(ufl + 4 * j4);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned i = i0 + 32 * k + lane;
                    if (i < V) {
                        const float4 ww = *reinterpret_cast<const float4 *>(Wt + (size_t)i * p.WS + 4 * j4);
                        z[k] = __fadd_rn(z[k], __fmul_rn(ww.x, uu.x));
                        z[k] = __fadd_rn(z[k], __fmul_rn(ww.y, uu.y));
                        z[k] = __fadd_rn(z[k], __fmul_rn(ww.z, uu.z));
                        z[k] = __fadd_rn(z[k], __fmul_rn(ww.w, uu.w));
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned i = i0 + 32 * k + lane;
                if (i < V) {
                    zbuf[i] = z[k];
                    zmax = fmaxf(zmax, z[k]);
                    if (DEBUG && p.dbg.dev_z) p.dbg.dev_z[(size_t)story * V + i] = z[k];
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
        __syncwarp();
        // e_i = __expf(z_i - max).  h_i = fl(e_i / total) is monotone in e_i, so only slots whose e is
        // within 2^-20 of the largest can share the maximal probability; the total is needed only
        // to break such near-ties exactly, or when probabilities are requested.
        unsigned n_cand = 0, cand_idx = 0;
        for (unsigned i0 = 0; i0 < V; i0 += 32) {
            const unsigned i = i0 + lane;
            bool cand = false;
            if (i < V) {
                const float e = __expf(zbuf[i] - zmax);
                zbuf[i] = e;
                cand = (e >= 0.99999905f);
            }
            const unsigned b = __ballot_sync(0xffffffffu, cand);
            if (b) { n_cand += __popc(b); cand_idx = i0 + 31 - __clz(b); }
        }
        __syncwarp();
        unsigned pred_i = cand_idx;
        const bool need_total = (n_cand > 1) || p.want_h || (DEBUG && p.dbg.dev_h);
        float h_true_v = 0.0f;
        if (need_total) {
            double total = 0.0;
            for (unsigned i = 0; i < V; i++) total += (double)zbuf[i];
            float best = -INFINITY;
            unsigned best_i = 0;
            for (unsigned i0 = 0; i0 < V; i0 += 32) {
                const unsigned i = i0 + lane;
                if (i < V) {
                    const float hv = (float)((double)zbuf[i] / total);
                    if (DEBUG && p.dbg.dev_h) p.dbg.dev_h[(size_t)story * V + i] = hv;
                    if (!(best > hv)) { best = hv; best_i = i; }
                    if (i == ans_idx) h_true_v = hv;
                }
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const unsigned oi = __shfl_xor_sync(0xffffffffu, best_i, o);
                if (ov > best || (ov == best && oi > best_i)) { best = ov; best_i = oi; }
            }
            pred_i = best_i;
            // exactly one lane holds h[y] (probabilities are >= 0, so the uint order is the float order)
            h_true_v = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(h_true_v)));
        }
        if (lane == 0) {
            if (p.pred) p.pred[story] = pred_i;
            if (p.h_true) p.h_true[story] = h_true_v;
            if (p.match && ans_idx != ANS_NONE && pred_i == ans_idx) atomicAdd(p.match, 1u);
        }
        __syncwarp();
    }
}

}  // namespace

// =============================================================================================
// host objects
// =============================================================================================
struct qmann_model {
    qmann_config cfg;
    int device;
    int sm_count;
    unsigned char *dev_img;
    FwdParams base;                // everything but the per-call fields
    unsigned LPR, NW, smem_bytes;
    unsigned rec_stride, off_rend, off_ent, lcap;
    // chunk scratch
    unsigned chunk_cap;
    unsigned char *dev_rec;
    uint2 *dev_heap;
    unsigned long long heap_cap;
    unsigned long long *dev_heap_used;
    unsigned *dev_counter, *dev_err;
    // qmann_infer_host staging (grow-only device arenas, two streams)
    float *e2e_m = nullptr, *e2e_q = nullptr, *e2e_a = nullptr, *e2e_h = nullptr;
    uint32_t *e2e_pred = nullptr, *e2e_match = nullptr;
    size_t e2e_m_cap = 0, e2e_n_cap = 0;
    cudaStream_t e2e_compute = nullptr, e2e_copy = nullptr;
    std::vector<cudaEvent_t> e2e_events;
    // optional per-kernel timing (qmann_profile_*)
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;     // triples: before compact, between, after forward
    size_t prof_used = 0;
};

struct qmann_batch {
    unsigned N;
    unsigned long long sum_sen;
    unsigned long long *dev_sen_off;
    std::vector<unsigned long long> sen_off;
    unsigned max_sen;
};

namespace {
unsigned round_up(unsigned v, unsigned m) { return (v + m - 1) / m * m; }

template <int LPR, int MODE>
int launch_forward_t(const qmann_model *m, const FwdParams &p, bool debug, cudaStream_t st)
{
    const unsigned grid = (unsigned)m->sm_count, block = m->NW * 32;
    if (debug) {
        QCUDA(cudaFuncSetAttribute(k_forward<LPR, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_bytes));
        k_forward<LPR, MODE, true><<<grid, block, m->smem_bytes, st>>>(p);
    } else {
        QCUDA(cudaFuncSetAttribute(k_forward<LPR, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_bytes));
        k_forward<LPR, MODE, false><<<grid, block, m->smem_bytes, st>>>(p);
    }
    count_launch();
    QCUDA(cudaPeekAtLastError());
    return QMANN_OK;
}
template <int LPR>
int launch_forward_l(const qmann_model *m, const FwdParams &p, bool debug, cudaStream_t st)
{
    return m->cfg.mode == 3 ? launch_forward_t<LPR, 3>(m, p, debug, st) : launch_forward_t<LPR, 2>(m, p, debug, st);
}
int launch_forward(const qmann_model *m, const FwdParams &p, bool debug, cudaStream_t st)
{
    switch (m->LPR) {
        case 4: return launch_forward_l<4>(m, p, debug, st);
        case 8: return launch_forward_l<8>(m, p, debug, st);
        case 16: return launch_forward_l<16>(m, p, debug, st);
        default: return launch_forward_l<32>(m, p, debug, st);
    }
}
}  // namespace

extern "C" {

const char *qmann_last_error(void) { return g_err.c_str(); }
const char *qmann_version(void) { return "qmann_b200 0.1 (sm_100a)"; }
uint64_t qmann_launch_count(void) { return g_launches.load(); }

int qmann_model_create(qmann_model **out, const qmann_config *cfg, const qmann_weights *w)
{
    if (!out || !cfg || !w) return fail(QMANN_E_ARG, "null argument");
    *out = nullptr;
    const qmann_config &c = *cfg;
    if (c.H == 0 || c.H > MAXH) return fail(QMANN_E_ARG, "H must be in 1..8");
    if (c.mode != 2 && c.mode != 3) return fail(QMANN_E_ARG, "attention mode must be 2 (fixed-point dot) or 3 (Hamming/approximate)");
    if (c.V == 0 || c.V > 65535) return fail(QMANN_E_ARG, "V must be in 1..65535");
    if (c.d == 0 || c.d > 512) return fail(QMANN_E_ARG, "d must be in 1..512");
    if (c.S_max == 0 || c.S_max > 4096) return fail(QMANN_E_ARG, "S_max must be in 1..4096");
    auto okfmt = [](unsigned i, unsigned f) { return i + f >= 1 && i + f <= 7; };
    for (unsigned h = 0; h < c.H; h++)
        if (!okfmt(c.iwl[h], c.frac[h]) || !okfmt(c.iwl_w[h], c.frac_w[h]) || !okfmt(c.iwl_att[h], c.frac_att[h]))
            return fail(QMANN_E_ARG, "formats must satisfy 1 <= iwl+frac <= 7 (8-bit word length, BW_WL)");
    if (!okfmt(c.iwl_bin, c.frac_bin)) return fail(QMANN_E_ARG, "bin format must satisfy 1 <= iwl+frac <= 7");
    if (c.mode == 3) {
        if (c.const_scale > 0 || c.const_scale < -16) return fail(QMANN_E_ARG, "const_scale must be in -16..0");
        for (unsigned h = 0; h < c.H; h++)
            if (c.iwl_att[h] < 1) return fail(QMANN_E_ARG, "mode 3 needs iwl_att >= 1 (the reference's 1<<31 encode is undefined at iwl 0)");
    }
    if (!w->dev_B || !w->dev_W) return fail(QMANN_E_ARG, "missing weight pointer");
    for (unsigned h = 0; h < c.H; h++)
        if (!w->dev_A[h] || !w->dev_C[h] || (c.lin_map && !w->dev_Hm[h])) return fail(QMANN_E_ARG, "missing per-hop weight pointer");

    qmann_model *m = new qmann_model();
    m->cfg = c;
    QCUDA(cudaGetDevice(&m->device));
    cudaDeviceProp prop;
    QCUDA(cudaGetDeviceProperties(&prop, m->device));
    m->sm_count = prop.multiProcessorCount;

    // ---- image layout ----
    unsigned LPR = 4;
    while (16 * LPR < c.d) LPR *= 2;
    m->LPR = LPR;
    const unsigned DP = 16 * LPR;
    unsigned HS = round_up(c.d, 4);
    if (((HS / 4) & 1u) == 0) HS += 4;             // odd word stride: lane-per-row reads are conflict-free
    unsigned WS = round_up(c.d, 4);
    if (((WS / 4) & 1u) == 0) WS += 4;             // (WS/4) odd: conflict-free 128-bit lane-per-row reads
    FwdParams &p = m->base;
    memset(&p, 0, sizeof(p));
    unsigned off = 0;
    auto take = [&](unsigned bytes) { unsigned o = off; off += round_up(bytes, 16); return o; };
    p.offB = take(c.V * DP);
    for (unsigned h = 0; h < c.H; h++) { p.offA[h] = take(c.V * DP); p.offC[h] = take(c.V * DP); }
    for (unsigned h = 0; h < c.H; h++) p.offH[h] = c.lin_map ? take(c.d * HS) : 0;
    p.offW = take(c.V * WS * 4);
    p.img_bytes = off;
    p.tables_bytes = off;
    p.V = c.V; p.d = c.d; p.S_max = c.S_max; p.H = c.H; p.lin_map = c.lin_map; p.const_scale = c.const_scale;
    p.DP = DP; p.HS = HS; p.WS = WS;
    for (unsigned h = 0; h < c.H; h++) {
        p.fw[h] = c.frac_w[h]; p.iw[h] = c.iwl_w[h]; p.lw[h] = fixed_max(c.iwl_w[h], c.frac_w[h]);
        p.fa[h] = c.frac_att[h]; p.ia[h] = c.iwl_att[h]; p.la[h] = fixed_max(c.iwl_att[h], c.frac_att[h]);
        p.ff[h] = c.frac[h]; p.iff[h] = c.iwl[h]; p.lf[h] = fixed_max(c.iwl[h], c.frac[h]);
    }
    p.fb = c.frac_bin; p.lb = fixed_max(c.iwl_bin, c.frac_bin);

    // ---- per-warp scratch layout ----
    const unsigned S_pad = round_up(c.S_max, 32);
    unsigned o = 0;
    auto wtake = [&](unsigned bytes) { unsigned r = o; o += round_up(bytes, 16); return r; };
    const unsigned ent_fixed = round_up(c.V, 32) * 4;            // the region is reused as zbuf[V]
    p.o_rend = 0;  // placeholder, set below
    int max_smem = 0;
    QCUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
    unsigned NW = 0, LW = 0;
    unsigned fixed_warp = 0;
    {
        // everything except the entry list
        o = 0;
        wtake(0);
        unsigned o_rend = wtake((c.S_max + 2) * 2), o_sc = wtake(S_pad * 4), o_ex = wtake(S_pad * 4), o_pq = wtake(S_pad);
        unsigned o_uvec = wtake(DP), o_ubvec = wtake(DP), o_ovec = wtake(DP), o_ufl = wtake(DP * 4), o_exc = wtake(MAX_EXC * 4);
        fixed_warp = o;
        p.o_rend = o_rend; p.o_sc = o_sc; p.o_ex = o_ex; p.o_pq = o_pq; p.o_uvec = o_uvec; p.o_ubvec = o_ubvec;
        p.o_ovec = o_ovec; p.o_ufl = o_ufl; p.o_exc = o_exc;
    }
    const unsigned want_LW = std::min(65535u, round_up(8 * (c.S_max + 1), 32));
    for (unsigned nw : {16u, 12u, 8u, 6u, 4u, 2u, 1u}) {
        const long long avail = (long long)max_smem - (long long)p.tables_bytes - 1024;
        if (avail <= 0) break;
        const long long per_warp = avail / nw;
        const long long ent_bytes = per_warp - (long long)fixed_warp;
        if (ent_bytes < (long long)ent_fixed || ent_bytes < 32 * 4) continue;
        unsigned lw_ = (unsigned)std::min<long long>(ent_bytes / 4, want_LW);
        lw_ = lw_ / 4 * 4;
        NW = nw; LW = lw_;
        break;
    }
    if (NW == 0) {
        delete m;
        return fail(QMANN_E_NOMEM, "model tables (" + std::to_string(p.tables_bytes) + " B) plus per-warp scratch do not fit the SM's " +
                                       std::to_string(max_smem) + " B of shared memory");
    }
    const unsigned ent_region = round_up(std::max(LW * 4, ent_fixed), 16);
    // entry list sits first in the warp scratch; shift the other offsets behind it
    p.o_rend += ent_region; p.o_sc += ent_region; p.o_ex += ent_region; p.o_pq += ent_region; p.o_uvec += ent_region;
    p.o_ubvec += ent_region; p.o_ovec += ent_region; p.o_ufl += ent_region; p.o_exc += ent_region;
    p.warp_bytes = ent_region + fixed_warp;
    p.LW = LW; p.S_pad = S_pad;
    m->NW = NW;
    m->smem_bytes = p.tables_bytes + NW * p.warp_bytes;

    // ---- quantise the weights ----
    QCUDA(cudaMalloc((void **)&m->dev_img, p.img_bytes));
    QCUDA(cudaMemset(m->dev_img, 0, p.img_bytes));
    auto prep_emb = [&](const float *src, unsigned o_, int iwl, int frac) {
        k_prep_emb<<<64, 256>>>(src, reinterpret_cast<signed char *>(m->dev_img + o_), c.V, c.d, DP, iwl, frac);
        count_launch();
    };
    prep_emb(w->dev_B, p.offB, c.iwl_w[0], c.frac_w[0]);
    for (unsigned h = 0; h < c.H; h++) {
        prep_emb(w->dev_A[h], p.offA[h], c.iwl_w[h], c.frac_w[h]);
        prep_emb(w->dev_C[h], p.offC[h], c.iwl_w[h], c.frac_w[h]);
        if (c.lin_map) {
            k_prep_lin<<<16, 256>>>(w->dev_Hm[h], reinterpret_cast<signed char *>(m->dev_img + p.offH[h]), c.d, HS, c.iwl_w[h], c.frac_w[h]);
            count_launch();
        }
    }
    k_prep_ans<<<64, 256>>>(w->dev_W, reinterpret_cast<float *>(m->dev_img + p.offW), c.V, c.d, WS);
    count_launch();
    QCUDA(cudaPeekAtLastError());
    QCUDA(cudaDeviceSynchronize());
    p.img = m->dev_img;

    // ---- compact-record geometry and chunk scratch ----
    m->lcap = std::min(65535u, round_up(12 * (c.S_max + 1), 32));
    m->off_rend = REC_HDR_BYTES;
    m->off_ent = round_up(REC_HDR_BYTES + (c.S_max + 2) * 2, 16);
    m->rec_stride = m->off_ent + m->lcap * 8;
    m->chunk_cap = 32768;
    QCUDA(cudaMalloc((void **)&m->dev_rec, (size_t)m->chunk_cap * m->rec_stride));
    m->heap_cap = 8ull << 20;                                        // 8 Mi entries = 64 MiB
    QCUDA(cudaMalloc((void **)&m->dev_heap, m->heap_cap * sizeof(uint2)));
    QCUDA(cudaMalloc((void **)&m->dev_heap_used, sizeof(unsigned long long)));
    QCUDA(cudaMalloc((void **)&m->dev_counter, sizeof(unsigned)));
    QCUDA(cudaMalloc((void **)&m->dev_err, sizeof(unsigned)));
    QCUDA(cudaMemset(m->dev_err, 0, sizeof(unsigned)));
    p.rec = m->dev_rec; p.rec_stride = m->rec_stride; p.off_rend = m->off_rend; p.off_ent = m->off_ent;
    p.heap = m->dev_heap; p.counter = m->dev_counter; p.err_flag = m->dev_err;
    *out = m;
    return QMANN_OK;
}

void qmann_model_destroy(qmann_model *m)
{
    if (!m) return;
    cudaFree(m->dev_img); cudaFree(m->dev_rec); cudaFree(m->dev_heap); cudaFree(m->dev_heap_used);
    cudaFree(m->dev_counter); cudaFree(m->dev_err);
    cudaFree(m->e2e_m); cudaFree(m->e2e_q); cudaFree(m->e2e_a); cudaFree(m->e2e_h); cudaFree(m->e2e_pred); cudaFree(m->e2e_match);
    if (m->e2e_compute) cudaStreamDestroy(m->e2e_compute);
    if (m->e2e_copy) cudaStreamDestroy(m->e2e_copy);
    for (auto e : m->e2e_events) cudaEventDestroy(e);
    for (auto e : m->prof_events) cudaEventDestroy(e);
    delete m;
}

int qmann_batch_create(qmann_batch **out, const uint32_t *n_sen, uint32_t N)
{
    if (!out || (!n_sen && N)) return fail(QMANN_E_ARG, "null argument");
    qmann_batch *b = new qmann_batch();
    b->N = N;
    b->sen_off.resize((size_t)N + 1);
    b->sen_off[0] = 0;
    b->max_sen = 0;
    for (uint32_t i = 0; i < N; i++) {
        b->sen_off[i + 1] = b->sen_off[i] + n_sen[i];
        b->max_sen = std::max(b->max_sen, n_sen[i]);
    }
    b->sum_sen = b->sen_off[N];
    b->dev_sen_off = nullptr;
    cudaError_t e = cudaMalloc((void **)&b->dev_sen_off, ((size_t)N + 1) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpy(b->dev_sen_off, b->sen_off.data(), ((size_t)N + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete b; return fail(QMANN_E_CUDA, cudaGetErrorString(e)); }
    *out = b;
    return QMANN_OK;
}

void qmann_batch_destroy(qmann_batch *b)
{
    if (!b) return;
    cudaFree(b->dev_sen_off);
    delete b;
}

// Forward of stories [first, first+count) of the batch; data pointers are the FULL arenas.
static int forward_range(qmann_model *m, const qmann_batch *b, uint32_t first, uint32_t count, const float *dev_m, const float *dev_q,
                         const float *dev_a, uint32_t *dev_pred, float *dev_h_true, uint32_t *dev_match, const qmann_debug *dbg, cudaStream_t st)
{
    const bool debug = dbg != nullptr;
    const bool vec4 = (m->cfg.V % 4 == 0) && (((uintptr_t)dev_m | (uintptr_t)dev_q) % 16 == 0);
    for (uint32_t s0 = first; s0 < first + count; s0 += m->chunk_cap) {
        const uint32_t n = std::min<uint32_t>(m->chunk_cap, first + count - s0);
        CompactParams cp;
        cp.m = dev_m; cp.q = dev_q; cp.a = dev_a; cp.sen_off = b->dev_sen_off; cp.V = m->cfg.V; cp.S_max = m->cfg.S_max;
        cp.story0 = s0; cp.n_stories = n; cp.rec = m->dev_rec; cp.rec_stride = m->rec_stride; cp.off_rend = m->off_rend;
        cp.off_ent = m->off_ent; cp.lcap = m->lcap; cp.heap = m->dev_heap; cp.heap_cap = m->heap_cap; cp.heap_used = m->dev_heap_used;
        cp.vec4 = vec4;
        QCUDA(cudaMemsetAsync(m->dev_heap_used, 0, sizeof(unsigned long long), st));
        QCUDA(cudaMemsetAsync(m->dev_counter, 0, sizeof(unsigned), st));
        const unsigned cblocks = std::min<unsigned>((n + 7) / 8, (unsigned)m->sm_count * 8);
        cudaEvent_t *pe = nullptr;
        if (m->profile) {
            if (m->prof_used + 3 > m->prof_events.size()) {
                for (int k = 0; k < 3; k++) { cudaEvent_t e; QCUDA(cudaEventCreate(&e)); m->prof_events.push_back(e); }
            }
            pe = &m->prof_events[m->prof_used];
            m->prof_used += 3;
            QCUDA(cudaEventRecord(pe[0], st));
        }
        if (vec4) k_compact<true><<<cblocks, 256, 0, st>>>(cp);
        else      k_compact<false><<<cblocks, 256, 0, st>>>(cp);
        count_launch();
        QCUDA(cudaPeekAtLastError());
        if (pe) QCUDA(cudaEventRecord(pe[1], st));

        FwdParams p = m->base;
        p.sen_off = b->dev_sen_off; p.story0 = s0; p.n_stories = n; p.n_total = b->N; p.sum_sen = b->sum_sen;
        p.pred = dev_pred; p.h_true = dev_h_true; p.match = dev_match; p.want_h = (dev_h_true != nullptr);
        if (dbg) p.dbg = *dbg;
        int rc = launch_forward(m, p, debug, st);
        if (rc) return rc;
        if (pe) QCUDA(cudaEventRecord(pe[2], st));
    }
    return QMANN_OK;
}

int qmann_forward_batch(qmann_model *m, const qmann_batch *b, const float *dev_m, const float *dev_q, const float *dev_a,
                        uint32_t *dev_pred, float *dev_h_true, uint32_t *dev_match, const qmann_debug *dbg, void *stream)
{
    if (!m || !b || !dev_q || (!dev_m && b->sum_sen)) return fail(QMANN_E_ARG, "null argument");
    if (b->max_sen > m->cfg.S_max) return fail(QMANN_E_ARG, "a story has more sentences than S_max");
    if (b->N == 0) return QMANN_OK;
    return forward_range(m, b, 0, b->N, dev_m, dev_q, dev_a, dev_pred, dev_h_true, dev_match, dbg, (cudaStream_t)stream);
}

int qmann_shard_plan(const uint32_t *n_sen, uint32_t N, uint32_t world, uint32_t rank, uint32_t *first, uint32_t *count)
{
    if (!first || !count || world == 0 || rank >= world || (!n_sen && N)) return fail(QMANN_E_ARG, "bad argument");
    // contiguous ranges balanced by work ~ (sentences + 1 question row) per story
    unsigned long long total = 0;
    for (uint32_t i = 0; i < N; i++) total += (unsigned long long)n_sen[i] + 1;
    auto boundary = [&](uint32_t r) -> uint32_t {
        if (r == 0) return 0;
        if (r >= world) return N;
        const unsigned long long target = total * r / world;
        unsigned long long acc = 0;
        uint32_t i = 0;
        while (i < N && acc + (n_sen[i] + 1ull) / 2 <= target) { acc += n_sen[i] + 1ull; i++; }
        return i;
    };
    const uint32_t lo = boundary(rank), hi = boundary(rank + 1);
    *first = lo;
    *count = hi > lo ? hi - lo : 0;
    return QMANN_OK;
}

int qmann_infer_host(qmann_model *m, const float *m_host, const float *q_host, const float *a_host, const uint32_t *n_sen, uint32_t N,
                     uint32_t *pred_host, uint32_t *match, float *cost)
{
    if (!m || !q_host || !n_sen || !pred_host) return fail(QMANN_E_ARG, "null argument");
    qmann_batch *b = nullptr;
    int rc = qmann_batch_create(&b, n_sen, N);
    if (rc) return rc;
    if (b->max_sen > m->cfg.S_max) { qmann_batch_destroy(b); return fail(QMANN_E_ARG, "a story has more sentences than S_max"); }
    const size_t V = m->cfg.V;
#define QC2(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { qmann_batch_destroy(b); return fail(QMANN_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
    // grow-only device arenas owned by the model: repeated calls reuse them
    if (b->sum_sen > m->e2e_m_cap) {
        cudaFree(m->e2e_m); m->e2e_m = nullptr; m->e2e_m_cap = 0;
        QC2(cudaMalloc((void **)&m->e2e_m, std::max<size_t>(1, b->sum_sen * V) * sizeof(float)));
        m->e2e_m_cap = b->sum_sen;
    }
    if (N > m->e2e_n_cap) {
        cudaFree(m->e2e_q); cudaFree(m->e2e_a); cudaFree(m->e2e_h); cudaFree(m->e2e_pred);
        m->e2e_q = m->e2e_a = m->e2e_h = nullptr; m->e2e_pred = nullptr; m->e2e_n_cap = 0;
        QC2(cudaMalloc((void **)&m->e2e_q, (size_t)N * V * sizeof(float)));
        QC2(cudaMalloc((void **)&m->e2e_a, (size_t)N * V * sizeof(float)));
        QC2(cudaMalloc((void **)&m->e2e_h, (size_t)N * sizeof(float)));
        QC2(cudaMalloc((void **)&m->e2e_pred, (size_t)N * sizeof(uint32_t)));
        m->e2e_n_cap = N;
    }
    if (!m->e2e_match) QC2(cudaMalloc((void **)&m->e2e_match, sizeof(uint32_t)));
    if (!m->e2e_compute) QC2(cudaStreamCreateWithFlags(&m->e2e_compute, cudaStreamNonBlocking));
    if (!m->e2e_copy) QC2(cudaStreamCreateWithFlags(&m->e2e_copy, cudaStreamNonBlocking));
    cudaStream_t sc = m->e2e_compute, sx = m->e2e_copy;
    float *dm = m->e2e_m, *dq = m->e2e_q, *da = a_host ? m->e2e_a : nullptr, *dh = (a_host && cost) ? m->e2e_h : nullptr;
    QC2(cudaMemsetAsync(m->e2e_match, 0, sizeof(uint32_t), sc));
    // the copy stream runs ahead chunk by chunk; the compute stream waits per chunk, so the H2D of
    // chunk k+1 overlaps the kernels of chunk k
    const uint32_t CH = 4096;
    size_t ev_i = 0;
    for (uint32_t s0 = 0; s0 < N; s0 += CH) {
        const uint32_t n = std::min<uint32_t>(CH, N - s0);
        const size_t r0 = b->sen_off[s0], r1 = b->sen_off[s0 + n];
        if (r1 > r0) QC2(cudaMemcpyAsync(dm + r0 * V, m_host + r0 * V, (r1 - r0) * V * sizeof(float), cudaMemcpyHostToDevice, sx));
        QC2(cudaMemcpyAsync(dq + (size_t)s0 * V, q_host + (size_t)s0 * V, (size_t)n * V * sizeof(float), cudaMemcpyHostToDevice, sx));
        if (da) QC2(cudaMemcpyAsync(da + (size_t)s0 * V, a_host + (size_t)s0 * V, (size_t)n * V * sizeof(float), cudaMemcpyHostToDevice, sx));
        if (ev_i >= m->e2e_events.size()) {
            cudaEvent_t ev;
            QC2(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            m->e2e_events.push_back(ev);
        }
        cudaEvent_t ev = m->e2e_events[ev_i++];
        QC2(cudaEventRecord(ev, sx));
        QC2(cudaStreamWaitEvent(sc, ev, 0));
        rc = forward_range(m, b, s0, n, dm, dq, da, m->e2e_pred, dh, da ? m->e2e_match : nullptr, nullptr, sc);
        if (rc) { qmann_batch_destroy(b); return rc; }
    }
    QC2(cudaMemcpyAsync(pred_host, m->e2e_pred, (size_t)N * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
    uint32_t mt = 0;
    QC2(cudaMemcpyAsync(&mt, m->e2e_match, sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
    std::vector<float> ht;
    if (dh) { ht.resize(N); QC2(cudaMemcpyAsync(ht.data(), dh, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, sc)); }
    unsigned err = 0;
    QC2(cudaMemcpyAsync(&err, m->dev_err, sizeof(unsigned), cudaMemcpyDeviceToHost, sc));
    QC2(cudaStreamSynchronize(sc));
#undef QC2
    qmann_batch_destroy(b);
    if (match) *match = mt;
    if (cost && dh) {
        // the reference accumulates cost += -h[y] story by story in fp32 (layer_cuda.cu:2198)
        float cacc = *cost;
        for (uint32_t i = 0; i < N; i++) cacc = (float)((double)cacc + -1.0 * (double)ht[i]);
        *cost = cacc;
    }
    if (err) return fail(QMANN_E_NOMEM, "a story overflowed the compaction heap");
    return QMANN_OK;
}

int qmann_profile_enable(qmann_model *m, int enable)
{
    if (!m) return fail(QMANN_E_ARG, "null model");
    m->profile = enable != 0;
    m->prof_used = 0;
    return QMANN_OK;
}

int qmann_profile_read(qmann_model *m, float *ms_compact, float *ms_forward, uint32_t *n_pairs)
{
    if (!m) return fail(QMANN_E_ARG, "null model");
    double tc = 0.0, tf = 0.0;
    for (size_t i = 0; i + 3 <= m->prof_used; i += 3) {
        float a = 0.f, b2 = 0.f;
        QCUDA(cudaEventSynchronize(m->prof_events[i + 2]));
        QCUDA(cudaEventElapsedTime(&a, m->prof_events[i], m->prof_events[i + 1]));
        QCUDA(cudaEventElapsedTime(&b2, m->prof_events[i + 1], m->prof_events[i + 2]));
        tc += a; tf += b2;
    }
    if (ms_compact) *ms_compact = (float)tc;
    if (ms_forward) *ms_forward = (float)tf;
    if (n_pairs) *n_pairs = (uint32_t)(m->prof_used / 3);
    m->prof_used = 0;
    return QMANN_OK;
}

}  // extern "C"

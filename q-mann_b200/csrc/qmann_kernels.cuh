// qmann_kernels.cuh -- device code of the batched forward (included by qmann_forward.cu only).
//
//   k_compact  : streams the dense fp32 bag-of-words arenas (the reference's boundary format,
//                MemN2N.c:2294-2350) once, HBM-bound, and writes per story an ordered list of its
//                non-zero (column, value) entries with per-row end offsets.
//   k_forward  : one warp per story, persistent CTAs, all quantised weight tables resident in
//                shared memory as int8 codes.
//
// Arithmetic follows SURVEY.md Appendix A (integer forms proven against the reference by
// tests/golden/kat_*.npz); every formula cites the reference line it reproduces.
#pragma once
#include "qmann_fixed.cuh"
#include "../../include/qmann_abi.h"

namespace {

using namespace qmann;

constexpr int MAXH = QMANN_MAX_HOP;
constexpr unsigned REC_HDR_BYTES = 32;          // n_ent, flags, ans_idx, heap_off, n_exc, pad[3]
constexpr unsigned FLAG_HEAP = 1u, FLAG_ERROR = 2u;
constexpr unsigned MAX_EXC = 32;                // exception entries (values that are not 1.0) kept per story in the record
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
    unsigned rec_stride, off_rend, off_exc, off_ent, lcap;
    uint2 *heap;                       // overflow entries {column, fp32 bits}
    unsigned long long heap_cap;
    unsigned long long *heap_used;
    const unsigned char *colmax;       // [V] max |code| of a column over all embedding tables and dims
    unsigned nmax;                     // largest count n with n * 2^frac_w representable in every weight format (0: no splitting)
    unsigned split_lim;                // min over the hops of the weight-format code limit lw[h]: n copies of a unit entry equal the
                                       // reference's per-product clamp Q_w(Q_w(n) * Q_w(T)) only while n * max|code| <= lw[h] for every hop
    const unsigned *work_list;         // optional: chunk indices of the stories to compact (NULL: all of the chunk); the launch
    const unsigned *work_count;        // handles list positions work_off .. work_off + work_cap and writes record (position - work_off)
    unsigned work_off, work_cap;
};

// Scratch layout and shared-memory image of the production kernel k_story (qmann_fast.cuh); byte offsets.
struct FastLayout {
    unsigned sA[MAXH], sCM, sTAU, sW8;     // shared-memory image: A_h tables, packed column maxima, tau, int8 image of W
    unsigned tables_bytes;
    unsigned warp_bytes, LW;               // per-warp scratch; capacity of the entry list (16-bit premultiplied offsets)
    unsigned o_rend, o_sc, o_ex, o_pq, o_uvec, o_ub32, o_ovec, o_ufl, o_zent, o_brow, o_perm, o_cnt, o_bar, o_stage;
    unsigned NB, R, buf_bytes;             // dense stream: staging buffers per warp, rows per chunk, bytes per buffer
};

struct FwdParams {
    // quantised images (global) and their layout, copied into shared memory per CTA
    const unsigned char *img;
    unsigned img_bytes;
    unsigned offB, offA[MAXH], offC[MAXH], offH[MAXH], offW;
    unsigned offCM[MAXH], offTAU;          // packed path of k_forward_fast: per-column max |code| of A_h, saturation thresholds
    // answer-projection prefilter of k_forward_fast: int8 image of W (global offset offW8, rows of W8S bytes; it takes the
    // place of the fp32 rows in that kernel's shared memory), ok flag, and 1e-5 in integer-dot units
    unsigned offW8, W8S, w8_bytes;
    int w8_ok, ans_margin;
    unsigned V, d, S_max, H, lin_map;
    int const_scale;
    unsigned DP, HS, WS;
    // formats: fractional bits, integer bits and code limits
    int fw[MAXH], lw[MAXH], iw[MAXH];      // weight layers
    int fa[MAXH], la[MAXH], ia[MAXH];      // addressing
    int ff[MAXH], lf[MAXH], iff[MAXH];     // read + update
    int fb, lb;                            // u operand of scorer / linear map
    // compact records
    const unsigned char *rec;
    unsigned rec_stride, off_rend, off_exc, off_ent;
    const uint2 *heap;
    const unsigned long long *sen_off;
    unsigned story0, n_stories, n_total;
    unsigned long long sum_sen;
    // per-warp shared-memory scratch layout (byte offsets inside the warp's scratch)
    unsigned warp_bytes, LW, S_pad, o_rend, o_sc, o_ex, o_pq, o_uvec, o_ub32, o_ovec, o_ufl, o_exc, o_zent, o_brow, o_perm, o_cnt, tables_bytes;
    // outputs
    unsigned *pred;
    float *h_true;
    unsigned *match;
    unsigned *counter;
    unsigned *err_flag;
    int want_h;
    // two-pass dispatch: k_forward_fast appends the chunk indices of the stories it does not handle to
    // slow_list; the general kernel then takes its work from work_list[0 .. *work_count)
    unsigned *slow_list, *slow_count;
    const unsigned *work_list, *work_count;
    unsigned work_off, work_cap;           // general kernel: list positions work_off .. work_off + work_cap of work_list
    unsigned rec_by_pos;                   // general kernel: record index = list position - work_off (else the chunk index)
    // linear-map product table (global, L2-resident): lut[offL[h] + (j*255 + v + 127)*DP + i] =
    // Q_w(Q_w(Hm[i][j]) * v) for every code v of Q_bin(u[j]); NULL: compute the products (large d)
    const signed char *lut;
    unsigned offL[MAXH];
    qmann_debug dbg;
    // ---- production kernel k_story ----
    FastLayout fl;
    const float *dm, *dq, *da;             // dense arenas (DENSE source): streamed by the kernel itself with bulk copies
    unsigned long long m_bytes, q_bytes;   // their sizes in bytes (copies never read past them)
    const unsigned char *colmax;           // count splitting, as in CompactParams
    unsigned nmax, split_lim;
    int fast_softmax;                      // attention codes from a float total when no weight is near a truncation boundary
    unsigned pf_dist, pf_mode;             // L2 prefetch of the story pf_dist claims ahead (0: off, 1: bulk prefetch, 2: per line)
    unsigned long long *path_count;        // [3] stories that entered the packed, unpacked and general tier
    // optional layers (general kernel only; the production tiers are switched off when either is on)
    int en_sc_att;                         // EN_SC_ATT: s' = s * sc_w[h] in fp32 between scorer and softmax (lib/layer_cuda.cu:4805)
    float sc_w[MAXH];
    int en_non_lin;                        // EN_NON_LINEARITY: RELU + Q_f after the hop update (lib/layer_cuda.cu:4548)
};

// =============================================================================================
// weight preparation: fp32 [dim_out][dim_in] -> int8 codes, transposed to [dim_in][row_stride]
// =============================================================================================
// emb tables (B, A_h, C_h): img[v*DP + c] = code(w[c][v]) -- CUDA_FLOAT_QUANT of the weight inside
// FIXED_MUL, reference lib/layer_cuda.cu:120 with formats from MemN2N.c:826-838.  Row V is an
// all-zero row: idle lanes of the gather loop read it instead of branching.
__global__ void k_prep_emb(const float *__restrict__ w, signed char *__restrict__ img, unsigned V, unsigned d, unsigned DP, int iwl, int frac)
{
    const size_t n = (size_t)(V + 1) * DP;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned v = (unsigned)(i / DP), c = (unsigned)(i % DP);
        img[i] = (c < d && v < V) ? (signed char)qi_encode(w[(size_t)c * V + v], iwl, frac) : (signed char)0;
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
// product table of the linear map: one int8 row of DP outputs per (input dim j, input code v)
__global__ void k_prep_lut(const float *__restrict__ w, signed char *__restrict__ lut, unsigned d, unsigned DP, int iwl, int frac, int fb)
{
    const size_t n = (size_t)d * 255 * DP;
    const int lim = fixed_max(iwl, frac);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned c = (unsigned)(i % DP);
        const unsigned jv = (unsigned)(i / DP);
        const unsigned j = jv / 255;
        const int v = (int)(jv % 255) - 127;
        lut[i] = (c < d) ? (signed char)qi_mul(qi_encode(w[(size_t)c * d + j], iwl, frac), v, lim, fb) : (signed char)0;
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

// Packed path of k_forward_fast: cm[v] = max_c |A_h code[v][c]| (packed for up to three hops in the 10-bit fields of
// cm10[v]) and the A_h table rewritten as code + cm[v] (all DP bytes of a row, so that padding dims unbias to 0); row V
// stays all-zero with cm = 0.
__global__ void k_prep_bias(signed char *__restrict__ tab, unsigned *__restrict__ cm10, unsigned hop, unsigned V, unsigned DP)
{
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v <= V; v += gridDim.x * blockDim.x) {
        int mx = 0;
        for (unsigned c = 0; c < DP; c++) mx = max(mx, abs((int)tab[(size_t)v * DP + c]));
        cm10[v] |= (unsigned)mx << (10u * hop);           // 10-bit field per hop (launches of the hops are serialised)
        unsigned char *ut = reinterpret_cast<unsigned char *>(tab);
        for (unsigned c = 0; c < DP; c++) ut[(size_t)v * DP + c] = (unsigned char)((int)tab[(size_t)v * DP + c] + mx);
    }
}
// tau[|u|] = 0x80 - ceil(512 / |u|) (0 for |u| <= 4): a byte |y| + tau[|u|] has bit 7 set iff |y * u| >= 512, i.e. iff the
// product saturates Q_att (la = 127, two fractional bits in u)
__global__ void k_prep_tau(unsigned char *__restrict__ tau)
{
    const unsigned u = threadIdx.x;
    if (u < 128) tau[u] = (u <= 4) ? 0 : (unsigned char)(128u - (512u + u - 1u) / u);
}

// =============================================================================================
// k_compact: dense fp32 BoW -> ordered (column, value) lists.  One warp per story.
// Record layout (rec_stride bytes per story):
//   +0   u32 n_ent      total entries (question row + all sentence rows)
//   +4   u32 flags      FLAG_HEAP: entries live in the overflow heap at heap_off as {column, fp32 bits};
//                       FLAG_ERROR: heap exhausted
//   +8   u32 ans_idx    index of the 1.0 in the answer row (last one), ANS_NONE without answers
//   +12  u32 heap_off
//   +16  u32 n_exc      number of exception entries (<= MAX_EXC unless FLAG_HEAP)
//   +off_rend  u16 rend[S_max+2]   rend[k] = end (exclusive) of row k; row 0 is the question, rows 1..S the sentences
//   +off_exc   uint2 exc[MAX_EXC]  {row | column << 16, fp32 bits}: values that are not 1.0 and cannot be split (below)
//   +off_ent   u32 ent[lcap]       columns of the entries whose value is 1.0 ("unit" entries)
// Entries of one row are contiguous; their order inside a row is irrelevant (the per-row sums of
// SURVEY A.2 are exact integers).
// Count splitting: a word repeated n times in a sentence has value n.  Its embedding term is
// Q_w(Q_w(n) * Q_w(T)) = clamp(n * t); when n * max|t| over every table and dimension of that column
// is <= 127 and n * 2^frac_w does not saturate, that equals t + ... + t, so the entry is emitted as
// n unit entries (colmax[] and nmax come from the model).  Everything else is an exception entry.
// =============================================================================================
__device__ __forceinline__ float4 ldg_stream4(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldg_stream2(const float2 *p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float *p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned nzbit(float x) { return ((__float_as_uint(x) << 1) != 0u) ? 1u : 0u; }

// ---- pieces of the row scan --------------------------------------------------------------------
// True when some lane of the warp holds a value other than 0.0 / 1.0 among its 2*W values
// (x*x - x == 0 exactly iff x is 0 or 1; one FFMA per value on the FMA pipes, which this kernel leaves idle).
template <int W>
__device__ __forceinline__ bool chunk_irregular(const float (&v)[2 * W])
{
    unsigned bad = 0;
#pragma unroll
    for (int j = 0; j < 2 * W; j++) bad |= __float_as_uint(__fmaf_rn(v[j], v[j], -v[j]));
    return __any_sync(0xffffffffu, (bad << 1) != 0u);
}

// Bit j set iff v[j] == 1.0, for values known to be 0.0 or 1.0: sum_j 2^j v[j] formed on top of 2^23, so the
// integer lands in the low mantissa bits (two FFMA chains, no conversion instruction).
template <int W>
__device__ __forceinline__ unsigned unit_mask(const float (&v)[2 * W])
{
    if (W == 4) {
        float s0 = __fmaf_rn(v[0], 1.0f, 8388608.0f), s1 = __fmaf_rn(v[4], 16.0f, 8388608.0f);
        s0 = __fmaf_rn(v[1], 2.0f, s0);  s1 = __fmaf_rn(v[5], 32.0f, s1);
        s0 = __fmaf_rn(v[2], 4.0f, s0);  s1 = __fmaf_rn(v[6], 64.0f, s1);
        s0 = __fmaf_rn(v[3], 8.0f, s0);  s1 = __fmaf_rn(v[7], 128.0f, s1);
        return (__float_as_uint(s0) | __float_as_uint(s1)) & 0xFFu;
    }
    if (W == 2) {
        float s = __fmaf_rn(v[1], 2.0f, v[0] + 8388608.0f);
        s = __fmaf_rn(v[2], 4.0f, s);
        s = __fmaf_rn(v[3], 8.0f, s);
        return __float_as_uint(s) & 0xFu;
    }
    const float s = __fmaf_rn(v[1], 2.0f, v[0] + 8388608.0f);
    return __float_as_uint(s) & 0x3u;
}

// Appends one unit entry per set bit of mm (branch-free: idle lanes run the same instructions predicated off).
template <int W>
__device__ __forceinline__ unsigned emit_units(unsigned mm, unsigned ca, unsigned cb, unsigned *__restrict__ ent, unsigned cap, unsigned base, unsigned lt)
{
    unsigned any;
    while ((any = __ballot_sync(0xffffffffu, mm != 0u)) != 0u) {
        const unsigned k = (unsigned)(__ffs((int)mm) - 1);               // garbage when mm == 0, unused
        const unsigned col = ((k >= (unsigned)W) ? cb : ca) * W + (k & (W - 1));
        const unsigned pos = base + __popc(any & lt);
        const unsigned ok = (mm != 0u) & (pos < cap);
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.global.u32 [%1], %2;\n\t}" ::"r"(ok), "l"(ent + pos), "r"(col) : "memory");
        mm &= mm - 1u;                                                     // 0 stays 0
        base += __popc(any);
    }
    return base;
}

// General emission of one chunk (MODE 0: unit entries, count splitting, exception entries; MODE 2: {column, bits}
// pairs into the heap).  Each pass emits one entry per lane that still has something pending.
template <int W, int MODE>
__device__ __forceinline__ unsigned emit_general(const CompactParams &p, const float (&v)[2 * W], unsigned ca, unsigned cb, unsigned rix,
                                                 unsigned *__restrict__ ent, uint2 *__restrict__ heap_dst, unsigned cap, unsigned base, unsigned lt,
                                                 uint2 *__restrict__ exc, unsigned &n_exc)
{
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < 2 * W; j++) m |= nzbit(v[j]) << j;
    unsigned rep = 0;                            // unit entries still owed for a split count
    unsigned col = 0;
    unsigned any;
    while ((any = __ballot_sync(0xffffffffu, (m | rep) != 0u)) != 0u) {
        const bool had = ((m | rep) != 0u);
        const unsigned pos = base + __popc(any & lt);
        bool unit = true;
        float x = 1.0f;
        if (rep) {
            rep--;                               // another copy of the same column
        } else if (had) {
            const unsigned k = (unsigned)(__ffs((int)m) - 1);
            m &= m - 1u;
            x = v[0];
#pragma unroll
            for (int j = 1; j < 2 * W; j++) x = (k == (unsigned)j) ? v[j] : x;
            col = ((k >= (unsigned)W) ? cb : ca) * W + (k & (W - 1));
            unit = (__float_as_uint(x) == 0x3F800000u);
        }
        if (MODE == 2) {
            if (had && pos < cap) heap_dst[pos] = make_uint2(col, __float_as_uint(x));
        } else {
            const unsigned bnu = __ballot_sync(0xffffffffu, had && !unit);
            bool is_exc = false;
            if (bnu) {                           // some lane holds a value that is not 1.0
                if (had && !unit) {
                    const float n = truncf(x);
                    if (n == x && x >= 2.0f && x <= (float)p.nmax && (unsigned)n * (unsigned)p.colmax[col] <= p.split_lim) rep = (unsigned)n - 1u;
                    else is_exc = true;
                }
                const unsigned bex = __ballot_sync(0xffffffffu, is_exc);
                if (bex) {
                    if (is_exc) {
                        const unsigned xi = n_exc + __popc(bex & lt);
                        if (xi < MAX_EXC) exc[xi] = make_uint2(rix | (col << 16), __float_as_uint(x));
                    }
                    n_exc += __popc(bex);
                    // exception entries do not occupy the unit list: close the gap they would leave
                    const unsigned live = any & ~bex;
                    if (had && !is_exc) { const unsigned pos2 = base + __popc(live & lt); if (pos2 < cap) ent[pos2] = col; }
                    base += __popc(live);
                    continue;
                }
            }
            if (had && pos < cap) ent[pos] = col;
        }
        base += __popc(any);
    }
    return base;
}

// Scans one row of V floats and appends its entries at `base` (warp-uniform running count).
// W = floats per load (4: 128-bit loads, needs V % 4 == 0 and 16-byte alignment; 2: 64-bit loads, V even and 8-byte
// alignment; 1: scalar).
// MODE 0: compact record (unit entries, count splitting, exception entries)
// MODE 1: count the non-zero values only
// MODE 2: {column, bits} pairs into the heap
// Stores beyond `cap` are dropped (the caller falls back to the heap).  Returns the new count.
template <int W, int MODE>
__device__ __forceinline__ unsigned scan_row(const CompactParams &p, const float *__restrict__ row, unsigned rix, unsigned *__restrict__ ent,
                                             uint2 *__restrict__ heap_dst, unsigned cap, unsigned base, unsigned lane,
                                             uint2 *__restrict__ exc, unsigned &n_exc)
{
    const unsigned lt = (1u << lane) - 1u;
    const unsigned VW = p.V / W;                     // loads per row
    for (unsigned c0 = 0; c0 < VW; c0 += 64) {
        // two independent loads in flight per lane
        const unsigned ca = c0 + lane, cb = c0 + 32 + lane;
        float v[2 * W];
#pragma unroll
        for (int j = 0; j < 2 * W; j++) v[j] = 0.0f;
        if (W == 4) {
            const float4 *r4 = reinterpret_cast<const float4 *>(row);
            if (ca < VW) { const float4 t = ldg_stream4(r4 + ca); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
            if (cb < VW) { const float4 t = ldg_stream4(r4 + cb); v[4] = t.x; v[5] = t.y; v[6] = t.z; v[7] = t.w; }
        } else if (W == 2) {
            const float2 *r2 = reinterpret_cast<const float2 *>(row);
            if (ca < VW) { const float2 t = ldg_stream2(r2 + ca); v[0] = t.x; v[1] = t.y; }
            if (cb < VW) { const float2 t = ldg_stream2(r2 + cb); v[2] = t.x; v[3] = t.y; }
        } else {
            if (ca < VW) v[0] = ldg_stream1(row + ca);
            if (cb < VW) v[1] = ldg_stream1(row + cb);
        }
        if (MODE == 1) {
            unsigned m = 0;
#pragma unroll
            for (int j = 0; j < 2 * W; j++) m |= nzbit(v[j]) << j;
            base += __reduce_add_sync(0xffffffffu, (unsigned)__popc(m));
        } else if (MODE == 0 && !chunk_irregular<W>(v)) {
            base = emit_units<W>(unit_mask<W>(v), ca, cb, ent, cap, base, lt);     // the common case
        } else {
            base = emit_general<W, MODE>(p, v, ca, cb, rix, ent, heap_dst, cap, base, lt, exc, n_exc);
        }
    }
    return base;
}

// MODE-0 scan of a whole story whose rows fit one chunk (V <= 256) with 128-bit loads, software-pipelined: the
// loads of row r+1 are in flight while row r is classified and emitted (two register stages, unrolled by two so
// that no register is copied).  FULL: V/4 == 64, no bounds checks.
template <bool FULL>
__device__ __forceinline__ unsigned scan_story_pipelined(const CompactParams &p, unsigned story, unsigned S, unsigned long long soff,
                                                         unsigned *__restrict__ ent, unsigned cap, unsigned short *__restrict__ rend,
                                                         uint2 *__restrict__ exc, unsigned &n_exc, unsigned lane)
{
    const unsigned lt = (1u << lane) - 1u;
    const unsigned VW = p.V >> 2;
    const unsigned ca = lane, cb = lane + 32u;
    const float4 *qrow = reinterpret_cast<const float4 *>(p.q + (size_t)story * p.V) + lane;
    const float4 *mrow = reinterpret_cast<const float4 *>(p.m + (size_t)soff * p.V) + lane;     // row r >= 1 is mrow + (r-1)*VW
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load = [&](const float4 *row, float4 &a, float4 &b) {
        if (FULL) { a = ldg_stream4(row); b = ldg_stream4(row + 32); }
        else { a = (ca < VW) ? ldg_stream4(row) : zero4; b = (cb < VW) ? ldg_stream4(row + 32) : zero4; }
    };
    unsigned base = 0;
    auto process = [&](const float4 &a, const float4 &b, unsigned r) {
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (!chunk_irregular<4>(v)) base = emit_units<4>(unit_mask<4>(v), ca, cb, ent, cap, base, lt);
        else base = emit_general<4, 0>(p, v, ca, cb, r, ent, nullptr, cap, base, lt, exc, n_exc);
        if (lane == 0) rend[r] = (unsigned short)min(base, 0xFFFFu);
    };
    float4 a0, b0, a1, b1;
    load(qrow, a0, b0);
    unsigned r = 0;
    for (;;) {
        if (r < S) load(mrow, a1, b1);            // row r+1
        process(a0, b0, r);
        if (++r > S) break;
        mrow += VW;
        if (r < S) load(mrow, a0, b0);            // row r+1
        process(a1, b1, r);
        if (++r > S) break;
        mrow += VW;
    }
    return base;
}

template <int W, int MODE>
__device__ __forceinline__ unsigned scan_story(const CompactParams &p, unsigned story, unsigned S, unsigned long long soff, unsigned *ent,
                                               uint2 *heap_dst, unsigned cap, unsigned short *rend, uint2 *exc, unsigned &n_exc, unsigned lane)
{
    unsigned cnt = scan_row<W, MODE>(p, p.q + (size_t)story * p.V, 0u, ent, heap_dst, cap, 0u, lane, exc, n_exc);
    if (MODE != 1 && lane == 0) rend[0] = (unsigned short)min(cnt, 0xFFFFu);
    const float *mrow = p.m + (size_t)soff * p.V;
    for (unsigned r = 0; r < S; r++) {
        cnt = scan_row<W, MODE>(p, mrow + (size_t)r * p.V, r + 1u, ent, heap_dst, cap, cnt, lane, exc, n_exc);
        if (MODE != 1 && lane == 0) rend[r + 1] = (unsigned short)min(cnt, 0xFFFFu);
    }
    return cnt;
}

template <int W>
__global__ void __launch_bounds__(256) k_compact(const CompactParams p)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned warps = (gridDim.x * blockDim.x) >> 5;
    unsigned n_work = p.n_stories;
    if (p.work_list) { const unsigned c = *p.work_count; n_work = (c > p.work_off) ? min(c - p.work_off, p.work_cap) : 0u; }
    for (unsigned wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < n_work; wi += warps) {
        const unsigned w = p.work_list ? p.work_list[p.work_off + wi] : wi;
        const unsigned story = p.story0 + w;
        const unsigned long long soff = p.sen_off[story];
        const unsigned S = (unsigned)(p.sen_off[story + 1] - soff);
        unsigned char *rec = p.rec + (size_t)wi * p.rec_stride;
        unsigned *hdr = reinterpret_cast<unsigned *>(rec);
        unsigned short *rend = reinterpret_cast<unsigned short *>(rec + p.off_rend);
        uint2 *exc = reinterpret_cast<uint2 *>(rec + p.off_exc);
        unsigned *ent = reinterpret_cast<unsigned *>(rec + p.off_ent);

        unsigned n_exc = 0;
        unsigned n;
        if (W == 4 && p.V <= 256u) {
            n = (p.V == 256u) ? scan_story_pipelined<true>(p, story, S, soff, ent, p.lcap, rend, exc, n_exc, lane)
                              : scan_story_pipelined<false>(p, story, S, soff, ent, p.lcap, rend, exc, n_exc, lane);
        } else {
            n = scan_story<W, 0>(p, story, S, soff, ent, nullptr, p.lcap, rend, exc, n_exc, lane);
        }
        unsigned flags = 0, heap_off = 0;
        if (n > p.lcap || n > 0xFFFFu || n_exc > MAX_EXC) {
            // rare: denser than the fixed slot, or too many exception entries.  Count the raw
            // non-zeros, reserve exactly that many {column, value} pairs in the heap and rescan.
            unsigned dummy = 0;
            n = scan_story<W, 1>(p, story, S, soff, nullptr, nullptr, 0u, nullptr, nullptr, dummy, lane);
            unsigned long long off = 0;
            if (lane == 0) off = atomicAdd(p.heap_used, (unsigned long long)n);
            off = __shfl_sync(0xffffffffu, off, 0);
            if (n > 0xFFFFu || off + n > p.heap_cap || off + n > 0xFFFFFFFFull) flags = FLAG_ERROR;
            else {
                flags = FLAG_HEAP;
                heap_off = (unsigned)off;
                scan_story<W, 2>(p, story, S, soff, nullptr, p.heap + off, n, rend, nullptr, dummy, lane);
            }
            n_exc = 0;
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
        if (lane == 0) { hdr[0] = n; hdr[1] = flags; hdr[2] = ans; hdr[3] = heap_off; hdr[4] = n_exc; }
    }
}

// =============================================================================================
// k_ids_compact: word-id lists -> the same compact records.
// The input is what the reference holds BEFORE sample_vectorization() scatters it into the dense arenas
// (MemN2N/sample.c:413-575): per row (question, then the story's sentences) the list of column ids, every
// occurrence adding 1.0 to its column (sample.c:547, :560, :568; a sentence's last id is its time column
// V_dict + n_sen-1-j, sample.c:474).  An id that occurs n times in a row is the dense value n and is classified
// exactly like k_compact does: n unit entries when n * colmax <= 127 and n <= nmax, otherwise one exception entry
// {row | column << 16, (float)n}.  One warp per story; ~0.6 KB read per story instead of 52 KB.
// =============================================================================================
struct IdsParams {
    const unsigned short *ids;         // all rows back to back
    const unsigned *row_off;           // [N + sum_sen + 1] prefix offsets into ids; story i owns rows sen_off[i]+i .. (question first)
    const unsigned *ans;               // [N] answer column or NULL
    const unsigned long long *sen_off;
    unsigned V;
    unsigned story0, n_stories;
    unsigned char *rec;
    unsigned rec_stride, off_rend, off_exc, off_ent, lcap;
    uint2 *heap;
    unsigned long long heap_cap;
    unsigned long long *heap_used;
    const unsigned char *colmax;
    unsigned nmax, split_lim;
};

// Classification and emission of the (row, id) occurrences held one per lane (`have` lanes, in row-major order).
// cnt = occurrences of the id in its row, earlier = an occurrence sits in a lower position of the row.
// MODE 0: compact record; MODE 1: count the distinct ids per row (= non-zero dense values); MODE 2: {column, fp32 count}
// pairs into the heap.  Returns the ballot of the lanes that emitted a list entry; base is advanced past them.
template <int MODE>
__device__ __forceinline__ unsigned ids_emit(const IdsParams &p, bool have, unsigned id, unsigned row, unsigned cnt, bool earlier,
                                             unsigned *__restrict__ ent, uint2 *__restrict__ heap_dst, unsigned cap, unsigned &base,
                                             uint2 *__restrict__ exc, unsigned &n_exc, bool &bad, unsigned lt)
{
    const bool oob = have && id >= p.V;
    bad |= __any_sync(0xffffffffu, oob);
    const bool ok = have && !oob;
    unsigned bl;
    if (MODE == 0) {
        const bool unit = ok && (cnt == 1u || (cnt <= p.nmax && cnt * (unsigned)p.colmax[id] <= p.split_lim));
        const bool isx = ok && !unit && !earlier;
        bl = __ballot_sync(0xffffffffu, unit);
        const unsigned pos = base + __popc(bl & lt);
        if (unit && pos < cap) ent[pos] = id;
        const unsigned bx = __ballot_sync(0xffffffffu, isx);
        if (bx) {
            const unsigned xi = n_exc + __popc(bx & lt);
            if (isx && xi < MAX_EXC) exc[xi] = make_uint2(row | (id << 16), __float_as_uint((float)cnt));
            n_exc += __popc(bx);
        }
    } else {
        const bool first = ok && !earlier;
        bl = __ballot_sync(0xffffffffu, first);
        const unsigned pos = base + __popc(bl & lt);
        if (MODE == 2 && first && pos < cap) heap_dst[pos] = make_uint2(id, __float_as_uint((float)cnt));
    }
    base += __popc(bl);
    return bl;
}

// One story.  Rows are packed into tiles: as many whole consecutive rows as hold at most 32 ids together, one id
// per lane, so that a bAbI-shaped story (rows of ~5 ids) takes ~9 warp steps instead of one per row; duplicates
// inside a row are found with one match on (row, id).  A row longer than 32 ids is handled alone.
template <int MODE>
__device__ __forceinline__ unsigned ids_scan_story(const IdsParams &p, unsigned long long row0, unsigned S, unsigned *__restrict__ ent,
                                                   uint2 *__restrict__ heap_dst, unsigned cap, unsigned short *__restrict__ rend,
                                                   uint2 *__restrict__ exc, unsigned &n_exc, bool &bad, unsigned lane)
{
    const unsigned lt = (1u << lane) - 1u;
    unsigned base = 0;
    unsigned ra = 0;                                   // next row (0 = question)
    while (ra <= S) {
        const unsigned r = ra + lane;
        const bool valid = r <= S;
        const unsigned o_lo = valid ? p.row_off[row0 + r] : 0u;
        const unsigned o_hi = valid ? p.row_off[row0 + r + 1] : o_lo;
        const unsigned len = o_hi - o_lo;
        unsigned incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        const unsigned k = (unsigned)__popc(__ballot_sync(0xffffffffu, valid && incl <= 32u));      // whole rows in this tile
        const unsigned first_off = __shfl_sync(0xffffffffu, o_lo, 0);
        if (k == 0) {
            // row `ra` alone, 32 ids at a time, occurrences counted by walking the row
            const unsigned L = __shfl_sync(0xffffffffu, len, 0);
            for (unsigned c0 = 0; c0 < L; c0 += 32) {
                const unsigned i = c0 + lane;
                const bool have = i < L;
                const unsigned id = have ? (unsigned)p.ids[first_off + i] : 0u;
                unsigned cnt = 0;
                bool earlier = false;
                for (unsigned j = 0; j < L; j++) {
                    const bool same = (unsigned)p.ids[first_off + j] == id;
                    cnt += same ? 1u : 0u;
                    earlier |= same && (j < i);
                }
                ids_emit<MODE>(p, have, id, ra, cnt, earlier, ent, heap_dst, cap, base, exc, n_exc, bad, lt);
            }
            if (MODE != 1 && lane == 0) rend[ra] = (unsigned short)min(base, 0xFFFFu);
            ra += 1;
            continue;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, k - 1);
        const bool have = lane < total;
        // row of this lane's id inside the tile: j = #{t < k : incl[t] <= lane}
        unsigned j = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const unsigned t = j + (unsigned)step;
            const unsigned v = __shfl_sync(0xffffffffu, incl, min(t, 32u) - 1u);
            if (t <= k && v <= lane) j = t;
        }
        const unsigned id = have ? (unsigned)p.ids[first_off + lane] : 0u;
        const unsigned peers = __match_any_sync(0xffffffffu, have ? ((j << 16) | id) : (0x80000000u | lane));
        const unsigned base0 = base;
        const unsigned bl = ids_emit<MODE>(p, have, id, ra + j, (unsigned)__popc(peers), (peers & lt) != 0u, ent, heap_dst, cap, base, exc, n_exc, bad, lt);
        if (MODE != 1 && lane < k) {
            const unsigned below = (incl >= 32u) ? 0xFFFFFFFFu : ((1u << incl) - 1u);                 // lanes of rows ra .. ra+lane
            rend[ra + lane] = (unsigned short)min(base0 + (unsigned)__popc(bl & below), 0xFFFFu);
        }
        ra += k;
    }
    return base;
}

__global__ void __launch_bounds__(256) k_ids_compact(const IdsParams p)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < p.n_stories; w += warps) {
        const unsigned story = p.story0 + w;
        const unsigned long long soff = p.sen_off[story];
        const unsigned S = (unsigned)(p.sen_off[story + 1] - soff);
        const unsigned long long row0 = soff + story;
        unsigned char *rec = p.rec + (size_t)w * p.rec_stride;
        unsigned *hdr = reinterpret_cast<unsigned *>(rec);
        unsigned short *rend = reinterpret_cast<unsigned short *>(rec + p.off_rend);
        uint2 *exc = reinterpret_cast<uint2 *>(rec + p.off_exc);
        unsigned *ent = reinterpret_cast<unsigned *>(rec + p.off_ent);
        unsigned n_exc = 0;
        bool bad = false;
        unsigned n = ids_scan_story<0>(p, row0, S, ent, nullptr, p.lcap, rend, exc, n_exc, bad, lane);
        unsigned flags = 0, heap_off = 0;
        if (!bad && (n > p.lcap || n > 0xFFFFu || n_exc > MAX_EXC)) {
            unsigned dummy = 0;
            n = ids_scan_story<1>(p, row0, S, nullptr, nullptr, 0u, nullptr, nullptr, dummy, bad, lane);
            unsigned long long off = 0;
            if (lane == 0) off = atomicAdd(p.heap_used, (unsigned long long)n);
            off = __shfl_sync(0xffffffffu, off, 0);
            if (n > 0xFFFFu || off + n > p.heap_cap || off + n > 0xFFFFFFFFull) flags = FLAG_ERROR;
            else {
                flags = FLAG_HEAP;
                heap_off = (unsigned)off;
                ids_scan_story<2>(p, row0, S, nullptr, p.heap + off, n, rend, nullptr, dummy, bad, lane);
            }
            n_exc = 0;
        }
        if (bad) flags = FLAG_ERROR;
        unsigned ans = ANS_NONE;
        if (p.ans) {
            ans = p.ans[story];
            if (ans >= p.V) ans = ANS_NONE;
        }
        if (lane == 0) { hdr[0] = n; hdr[1] = flags; hdr[2] = ans; hdr[3] = heap_off; hdr[4] = n_exc; }
    }
}

// per-column max |code| over an embedding table image (rows of DP int8), folded into colmax[] with max
__global__ void k_colmax(const signed char *__restrict__ img, unsigned V, unsigned DP, unsigned char *__restrict__ colmax)
{
    for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        int mx = colmax[v];
        for (unsigned c = 0; c < DP; c++) mx = max(mx, abs((int)img[(size_t)v * DP + c]));
        colmax[v] = (unsigned char)mx;
    }
}

// =============================================================================================
// k_forward
// =============================================================================================
extern __shared__ __align__(16) unsigned char smem[];

__device__ __forceinline__ int sbyte(unsigned w, int b) { return (int)(signed char)((w >> (8 * b)) & 0xFFu); }

// byte b of w, sign-extended, in one PRMT (selector nibble 8|b replicates the sign of byte b)
template <int B>
__device__ __forceinline__ int sbyte_prmt(unsigned w)
{
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "r"((unsigned)(B | ((8 | B) << 4) | ((8 | B) << 8) | ((8 | B) << 12))));
    return (int)r;
}

// v / 2^sh toward zero with a precomputed mask = 2^sh - 1: (v + (v < 0 ? mask : 0)) >> sh
// (written as PTX: the compiler otherwise strength-reduces the multiply into SHF + LOP3 + IADD, one
// instruction more on the already saturated ALU pipe)
__device__ __forceinline__ int shr0m(int v, int sh, int mask)
{
    int t;
    asm("{\n\t.reg .u32 s;\n\tshr.u32 s, %1, 31;\n\tmad.lo.s32 %0, s, %2, %1;\n\t}" : "=r"(t) : "r"(v), "r"(mask));
    return t >> sh;
}
// clamp(v, -L, L) + L in one instruction (VIADDMNMX.RELU): max(min(v + L, 2L), 0)
__device__ __forceinline__ int clamp_biased(int v, int L, int L2) { return __viaddmin_s32_relu(v, L, L2); }

struct Story {
    unsigned ws;                   // shared-memory byte offset of this warp's scratch
    bool in_smem;                  // unit entries (columns) are in shared memory; exception entries beside them
    unsigned n_exc;                // exception entries {row | column << 16, value bits} in shared memory
    const uint2 *ent_g;            // heap story: every entry as a {column, value} pair in global memory
    unsigned o_rend, o_exc, o_zent;
};

// Gather-and-sum embedding of up to 32/LPR rows at once (one row per LPR-lane group; lane q of a
// group owns dims 16q..16q+15).  acc[j] = sum over the row's entries of Q_w(Q_w(x) * Q_w(T[c][id]))
// -- the per-product quantise + clamp of the reference (lib/layer_cuda.cu:120) is kept, which is why
// this is not a dp4a dot product over ids.  For x == 1.0 the term is the table code itself
// (Q_w(1.0) = 2^frac_w when iwl_w >= 1) and dp4a with a one-hot selector does the sign-extending
// byte accumulate.  `row` is the record row index (0 = question) or -1 for an idle group.
// `tab` is the shared-memory byte offset of the table, whose row V is all zero.
template <int LPR>
__device__ __forceinline__ void embed_rows(const FwdParams &p, const Story &st, unsigned lane, unsigned tab, int row, int iwl_w, int frac_w,
                                           int lim_w, int acc[16], const int sel[4])
{
    const unsigned q = lane % LPR;
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0;
    unsigned beg = 0, len = 0;
    if (row >= 0) {
        const unsigned short *rend = reinterpret_cast<const unsigned short *>(smem + st.ws + st.o_rend);
        beg = row ? rend[row - 1] : 0u;
        len = rend[row] - beg;
    }
    const unsigned maxlen = __reduce_max_sync(0xffffffffu, len);
    const unsigned tabq = tab + 16u * q;
    if (st.in_smem && iwl_w >= 1) {
        // common case: branch-free; lanes past their row's end gather the zero row
        unsigned ea = st.ws + 4u * beg;                   // the entry list starts the warp scratch
        const unsigned zaddr = st.ws + st.o_zent;         // a pseudo entry naming the all-zero row V
        for (unsigned k = 0; k < maxlen; k++, ea += 4u) {
            const unsigned col = *reinterpret_cast<const unsigned *>(smem + ((k < len) ? ea : zaddr));
            const uint4 t = *reinterpret_cast<const uint4 *>(smem + tabq + col * p.DP);
            acc[0] = __dp4a((int)t.x, sel[0], acc[0]);   acc[1] = __dp4a((int)t.x, sel[1], acc[1]);
            acc[2] = __dp4a((int)t.x, sel[2], acc[2]);   acc[3] = __dp4a((int)t.x, sel[3], acc[3]);
            acc[4] = __dp4a((int)t.y, sel[0], acc[4]);   acc[5] = __dp4a((int)t.y, sel[1], acc[5]);
            acc[6] = __dp4a((int)t.y, sel[2], acc[6]);   acc[7] = __dp4a((int)t.y, sel[3], acc[7]);
            acc[8] = __dp4a((int)t.z, sel[0], acc[8]);   acc[9] = __dp4a((int)t.z, sel[1], acc[9]);
            acc[10] = __dp4a((int)t.z, sel[2], acc[10]); acc[11] = __dp4a((int)t.z, sel[3], acc[11]);
            acc[12] = __dp4a((int)t.w, sel[0], acc[12]); acc[13] = __dp4a((int)t.w, sel[1], acc[13]);
            acc[14] = __dp4a((int)t.w, sel[2], acc[14]); acc[15] = __dp4a((int)t.w, sel[3], acc[15]);
        }
        // exception entries of these rows: values that are not 1.0 (repeated words that cannot be
        // split, position-encoding-like weights): the general quantised product
        for (unsigned e = 0; e < st.n_exc; e++) {
            const uint2 xe = *reinterpret_cast<const uint2 *>(smem + st.ws + st.o_exc + 8u * e);
            if (row >= 0 && (xe.x & 0xFFFFu) == (unsigned)row) {
                const uint4 t = *reinterpret_cast<const uint4 *>(smem + tabq + (xe.x >> 16) * p.DP);
                const unsigned tw[4] = {t.x, t.y, t.z, t.w};
                const int xq = qi_encode(__uint_as_float(xe.y), iwl_w, frac_w);
#pragma unroll
                for (int w = 0; w < 4; w++)
#pragma unroll
                    for (int b = 0; b < 4; b++) acc[4 * w + b] += qi_mul(xq, sbyte(tw[w], b), lim_w, frac_w);
            }
        }
        return;
    }
    // general path: a weight format without integer bits (Q_w(1.0) != 2^frac_w), or a heap story
    // whose entries are {column, value} pairs in global memory
    const unsigned *ent_s = reinterpret_cast<const unsigned *>(smem + st.ws);
    for (unsigned k = 0; k < maxlen; k++) {
        if (k < len) {
            unsigned col;
            float x = 1.0f;
            if (st.in_smem) {
                col = ent_s[beg + k];
            } else {
                const uint2 e = st.ent_g[beg + k];
                col = e.x;
                x = __uint_as_float(e.y);
            }
            const uint4 t = *reinterpret_cast<const uint4 *>(smem + tabq + col * p.DP);
            const unsigned tw[4] = {t.x, t.y, t.z, t.w};
            const int xq = qi_encode(x, iwl_w, frac_w);
#pragma unroll
            for (int w = 0; w < 4; w++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[4 * w + b] += qi_mul(xq, sbyte(tw[w], b), lim_w, frac_w);
        }
    }
    if (st.in_smem) {
        for (unsigned e = 0; e < st.n_exc; e++) {
            const uint2 xe = *reinterpret_cast<const uint2 *>(smem + st.ws + st.o_exc + 8u * e);
            if (row >= 0 && (xe.x & 0xFFFFu) == (unsigned)row) {
                const uint4 t = *reinterpret_cast<const uint4 *>(smem + tabq + (xe.x >> 16) * p.DP);
                const unsigned tw[4] = {t.x, t.y, t.z, t.w};
                const int xq = qi_encode(__uint_as_float(xe.y), iwl_w, frac_w);
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

// Fixed-point dot-product scorer for one memory row held as 16 unclamped int32 sums per lane:
//   s[r] = Q_att( sum_t Q_att( Q_att(M[r][t]) * Q_bin(u[t]) ) )                 layer_cuda.cu:105-141
// KA = sign of (frac_att - frac_w): how M (weight format) is re-quantised to the addressing format.
// Returns sum_t (term_t + la) over this lane's 16 dims (the bias is removed by the caller).
template <int KA>
__device__ __forceinline__ int score_part(const int acc[16], const int ub[16], int lw, int la, int ka, int fb, int mb)
{
    const int lw2 = 2 * lw, la2 = 2 * la;
    int part = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        int m_att;
        if (KA == 0) {
            m_att = clamp_biased(acc[j], lw, lw2) - lw;                       // Q_w then Q_att are the same clamp (la == lw checked by caller)
        } else if (KA > 0) {
            m_att = clamp_biased(acc[j] << ka, la, la2) - la;                 // clamp_att(clamp_w(a) << k) == clamp_att(a << k) since la <= lw << k
        } else {
            const int mw = clamp_biased(acc[j], lw, lw2) - lw;
            m_att = qi_clamp(shr0m(mw, -ka, (1 << (-ka)) - 1), la);
        }
        const int x = m_att * ub[j];
        part += clamp_biased(shr0m(x, fb, mb), la, la2);
    }
    return part;
}

template <int LPR, int MODE, bool DEBUG>
__global__ void __launch_bounds__(512, 1) k_forward(const __grid_constant__ FwdParams p)
{
    constexpr int G = 32 / LPR;                 // rows embedded concurrently by one warp
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned g = lane / LPR, q = lane % LPR;
    unsigned n_work = p.n_stories;
    if (p.work_count) { const unsigned c = *p.work_count; n_work = (c > p.work_off) ? min(c - p.work_off, p.work_cap) : 0u; }
    if (n_work == 0) return;                    // nothing left over by the fast kernel
    if (p.path_count && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.path_count + 2, (unsigned long long)n_work);       // stories entering the general tier

    // ---- stage the quantised tables into shared memory (once per CTA) ----
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.img);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (unsigned i = threadIdx.x; i < p.tables_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    Story st;
    st.ws = p.tables_bytes + wid * p.warp_bytes;
    st.o_rend = p.o_rend;
    st.o_exc = p.o_exc;
    st.o_zent = p.o_zent;
    unsigned char *ws = smem + st.ws;
    if (lane == 0) *reinterpret_cast<unsigned *>(ws + p.o_zent) = p.V;
    unsigned *ent_s = reinterpret_cast<unsigned *>(ws);
    unsigned short *rend_s = reinterpret_cast<unsigned short *>(ws + p.o_rend);
    int *sc = reinterpret_cast<int *>(ws + p.o_sc);
    float *ex = reinterpret_cast<float *>(ws + p.o_ex);
    unsigned char *pq = ws + p.o_pq;
    signed char *uvec = reinterpret_cast<signed char *>(ws + p.o_uvec);
    int *ub32 = reinterpret_cast<int *>(ws + p.o_ub32);
    signed char *ovec = reinterpret_cast<signed char *>(ws + p.o_ovec);
    float *ufl = reinterpret_cast<float *>(ws + p.o_ufl);
    uint2 *excs = reinterpret_cast<uint2 *>(ws + p.o_exc);
    float *zbuf = reinterpret_cast<float *>(ws);      // aliases the entry list (dead by the answer phase)
    const unsigned d = p.d, DP = p.DP, V = p.V;
    // one-hot dp4a selectors kept in registers (opaque to constant propagation)
    int sel[4];
    asm volatile("mov.u32 %0, 0x00000001;" : "=r"(sel[0]));
    asm volatile("mov.u32 %0, 0x00000100;" : "=r"(sel[1]));
    asm volatile("mov.u32 %0, 0x00010000;" : "=r"(sel[2]));
    asm volatile("mov.u32 %0, 0x01000000;" : "=r"(sel[3]));

    for (;;) {
        unsigned w = 0;
        if (lane == 0) w = atomicAdd(p.counter, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_work) break;
        const unsigned pos = w;
        if (p.work_list) w = p.work_list[p.work_off + pos];
        const unsigned story = p.story0 + w;
        const unsigned long long soff = p.sen_off[story];
        const unsigned S = (unsigned)(p.sen_off[story + 1] - soff);

        // ---- load this story's compact record ----
        const unsigned char *rec = p.rec + (size_t)(p.rec_by_pos ? pos : w) * p.rec_stride;
        const unsigned *hdr = reinterpret_cast<const unsigned *>(rec);
        const unsigned n_ent = hdr[0], flags = hdr[1], ans_idx = hdr[2], heap_off = hdr[3], n_exc = hdr[4];
        if (flags & FLAG_ERROR) {
            if (lane == 0) { atomicExch(p.err_flag, 1u); if (p.pred) p.pred[story] = ANS_NONE; }
            continue;
        }
        {
            const unsigned short *rend_g = reinterpret_cast<const unsigned short *>(rec + p.off_rend);
            for (unsigned r = lane; r < S + 1; r += 32) rend_s[r] = rend_g[r];
        }
        st.in_smem = !(flags & FLAG_HEAP);
        st.n_exc = st.in_smem ? n_exc : 0u;
        st.ent_g = p.heap + heap_off;
        if (st.in_smem) {
            if (n_ent > p.LW) {          // unreachable: k_compact sends such stories to the heap (lcap == LW)
                if (lane == 0) { atomicExch(p.err_flag, 1u); if (p.pred) p.pred[story] = ANS_NONE; }
                continue;
            }
            const unsigned *ent_g = reinterpret_cast<const unsigned *>(rec + p.off_ent);
            for (unsigned k = lane; k < n_ent; k += 32) ent_s[k] = ent_g[k];
            if (n_exc) {
                const uint2 *exc_g = reinterpret_cast<const uint2 *>(rec + p.off_exc);
                if (lane < n_exc) excs[lane] = exc_g[lane];
            }
        }
        __syncwarp();

        int acc[16];
        // ---- question embedding: u0 = Q_w0( sum_j Q_w0(Q_w0(B[i][j]) * Q_w0(q[j])) )   MemN2N.c:826, layer_cuda.cu:49 ----
        embed_rows<LPR>(p, st, lane, p.offB, (g == 0) ? 0 : -1, p.iw[0], p.fw[0], p.lw[0], acc, sel);
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
            const int mb = (1 << fb) - 1;

            // u operand: Q_bin(u) for the scorer (mode 2) and the linear map, as int32  MemN2N.c:847,873
            for (unsigned j = lane; j < DP; j += 32) ub32[j] = (j < d) ? qi_requant((int)uvec[j], fu, lb, fb) : 0;
            __syncwarp();
            int ub[16];
            unsigned au[16];
            unsigned su_bits = 0;
            if (MODE == 3) {
                const uint4 t = *reinterpret_cast<const uint4 *>(uvec + 16 * q);
                const unsigned tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    // the Hamming scorer quantises BOTH operands with the addressing format at
                    // 31-iwl fractional bits (layer.c:215-233, layer_cuda.cu:2515)
                    unsigned s_, m_;
                    appx_encode(sbyte(tw[j >> 2], j & 3), fu, p.ia[h], s_, m_);
                    au[j] = m_;
                    su_bits |= (s_ >> 31) << j;
                    ub[j] = 0;
                }
            } else {
#pragma unroll
                for (int w4 = 0; w4 < 4; w4++) {
                    const int4 t = *reinterpret_cast<const int4 *>(ub32 + 16 * q + 4 * w4);
                    ub[4 * w4 + 0] = t.x; ub[4 * w4 + 1] = t.y; ub[4 * w4 + 2] = t.z; ub[4 * w4 + 3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < 16; j++) au[j] = 0;
            }
            const int ka = fa - fw;
            // the single-clamp shortcuts of score_part need these (true for every reference config)
            const bool simple = (ka == 0) ? (la == lw) : (ka > 0 ? (la <= (lw << ka)) : true);

            // ---- memory embedding + addressing, G rows per pass ----
            for (unsigned r0 = 0; r0 < S; r0 += G) {
                const unsigned r = r0 + g;
                embed_rows<LPR>(p, st, lane, p.offA[h], (r < S) ? (int)(r + 1) : -1, iw, fw, lw, acc, sel);
                int part = 0;
                if (MODE == 3) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int m_w = qi_clamp(acc[j], lw);                   // M_h[r][t], weight format
                        unsigned sm, am;
                        appx_encode(m_w, fw, p.ia[h], sm, am);
                        const unsigned sv = ((su_bits >> j) & 1u) << 31;
                        // dims >= d hold zero codes on both sides (all 7 bits match, e = +127): masked out
                        part += (16u * q + j < d) ? appx_element_x128(sm, am, sv, au[j]) : 0;
                    }
                } else if (!simple) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int m_att = qi_requant(qi_clamp(acc[j], lw), fw, la, fa);
                        part += qi_mul(m_att, ub[j], la, fb);
                    }
                } else {
                    if (ka == 0) part = score_part<0>(acc, ub, lw, la, ka, fb, mb);
                    else if (ka > 0) part = score_part<1>(acc, ub, lw, la, ka, fb, mb);
                    else part = score_part<-1>(acc, ub, lw, la, ka, fb, mb);
                    part -= 16 * la;                                            // remove the clamp bias of the 16 terms
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
                if (DEBUG && p.dbg.dev_s) p.dbg.dev_s[(size_t)h * p.sum_sen + soff + r] = sv;      // the scorer's output (before the scale layer)
                if (p.en_sc_att) sv = __fmul_rn(sv, p.sc_w[h]);     // scale layer (EN_SC_ATT): out = in * w, fp32, not quantised (lib/layer_cuda.cu:4805-4822)
                ex[r] = sv;
                mx = fmaxf(mx, sv);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
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
                    if (DEBUG && p.dbg.dev_pcode) p.dbg.dev_pcode[(size_t)h * p.sum_sen + soff + r] = (unsigned char)code;
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
                    embed_rows<LPR>(p, st, lane, p.offC[h], (r < S) ? (int)(r + 1) : -1, iw, fw, lw, acc, sel);
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
                embed_rows<LPR>(p, st, lane, p.offC[h], (r >= 0) ? r + 1 : -1, iw, fw, lw, acc, sel);
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
            const int lw2 = 2 * lw;
            for (unsigned i0 = 0; i0 < d; i0 += 32) {
                const unsigned i = i0 + lane;
                int a_f = 0, g_w = 0;
                if (i < d) {
                    if (p.lin_map) {
                        const unsigned hrow = p.offH[h] + i * p.HS;
                        int s_ = 0;
                        const unsigned d4 = (d + 3) / 4;
                        for (unsigned j4 = 0; j4 < d4; j4++) {
                            const unsigned hw = *reinterpret_cast<const unsigned *>(smem + hrow + 4u * j4);
                            const int4 uu = *reinterpret_cast<const int4 *>(ub32 + 4 * j4);
                            s_ += clamp_biased(shr0m(sbyte_prmt<0>(hw) * uu.x, fb, mb), lw, lw2);
                            s_ += clamp_biased(shr0m(sbyte_prmt<1>(hw) * uu.y, fb, mb), lw, lw2);
                            s_ += clamp_biased(shr0m(sbyte_prmt<2>(hw) * uu.z, fb, mb), lw, lw2);
                            s_ += clamp_biased(shr0m(((int)hw >> 24) * uu.w, fb, mb), lw, lw2);
                        }
                        g_w = qi_clamp(s_ - (int)(4u * d4) * lw, lw);          // remove the clamp bias of the 4*d4 terms
                        a_f = qi_requant(g_w, fw, lf, ff);
                    } else {
                        g_w = (int)uvec[i];
                        a_f = qi_requant(g_w, fu, lf, ff);
                    }
                }
                __syncwarp();
                if (i < d) {
                    int un = qi_clamp(a_f + (int)ovec[i], lf);
                    if (p.en_non_lin) un = max(un, 0);              // activation layer (EN_NON_LINEARITY): Q_f(RELU(u')), on the grid already (lib/layer_cuda.cu:4548)
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
        const unsigned d4 = (d + 3) / 4;
        float zmax = -INFINITY;
        for (unsigned i0 = 0; i0 < V; i0 += 128) {
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            unsigned wrow[4];
#pragma unroll
            for (int k = 0; k < 4; k++) wrow[k] = p.offW + min(i0 + 32 * k + lane, V - 1) * (p.WS * 4u);
            for (unsigned j4 = 0; j4 < d4; j4++) {
                const float4 uu = *reinterpret_cast<const float4 *>(ufl + 4 * j4);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float4 ww = *reinterpret_cast<const float4 *>(smem + wrow[k] + 16u * j4);
                    z[k] = __fadd_rn(z[k], __fmul_rn(ww.x, uu.x));
                    z[k] = __fadd_rn(z[k], __fmul_rn(ww.y, uu.y));
                    z[k] = __fadd_rn(z[k], __fmul_rn(ww.z, uu.z));
                    z[k] = __fadd_rn(z[k], __fmul_rn(ww.w, uu.w));
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned i = i0 + 32 * k + lane;
                if (i < V) {
                    zbuf[i] = z[k];
                    zmax = fmaxf(zmax, z[k]);
                    if (DEBUG && p.dbg.dev_z) p.dbg.dev_z[(size_t)story * V + i] = z[k];
                    if (DEBUG && p.dbg.dev_cand) p.dbg.dev_cand[(size_t)story * V + i] = 2;
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
            if (DEBUG && p.dbg.dev_path) p.dbg.dev_path[story] = 3;
        }
        __syncwarp();
    }
}

}  // namespace

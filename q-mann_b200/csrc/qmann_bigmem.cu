// qmann_bigmem.cu -- memory addressing over ONE very large pre-embedded memory (BASELINE config 5:
// >= 1M slots, d = 256, 3 hops), slot-sharded across GPUs (include/qmann_abi.h part 3).
//
// The reference cannot run this shape (its <<<d, S>>> launches and temp[1024] cap every dimension at
// 1024, lib/layer_cuda.cu:547-559, lib/layer_cuda.h:11); the arithmetic is the reference's, applied to
// a memory whose rows M_h[r][:], C_h[r][:] are given as the int8 codes emb_m[h]/emb_c[h] would output
// (MemN2N/MemN2N.c:835-838), for Q queries at once.
//
// Per hop, per rank (this rank owns S_local contiguous slots):
//   k_big_scores   streams M_h once (HBM), lane-per-slot, QB queries register-blocked per pass:
//                  s[r] = Q_att(sum_t Q_att(Q_att(M[r][t]) * Q_bin(u[t])))        layer_cuda.cu:105-141
//                  or the Hamming/approximate similarity                          layer_cuda.cu:355-541
//                  -> score bins [Q][S_local] (bytes in mode 2, 16-bit in mode 3)
//                  fast forms of the mode-2 scorer: k_big_scores_fast (Q < 4, HBM-bound), k_big_scores_tq / _tc (tcgen05.mma
//                  kind::i8, query planes in tensor memory / shared memory), k_big_scores_mma (mma.sync)
//   k_big_hist_lanes / k_big_hist   per-query histogram of the score bins (lane-private counters / shared-memory atomics)
//        ---- all-reduce(SUM, u32) of the histograms across ranks (NCCL, by the caller) ----
//   k_big_softmax  every rank rebuilds the SAME max and double-precision total from the global
//                  histogram in a fixed order (bins ascending, 256 contiguous ranges, range partials
//                  added ascending), p_bin = fl32(__expf(v_bin - max) / total), Q_f(p_bin)
//                                                                                layer_cuda.cu:1969-2060, :561
//   k_big_read     scans the score bins; only slots with Q_f(p) != 0 (at most 2^frac of them, because
//                  sum p = 1) contribute Q_f(Q_f(p) * Q_f(C[r][c])) to the read: a sparse gather of C
//                  rows into an int32 partial read                                layer_cuda.cu:547-579
//        ---- all-reduce(SUM, i32) of the partial reads [Q][d] ----
//   k_big_update   o = Q_f(sum), g = linear map, u' = Q_f(Q_f(g) + Q_f(o))        layer_cuda.cu:49-68, 1535
// and after the last hop k_big_answer (fp32 projection in index order, softmax, argmax_last).
//
// A log-sum-exp merge of locally normalised reads would NOT be exact: Q_f(p) truncates, so every
// shard must quantise with the global max and total (hard part 3).  The integer
// histogram makes that a single exact collective and the result independent of the number of shards.
#include "qmann_fixed.cuh"
#include "qmann_common.h"
#include "qmann_tc.cuh"
#include "../../include/qmann_abi.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

using namespace qmann;

namespace {

thread_local std::string g_berr;
int bfail(int code, const std::string &msg) { g_berr = msg; return code; }
#define BCUDA(expr)                                                                                \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) return bfail(QMANN_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

constexpr int MAXH = QMANN_MAX_HOP;
constexpr unsigned SOFTMAX_RANGES = 256;       // fixed summation order of k_big_softmax

struct HopFmt {
    int fw, lw, fa, la, ia, ff, lf, iff, fb, lb;
};

// -------------------------------------------------------------------------------------------------
// query preparation: u (int8 codes, fu fractional bits) -> operands of the scorer
//   mode 2: ub[q][t] = Q_bin(u) as int32                                          MemN2N.c:847
//   mode 3: av[q][t] = 31-bit magnitude, sv[q][t] = sign word                      layer.c:215-233
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_big_prep_query(const signed char *__restrict__ u, int fu, unsigned Q, unsigned d, int mode, HopFmt f,
                                                        int *__restrict__ ub, unsigned *__restrict__ av, unsigned *__restrict__ sv,
                                                        signed char *__restrict__ ub8, unsigned *__restrict__ umax, int *__restrict__ u9,
                                                        unsigned *__restrict__ quirk, int sh_u)
{
    // one CTA per query; ub8/umax feed k_big_scores_fast (|Q_bin(u)| <= lb <= 127 fits a byte); mode 3 also writes the nine-bit
    // operand U9 = sat9(code << sh_u) of the fast Hamming form (see k_big_scores) and flags a query that holds the -2^iwl value
    __shared__ unsigned s_max;
    const unsigned q = blockIdx.x;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    unsigned mx = 0;
    int qk = 0;
    for (unsigned t = threadIdx.x; t < d; t += blockDim.x) {
        const size_t i = (size_t)q * d + t;
        const int c = (int)u[i];
        const int b = qi_requant(c, fu, f.lb, f.fb);
        ub[i] = b;
        ub8[i] = (signed char)b;
        mx = max(mx, (unsigned)abs(b));
        if (mode == 3) {
            unsigned s_, m_;
            appx_encode(c, fu, f.ia, s_, m_);
            av[i] = m_;
            sv[i] = s_;
            if (sh_u >= 0) {
                const int tu = c << sh_u;
                qk |= (tu == -256) ? 1 : 0;
                u9[i] = max(-255, min(tu, 255));
            }
        }
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, mx);
    const int any_qk = __syncthreads_or(qk);
    if (threadIdx.x == 0) {
        umax[q] = s_max;
        if (mode == 3) quirk[q] = any_qk ? 1u : 0u;
    }
}

// -------------------------------------------------------------------------------------------------
// k_big_scores: lane-per-slot.  A warp stages 32 memory rows (32*d bytes, coalesced 128-bit loads)
// into its shared-memory tile with a row stride of d+16 bytes (conflict-free 128-bit lane-per-row
// reads), then every lane walks its own row in 16-dim chunks against QB queries whose operands are
// broadcast from shared memory.
// -------------------------------------------------------------------------------------------------
struct ScoreParams {
    const signed char *M;           // [S_local][d] codes of this hop, weight format
    unsigned long long S_local;
    unsigned d, Q;
    HopFmt f;
    int const_scale;
    const int *ub;                  // [Q][d]
    const unsigned *av, *sv;        // [Q][d] (mode 3)
    const int *u9;                  // [Q][d] nine-bit operands (mode 3, fast form)
    const unsigned *quirk;          // [Q] 1: the query holds the -2^iwl value (literal path for its block)
    int nine, sh_m;                 // fast form usable (both shifts in 0..8), memory-side shift
    int only_quirk;                 // 1: this launch serves only the query blocks k_big_scores_ham declined (a query holds -2^iwl)
    void *bins;                     // [Q][S_local] score bin = code + bias (uint8 in mode 2, uint16 in mode 3)
    int bin8;
    unsigned bias;                  // la (mode 2) or 127*d (mode 3)
};

// score bins are bytes when every bin index fits one (mode 2: 2*127+1 bins), 16-bit otherwise
__device__ __forceinline__ void store_bin(void *bins, int bin8, size_t idx, unsigned v)
{
    if (bin8) reinterpret_cast<unsigned char *>(bins)[idx] = (unsigned char)v;
    else reinterpret_cast<unsigned short *>(bins)[idx] = (unsigned short)v;
}

__device__ __forceinline__ int sx8(unsigned w, int b) { return (int)(signed char)((w >> (8 * b)) & 0xFFu); }

template <int MODE, int QB>
__global__ void __launch_bounds__(256) k_big_scores(const ScoreParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned d = p.d, rs = d + 16;                    // tile row stride (bytes)
    const unsigned q0 = blockIdx.y * QB;
    const unsigned nq = min((unsigned)QB, p.Q - q0);
    // query operands of this block of queries: [QB][d] int32 (+ [QB][d] magnitudes and signs in mode 3)
    int *ubs = reinterpret_cast<int *>(sm);
    unsigned *avs = reinterpret_cast<unsigned *>(sm + (size_t)QB * d * 4);
    unsigned *svs = avs + (size_t)QB * d;
    const size_t qbytes = (size_t)QB * d * 4 * (MODE == 3 ? 3 : 1);
    unsigned char *tile = sm + qbytes + (size_t)wid * 32 * rs;
    // Mode 3 in nine bits (the form of the batched forward, qmann_fast.cuh): the reference compares 31-bit magnitudes
    // |x| * 2^(31-iwl) (layer_cuda.cu:384-428) and keeps bits 30..24 of their difference or sum; every magnitude is a multiple of
    // 2^23 or the saturated 0x7FFFFFFF, so the element is exact on A = sat9(code << sh) in [-255, 255]:  w = |A_m - A_u|,
    // e = 127 - ((w >> 1) & 127), negative iff the signs differ and w < 256.  ~9 instructions per (slot, query, dim) instead of ~35.
    // A query holding the -2^iwl value sends its block down the literal path.
    bool nine = false;
    if (MODE == 3 && p.nine) {
        int qk = 0;
        for (unsigned q = threadIdx.x; q < nq; q += blockDim.x) qk |= (int)p.quirk[q0 + q];
        nine = !__syncthreads_or(qk);
        if (p.only_quirk && nine) return;                  // k_big_scores_ham has this block
    }
    const int sh_m = p.sh_m;
    for (unsigned i = threadIdx.x; i < QB * d; i += blockDim.x) {
        const unsigned q = i / d, t = i % d;
        const bool ok = q < nq;
        ubs[i] = ok ? ((MODE == 3 && nine) ? p.u9[(size_t)(q0 + q) * d + t] : p.ub[(size_t)(q0 + q) * d + t]) : 0;
        if (MODE == 3) {
            avs[i] = ok ? p.av[(size_t)(q0 + q) * d + t] : 0u;
            svs[i] = ok ? p.sv[(size_t)(q0 + q) * d + t] : 0u;
        }
    }
    __syncthreads();

    const HopFmt f = p.f;
    const int mb = (1 << f.fb) - 1;
    const unsigned c16 = d / 16;                            // 16-byte chunks per row
    const unsigned long long n_tiles = (p.S_local + 31) / 32;
    for (unsigned long long tix = (unsigned long long)blockIdx.x * nw + wid; tix < n_tiles; tix += (unsigned long long)gridDim.x * nw) {
        const unsigned long long slot0 = tix * 32;
        // ---- stage 32 rows ----
        const unsigned total16 = 32 * c16;
#pragma unroll 4
        for (unsigned i = lane; i < total16; i += 32) {
            const unsigned r = i / c16, c = i % c16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (slot0 + r < p.S_local) v = __ldg(reinterpret_cast<const uint4 *>(p.M + (slot0 + r) * d) + c);
            *reinterpret_cast<uint4 *>(tile + r * rs + c * 16) = v;
        }
        __syncwarp();
        // ---- lane-per-slot scoring ----
        int acc[QB];
#pragma unroll
        for (int q = 0; q < QB; q++) acc[q] = 0;
        const unsigned char *row = tile + lane * rs;
        for (unsigned c = 0; c < c16; c++) {
            const uint4 w4 = *reinterpret_cast<const uint4 *>(row + c * 16);
            const unsigned ww[4] = {w4.x, w4.y, w4.z, w4.w};
            int m[16];
            unsigned am[16], smb = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int code = sx8(ww[j >> 2], j & 3);
                if (MODE == 3 && nine) {
                    int t = code << sh_m;
                    t = (t == -256) ? 0 : max(-255, min(t, 255));          // -2^iwl encodes to 0 (it keeps its sign: smb); a caller's -128 too
                    m[j] = t;
                    am[j] = 0;
                    smb |= (code < 0 ? 1u : 0u) << j;
                } else if (MODE == 3) {
                    unsigned s_, m_;
                    appx_encode(code, f.fw, f.ia, s_, m_);
                    am[j] = m_;
                    smb |= (s_ >> 31) << j;
                    m[j] = 0;
                } else {
                    m[j] = qi_requant(code, f.fw, f.la, f.fa);         // Q_att(M)
                    am[j] = 0;
                }
            }
#pragma unroll
            for (int q = 0; q < QB; q++) {
                int part = 0;
                if (MODE == 3 && nine) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; j4++) {
                        const int4 u4 = *reinterpret_cast<const int4 *>(ubs + (size_t)q * d + c * 16 + j4 * 4);
                        const int uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int j = j4 * 4 + k;
                            const int w9 = (int)__sad(m[j], uu[k], 0u);
                            const int e = ~(w9 >> 1) & 0x7F;
                            const bool neg = (((int)(((smb >> j) & 1u) << 31) ^ uu[k]) < 0) && (w9 < 256);
                            part += neg ? -e : e;
                        }
                    }
                } else if (MODE == 3) {
#pragma unroll
                    for (int j4 = 0; j4 < 4; j4++) {
                        const uint4 a4 = *reinterpret_cast<const uint4 *>(avs + (size_t)q * d + c * 16 + j4 * 4);
                        const uint4 s4 = *reinterpret_cast<const uint4 *>(svs + (size_t)q * d + c * 16 + j4 * 4);
                        const unsigned aa[4] = {a4.x, a4.y, a4.z, a4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int j = j4 * 4 + k;
                            part += appx_element_x128(((smb >> j) & 1u) << 31, am[j], ss[k], aa[k]);
                        }
                    }
                } else {
#pragma unroll
                    for (int j4 = 0; j4 < 4; j4++) {
                        const int4 u4 = *reinterpret_cast<const int4 *>(ubs + (size_t)q * d + c * 16 + j4 * 4);
                        const int uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int x = m[j4 * 4 + k] * uu[k];
                            const int t = (int)((unsigned)x >> 31) * mb + x >> f.fb;      // trunc toward zero
                            part += __viaddmin_s32_relu(t, f.la, 2 * f.la);                 // clamp + la
                        }
                    }
                    part -= 16 * f.la;
                }
                acc[q] += part;
            }
        }
        __syncwarp();
        if (slot0 + lane < p.S_local) {
#pragma unroll
            for (int q = 0; q < QB; q++) {
                if ((unsigned)q < nq) {
                    // mode 2: Q_att of the sum; mode 3: the raw sum of e*128 (saturation is applied where the
                    // value is formed, in k_big_softmax)
                    const int code = (MODE == 3) ? acc[q] : qi_clamp(acc[q], f.la);
                    store_bin(p.bins, p.bin8, (size_t)(q0 + q) * p.S_local + slot0 + lane, (unsigned)(code + (int)p.bias));
                }
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// k_big_scores_ham (mode 3, d % 16 == 0): the approximate Hamming scorer on four packed bytes per instruction.
//
// The nine-bit element (see k_big_scores) only needs the magnitudes a = |A_m|, b = |A_u| in 0..255 and the two signs:
//   signs agree : e + 127 = 254 - (|a - b| >> 1)                                  one VABSDIFF4 (native on sm_100a) + 3 logic ops
//   signs differ: m = (a + b) >> 1 (halving add, a byte);  e + 127 = m  for m < 128,  382 - m  otherwise
// every step a per-byte operation without carries between bytes; the four biased elements of a word are added with one dp4a.
// ~17 integer instructions per four (slot, query, dim) elements instead of ~9 per element (nine-bit scalar form) or ~35 (literal).
// Checked on every pair of 8-bit codes in tests/test_identities.py (CPU) and against the literal kernel on the GPU.
// Lane-per-slot over a staged tile like k_big_scores; the memory word is encoded once per chunk and shared by the QB queries of the
// block, whose encoded operands sit in shared memory.  A block with a query that holds the -2^iwl value returns at once: the
// second launch (k_big_scores<3, QB> with only_quirk) serves it on the literal path.
// -------------------------------------------------------------------------------------------------
struct HamParams {
    const signed char *M;           // [S_local][d] codes of this hop, weight format
    unsigned long long S_local;
    unsigned d, Q;
    int sh_m, sh_u;                 // A = sat9(code << sh)
    const signed char *u8;          // [Q][d] query codes
    const unsigned *quirk;          // [Q]
    void *bins;
    int bin8;
    unsigned bias;
};

// four codes -> magnitudes a = sat8(|code| << sh) (the value whose shift is exactly -256 encodes to 0) and sign masks (0xFF: negative)
__device__ __forceinline__ void ham_encode(unsigned cw, int sh, unsigned &a, unsigned &sg)
{
    asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(sg) : "r"(cw));
    const unsigned mag = (cw ^ sg) + (sg & 0x01010101u);                      // |code| <= 128 per byte
    if (sh == 0) { a = mag; return; }
    const unsigned thr = 256u >> sh;                                          // |code| >= thr saturates
    const unsigned ov = (mag + (128u - thr) * 0x01010101u) & 0x80808080u;
    a = ((mag << sh) & (((0xFFu << sh) & 0xFFu) * 0x01010101u)) | ((ov >> 7) * 0xFFu);
    const unsigned z = mag ^ (thr * 0x01010101u);
    const unsigned nz = (((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & 0x80808080u;  // bit 7 of the bytes with |code| != thr
    const unsigned eqm = ((~nz & 0x80808080u) >> 7) * 0xFFu;
    a &= ~(eqm & sg);                                                          // code << sh == -256: magnitude 0, the sign stays
}

// e + 127 of four elements
__device__ __forceinline__ unsigned ham_word(unsigned a, unsigned sa, unsigned b, unsigned sb)
{
    const unsigned sd = sa ^ sb;
    const unsigned xs = 0xFEFEFEFEu - ((__vabsdiffu4(a, b) >> 1) & 0x7F7F7F7Fu);
    const unsigned m = (a & b) + (((a ^ b) & 0xFEFEFEFEu) >> 1);
    unsigned mk;
    asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(mk) : "r"(m));                  // 0xFF in the bytes with m >= 128
    const unsigned xd = (m ^ mk) + (mk & 0x7F7F7F7Fu);
    return (xd & sd) | (xs & ~sd);
}

template <int QB>
__global__ void __launch_bounds__(256) k_big_scores_ham(const HamParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned d = p.d, rs = d + 16;                    // tile row stride (bytes)
    const unsigned q0 = blockIdx.y * QB;
    const unsigned nq = min((unsigned)QB, p.Q - q0);
    {
        int qk = 0;
        for (unsigned q = threadIdx.x; q < nq; q += blockDim.x) qk |= (int)p.quirk[q0 + q];
        if (__syncthreads_or(qk)) return;                   // the literal kernel takes this block
    }
    unsigned *qa = reinterpret_cast<unsigned *>(sm);        // [QB][d / 4] magnitudes
    unsigned *qs = qa + (size_t)QB * d / 4;                 // [QB][d / 4] sign masks
    unsigned char *tile = sm + (size_t)QB * d * 2 + (size_t)wid * 32 * rs;
    for (unsigned i = threadIdx.x; i < QB * d / 4; i += blockDim.x) {
        const unsigned q = i / (d / 4), wi = i % (d / 4);
        unsigned a = 0, sg = 0;
        if (q < nq) ham_encode(*reinterpret_cast<const unsigned *>(p.u8 + (size_t)(q0 + q) * d + 4u * wi), p.sh_u, a, sg);
        qa[i] = a;
        qs[i] = sg;
    }
    __syncthreads();
    const unsigned c16 = d / 16;
    const unsigned long long n_tiles = (p.S_local + 31) / 32;
    for (unsigned long long tix = (unsigned long long)blockIdx.x * nw + wid; tix < n_tiles; tix += (unsigned long long)gridDim.x * nw) {
        const unsigned long long slot0 = tix * 32;
        const unsigned total16 = 32 * c16;
#pragma unroll 4
        for (unsigned i = lane; i < total16; i += 32) {
            const unsigned r = i / c16, c = i % c16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (slot0 + r < p.S_local) v = __ldg(reinterpret_cast<const uint4 *>(p.M + (slot0 + r) * d) + c);
            *reinterpret_cast<uint4 *>(tile + r * rs + c * 16) = v;
        }
        __syncwarp();
        unsigned acc[QB];
#pragma unroll
        for (int q = 0; q < QB; q++) acc[q] = 0;
        const unsigned char *row = tile + lane * rs;
        for (unsigned c = 0; c < c16; c++) {
            const uint4 w4 = *reinterpret_cast<const uint4 *>(row + c * 16);
            const unsigned ww[4] = {w4.x, w4.y, w4.z, w4.w};
            unsigned a[4], sa[4];
#pragma unroll
            for (int k = 0; k < 4; k++) ham_encode(ww[k], p.sh_m, a[k], sa[k]);
#pragma unroll
            for (int q = 0; q < QB; q++) {
                const uint4 b4 = *reinterpret_cast<const uint4 *>(qa + (size_t)q * (d / 4) + c * 4);
                const uint4 s4 = *reinterpret_cast<const uint4 *>(qs + (size_t)q * (d / 4) + c * 4);
                unsigned t = acc[q];
                t = __dp4a(ham_word(a[0], sa[0], b4.x, s4.x), 0x01010101u, t);
                t = __dp4a(ham_word(a[1], sa[1], b4.y, s4.y), 0x01010101u, t);
                t = __dp4a(ham_word(a[2], sa[2], b4.z, s4.z), 0x01010101u, t);
                t = __dp4a(ham_word(a[3], sa[3], b4.w, s4.w), 0x01010101u, t);
                acc[q] = t;
            }
        }
        __syncwarp();
        if (slot0 + lane < p.S_local) {
#pragma unroll
            for (int q = 0; q < QB; q++)
                if ((unsigned)q < nq)
                    store_bin(p.bins, p.bin8, (size_t)(q0 + q) * p.S_local + slot0 + lane, (unsigned)((int)acc[q] - 127 * (int)d + (int)p.bias));
        }
    }
}

// -------------------------------------------------------------------------------------------------
// k_big_scores_fast (mode 2, frac_bin == 2, d/16 a power of two): the HBM-bound scorer.
//
// Inputs prepared once per memory by k_big_prep_mem: Y = Q_att(M) as int8 (M itself when the re-quantisation is the
// identity) and rowmax[r] = max_t |Y[r][t]|.  With x_t = y_t * u_t the reference score is
//   s = Q_att( sum_t Q_att( trunc0(x_t / 4) ) )                                  layer_cuda.cu:105-141
// and, for a row where no product saturates (rowmax * max|u| <= 4 la + 3),
//   4 * sum_t trunc0(x_t / 4) = sum_t x_t - sum_t (x_t mod 4) + 4 * #{t : x_t < 0, x_t mod 4 != 0}
// (floor-mod; trunc0 = floor + 1 exactly for negative non-multiples).  sum_t x_t is one dp4a per four dims.
// x mod 4 depends on the two low bits of y and u only: bit0 = y0 & u0, bit1 = (y1 & u0) ^ (y0 & u1), evaluated on
// four packed bytes at once; the sign test is the XOR of the two sign bits.  Per byte the kernel forms
//   c = (x mod 4) + 3 - 4 [x < 0 and x mod 4 != 0]   in 0..6
// by clearing bit 2 of (x mod 4) + 3 (a value in 4..6 exactly when x mod 4 != 0), so that
//   4 * sum = sum_t x_t - sum_t c_t + 3 d.
// About 10 integer instructions per four products instead of ~28, which moves the kernel from the issue limit to
// the HBM limit.  Rows that may saturate are recomputed product by product in the same launch.
//
// Mapping: LPR = d/16 lanes per row, 32/LPR rows per 512-byte warp load (coalesced 128-bit loads straight from
// global memory, four in flight per lane, no shared-memory staging); the query operands of QB queries stay in
// registers for the whole kernel (a lane always sees the same 16 dims).  A warp owns 32 consecutive rows per tile
// and writes their 32 score bins as one 64-byte segment.
// -------------------------------------------------------------------------------------------------
struct FastScoreParams {
    const signed char *Y;           // [S_local][d] Q_att(M) codes
    const unsigned char *rowmax;    // [S_local]
    unsigned long long S_local;
    unsigned d, Q;
    int la, fb;
    const signed char *ub8;         // [Q][d] Q_bin(u)
    const unsigned *umax;           // [Q] max_t |Q_bin(u)|
    void *bins;                     // [Q][S_local], uint8 (mode 2: 255 bins)
    int bin8;
    unsigned bias;
};

__device__ __forceinline__ uint4 ldg_stream(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int LPR, int QB>
__global__ void __launch_bounds__(256) k_big_scores_fast(const FastScoreParams p)
{
    constexpr int RPL = 32 / LPR;                 // rows per warp load
    constexpr int UN = (LPR < 4) ? LPR : 4;       // loads in flight per lane
    const unsigned lane = threadIdx.x & 31;
    const unsigned sub = lane % LPR, rsub = lane / LPR;
    const unsigned q0 = blockIdx.y * QB;
    const unsigned d = p.d;
    const int la = p.la;
    const unsigned sat_lim = 4u * (unsigned)la + 3u;

    unsigned Uw[QB][4], U0[QB][4], U0s[QB][4], U1[QB][4], Us4[QB][4], umax[QB];
#pragma unroll
    for (int qi = 0; qi < QB; qi++) {
        const bool ok = q0 + qi < p.Q;
        uint4 t = make_uint4(0u, 0u, 0u, 0u);
        if (ok) t = *reinterpret_cast<const uint4 *>(p.ub8 + (size_t)(q0 + qi) * d + 16u * sub);
        umax[qi] = ok ? p.umax[q0 + qi] : 0u;
        const unsigned tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int w = 0; w < 4; w++) {
            Uw[qi][w] = tw[w];
            U0[qi][w] = tw[w] & 0x01010101u;
            U0s[qi][w] = (tw[w] & 0x01010101u) << 1;
            U1[qi][w] = tw[w] & 0x02020202u;
            Us4[qi][w] = (tw[w] >> 5) & 0x04040404u;
        }
    }

    const unsigned long long n_tiles = (p.S_local + 31) / 32;
    const unsigned long long wglobal = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long wstride = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long tix = wglobal; tix < n_tiles; tix += wstride) {
        const unsigned long long row0 = tix * 32;
        const unsigned rmv = (row0 + lane < p.S_local) ? (unsigned)p.rowmax[row0 + lane] : 0u;
        int keep[QB];
#pragma unroll
        for (int qi = 0; qi < QB; qi++) keep[qi] = 0;
        const bool full = row0 + 32 <= p.S_local;
        const signed char *tile_ptr = p.Y + (row0 + rsub) * d + 16u * sub;
#pragma unroll 1
        for (unsigned it = 0; it < (unsigned)LPR; it += UN) {
            uint4 y[UN];
            if (full) {
                // whole tile inside the shard: one 64-bit base per tile, 32-bit offsets per load, no bounds checks
#pragma unroll
                for (int k = 0; k < UN; k++) y[k] = ldg_stream(tile_ptr + (it + k) * (RPL * d));
            } else {
#pragma unroll
                for (int k = 0; k < UN; k++) {
                    const unsigned long long row = row0 + (unsigned long long)(it + k) * RPL + rsub;
                    y[k] = make_uint4(0u, 0u, 0u, 0u);
                    if (row < p.S_local) y[k] = ldg_stream(p.Y + row * d + 16u * sub);
                }
            }
#pragma unroll
            for (int k = 0; k < UN; k++) {
                const unsigned yw[4] = {y[k].x, y[k].y, y[k].z, y[k].w};
                unsigned ys[4], ys4[4];
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    ys[w] = yw[w] << 1;
                    ys4[w] = (yw[w] >> 5) & 0x04040404u;
                }
                const unsigned rm = __shfl_sync(0xffffffffu, rmv, (it + k) * RPL + rsub);
#pragma unroll
                for (int qi = 0; qi < QB; qi++) {
                    int D = 0;
                    unsigned cs = 0;
#pragma unroll
                    for (int w = 0; w < 4; w++) {
                        D = __dp4a((int)yw[w], (int)Uw[qi][w], D);
                        const unsigned t0 = ys[w] & U1[qi][w];
                        const unsigned t1 = (yw[w] & U0s[qi][w]) ^ t0;
                        const unsigned bm = (yw[w] & U0[qi][w]) | t1;            // x mod 4 per byte
                        const unsigned wv = bm + 0x03030303u;                    // 3..6, bit 2 set iff x mod 4 != 0
                        cs += wv & ~(ys4[w] ^ Us4[qi][w]);                       // clear bit 2 where the product is negative
                    }
                    int part = D - (int)__dp4a(cs, 0x01010101u, 0u) + 48;
#pragma unroll
                    for (int o = 1; o < LPR; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    int tot = part >> 2;                                         // exact: the numerator is a multiple of 4
                    const bool risky = rm * umax[qi] > sat_lim;
                    if (__any_sync(0xffffffffu, risky)) {
                        // some product of this row may saturate: the reference order of operations, product by product
                        int sp = 0;
#pragma unroll
                        for (int w = 0; w < 4; w++)
#pragma unroll
                            for (int b = 0; b < 4; b++) {
                                const int yy = (int)(signed char)((yw[w] >> (8 * b)) & 0xFFu);
                                const int uu = (int)(signed char)((Uw[qi][w] >> (8 * b)) & 0xFFu);
                                sp += qi_mul(yy, uu, la, p.fb);
                            }
#pragma unroll
                        for (int o = 1; o < LPR; o <<= 1) sp += __shfl_xor_sync(0xffffffffu, sp, o);
                        if (risky) tot = sp;
                    }
                    if (sub == it + k) keep[qi] = qi_clamp(tot, la);
                }
            }
        }
        const unsigned long long row = row0 + sub * RPL + rsub;
        if (row < p.S_local) {
#pragma unroll
            for (int qi = 0; qi < QB; qi++)
                if (q0 + qi < p.Q) store_bin(p.bins, p.bin8, (size_t)(q0 + qi) * p.S_local + row, (unsigned)(keep[qi] + (int)p.bias));
        }
    }
}

// -------------------------------------------------------------------------------------------------
// k_big_scores_mma (mode 2, frac_bin == 2, d % 64 == 0, Q >= 4): the scorer as four int8 tensor-core contractions.
//
// With trunc0(x/4) = (x - r)/4, r = sgn(x) * (|x| mod 4) and |x| mod 4 = ((|y| & 3) * (|u| & 3)) & 3 =: rho(a, b):
//   4 * sum_t trunc0(y_t u_t / 4) = sum_t y_t u_t - sum_{a=1..3} sum_t I_a(y_t) R_a(u_t)
//   I_a(y) = sgn(y) [ |y| & 3 == a ]  in {-1, 0, 1},   R_a(u) = sgn(u) rho(a, |u| & 3)  in {-3 .. 3}
// i.e. the [S x d] memory tile times four [d x Q] int8 query planes (u, -R_1, -R_2, -R_3), accumulated exactly in
// int32: a dense contraction of 4 * S * d * Q multiply-adds per hop (69 G at S = 2^20, d = 256, Q = 64), which is
// what the tensor cores are for; the CUDA cores only build the three indicator planes of each memory word
// (~18 integer instructions per four bytes, shared by all Q queries).  Rows that may saturate a product
// (rowmax * max|u| > 4 la + 3) are recomputed product by product.
//
// mma.sync.m16n8k32.s8: a warp owns 16 slots per tile and 64 queries (8 n-tiles, 32 accumulator registers).  The
// contraction index is permuted so that a lane's A words come from ONE 128-bit load per row and 64-byte step
// (lane (g, c) reads bytes 64 kk + 16 c .. +15 of rows g and g + 8); k_big_prep_bfrag stores the query planes in
// the same permutation, already in fragment order, and the kernel copies them linearly into shared memory
// (1 KB per k-step and n-tile: two conflict-free 128-bit loads per lane bring b0/b1 of all four planes).
// -------------------------------------------------------------------------------------------------
constexpr unsigned MMA_QB = 64;                 // queries per block (8 n-tiles)

struct MmaScoreParams {
    const signed char *Y;
    const unsigned char *rowmax;
    unsigned long long S_local;
    unsigned d, Q;
    int la, fb;
    const signed char *ub8;         // [Q][d] (exact recomputation of risky rows)
    const unsigned *umax;           // [Q]
    const uint4 *bfrag;             // [qblocks][d/32][8][2][32] fragment-ordered query planes
    void *bins;
    int bin8;
    unsigned bias;
};

__device__ __forceinline__ void mma_s8(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// the three indicator planes of four packed codes: I_a = sgn(y) [ |y| & 3 == a ]
__device__ __forceinline__ void indicator_planes(unsigned yw, unsigned &i1, unsigned &i2, unsigned &i3)
{
    unsigned fill;
    asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(fill) : "r"(yw));           // 0xFF in the bytes that are negative
    const unsigned a = (((yw ^ fill) & 0x03030303u) + (fill & 0x01010101u)) & 0x03030303u;    // |y| & 3 per byte
    const unsigned a0 = a & 0x01010101u, a1 = (a >> 1) & 0x01010101u;
    const unsigned sg = fill | 0x01010101u;                                // 0xFF (-1) or 0x01 (+1) per byte
    i1 = ((a0 & ~a1) * 0xFFu) & sg;
    i2 = ((a1 & ~a0) * 0xFFu) & sg;
    i3 = ((a0 & a1) * 0xFFu) & sg;
}

// 16-byte asynchronous copy global -> shared (zero-fill when !valid)
__device__ __forceinline__ void cp_async16(unsigned smem_addr, const void *gptr, bool valid)
{
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Shared memory: [query planes, d*256 B][umax, 256 B][per warp: two tile buffers of 16 rows x d bytes].
// A warp streams its tiles through the two buffers with cp.async (the copy of tile i+1 runs under the 256 IMMAs of tile
// i; no staging registers), rows are stored with their 16-byte chunks XOR-swizzled by the row parity (chunk ^ 4 for odd
// rows) so that the quarter-warp reads of two neighbouring rows hit disjoint banks.
__global__ void __launch_bounds__(512, 1) k_big_scores_mma(const MmaScoreParams p)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned g = lane >> 2, c = lane & 3;
    const unsigned d = p.d, KS = d / 32;                   // k-steps
    const unsigned q0 = blockIdx.y * MMA_QB;
    const unsigned frag_vec = KS * 8 * 2 * 32;             // uint4 per query block
    uint4 *bs = reinterpret_cast<uint4 *>(sm);
    unsigned *umax_s = reinterpret_cast<unsigned *>(sm + (size_t)frag_vec * 16);
    unsigned char *tile0 = sm + (size_t)frag_vec * 16 + 256 + (size_t)wid * 2 * 16 * d;
    {
        const uint4 *src = p.bfrag + (size_t)blockIdx.y * frag_vec;
        for (unsigned i = threadIdx.x; i < frag_vec; i += blockDim.x) bs[i] = src[i];
        for (unsigned i = threadIdx.x; i < MMA_QB; i += blockDim.x) umax_s[i] = (q0 + i < p.Q) ? p.umax[q0 + i] : 0u;
    }
    __syncthreads();
    const int la = p.la;
    const unsigned sat_lim = 4u * (unsigned)la + 3u;
    const unsigned long long n_tiles = (p.S_local + 15) / 16;
    const unsigned long long tstride = (unsigned long long)gridDim.x * nw;
    const unsigned c16 = d / 16;                            // 16-byte chunks per row
    const unsigned swz_bit = ((c16 & 7u) == 0u) ? 4u : 0u;  // chunk ^ 4 on odd rows needs rows of whole 8-chunk groups (d % 128 == 0)
    const unsigned tile_smem0 = (unsigned)__cvta_generic_to_shared(tile0);

    // asynchronous copy of tile `tix` into buffer `buf`: 16 rows x c16 chunks, lanes over consecutive chunks (coalesced)
    auto stage = [&](unsigned long long tix, unsigned buf) {
        const unsigned base = tile_smem0 + buf * 16u * d;
        for (unsigned i = lane; i < 16u * c16; i += 32) {
            const unsigned r = i / c16, j = i % c16;
            const unsigned long long row = tix * 16 + r;
            const bool ok = row < p.S_local;
            cp_async16(base + r * d + 16u * (j ^ ((r & 1u) * swz_bit)), p.Y + (ok ? row : 0ull) * d + 16u * j, ok);
        }
        cp_async_commit();
    };

    unsigned long long tix = (unsigned long long)blockIdx.x * nw + wid;
    unsigned buf = 0;
    if (tix < n_tiles) stage(tix, 0);
    for (; tix < n_tiles; tix += tstride, buf ^= 1u) {
        if (tix + tstride < n_tiles) { stage(tix + tstride, buf ^ 1u); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();
        const unsigned long long rowA = tix * 16 + g, rowB = rowA + 8;
        const bool okA = rowA < p.S_local, okB = rowB < p.S_local;
        int acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[nt][j] = 0;
        const unsigned char *ta = tile0 + (size_t)buf * 16 * d + (size_t)g * d;          // row g; row g+8 has the same parity
        const unsigned swz = (g & 1u) * swz_bit;
#pragma unroll 1
        for (unsigned kk = 0; kk < d / 64; kk++) {
            const unsigned off = 16u * ((4u * kk + c) ^ swz);
            const uint4 ya = *reinterpret_cast<const uint4 *>(ta + off);
            const uint4 yb = *reinterpret_cast<const uint4 *>(ta + 8u * d + off);
            const unsigned wa[4] = {ya.x, ya.y, ya.z, ya.w};
            const unsigned wb[4] = {yb.x, yb.y, yb.z, yb.w};
#pragma unroll
            for (int odd = 0; odd < 2; odd++) {
                // A fragments of this k-step: a0/a2 from row g (words 2 odd, 2 odd + 1), a1/a3 from row g + 8
                unsigned A[4][4];
                A[0][0] = wa[2 * odd]; A[0][2] = wa[2 * odd + 1]; A[0][1] = wb[2 * odd]; A[0][3] = wb[2 * odd + 1];
#pragma unroll
                for (int j = 0; j < 4; j++) indicator_planes(A[0][j], A[1][j], A[2][j], A[3][j]);
                const unsigned ks = 2 * kk + odd;
                const uint4 *bk = bs + (size_t)ks * (8 * 2 * 32) + lane;
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    const uint4 b01 = bk[nt * 64], b23 = bk[nt * 64 + 32];
                    mma_s8(acc[nt], A[0][0], A[0][1], A[0][2], A[0][3], b01.x, b01.y);
                    mma_s8(acc[nt], A[1][0], A[1][1], A[1][2], A[1][3], b01.z, b01.w);
                    mma_s8(acc[nt], A[2][0], A[2][1], A[2][2], A[2][3], b23.x, b23.y);
                    mma_s8(acc[nt], A[3][0], A[3][1], A[3][2], A[3][3], b23.z, b23.w);
                }
            }
        }
        __syncwarp();                                        // every lane is done with this buffer before it is refilled
        // epilogue: C fragment (row g | g+8, query nt*8 + 2c | +1); entries whose products may saturate are only marked
        const unsigned rmA = okA ? (unsigned)p.rowmax[rowA] : 0u, rmB = okB ? (unsigned)p.rowmax[rowB] : 0u;
        unsigned risky = 0;                                  // bit 4 nt + j
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned ql = nt * 8 + 2 * c + (j & 1);
                const unsigned q = q0 + ql;
                const bool hi = j >= 2;
                if (q >= p.Q || !(hi ? okB : okA)) continue;
                if ((hi ? rmB : rmA) * umax_s[ql] > sat_lim) { risky |= 1u << (4 * nt + j); continue; }
                store_bin(p.bins, p.bin8, (size_t)q * p.S_local + (hi ? rowB : rowA),
                          (unsigned)(qi_clamp(acc[nt][j] >> 2, la) + (int)p.bias));          // exact: a multiple of 4
            }
        }
        while (risky) {
            // some product of this (row, query) may saturate: the reference order of operations, product by product
            const unsigned k = (unsigned)(__ffs((int)risky) - 1);
            risky &= risky - 1u;
            const unsigned nt = k >> 2, j = k & 3u;
            const unsigned q = q0 + nt * 8 + 2 * c + (j & 1u);
            const unsigned long long row = (j >= 2) ? rowB : rowA;
            const signed char *yr = p.Y + row * d, *ur = p.ub8 + (size_t)q * d;
            int sp = 0;
#pragma unroll 4
            for (unsigned t = 0; t < d; t++) sp += qi_mul((int)yr[t], (int)ur[t], la, p.fb);
            store_bin(p.bins, p.bin8, (size_t)q * p.S_local + row, (unsigned)(qi_clamp(sp, la) + (int)p.bias));
        }
    }
}

// Query planes of k_big_scores_mma in fragment order.  For query block qb, k-step ks = 2 kk + odd, n-tile nt, half h,
// lane (g, c): one uint4 = { b0, b1 of plane 2h, b0, b1 of plane 2h+1 } with b0 = dims 64 kk + 16 c + 8 odd + 0..3 and
// b1 = the next four dims of query qb*64 + nt*8 + g; planes: 0 = u, a = 1..3: -sgn(u) rho(a, |u| & 3).
__global__ void __launch_bounds__(256) k_big_prep_bfrag(const signed char *__restrict__ ub8, unsigned Q, unsigned d, uint4 *__restrict__ bfrag)
{
    const unsigned KS = d / 32;
    const unsigned per_block = KS * 8 * 2 * 32;
    const unsigned qblocks = (Q + MMA_QB - 1) / MMA_QB;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < qblocks * per_block; i += gridDim.x * blockDim.x) {
        const unsigned qb = i / per_block, r = i % per_block;
        const unsigned ks = r / (8 * 2 * 32), nt = (r / 64) % 8, h = (r / 32) % 2, lane = r % 32;
        const unsigned g = lane >> 2, c = lane & 3, kk = ks >> 1, odd = ks & 1;
        const unsigned q = qb * MMA_QB + nt * 8 + g;
        unsigned w[4] = {0u, 0u, 0u, 0u};
        if (q < Q) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned plane = 2 * h + (j >> 1), which = j & 1;
                unsigned word = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int u = (int)ub8[(size_t)q * d + 64 * kk + 16 * c + 8 * odd + 4 * which + t];
                    int v = u;
                    if (plane) {
                        const int sg = (u > 0) - (u < 0);
                        v = -sg * (int)(((unsigned)plane * ((unsigned)abs(u) & 3u)) & 3u);
                    }
                    word |= ((unsigned)v & 0xFFu) << (8 * t);
                }
                w[j] = word;
            }
        }
        bfrag[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}


// -------------------------------------------------------------------------------------------------
// k_big_scores_tc: the same four int8 contractions on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in
// tensor memory), d % 128 == 0.
//
//   4 * sum_t trunc0(y_t u_t / 4) = sum_t y_t u_t + sum_{a=1..3} sum_t I_a(y_t) * (-R_a(u_t))          (see k_big_scores_mma)
//
// One persistent CTA per SM and query block of 64.  Per tile of 128 slots and 128-byte K chunk:
//   producer thread : TMA box (128 bytes x 128 rows, 128-byte swizzle) of Y -> stage;  Y itself is plane 0 of the A operand
//   8 builder warps : the three indicator planes of every 16-byte chunk, written at the SAME offset of three more tiles (the map is
//                     elementwise, so the swizzled layout carries over), then fence.proxy.async so that the tensor core sees them
//   MMA thread      : 4 planes x 4 K steps of tcgen05.mma M128 N64 K32 into ONE accumulator tile (the query planes of a = 1..3
//                     are stored negated); tcgen05.commit frees the stage, and after the last chunk hands the tile to the epilogue
//   8 epilogue warps: (two per quadrant of tensor memory, 32 queries each) tcgen05.ld, (acc >> 2) clamped -> score bin per
//                     (query, slot); entries whose products may saturate are recomputed product by product (same rule as the
//                     other fast scorers)
// The Y tiles sit in a ring of three (the TMA runs three deep: a 16 KB box takes ~2 300 cycles to arrive), the plane tiles in a ring
// of two, two accumulator tiles (2 x 64 TMEM columns).  The memory is streamed once per block of 64 queries.
// -------------------------------------------------------------------------------------------------
constexpr unsigned TCS_QB = 64, TCS_NY = 3, TCS_NP = 2, TCS_TILE = 128 * 128, TCS_BUILDERS = 8;
constexpr unsigned TCS_EPI = 8;                       // epilogue warps: two per TMEM quadrant, 32 accumulator columns each
// warp roles: the issue arbiter of an SM sub-partition favours the highest warp ids, so the epilogue (the critical path once the tensor
// core is fed) gets them: builders first, then the producer, the MMA issuer, and the eight epilogue warps
constexpr unsigned TCS_W_BUILD = 0, TCS_W_PROD = TCS_BUILDERS, TCS_W_MMA = TCS_BUILDERS + 1, TCS_W_EPI = TCS_BUILDERS + 2, TCS_WARPS = TCS_W_EPI + TCS_EPI;

struct TcScoreParams {
    alignas(64) CUtensorMap tmY;     // Y as bytes [S_local][d], box 128 x 128, 128-byte swizzle
    const signed char *Y;
    const unsigned char *rowmax;
    unsigned long long S_local;
    unsigned d, Q;
    int la, fb;
    const signed char *ub8;
    const unsigned *umax;
    const uint4 *bplanes;            // [qblocks][d/128][4 planes][64 rows x 128 B, swizzled]: the B operand's shared-memory image
    void *bins;
    int bin8;
    unsigned bias;
    unsigned long long *clk;         // QMANN_TC_TRACE builds: per-role cycle accumulators of CTA (0, 0)
};
#ifdef QMANN_TC_TRACE
#define SCK_DECL unsigned long long sck[4] = {0ull, 0ull, 0ull, 0ull}; long long sck_t = clock64();
#define SCK(i) do { const long long n_ = clock64(); sck[i] += (unsigned long long)(n_ - sck_t); sck_t = n_; } while (0)
#define SCK_FLUSH(role) do { if (p.clk && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0) for (int i_ = 0; i_ < 4; i_++) p.clk[(role) * 4 + i_] = sck[i_]; } while (0)
#else
#define SCK_DECL
#define SCK(i) do { } while (0)
#define SCK_FLUSH(role) do { } while (0)
#endif

// instruction descriptor of kind::i8: D int32 (bits 4-5 = 2), A and B signed 8-bit (bits 7-9, 10-12 = 1), K-major, N >> 3, M >> 4
__host__ __device__ constexpr unsigned umma_idesc_i8(unsigned M, unsigned N) { return (2u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }
__device__ __forceinline__ void umma_i8(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc, unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(TCS_WARPS * 32, 1) k_big_scores_tc(const __grid_constant__ TcScoreParams p)
{
    using namespace qtc;
    extern __shared__ __align__(16) unsigned char sm[];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned KC = p.d / 128;                                     // K chunks per tile
    const unsigned sraw = smem_u32(sm), sbase = (sraw + 1023u) & ~1023u;
    unsigned char *gbase = sm + (sbase - sraw);
    // layout: [B planes: KC x 4 x 8 KB][Y ring: 3 x 16 KB][plane ring: 2 x 3 x 16 KB][control, 512 B][two score-bin tiles of 64 x 128 bytes]
    const unsigned bsm = sbase, yrg = bsm + KC * 4u * 8192u, prg = yrg + TCS_NY * TCS_TILE, cb = prg + TCS_NP * 3u * TCS_TILE;
    unsigned char *cbg = gbase + (cb - sbase);
    const unsigned bar_yfull = cb, bar_yfree = cb + 24, bar_built = cb + 48, bar_pfree = cb + 64, bar_dfull = cb + 80, bar_dfree = cb + 96, tmem_slot = cb + 112;
    unsigned *umax_s = reinterpret_cast<unsigned *>(cbg + 128);
    const unsigned q0 = blockIdx.y * TCS_QB;
    {
        const uint4 *src = p.bplanes + (size_t)blockIdx.y * (KC * 4u * 8192u / 16u);
        uint4 *dst = reinterpret_cast<uint4 *>(gbase);
        for (unsigned i = threadIdx.x; i < KC * 4u * 8192u / 16u; i += blockDim.x) dst[i] = src[i];
        for (unsigned i = threadIdx.x; i < TCS_QB; i += blockDim.x) umax_s[i] = (q0 + i < p.Q) ? p.umax[q0 + i] : 0u;
    }
    if (threadIdx.x == 0) {
        for (unsigned s = 0; s < TCS_NY; s++) { mbar_init(bar_yfull + 8 * s, 1); mbar_init(bar_yfree + 8 * s, 1); }
        for (unsigned s = 0; s < TCS_NP; s++) { mbar_init(bar_built + 8 * s, TCS_BUILDERS); mbar_init(bar_pfree + 8 * s, 1); }
        for (unsigned s = 0; s < 2; s++) { mbar_init(bar_dfull + 8 * s, 1); mbar_init(bar_dfree + 8 * s, TCS_EPI); }
        mbar_fence_init();
        tma_prefetch_desc(&p.tmY);
    }
    if (warp == TCS_W_MMA) tmem_alloc(tmem_slot, 128);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the B planes written above are read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *reinterpret_cast<const unsigned *>(cbg + 112);
    const unsigned long long n_tiles = (p.S_local + 127ull) / 128ull;

    if (warp == TCS_W_PROD) {
        if (lane == 0) {
            unsigned it = 0;
            SCK_DECL
            for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (unsigned kc = 0; kc < KC; kc++, it++) {
                    const unsigned s = it % TCS_NY, ph = (it / TCS_NY) & 1u;
                    SCK(1);
                    mbar_wait(bar_yfree + 8 * s, ph ^ 1u);
                    SCK(0);
                    mbar_expect_tx(bar_yfull + 8 * s, TCS_TILE);
                    tma_load_2d(yrg + s * TCS_TILE, &p.tmY, (int)(128u * kc), (int)(tile * 128ull), bar_yfull + 8 * s);
                }
            SCK(1); SCK_FLUSH(0);
        }
    } else if (warp == TCS_W_MMA) {
        if (lane == 0) {
            const unsigned idesc = umma_idesc_i8(128, TCS_QB);
            unsigned it = 0, tc = 0;
            // descriptors of the A tiles (Y ring = plane 0, plane ring = planes 1..3) and of the B tiles are formed once; an MMA then
            // only steps the start-address field (+2 per 32 bytes of K)
            const unsigned long long ydesc0 = umma_desc_sw128(yrg), pdesc0 = umma_desc_sw128(prg), bdesc0 = umma_desc_sw128(bsm);
            SCK_DECL
            for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tc++) {
                const unsigned buf = tc & 1u;
                SCK(2);
                mbar_wait(bar_dfree + 8 * buf, ((tc >> 1) & 1u) ^ 1u);
                SCK(0);
                tc_fence_after();
                for (unsigned kc = 0; kc < KC; kc++, it++) {
                    const unsigned ys = it % TCS_NY, ps = it % TCS_NP, php = (it / TCS_NP) & 1u;
                    SCK(2);
                    mbar_wait(bar_built + 8 * ps, php);                     // implies the Y tile has arrived (the builders read it)
                    SCK(1);
                    tc_fence_after();
#pragma unroll
                    for (unsigned pl = 0; pl < 4; pl++) {
                        const unsigned long long ad = pl == 0 ? ydesc0 + (unsigned long long)(ys * (TCS_TILE >> 4))
                                                              : pdesc0 + (unsigned long long)((ps * 3u + pl - 1u) * (TCS_TILE >> 4));
                        const unsigned long long bd = bdesc0 + (unsigned long long)((kc * 4u + pl) * (8192u >> 4));
#pragma unroll
                        for (unsigned j = 0; j < 4; j++) umma_i8(tmem + buf * TCS_QB, ad + 2ull * j, bd + 2ull * j, idesc, (kc | pl | j) ? 1u : 0u);
                    }
                    umma_commit(bar_yfree + 8 * ys);
                    umma_commit(bar_pfree + 8 * ps);
                }
                umma_commit(bar_dfull + 8 * buf);
            }
            SCK(2); SCK_FLUSH(1);
        }
    } else if (warp < TCS_BUILDERS) {
        const unsigned bt = threadIdx.x;                                   // 0 .. 255
        unsigned it = 0;
        SCK_DECL
        for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (unsigned kc = 0; kc < KC; kc++, it++) {
                const unsigned ys = it % TCS_NY, phy = (it / TCS_NY) & 1u, ps = it % TCS_NP, php = (it / TCS_NP) & 1u;
                SCK(1);
                mbar_wait(bar_pfree + 8 * ps, php ^ 1u);                   // the MMAs that read this plane buffer have completed
                mbar_wait(bar_yfull + 8 * ys, phy);
                SCK(0);
                const unsigned char *y0 = gbase + (yrg - sbase) + (size_t)ys * TCS_TILE;
                unsigned char *p0 = gbase + (prg - sbase) + (size_t)ps * 3u * TCS_TILE;
#pragma unroll
                for (unsigned i = 0; i < (TCS_TILE / 16u + TCS_BUILDERS * 32u - 1u) / (TCS_BUILDERS * 32u); i++) {
                    const unsigned off = 16u * (bt + i * TCS_BUILDERS * 32u);
                    if (off >= TCS_TILE) break;
                    const uint4 y = *reinterpret_cast<const uint4 *>(y0 + off);
                    uint4 i1, i2, i3;
                    indicator_planes(y.x, i1.x, i2.x, i3.x);
                    indicator_planes(y.y, i1.y, i2.y, i3.y);
                    indicator_planes(y.z, i1.z, i2.z, i3.z);
                    indicator_planes(y.w, i1.w, i2.w, i3.w);
                    *reinterpret_cast<uint4 *>(p0 + off) = i1;
                    *reinterpret_cast<uint4 *>(p0 + TCS_TILE + off) = i2;
                    *reinterpret_cast<uint4 *>(p0 + 2u * TCS_TILE + off) = i3;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_built + 8 * ps);
            }
        SCK(1);
        if (warp == TCS_W_BUILD) SCK_FLUSH(2);
    } else {
        // epilogue warps: warp w reads TMEM lanes 32 (w % 4) .. + 31 = slots 32 (w % 4) + lane of the tile; the two warps of a
        // quadrant take accumulator columns 0-31 and 32-63 (queries of this block)
        const int la = p.la;
        const unsigned sat_lim = 4u * (unsigned)la + 3u;
        const unsigned ew = warp - TCS_W_EPI, et = threadIdx.x - TCS_W_EPI * 32u;          // epilogue warp / thread index
        const unsigned quad = warp & 3u, qh = 32u * (ew >> 2);
        const unsigned tq = tmem + ((32u * quad) << 16);
        const unsigned nq = (p.Q > q0 + qh) ? min(32u, p.Q - q0 - qh) : 0u;       // queries of this warp's half that exist
        unsigned umax_all = 0;
        for (unsigned i = 0; i < 32; i++) umax_all = max(umax_all, umax_s[qh + i]);
        unsigned tc = 0;
        SCK_DECL
        // the row maximum of the NEXT tile's slot is loaded one tile ahead (its DRAM latency would otherwise sit between the
        // accumulator load and the stores of every tile)
        unsigned rm_next = 0;
        {
            const unsigned long long r0 = (unsigned long long)blockIdx.x * 128ull + 32u * quad + lane;
            if (blockIdx.x < n_tiles && r0 < p.S_local) rm_next = (unsigned)p.rowmax[r0];
        }
        for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tc++) {
            const unsigned buf = tc & 1u;
            const unsigned long long row = tile * 128ull + 32u * quad + lane;
            const bool ok = row < p.S_local;
            const unsigned rm = rm_next;
            {
                const unsigned long long rn = row + (unsigned long long)gridDim.x * 128ull;
                rm_next = (tile + gridDim.x < n_tiles && rn < p.S_local) ? (unsigned)p.rowmax[rn] : 0u;
            }
            const bool lane_risky = rm * umax_all > sat_lim;                   // some query may saturate a product of this slot
            SCK(3);
            mbar_wait(bar_dfull + 8 * buf, (tc >> 1) & 1u);
            SCK(0);
            tc_fence_after();
            unsigned v[32];
            {
                unsigned v0[16], v1[16];
                tmem_ld16(tq + buf * TCS_QB + qh, v0);
                tmem_ld16(tq + buf * TCS_QB + qh + 16u, v1);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i++) { v[i] = v0[i]; v[16 + i] = v1[i]; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_dfree + 8 * buf);                  // the accumulators are in registers: the tile is free
            SCK(1);
            // Score bins of this warp's 32 slots x 32 queries go through a shared-memory tile [64 queries][128 slots] so that the
            // tile leaves as whole 128-byte lines (two 128-bit stores per thread) instead of 32 byte stores per lane; entries whose
            // products may saturate are recomputed product by product first (the reference order of operations).
            unsigned char *tileb = cbg + 512 + (tc & 1u) * 8192u;
            const bool vec_ok = p.bin8 && (p.S_local % 16ull == 0ull) && (tile * 128ull + 128ull <= p.S_local);
            unsigned risky = 0u;
            if (ok && lane_risky) {
#pragma unroll
                for (int i = 0; i < 32; i++)
                    if ((unsigned)i < nq && rm * umax_s[qh + i] > sat_lim) risky |= 1u << i;
            }
            if (vec_ok) {
#pragma unroll
                for (int i = 0; i < 32; i++) tileb[(qh + i) * 128u + 32u * quad + lane] = (unsigned char)(qi_clamp((int)v[i] >> 2, la) + (int)p.bias);      // exact: a multiple of 4
            } else if (ok) {
#pragma unroll 4
                for (unsigned i = 0; i < nq; i++)
                    if (!((risky >> i) & 1u)) store_bin(p.bins, p.bin8, (size_t)(q0 + qh + i) * p.S_local + row, (unsigned)(qi_clamp((int)v[i] >> 2, la) + (int)p.bias));
            }
            // (slot, query) pairs with a product that may saturate: the reference order of operations, product by product, by the
            // whole warp (8 products per lane and 256 dims, one reduction per pair) -- a lane alone would stall its tile for
            // thousands of instructions per pair
            unsigned any = __ballot_sync(0xffffffffu, risky != 0u);
            while (any) {
                const int src = __ffs((int)any) - 1;
                any &= any - 1u;
                unsigned m = __shfl_sync(0xffffffffu, risky, src);
                const signed char *yr = p.Y + (tile * 128ull + 32u * quad + (unsigned)src) * p.d;
                while (m) {
                    const unsigned i = (unsigned)(__ffs((int)m) - 1);
                    m &= m - 1u;
                    const signed char *ur = p.ub8 + (size_t)(q0 + qh + i) * p.d;
                    int sp = 0;
                    for (unsigned t0 = 8u * lane; t0 < p.d; t0 += 256u) {
                        const uint2 yv = *reinterpret_cast<const uint2 *>(yr + t0), uv = *reinterpret_cast<const uint2 *>(ur + t0);
#pragma unroll
                        for (int b8 = 0; b8 < 4; b8++) {
                            sp += qi_mul((int)(signed char)((yv.x >> (8 * b8)) & 0xFFu), (int)(signed char)((uv.x >> (8 * b8)) & 0xFFu), la, p.fb);
                            sp += qi_mul((int)(signed char)((yv.y >> (8 * b8)) & 0xFFu), (int)(signed char)((uv.y >> (8 * b8)) & 0xFFu), la, p.fb);
                        }
                    }
                    sp = __reduce_add_sync(0xffffffffu, sp);
                    const unsigned code = (unsigned)(qi_clamp(sp, la) + (int)p.bias);
                    if (lane == (unsigned)src) {
                        if (vec_ok) tileb[(qh + i) * 128u + 32u * quad + lane] = (unsigned char)code;
                        else store_bin(p.bins, p.bin8, (size_t)(q0 + qh + i) * p.S_local + row, code);
                    }
                }
            }
            SCK(2);
            if (vec_ok) {
                asm volatile("bar.sync 1, %0;" ::"n"(TCS_EPI * 32) : "memory");           // the eight epilogue warps
                unsigned char *dst = reinterpret_cast<unsigned char *>(p.bins) + tile * 128ull;
#pragma unroll
                for (unsigned k = 0; k < 2; k++) {
                    const unsigned idx = et + k * TCS_EPI * 32u;                            // 512 segments of 16 bytes
                    const unsigned ql = idx >> 3, seg = idx & 7u;
                    if (q0 + ql < p.Q)
                        *reinterpret_cast<uint4 *>(dst + (size_t)(q0 + ql) * p.S_local + 16u * seg) = *reinterpret_cast<const uint4 *>(tileb + ql * 128u + 16u * seg);
                }
            }
            SCK(3);
        }
        if (ew == 0) SCK_FLUSH(3);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == TCS_W_MMA) tmem_dealloc(tmem, 128);
}

// -------------------------------------------------------------------------------------------------
// k_big_scores_tq: k_big_scores_tc with the operand roles swapped -- the QUERIES are the A operand and live in tensor memory,
// the memory tile is the B operand (d = 128 or 256, byte bins).
//
//   D[128 queries][128 slots] (int32, TMEM) += A[128 queries][K] (TMEM: lane = query, four 8-bit K elements per column)
//                                              * B[128 slots][K]^T (shared memory, K-major, 128-byte swizzle)
//
// What that buys over k_big_scores_tc:
//   * the four query planes (4 x d bytes per query = 256 TMEM columns at d = 256) take no shared memory, so a CTA serves 128 queries
//     per pass over the memory instead of 64: the indicator planes of a tile -- the builders' work, the largest cost of the tc
//     kernel -- are built once per 128 queries, and one tcgen05.mma (M128 N128 K32) does twice the work per issue;
//   * an epilogue lane owns a QUERY and reads 64 consecutive slots of it from its TMEM lane: the score bins leave as 64 contiguous
//     bytes per lane without a shared-memory transpose.
// TMEM: accumulator tiles at columns [0,128) and [128,256), query planes at 256 + (kc * 4 + plane) * 32 + k / 4.
// Warp roles and the mbarrier protocol are those of k_big_scores_tc; the plane ring is three deep (the shared memory the query
// planes no longer take).
// (A variant with the score histogram fused into this epilogue was measured and dropped: with a lane per query every lane of a
// shared-memory atomic hits its own address, ~12 cycles per warp instruction, 6.4 k cycles per tile against 3.1 k of MMA issue;
// lane-exclusive counters without atomics need one epilogue warp per quadrant, whose dependent load-add-store chains took 9 k.)
// -------------------------------------------------------------------------------------------------
constexpr unsigned TQ_QB = 128;

struct TqScoreParams {
    alignas(64) CUtensorMap tmY;     // Y as bytes [S_local][d], box 128 x 128, 128-byte swizzle
    const signed char *Y;
    const unsigned char *rowmax;
    unsigned long long S_local;
    unsigned d, Q;
    int la, fb;
    const signed char *ub8;
    const unsigned *umax;
    unsigned char *bins;
    unsigned bias;
    unsigned qblk0;                  // first query block of this launch (blockIdx.y counts from it)
    unsigned long long *clk;         // QMANN_TC_TRACE builds: per-role cycle accumulators of CTA (0, 0)
};

__device__ __forceinline__ void umma_i8_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long b_desc, unsigned idesc, unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// plane a = 1..3 of four packed query codes: -sgn(u) ((a (|u| & 3)) & 3); plane 0 is u itself
__device__ __forceinline__ unsigned query_plane_word(unsigned uw, unsigned pl)
{
    if (pl == 0) return uw;
    unsigned fill;
    asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(fill) : "r"(uw));           // 0xFF in the bytes that are negative
    const unsigned a = (((uw ^ fill) & 0x03030303u) + (fill & 0x01010101u)) & 0x03030303u;    // |u| & 3 per byte
    const unsigned a0 = a & 0x01010101u;
    const unsigned m = (pl == 1) ? a : (pl == 2 ? (a0 << 1) : ((a + (a0 << 1)) & 0x03030303u));
    return (fill & m) | (~fill & __vneg4(m));
}

// TQ_NP: depth of the plane ring (3; 2 when the histogram kernel has to fit next to this CTA, see qmann_bigmem_hop_scores)
template <unsigned TQ_NP>
__global__ void __launch_bounds__(TCS_WARPS * 32, 1) k_big_scores_tq(const __grid_constant__ TqScoreParams p)
{
    using namespace qtc;
    extern __shared__ __align__(16) unsigned char sm[];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned KC = p.d / 128;                                     // K chunks per tile
    const unsigned sraw = smem_u32(sm), sbase = (sraw + 1023u) & ~1023u;
    unsigned char *gbase = sm + (sbase - sraw);
    // layout: [Y ring: 3 x 16 KB][plane ring: 3 x 3 x 16 KB][control, 1 KB]
    const unsigned yrg = sbase, prg = yrg + TCS_NY * TCS_TILE, cb = prg + TQ_NP * 3u * TCS_TILE;
    unsigned char *cbg = gbase + (cb - sbase);
    const unsigned bar_yfull = cb, bar_yfree = cb + 24, bar_built = cb + 48, bar_pfree = cb + 72, bar_dfull = cb + 96, bar_dfree = cb + 112, tmem_slot = cb + 128;
    unsigned *umax_s = reinterpret_cast<unsigned *>(cbg + 256);        // [128]
    const unsigned q0 = (p.qblk0 + blockIdx.y) * TQ_QB;
    for (unsigned i = threadIdx.x; i < TQ_QB; i += blockDim.x) umax_s[i] = (q0 + i < p.Q) ? p.umax[q0 + i] : 0u;
    if (threadIdx.x == 0) {
        for (unsigned s = 0; s < TCS_NY; s++) { mbar_init(bar_yfull + 8 * s, 1); mbar_init(bar_yfree + 8 * s, 1); }
        for (unsigned s = 0; s < TQ_NP; s++) { mbar_init(bar_built + 8 * s, TCS_BUILDERS); mbar_init(bar_pfree + 8 * s, 1); }
        for (unsigned s = 0; s < 2; s++) { mbar_init(bar_dfull + 8 * s, 1); mbar_init(bar_dfree + 8 * s, TCS_EPI); }
        mbar_fence_init();
        tma_prefetch_desc(&p.tmY);
    }
    if (warp == TCS_W_MMA) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *reinterpret_cast<const unsigned *>(cbg + 128);
    if (warp < 16) {
        // the A operand: this thread's query (TMEM lane 32 (warp % 4) + lane), 4 planes x d bytes; the four warps that can reach a
        // TMEM quadrant split the 16-byte chunks of a K chunk between them
        const unsigned aq = warp & 3u, part = warp >> 2;
        const unsigned q = q0 + 32u * aq + lane;
        const unsigned tq = tmem + ((32u * aq) << 16) + 256u;
        for (unsigned kc = 0; kc < KC; kc++)
            for (unsigned c16 = 2u * part; c16 < 2u * part + 2u; c16++) {
                uint4 u = make_uint4(0u, 0u, 0u, 0u);
                if (q < p.Q) u = *reinterpret_cast<const uint4 *>(p.ub8 + (size_t)q * p.d + 128u * kc + 16u * c16);
#pragma unroll
                for (unsigned pl = 0; pl < 4; pl++) {
                    const unsigned w4[4] = {query_plane_word(u.x, pl), query_plane_word(u.y, pl), query_plane_word(u.z, pl), query_plane_word(u.w, pl)};
                    tmem_st4(tq + (kc * 4u + pl) * 32u + 4u * c16, w4);
                }
            }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned long long n_tiles = (p.S_local + 127ull) / 128ull;

    if (warp == TCS_W_PROD) {
        if (lane == 0) {
            unsigned it = 0;
            SCK_DECL
            for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (unsigned kc = 0; kc < KC; kc++, it++) {
                    const unsigned s = it % TCS_NY, ph = (it / TCS_NY) & 1u;
                    SCK(1);
                    mbar_wait(bar_yfree + 8 * s, ph ^ 1u);
                    SCK(0);
                    mbar_expect_tx(bar_yfull + 8 * s, TCS_TILE);
                    tma_load_2d(yrg + s * TCS_TILE, &p.tmY, (int)(128u * kc), (int)(tile * 128ull), bar_yfull + 8 * s);
                }
            SCK(1); SCK_FLUSH(0);
        }
    } else if (warp == TCS_W_MMA) {
        if (lane == 0) {
            const unsigned idesc = umma_idesc_i8(TQ_QB, 128);
            unsigned it = 0, tc = 0;
            const unsigned long long ydesc0 = umma_desc_sw128(yrg), pdesc0 = umma_desc_sw128(prg);
            SCK_DECL
            for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tc++) {
                const unsigned buf = tc & 1u;
                SCK(2);
                mbar_wait(bar_dfree + 8 * buf, ((tc >> 1) & 1u) ^ 1u);
                SCK(0);
                tc_fence_after();
                for (unsigned kc = 0; kc < KC; kc++, it++) {
                    const unsigned ys = it % TCS_NY, ps = it % TQ_NP, php = (it / TQ_NP) & 1u;
                    SCK(2);
                    mbar_wait(bar_built + 8 * ps, php);                     // implies the Y tile has arrived (the builders read it)
                    SCK(1);
                    tc_fence_after();
#pragma unroll
                    for (unsigned pl = 0; pl < 4; pl++) {
                        const unsigned long long bd = pl == 0 ? ydesc0 + (unsigned long long)(ys * (TCS_TILE >> 4))
                                                              : pdesc0 + (unsigned long long)((ps * 3u + pl - 1u) * (TCS_TILE >> 4));
                        const unsigned at = tmem + 256u + (kc * 4u + pl) * 32u;
#pragma unroll
                        for (unsigned j = 0; j < 4; j++) umma_i8_ts(tmem + buf * 128u, at + 8u * j, bd + 2ull * j, idesc, (kc | pl | j) ? 1u : 0u);
                    }
                    umma_commit(bar_yfree + 8 * ys);
                    umma_commit(bar_pfree + 8 * ps);
                }
                umma_commit(bar_dfull + 8 * buf);
            }
            SCK(2); SCK_FLUSH(1);
        }
    } else if (warp < TCS_BUILDERS) {
        const unsigned bt = threadIdx.x;                                   // 0 .. 255
        unsigned it = 0;
        SCK_DECL
        for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (unsigned kc = 0; kc < KC; kc++, it++) {
                const unsigned ys = it % TCS_NY, phy = (it / TCS_NY) & 1u, ps = it % TQ_NP, php = (it / TQ_NP) & 1u;
                SCK(1);
                mbar_wait(bar_pfree + 8 * ps, php ^ 1u);                   // the MMAs that read this plane buffer have completed
                SCK(2);
                mbar_wait(bar_yfull + 8 * ys, phy);
                SCK(0);
                const unsigned char *y0 = gbase + (yrg - sbase) + (size_t)ys * TCS_TILE;
                unsigned char *p0 = gbase + (prg - sbase) + (size_t)ps * 3u * TCS_TILE;
#pragma unroll
                for (unsigned i = 0; i < TCS_TILE / 16u / (TCS_BUILDERS * 32u); i++) {
                    const unsigned off = 16u * (bt + i * TCS_BUILDERS * 32u);
                    const uint4 y = *reinterpret_cast<const uint4 *>(y0 + off);
                    uint4 i1, i2, i3;
                    indicator_planes(y.x, i1.x, i2.x, i3.x);
                    indicator_planes(y.y, i1.y, i2.y, i3.y);
                    indicator_planes(y.z, i1.z, i2.z, i3.z);
                    indicator_planes(y.w, i1.w, i2.w, i3.w);
                    *reinterpret_cast<uint4 *>(p0 + off) = i1;
                    *reinterpret_cast<uint4 *>(p0 + TCS_TILE + off) = i2;
                    *reinterpret_cast<uint4 *>(p0 + 2u * TCS_TILE + off) = i3;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_built + 8 * ps);
            }
        SCK(1);
        if (warp == TCS_W_BUILD) SCK_FLUSH(2);
    } else {
        // epilogue warps: warp w reads TMEM lanes 32 (w % 4) .. + 31 = queries q0 + 32 (w % 4) + lane; the two warps of a quadrant take
        // accumulator columns (slots of the tile) 0-63 and 64-127, in chunks of 32
        const int la = p.la;
        const unsigned sat_lim = 4u * (unsigned)la + 3u;
        const unsigned ew = warp - TCS_W_EPI;
        const unsigned quad = warp & 3u, half = ew >> 2;
        const unsigned tq = tmem + ((32u * quad) << 16);
        const unsigned ql = 32u * quad + lane;
        const bool qok = q0 + ql < p.Q;
        const unsigned umax_q = umax_s[ql];
        unsigned umax_all = 0;
        for (unsigned i = 0; i < TQ_QB; i++) umax_all = max(umax_all, umax_s[i]);
        unsigned char *brow = p.bins + (size_t)(q0 + ql) * p.S_local;
        const signed char *urow = p.ub8 + (size_t)(q0 + 32u * quad) * p.d;   // + src * d
        const bool al16 = (p.S_local % 16ull) == 0ull;
        unsigned tc = 0;
        // row maxima of the 64 slots of this half: lane holds slots lane and 32 + lane, loaded one tile ahead
        unsigned rm_next[2] = {0u, 0u};
        SCK_DECL
        {
            const unsigned long long r0 = (unsigned long long)blockIdx.x * 128ull + 64u * half + lane;
            if (blockIdx.x < n_tiles) {
                if (r0 < p.S_local) rm_next[0] = (unsigned)p.rowmax[r0];
                if (r0 + 32u < p.S_local) rm_next[1] = (unsigned)p.rowmax[r0 + 32u];
            }
        }
        for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, tc++) {
            const unsigned buf = tc & 1u;
            const unsigned long long slot0 = tile * 128ull + 64u * half;         // first slot of this warp's columns
            const unsigned rm[2] = {rm_next[0], rm_next[1]};
            {
                const unsigned long long rn = slot0 + (unsigned long long)gridDim.x * 128ull + lane;
                const bool more = tile + gridDim.x < n_tiles;
                rm_next[0] = (more && rn < p.S_local) ? (unsigned)p.rowmax[rn] : 0u;
                rm_next[1] = (more && rn + 32u < p.S_local) ? (unsigned)p.rowmax[rn + 32u] : 0u;
            }
            const unsigned nvalid = (slot0 >= p.S_local) ? 0u : (unsigned)min(64ull, p.S_local - slot0);
            SCK(3);
            mbar_wait(bar_dfull + 8 * buf, (tc >> 1) & 1u);
            SCK(0);
            tc_fence_after();
#pragma unroll
            for (unsigned c = 0; c < 2; c++) {
                unsigned v[32];
                {
                    unsigned v0[16], v1[16];
                    tmem_ld16(tq + buf * 128u + 64u * half + 32u * c, v0);
                    tmem_ld16(tq + buf * 128u + 64u * half + 32u * c + 16u, v1);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; i++) { v[i] = v0[i]; v[16 + i] = v1[i]; }
                }
                if (c == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_dfree + 8 * buf);          // the accumulators are in registers: the tile is free
                }
                SCK(1);
                // (slot, query) pairs with a product that may saturate: the reference order of operations, product by product, by
                // the whole warp; the exact sum replaces the accumulator (times 4: the accumulators are 4 x the score)
                unsigned risky_slots = __ballot_sync(0xffffffffu, rm[c] * umax_all > sat_lim);
                while (risky_slots) {
                    const unsigned sidx = (unsigned)(__ffs((int)risky_slots) - 1);
                    risky_slots &= risky_slots - 1u;
                    const unsigned rms = __shfl_sync(0xffffffffu, rm[c], (int)sidx);
                    const bool mine = qok && rms * umax_q > sat_lim;
                    unsigned lm = __ballot_sync(0xffffffffu, mine);
                    const signed char *yr = p.Y + (slot0 + 32u * c + sidx) * p.d;
                    unsigned val = 0u;
                    while (lm) {
                        const unsigned src = (unsigned)(__ffs((int)lm) - 1);
                        lm &= lm - 1u;
                        const signed char *ur = urow + (size_t)src * p.d;
                        int sp = 0;
                        for (unsigned t0 = 8u * lane; t0 < p.d; t0 += 256u) {
                            const uint2 yv = *reinterpret_cast<const uint2 *>(yr + t0), uv = *reinterpret_cast<const uint2 *>(ur + t0);
#pragma unroll
                            for (int b8 = 0; b8 < 4; b8++) {
                                sp += qi_mul((int)(signed char)((yv.x >> (8 * b8)) & 0xFFu), (int)(signed char)((uv.x >> (8 * b8)) & 0xFFu), la, p.fb);
                                sp += qi_mul((int)(signed char)((yv.y >> (8 * b8)) & 0xFFu), (int)(signed char)((uv.y >> (8 * b8)) & 0xFFu), la, p.fb);
                            }
                        }
                        sp = __reduce_add_sync(0xffffffffu, sp);
                        if (lane == src) val = (unsigned)(qi_clamp(sp, la) * 4);
                    }
#pragma unroll
                    for (unsigned i = 0; i < 32; i++)
                        if (i == sidx && mine) v[i] = val;
                }
                SCK(2);
                // bins of 32 slots: 32 contiguous bytes of this query's row
                const unsigned nv = (nvalid > 32u * c) ? min(32u, nvalid - 32u * c) : 0u;
                unsigned w8[8];
#pragma unroll
                for (unsigned i = 0; i < 32; i++) {
                    const unsigned code = (unsigned)(qi_clamp((int)v[i] >> 2, la) + (int)p.bias);      // exact: a multiple of 4
                    if ((i & 3u) == 0u) w8[i >> 2] = code; else w8[i >> 2] |= code << (8u * (i & 3u));
                }
                if (qok) {
                    unsigned char *dst = brow + slot0 + 32u * c;
                    if (al16 && nv == 32u) {
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                        *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
                    } else {
#pragma unroll
                        for (unsigned i = 0; i < 32; i++)
                            if (i < nv) dst[i] = (unsigned char)((w8[i >> 2] >> (8u * (i & 3u))) & 0xFFu);
                    }
                }
                SCK(3);
            }
        }
        if (ew == 0) SCK_FLUSH(3);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == TCS_W_MMA) tmem_dealloc(tmem, 512);
}

// Query planes of k_big_scores_tc as the shared-memory image of its B operand: for query block qb, K chunk kc, plane pl: 64 rows
// (queries) x 128 bytes with the 16-byte chunks XOR-swizzled by (row & 7) (the 128-byte swizzle of the K-major descriptor).
// Plane 0 = Q_bin(u); planes a = 1..3 = -sgn(u) ((a (|u| & 3)) & 3) (negated: all four contractions accumulate into one tile).
__global__ void __launch_bounds__(256) k_big_prep_bplanes(const signed char *__restrict__ ub8, unsigned Q, unsigned d, uint4 *__restrict__ out)
{
    const unsigned KC = d / 128, per_block = KC * 4u * 512u;              // uint4 per query block
    const unsigned qblocks = (Q + TCS_QB - 1) / TCS_QB;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < qblocks * per_block; i += gridDim.x * blockDim.x) {
        const unsigned qb = i / per_block, r = i % per_block;
        const unsigned kc = r / 2048u, pl = (r / 512u) % 4u, row = (r / 8u) % 64u, pos = r % 8u;
        const unsigned chunk = pos ^ (row & 7u);                         // logical 16-byte chunk stored at this position
        const unsigned q = qb * TCS_QB + row;
        unsigned w[4] = {0u, 0u, 0u, 0u};
        if (q < Q) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                unsigned word = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int u = (int)ub8[(size_t)q * d + 128u * kc + 16u * chunk + 4u * j + t];
                    int v = u;
                    if (pl) {
                        const int sg = (u > 0) - (u < 0);
                        v = -sg * (int)(((unsigned)pl * ((unsigned)abs(u) & 3u)) & 3u);
                    }
                    word |= ((unsigned)v & 0xFFu) << (8 * t);
                }
                w[j] = word;
            }
        }
        out[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// One warp per row: Y[r][t] = Q_att(M[r][t]) (written when Y != nullptr), rowmax[r] = max_t |Y[r][t]|, and *mismatch
// is set when some Y differs from its M code (then the scorer needs the copy).
__global__ void __launch_bounds__(256) k_big_prep_mem(const signed char *__restrict__ M, unsigned long long S, unsigned d, HopFmt f,
                                                      signed char *__restrict__ Y, unsigned char *__restrict__ rowmax, unsigned *__restrict__ mismatch)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long w0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long ws = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    bool mism = false;
    for (unsigned long long r = w0; r < S; r += ws) {
        unsigned mx = 0;
        for (unsigned c = lane; c < d / 16; c += 32) {
            const uint4 v = *(reinterpret_cast<const uint4 *>(M + r * d) + c);
            const unsigned vw[4] = {v.x, v.y, v.z, v.w};
            unsigned ow[4];
#pragma unroll
            for (int w = 0; w < 4; w++) {
                unsigned o = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int code = (int)(signed char)((vw[w] >> (8 * b)) & 0xFFu);
                    const int y = qi_requant(code, f.fw, f.la, f.fa);
                    mism |= (y != code);
                    mx = max(mx, (unsigned)abs(y));
                    o |= ((unsigned)y & 0xFFu) << (8 * b);
                }
                ow[w] = o;
            }
            if (Y) *(reinterpret_cast<uint4 *>(Y + r * d) + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) rowmax[r] = (unsigned char)mx;
    }
    if (__any_sync(0xffffffffu, mism) && lane == 0) atomicOr(mismatch, 1u);
}

// -------------------------------------------------------------------------------------------------
// k_big_hist: hist[q][bin] += count over this rank's slots.  Small bin counts (mode 2: 255) live in
// shared memory; large ones (mode 3: 2*127*d+1) use warp-aggregated global atomics.
// -------------------------------------------------------------------------------------------------
template <typename BinT>
__global__ void __launch_bounds__(256) k_big_hist(const BinT *__restrict__ bins, unsigned long long S_local, unsigned NB,
                                                  unsigned *__restrict__ hist, int use_smem)
{
    extern __shared__ unsigned hs[];
    constexpr unsigned PER = 16 / sizeof(BinT);              // bins per 128-bit load
    constexpr unsigned BITS = 8 * sizeof(BinT), MASK = (1u << BITS) - 1u;
    const unsigned q = blockIdx.y;
    const BinT *b = bins + (size_t)q * S_local;
    unsigned *hq = hist + (size_t)q * NB;
    if (use_smem) {
        for (unsigned i = threadIdx.x; i < NB; i += blockDim.x) hs[i] = 0;
        __syncthreads();
    }
    auto add = [&](unsigned v, unsigned n) {
        if (use_smem) atomicAdd(&hs[v], n);
        else atomicAdd(&hq[v], n);
    };
    // one 128-bit load per thread and iteration; equal neighbours (the scores of a query concentrate in a few bins) share
    // one atomic
    const bool vec = (S_local % PER == 0) && ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
    const unsigned long long nv = vec ? S_local / PER : 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint4 v4 = __ldg(reinterpret_cast<const uint4 *>(b) + i);
        const unsigned w[4] = {v4.x, v4.y, v4.z, v4.w};
        unsigned cur = w[0] & MASK, cnt = 0;
#pragma unroll
        for (unsigned k = 0; k < PER; k++) {
            const unsigned v = (w[k * BITS / 32] >> ((k * BITS) % 32)) & MASK;
            if (v != cur) { add(cur, cnt); cur = v; cnt = 0; }
            cnt++;
        }
        add(cur, cnt);
    }
    for (unsigned long long i = nv * PER + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < S_local; i += (unsigned long long)gridDim.x * blockDim.x)
        add((unsigned)b[i], 1u);
    if (use_smem) {
        __syncthreads();
        for (unsigned i = threadIdx.x; i < NB; i += blockDim.x)
            if (hs[i]) atomicAdd(&hq[i], hs[i]);
    }
}

// -------------------------------------------------------------------------------------------------
// k_big_hist_lanes (byte bins): the same histogram without atomics in the inner loop.  A warp takes (query, run of slots) tasks;
// every LANE keeps its own 16-bit counter per bin in shared memory ([bin][lane], 16 KB per warp), so an update is a plain
// load-add-store of an address no other thread touches.  Four bins are handled together: the four counters are loaded first, a
// later value of the group that repeats an earlier bin takes the earlier result, the stores follow in order -- the dependent
// load -> add -> store chain is paid once per four values.  At the end of a task the 32 lane copies of every bin are summed
// (one warp reduction per bin) and added to the global histogram.  k_big_hist's shared-memory atomics cost ~7 cycles per warp
// instruction; this form is bounded by plain shared-memory traffic.
// -------------------------------------------------------------------------------------------------
constexpr unsigned HL_WARPS = 12;

// (a bank-conflict-free counter layout -- the two bins of a pair in the lane's own 32-bit word -- was measured 15 % slower: the
// kernel is bound by instruction issue, and that layout costs two more integer instructions per value)
__device__ __forceinline__ void hl_count4(unsigned w, unsigned short *cl)
{
    const unsigned c0 = w & 0xFFu, c1 = (w >> 8) & 0xFFu, c2 = (w >> 16) & 0xFFu, c3 = w >> 24;
    unsigned short *a0 = cl + c0 * 32u, *a1 = cl + c1 * 32u, *a2 = cl + c2 * 32u, *a3 = cl + c3 * 32u;
    unsigned n0 = *a0, n1 = *a1, n2 = *a2, n3 = *a3;
    n0 += 1u;
    n1 = ((c1 == c0) ? n0 : n1) + 1u;
    n2 = ((c2 == c1) ? n1 : ((c2 == c0) ? n0 : n2)) + 1u;
    n3 = ((c3 == c2) ? n2 : ((c3 == c1) ? n1 : ((c3 == c0) ? n0 : n3))) + 1u;
    *a0 = (unsigned short)n0;
    *a1 = (unsigned short)n1;
    *a2 = (unsigned short)n2;
    *a3 = (unsigned short)n3;
}

__global__ void __launch_bounds__(HL_WARPS * 32, 1) k_big_hist_lanes(const unsigned char *__restrict__ bins, unsigned long long S_local, unsigned Q, unsigned NB,
                                                                     unsigned chunk, unsigned *__restrict__ hist)
{
    extern __shared__ __align__(16) unsigned char hsm[];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned short *cnt = reinterpret_cast<unsigned short *>(hsm) + (size_t)warp * (256u * 32u);      // [bin][lane]
    for (unsigned i = lane; i < 256u * 32u / 2u; i += 32) reinterpret_cast<unsigned *>(cnt)[i] = 0u;
    __syncwarp();
    unsigned short *cl = cnt + lane;
    const unsigned long long nch = (S_local + chunk - 1) / chunk, ntask = nch * Q;       // chunk: a multiple of 16, at most 2^20 (16-bit counters)
    const bool vec = (S_local % 16ull) == 0ull && (reinterpret_cast<uintptr_t>(bins) & 15) == 0;
    const unsigned nwarps = blockDim.x >> 5;                     // 12, or 5 when the CTA shares the SM with the scorer
    for (unsigned long long task = (unsigned long long)blockIdx.x * nwarps + warp; task < ntask; task += (unsigned long long)gridDim.x * nwarps) {
        const unsigned q = (unsigned)(task / nch);
        const unsigned long long s0 = (task % nch) * chunk, s1 = min(S_local, s0 + (unsigned long long)chunk);
        const unsigned char *b = bins + (size_t)q * S_local;
        const unsigned nvec = vec ? (unsigned)((s1 - s0) / 16ull) : 0u;
        const uint4 *vb = reinterpret_cast<const uint4 *>(b + s0);
        // groups of 128 vectors (four per lane), the next group's loads issued before the current one is counted: eight 128-bit
        // loads in flight per lane
        const unsigned ngrp = nvec / 128u;
        uint4 cur[4], nxt[4];
#pragma unroll
        for (unsigned k = 0; k < 4; k++) {
            cur[k] = make_uint4(0u, 0u, 0u, 0u);
            if (ngrp) cur[k] = __ldg(vb + lane + 32u * k);
        }
        for (unsigned g = 0; g < ngrp; g++) {
#pragma unroll
            for (unsigned k = 0; k < 4; k++) {
                nxt[k] = make_uint4(0u, 0u, 0u, 0u);
                if (g + 1u < ngrp) nxt[k] = __ldg(vb + (size_t)(g + 1u) * 128u + lane + 32u * k);
            }
#pragma unroll
            for (unsigned k = 0; k < 4; k++) {
                hl_count4(cur[k].x, cl);
                hl_count4(cur[k].y, cl);
                hl_count4(cur[k].z, cl);
                hl_count4(cur[k].w, cl);
            }
#pragma unroll
            for (unsigned k = 0; k < 4; k++) cur[k] = nxt[k];
        }
        for (unsigned i = ngrp * 128u + lane; i < nvec; i += 32) {
            const uint4 v = __ldg(vb + i);
            hl_count4(v.x, cl);
            hl_count4(v.y, cl);
            hl_count4(v.z, cl);
            hl_count4(v.w, cl);
        }
        for (unsigned long long s = s0 + 16ull * nvec + lane; s < s1; s += 32) cl[(unsigned)b[s] * 32u] += 1;
        __syncwarp();
        unsigned *hq = hist + (size_t)q * NB;
        for (unsigned bin0 = 0; bin0 < NB; bin0 += 8) {
            unsigned c[8], mine = 0u, any = 0u;
#pragma unroll
            for (unsigned k = 0; k < 8; k++) {                  // cnt has 256 rows: bins NB .. 255 stay zero
                c[k] = cl[(bin0 + k) * 32u];
                any |= c[k];
            }
            if (!__any_sync(0xffffffffu, any != 0u)) continue;  // the scores of a query occupy a few dozen neighbouring bins
#pragma unroll
            for (unsigned k = 0; k < 8; k++) cl[(bin0 + k) * 32u] = 0;
#pragma unroll
            for (unsigned k = 0; k < 8; k++) {
                const unsigned t = __reduce_add_sync(0xffffffffu, c[k]);
                if (lane == k) mine = t;
            }
            if (lane < 8u && bin0 + lane < NB && mine) atomicAdd(hq + bin0 + lane, mine);
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------------
// k_big_softmax: one CTA per query.  From the GLOBAL histogram: max bin, total, Q_f(p) per bin.
// value(bin): mode 2 (bin - la) / 2^fa; mode 3 Q_(iwl,31-iwl)((bin - 127 d) / 2^(7 - const_scale))
// with the +-2^iwl saturation and the -2^iwl -> 0 quirk.
// -------------------------------------------------------------------------------------------------
struct SoftmaxParams {
    const unsigned *hist;      // [Q][NB] global counts
    unsigned NB, bias;
    int mode, fa, ia, const_scale, iff, ff;
    unsigned char *pq;         // [Q][NB] Q_f(p) code per bin (0 where the bin is empty)
    unsigned *thr;             // [Q] lowest bin with a non-zero code (NB if none)
    float *pbin;               // optional [Q][NB] p per bin (parity checks)
    double *total_out;         // optional [Q]
};

__device__ __forceinline__ float big_bin_value(const SoftmaxParams &p, unsigned bin)
{
    const int n = (int)bin - (int)p.bias;
    if (p.mode == 3) {
        const float v = (float)n / (float)(1 << (7 - p.const_scale));
        const float lim = (float)(1 << p.ia);
        return (v >= lim) ? lim : (v < -lim ? -lim : (v == -lim ? 0.0f : v));
    }
    return (float)n / (float)(1 << p.fa);
}

__global__ void __launch_bounds__(SOFTMAX_RANGES) k_big_softmax(const SoftmaxParams p)
{
    __shared__ float s_max[SOFTMAX_RANGES];
    __shared__ double s_part[SOFTMAX_RANGES];
    __shared__ double s_total;
    __shared__ unsigned s_thr;
    const unsigned q = blockIdx.x, t = threadIdx.x;
    const unsigned *h = p.hist + (size_t)q * p.NB;
    const unsigned per = (p.NB + SOFTMAX_RANGES - 1) / SOFTMAX_RANGES;
    const unsigned b0 = min(p.NB, t * per), b1 = min(p.NB, b0 + per);
    // max over non-empty bins (mode 3 values are not monotone in the bin at the saturation edge, so take
    // the max of the VALUES, like _cuda_max does on the score vector, layer_cuda.cu:1895-1916)
    float mx = -INFINITY;
    for (unsigned b = b0; b < b1; b++)
        if (h[b]) mx = fmaxf(mx, big_bin_value(p, b));
    s_max[t] = mx;
    __syncthreads();
    if (t == 0) {
        float m = -INFINITY;
        for (unsigned i = 0; i < SOFTMAX_RANGES; i++) m = fmaxf(m, s_max[i]);
        s_max[0] = m;
        s_thr = p.NB;
    }
    __syncthreads();
    mx = s_max[0];
    // total = sum_bins count * __expf(v - max): each range ascending, then the range partials ascending
    double part = 0.0;
    for (unsigned b = b0; b < b1; b++)
        if (h[b]) part += (double)h[b] * (double)__expf(big_bin_value(p, b) - mx);
    s_part[t] = part;
    __syncthreads();
    if (t == 0) {
        double tot = 0.0;
        for (unsigned i = 0; i < SOFTMAX_RANGES; i++) tot += s_part[i];
        s_total = tot;
        if (p.total_out) p.total_out[q] = tot;
    }
    __syncthreads();
    const double total = s_total;
    unsigned lo = p.NB;
    for (unsigned b = b0; b < b1; b++) {
        unsigned code = 0;
        float pr = 0.0f;
        if (h[b]) {
            pr = (float)((double)__expf(big_bin_value(p, b) - mx) / total);
            code = (unsigned)qi_encode(pr, p.iff, p.ff);                       // layer_cuda.cu:561
            if (code && b < lo) lo = b;
        }
        p.pq[(size_t)q * p.NB + b] = (unsigned char)code;
        if (p.pbin) p.pbin[(size_t)q * p.NB + b] = pr;
    }
    if (lo < p.NB) atomicMin(&s_thr, lo);
    __syncthreads();
    if (t == 0) p.thr[q] = s_thr;
}

// -------------------------------------------------------------------------------------------------
// k_big_read: partial[q][c] += Q_f( Q_f(p[r]) * Q_f(C[r][c]) ) over this rank's slots with Q_f(p) != 0.
// The test is on the code itself (pq[bin] != 0; in mode 3 the value is not monotone in the bin at the
// -2^iwl edge); thr[] = lowest bin with a non-zero code skips the table lookup for almost every slot.
// -------------------------------------------------------------------------------------------------
struct ReadParams {
    const void *bins;
    const unsigned char *pq;
    const unsigned *thr;
    const signed char *C;
    unsigned long long S_local;
    unsigned NB, d;
    HopFmt f;
    int *partial;               // [Q][d]
    unsigned *nsel;             // optional [Q]: number of selected slots (diagnostics)
};

template <typename BinT>
__global__ void __launch_bounds__(256) k_big_read(const ReadParams p)
{
    constexpr unsigned PER = 16 / sizeof(BinT);              // bins per 128-bit load
    constexpr unsigned BITS = 8 * sizeof(BinT), MASK = (1u << BITS) - 1u;
    const unsigned q = blockIdx.y, lane = threadIdx.x & 31;
    const BinT *b = reinterpret_cast<const BinT *>(p.bins) + (size_t)q * p.S_local;
    const unsigned char *pq = p.pq + (size_t)q * p.NB;
    const unsigned thr = p.thr[q];
    if (thr >= p.NB) return;
    // the selected slots of one query are at most 2^frac: every lane scans PER bins per 128-bit load, rejects them with
    // one packed compare per word when the bins are bytes, and the warp gathers the C row of each (rare) hit together
    const bool vec = (p.S_local % PER == 0) && ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
    const unsigned long long n_items = vec ? p.S_local / PER : p.S_local;      // work items: vectors or single bins
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long n_iter = (n_items + stride - 1) / stride;
    // byte compare v >= thr on four packed bytes: high bit of ((v & 0x7F) + K) combined with the byte's own high bit
    const unsigned K = (thr <= 128u) ? (128u - thr) * 0x01010101u : (256u - thr) * 0x01010101u;
    // four 128-bit loads per thread are issued before the first is examined (the scan is a pure stream: a thread with one load in
    // flight per iteration left the kernel latency-bound at 2.7 TB/s)
    for (unsigned long long it0 = 0; it0 < n_iter; it0 += 4) {
      uint4 pre[4];
#pragma unroll
      for (unsigned u = 0; u < 4; u++) {
          const unsigned long long i = (it0 + u) * stride + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
          pre[u] = make_uint4(0u, 0u, 0u, 0u);
          if (vec && it0 + u < n_iter && i < n_items) pre[u] = __ldg(reinterpret_cast<const uint4 *>(b) + i);
      }
#pragma unroll
      for (unsigned u = 0; u < 4; u++) {
        if (it0 + u >= n_iter) break;                         // block-uniform
        const unsigned long long i = (it0 + u) * stride + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        unsigned w[4] = {0u, 0u, 0u, 0u};
        unsigned hitmask = 0;                                 // bit k: bin k of this item is selected
        if (i < n_items) {
            if (vec) {
                const uint4 v4 = pre[u];
                w[0] = v4.x; w[1] = v4.y; w[2] = v4.z; w[3] = v4.w;
                bool maybe = true;
                if (sizeof(BinT) == 1) {
                    unsigned any = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const unsigned lo = (w[k] & 0x7F7F7F7Fu) + K;
                        any |= (thr <= 128u) ? (lo | w[k]) : (lo & w[k]);
                    }
                    maybe = (any & 0x80808080u) != 0u;
                }
                if (maybe) {
#pragma unroll
                    for (unsigned k = 0; k < PER; k++) {
                        const unsigned v = (w[k * BITS / 32] >> ((k * BITS) % 32)) & MASK;
                        if (v >= thr && pq[v]) hitmask |= 1u << k;
                    }
                }
            } else {
                w[0] = (unsigned)b[i];
                if (w[0] >= thr && pq[w[0]]) hitmask = 1u;
            }
        }
        unsigned lanes = __ballot_sync(0xffffffffu, hitmask != 0u);
        while (lanes) {
            const unsigned src = (unsigned)(__ffs((int)lanes) - 1);
            lanes &= lanes - 1u;
            unsigned hm = __shfl_sync(0xffffffffu, hitmask, src);
            const unsigned long long item = __shfl_sync(0xffffffffu, i, src);
            while (hm) {
                const unsigned k = (unsigned)(__ffs((int)hm) - 1);
                hm &= hm - 1u;
                // the bin value sits in the source lane's registers
                const unsigned wi = k * BITS / 32;            // warp-uniform
                const unsigned word = __shfl_sync(0xffffffffu, wi == 0 ? w[0] : (wi == 1 ? w[1] : (wi == 2 ? w[2] : w[3])), src);
                const unsigned v = (word >> ((k * BITS) % 32)) & MASK;
                const int pc = (int)pq[v];
                const unsigned long long slot = vec ? item * PER + k : item;
                const signed char *crow = p.C + slot * p.d;
                for (unsigned c = lane; c < p.d; c += 32) {
                    const int c_f = qi_requant((int)crow[c], p.f.fw, p.f.lf, p.f.ff);
                    const int term = qi_mul(pc, c_f, p.f.lf, p.f.ff);
                    if (term) atomicAdd(&p.partial[(size_t)q * p.d + c], term);
                }
                if (p.nsel && lane == 0) atomicAdd(&p.nsel[q], 1u);
            }
        }
      }
    }
}

// -------------------------------------------------------------------------------------------------
// k_big_update: one CTA per query.  o = Q_f(partial sum); g = Q_w(sum_j Q_w(Q_w(Hm[i][j]) * Q_bin(u[j])));
// u' = Q_f(Q_f(g) + Q_f(o))                                   MemN2N.c:873,889; layer_cuda.cu:49-68, 1535
// -------------------------------------------------------------------------------------------------
struct UpdateParams {
    const int *partial;          // [Q][d] (summed over ranks)
    const signed char *u_in;     // [Q][d], fu fractional bits
    signed char *u_out;          // [Q][d], ff fractional bits
    const int *ub;               // [Q][d] Q_bin(u)
    const signed char *Hq;       // [d][d] codes of Hm in the hop's weight format (NULL: no linear map)
    unsigned d;
    int fu;
    HopFmt f;
    signed char *dbg_o, *dbg_g;  // optional [Q][d]
};

__global__ void __launch_bounds__(256) k_big_update(const UpdateParams p)
{
    // one CTA per query; a warp per output dim walks its Hm row with coalesced byte loads (the sum of the quantised
    // products is an exact integer, so the lanes' partial sums may be added in any order)
    extern __shared__ int ub_s[];
    const unsigned q = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const HopFmt f = p.f;
    const unsigned d = p.d;
    if (p.Hq) {
        for (unsigned j = threadIdx.x; j < d; j += blockDim.x) ub_s[j] = p.ub[(size_t)q * d + j];
        __syncthreads();
    }
    for (unsigned i = blockIdx.y * nw + wid; i < d; i += gridDim.y * nw) {
        int g_w, a_f;
        if (p.Hq) {
            int s = 0;
            const signed char *hrow = p.Hq + (size_t)i * d;
            for (unsigned j = lane; j < d; j += 32) s += qi_mul((int)hrow[j], ub_s[j], f.lw, f.fb);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            g_w = qi_clamp(s, f.lw);
            a_f = qi_requant(g_w, f.fw, f.lf, f.ff);
        } else {
            g_w = (int)p.u_in[(size_t)q * d + i];
            a_f = qi_requant(g_w, p.fu, f.lf, f.ff);
        }
        if (lane == 0) {
            const int o = qi_clamp(p.partial[(size_t)q * d + i], f.lf);
            p.u_out[(size_t)q * d + i] = (signed char)qi_clamp(a_f + o, f.lf);
            if (p.dbg_o) p.dbg_o[(size_t)q * d + i] = (signed char)o;
            if (p.dbg_g) p.dbg_g[(size_t)q * d + i] = (signed char)g_w;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// k_big_update_fast (frac_bin == 2, d % 16 == 0, linear map on): the linear map with the packed identity of k_big_scores_fast --
// one dp4a and ~12 integer instructions per four products instead of ~8 per product.  A thread owns one (query, output dim):
// it walks its Hm row and the query (shared memory) as 128-bit vectors.  A 16-dim vector in which some row of the warp may have a
// saturating product (rowmax(Hm[i]) * max|Q_bin(u)| over the vector > 4 lw + 3) is evaluated product by product from the same registers.  Every rank repeats this kernel for all queries,
// so on 8 GPUs it was 13 % of a 1024-query step.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_big_update_fast(const UpdateParams p, const signed char *__restrict__ ub8, const unsigned char *__restrict__ hmax)
{
    extern __shared__ uint4 us4[];                              // Q_bin(u) of this query, d bytes; then max|u| per 16-byte vector
    const unsigned q = blockIdx.x, d = p.d, nv = d / 16;
    unsigned *um_s = reinterpret_cast<unsigned *>(us4 + nv);
    const HopFmt f = p.f;
    for (unsigned k = threadIdx.x; k < nv; k += blockDim.x) {
        const uint4 v = *reinterpret_cast<const uint4 *>(ub8 + (size_t)q * d + 16u * k);
        us4[k] = v;
        const unsigned a4 = __vmaxu4(__vmaxu4(__vabs4(v.x), __vabs4(v.y)), __vmaxu4(__vabs4(v.z), __vabs4(v.w)));
        um_s[k] = max(max(a4 & 0xFFu, (a4 >> 8) & 0xFFu), max((a4 >> 16) & 0xFFu, a4 >> 24));
    }
    __syncthreads();
    const unsigned sat_lim = 4u * (unsigned)f.lw + 3u;
    const unsigned i = blockIdx.y * blockDim.x + threadIdx.x;
    const bool valid = i < d;
    const unsigned ir = valid ? i : d - 1u;
    const unsigned hm = valid ? (unsigned)hmax[ir] : 0u;
    const uint4 *hrow = reinterpret_cast<const uint4 *>(p.Hq + (size_t)ir * d);
    int acc4 = 0;                                               // 4 x the sum of the truncated products
#pragma unroll 2
    for (unsigned k = 0; k < nv; k++) {
        const uint4 hv = __ldg(hrow + k), uv = us4[k];
        const unsigned yw[4] = {hv.x, hv.y, hv.z, hv.w}, uw[4] = {uv.x, uv.y, uv.z, uv.w};
        // a product of this vector may saturate for some row of the warp: the reference order of operations for the whole warp (the
        // exact form is valid for every lane; deciding per warp keeps the branch uniform)
        if (__any_sync(0xffffffffu, hm * um_s[k] > sat_lim)) {
            int sp = 0;
#pragma unroll
            for (int w = 0; w < 4; w++)
#pragma unroll
                for (int b8 = 0; b8 < 4; b8++) sp += qi_mul(sx8(yw[w], b8), sx8(uw[w], b8), f.lw, f.fb);
            acc4 += 4 * sp;
        } else {
            int D = 0;
            unsigned cs = 0;
#pragma unroll
            for (int w = 0; w < 4; w++) {
                D = __dp4a((int)yw[w], (int)uw[w], D);
                const unsigned u0 = uw[w] & 0x01010101u;
                const unsigned t0 = (yw[w] << 1) & (uw[w] & 0x02020202u);
                const unsigned t1 = (yw[w] & (u0 << 1)) ^ t0;
                const unsigned bm = (yw[w] & u0) | t1;                            // x mod 4 per byte
                const unsigned wv = bm + 0x03030303u;                             // 3..6, bit 2 set iff x mod 4 != 0
                cs += wv & ~(((yw[w] ^ uw[w]) >> 5) & 0x04040404u);               // clear bit 2 where the product is negative
            }
            acc4 += D - (int)__dp4a(cs, 0x01010101u, 0u) + 48;
        }
    }
    if (!valid) return;
    const int g_w = qi_clamp(acc4 >> 2, f.lw);                                    // exact: a multiple of 4
    const int a_f = qi_requant(g_w, f.fw, f.lf, f.ff);
    const int o = qi_clamp(p.partial[(size_t)q * d + i], f.lf);
    p.u_out[(size_t)q * d + i] = (signed char)qi_clamp(a_f + o, f.lf);
    if (p.dbg_o) p.dbg_o[(size_t)q * d + i] = (signed char)o;
    if (p.dbg_g) p.dbg_g[(size_t)q * d + i] = (signed char)g_w;
}

// hmax[i] = max_j |Hq[i][j]|: one warp per row
__global__ void __launch_bounds__(256) k_big_rowmax_H(const signed char *__restrict__ Hq, unsigned d, unsigned char *__restrict__ hmax)
{
    const unsigned row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= d) return;
    unsigned mx = 0;
    for (unsigned j = lane; j < d; j += 32) mx = max(mx, (unsigned)abs((int)Hq[(size_t)row * d + j]));
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) hmax[row] = (unsigned char)min(mx, 255u);
}

__global__ void k_big_quant_H(const float *__restrict__ w, signed char *__restrict__ out, unsigned n, int iwl, int frac)
{
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = (signed char)qi_encode(w[i], iwl, frac);
}

// -------------------------------------------------------------------------------------------------
// k_big_answer: one warp per query.  z[i] = sum_j fl(W[i][j] * u[j]) sequential fp32 (MemN2N.c:902-906,
// layer_cuda.cu:69-82), softmax with the sequential double total, argmax_last on the probabilities
// (layer_cuda.cu:1918-1939).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_big_answer(const float *__restrict__ W, const signed char *__restrict__ u, int fu, unsigned Q, unsigned V,
                                                    unsigned d, float *__restrict__ zbuf, unsigned *__restrict__ pred, float *__restrict__ hout)
{
    // one warp per query (4 per CTA); u as fp32 in shared memory, W rows read with 128-bit loads (d % 16 == 0), the
    // products added in index order exactly like the reference's thread 0 does
    extern __shared__ __align__(16) float us_all[];
    const unsigned wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned q = blockIdx.x * (blockDim.x >> 5) + wid;
    if (q >= Q) return;
    float *us = us_all + (size_t)wid * d;
    const float inv = 1.0f / (float)(1 << fu);
    for (unsigned j = lane; j < d; j += 32) us[j] = (float)u[(size_t)q * d + j] * inv;
    __syncwarp();
    float *z = zbuf + (size_t)q * V;
    float zmax = -INFINITY;
    for (unsigned i = lane; i < V; i += 32) {
        float acc = 0.0f;
        const float4 *wr = reinterpret_cast<const float4 *>(W + (size_t)i * d);
#pragma unroll 16
        for (unsigned j4 = 0; j4 < d / 4; j4++) {
            const float4 w4 = __ldg(wr + j4);
            const float4 u4 = *reinterpret_cast<const float4 *>(us + 4 * j4);
            acc = __fadd_rn(acc, __fmul_rn(w4.x, u4.x));
            acc = __fadd_rn(acc, __fmul_rn(w4.y, u4.y));
            acc = __fadd_rn(acc, __fmul_rn(w4.z, u4.z));
            acc = __fadd_rn(acc, __fmul_rn(w4.w, u4.w));
        }
        z[i] = acc;
        zmax = fmaxf(zmax, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
    __syncwarp();
    double total = 0.0;
    for (unsigned i = 0; i < V; i++) total += (double)__expf(z[i] - zmax);
    float best = -INFINITY;
    unsigned best_i = 0;
    for (unsigned i = lane; i < V; i += 32) {
        const float hv = (float)((double)__expf(z[i] - zmax) / total);
        if (hout) hout[(size_t)q * V + i] = hv;
        if (!(best > hv)) { best = hv; best_i = i; }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const unsigned oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi > best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0 && pred) pred[q] = best_i;
}

}  // namespace

// =================================================================================================
// host object
// =================================================================================================
struct qmann_bigmem {
    qmann_config cfg;
    int sm_count;
    const signed char *M[MAXH], *C[MAXH];
    unsigned long long S_total, slot0, S_local;
    unsigned Q_max, Q, NB, bias;
    HopFmt f[MAXH];
    signed char *dev_H[MAXH];
    int *u9;                    // mode 3: nine-bit query operands, per-query quirk flags
    unsigned *quirk;
    bool nine_on, ham_off;      // QMANN_BIGMEM_FAST=0: literal mode-3 kernel; QMANN_BIGMEM_HAM=0: nine-bit scalar form instead of packed bytes
    unsigned char *dev_hmax[MAXH];      // row maxima of dev_H (k_big_update_fast)
    bool fast_update;
    const float *dev_W;
    // work buffers
    signed char *u_a, *u_b;          // ping-pong [Q_max][d]
    int fu;
    int *ub;
    unsigned *av, *sv;
    signed char *ub8;                // [Q_max][d] Q_bin(u) as bytes (k_big_scores_fast)
    unsigned *umax;                  // [Q_max] max |Q_bin(u)|
    uint4 *bfrag;                    // k_big_scores_mma query planes, fragment order
    bool mma_ok;
    // qmann_bigmem_forward_sharded: exchange buffers and the captured graph of one whole forward
    uint32_t *xhist;
    int32_t *xpartial;
    cudaGraphExec_t gexec;
    const void *g_u0, *g_pred, *g_comm;
    unsigned g_Q;
    const void *warm_comm;
    bool warm;
    // k_big_scores_tc (tcgen05): query planes as the B operand's shared-memory image, one tensor map of Y per hop
    uint4 *bplanes;
    bool tc_ok;
    cudaStream_t s2;            // histogram of query block b under the scorer of block b + 1 (qmann_bigmem_hop_scores)
    cudaEvent_t ev_blk, ev_join;
    int overlap;                // QMANN_BIGMEM_OVERLAP: 0 off (default), 1 large shards, 2 always
    bool tq_ok, tq_wide;        // tq_wide (QMANN_BIGMEM_TQ_WIDE=1): always one CTA per SM and query block                 // k_big_scores_tq: queries in tensor memory, 128 per pass, histogram fused (d <= 256, byte bins)
    CUtensorMap tmY[MAXH];
    // k_big_scores_fast inputs per hop: Y = Q_att(M) (== M when the re-quantisation is the identity), row maxima
    const signed char *Y[MAXH];
    signed char *Y_own[MAXH];
    unsigned char *rowmax[MAXH];
    bool fast[MAXH];
    void *bins;                      // [Q_max][S_local] uint8 (mode 2) or uint16 (mode 3)
    int bin8;
    unsigned char *pq;               // [Q_max][NB]
    unsigned *thr, *nsel;
    float *zbuf;
    int smem_optin;
    // optional timing of k_big_scores (qmann_bigmem_profile_*)
    bool profile;
    cudaEvent_t pev[2 * MAXH];
    unsigned pused;
    double pms;
    unsigned pcount;
};

template <int MODE, int QB>
static int launch_scores(qmann_bigmem *b, const ScoreParams &sp, cudaStream_t st)
{
    const unsigned d = sp.d, warps = 8;
    const size_t smem = (size_t)QB * d * 4 * (MODE == 3 ? 3 : 1) + (size_t)warps * 32 * (d + 16);
    if (smem > (size_t)b->smem_optin) return bfail(QMANN_E_NOMEM, "score tile does not fit shared memory");
    BCUDA(cudaFuncSetAttribute(k_big_scores<MODE, QB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned long long tiles = (sp.S_local + 31) / 32;
    const unsigned qblocks = (sp.Q + QB - 1) / QB;
    unsigned gx = (unsigned)std::min<unsigned long long>((tiles + warps - 1) / warps, (unsigned long long)b->sm_count * 2);
    gx = std::max(1u, gx);
    k_big_scores<MODE, QB><<<dim3(gx, qblocks), warps * 32, smem, st>>>(sp);
    count_launch();
    BCUDA(cudaPeekAtLastError());
    return QMANN_OK;
}

template <int QB>
static int launch_ham(qmann_bigmem *b, const HamParams &hp, cudaStream_t st)
{
    const unsigned d = hp.d, warps = 8;
    const size_t smem = (size_t)QB * d * 2 + (size_t)warps * 32 * (d + 16);
    if (smem > (size_t)b->smem_optin) return bfail(QMANN_E_NOMEM, "score tile does not fit shared memory");
    BCUDA(cudaFuncSetAttribute(k_big_scores_ham<QB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned long long tiles = (hp.S_local + 31) / 32;
    unsigned gx = (unsigned)std::min<unsigned long long>((tiles + warps - 1) / warps, (unsigned long long)b->sm_count * 2);
    gx = std::max(1u, gx);
    k_big_scores_ham<QB><<<dim3(gx, (hp.Q + QB - 1) / QB), warps * 32, smem, st>>>(hp);
    count_launch();
    BCUDA(cudaPeekAtLastError());
    return QMANN_OK;
}

template <int LPR>
static int launch_scores_fast(qmann_bigmem *b, const FastScoreParams &fp, cudaStream_t st)
{
    const unsigned long long tiles = (fp.S_local + 31) / 32;
    const bool q4 = fp.Q >= 3;
    const unsigned qblocks = q4 ? (fp.Q + 3) / 4 : fp.Q;
    // 8 warps per CTA, one 32-row tile per warp and iteration; enough CTAs to fill every SM several times over
    unsigned gx = (unsigned)std::min<unsigned long long>((tiles + 7) / 8, (unsigned long long)b->sm_count * 16);
    gx = std::max(1u, gx);
    if (q4) k_big_scores_fast<LPR, 4><<<dim3(gx, qblocks), 256, 0, st>>>(fp);
    else    k_big_scores_fast<LPR, 1><<<dim3(gx, qblocks), 256, 0, st>>>(fp);
    count_launch();
    BCUDA(cudaPeekAtLastError());
    return QMANN_OK;
}

static int dispatch_scores_fast(qmann_bigmem *b, const FastScoreParams &fp, cudaStream_t st)
{
    switch (fp.d / 16) {
    case 1: return launch_scores_fast<1>(b, fp, st);
    case 2: return launch_scores_fast<2>(b, fp, st);
    case 4: return launch_scores_fast<4>(b, fp, st);
    case 8: return launch_scores_fast<8>(b, fp, st);
    case 16: return launch_scores_fast<16>(b, fp, st);
    case 32: return launch_scores_fast<32>(b, fp, st);
    }
    return bfail(QMANN_E_ARG, "k_big_scores_fast: d/16 must be a power of two");
}

#ifdef QMANN_TC_TRACE
static unsigned long long *g_tcs_clk = nullptr;
extern "C" void qmann_bigmem_tc_trace_dump(void)
{
    if (!g_tcs_clk) return;
    const char *roles[4] = {"producer (wait empty | issue)", "mma (wait dfree | wait built | issue)", "builder (wait full | build)", "epilogue (wait dfull | tmem ld | stores | risky+loop)"};
    for (int r = 0; r < 4; r++) fprintf(stderr, "%-40s %12llu %12llu %12llu %12llu\n", roles[r], g_tcs_clk[4 * r], g_tcs_clk[4 * r + 1], g_tcs_clk[4 * r + 2], g_tcs_clk[4 * r + 3]);
}
#endif
namespace {
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda): byte matrix [rows][cols],
// boxes of 128 bytes x 128 rows, 128-byte swizzle -- the A operand tiles of k_big_scores_tc
bool tmap_bytes_rows(CUtensorMap *out, const void *base, unsigned long long rows, unsigned cols)
{
    typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static enc_fn enc = []() -> enc_fn {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        return (enc_fn)f;
    }();
    if (!enc || rows == 0 || ((uintptr_t)base % 16)) return false;
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols};
    cuuint32_t box[2] = {128, 128}, es[2] = {1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

extern "C" {

const char *qmann_bigmem_last_error(void) { return g_berr.c_str(); }

// inside qmann_bigmem_create, after the object exists: a CUDA failure releases whatever was allocated so far (an out-of-memory failure
// while creating a large shard must not leak GBs)
#define BCUDA_B(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) { qmann_bigmem_destroy(b); return bfail(QMANN_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } \
    } while (0)
void qmann_bigmem_destroy(qmann_bigmem *b);
int qmann_bigmem_create(qmann_bigmem **out, const qmann_config *cfg, const qmann_weights *w, const int8_t *const *dev_M,
                        const int8_t *const *dev_C, uint64_t S_total, uint64_t slot0, uint64_t S_local, uint32_t Q_max)
{
    if (!out || !cfg || !w || !dev_M || !dev_C) return bfail(QMANN_E_ARG, "null argument");
    *out = nullptr;
    const qmann_config &c = *cfg;
    if (c.H == 0 || c.H > MAXH) return bfail(QMANN_E_ARG, "H must be in 1..8");
    if (c.mode != 2 && c.mode != 3) return bfail(QMANN_E_ARG, "attention mode must be 2 or 3");
    if (c.d == 0 || c.d % 16 || c.d > 512) return bfail(QMANN_E_ARG, "d must be a multiple of 16 in 16..512");
    if (Q_max == 0 || Q_max > 65535) return bfail(QMANN_E_ARG, "Q_max must be in 1..65535");
    if (slot0 + S_local > S_total) return bfail(QMANN_E_ARG, "shard exceeds the memory");
    if (w->dev_W && ((uintptr_t)w->dev_W % 16)) return bfail(QMANN_E_ARG, "dev_W must be 16-byte aligned");
    auto okfmt = [](unsigned i, unsigned f) { return i + f >= 1 && i + f <= 7; };
    for (unsigned h = 0; h < c.H; h++)
        if (!okfmt(c.iwl[h], c.frac[h]) || !okfmt(c.iwl_w[h], c.frac_w[h]) || !okfmt(c.iwl_att[h], c.frac_att[h]))
            return bfail(QMANN_E_ARG, "formats must satisfy 1 <= iwl+frac <= 7");
    if (!okfmt(c.iwl_bin, c.frac_bin)) return bfail(QMANN_E_ARG, "bin format must satisfy 1 <= iwl+frac <= 7");
    if (c.mode == 3 && (c.const_scale > 0 || c.const_scale < -16)) return bfail(QMANN_E_ARG, "const_scale must be in -16..0");
    for (unsigned h = 0; h < c.H; h++)
        if ((S_local && (!dev_M[h] || !dev_C[h])) || (c.lin_map && !w->dev_Hm[h])) return bfail(QMANN_E_ARG, "missing per-hop pointer");

    qmann_bigmem *b = new qmann_bigmem();
    memset(b, 0, sizeof(*b));
    b->cfg = c;
    int dev = 0;
    BCUDA_B(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    BCUDA_B(cudaGetDeviceProperties(&prop, dev));
    b->sm_count = prop.multiProcessorCount;
    BCUDA_B(cudaDeviceGetAttribute(&b->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    b->S_total = S_total; b->slot0 = slot0; b->S_local = S_local; b->Q_max = Q_max;
    unsigned la_max = 0;
    for (unsigned h = 0; h < c.H; h++) {
        HopFmt &f = b->f[h];
        f.fw = c.frac_w[h]; f.lw = fixed_max(c.iwl_w[h], c.frac_w[h]);
        f.fa = c.frac_att[h]; f.ia = c.iwl_att[h]; f.la = fixed_max(c.iwl_att[h], c.frac_att[h]);
        f.ff = c.frac[h]; f.iff = c.iwl[h]; f.lf = fixed_max(c.iwl[h], c.frac[h]);
        f.fb = c.frac_bin; f.lb = fixed_max(c.iwl_bin, c.frac_bin);
        la_max = std::max(la_max, (unsigned)f.la);
        b->M[h] = dev_M[h]; b->C[h] = dev_C[h];
    }
    b->bias = (c.mode == 3) ? 127u * c.d : la_max;
    b->NB = 2 * b->bias + 1;
    if (b->NB > 65536) { qmann_bigmem_destroy(b); return bfail(QMANN_E_ARG, "score range does not fit 16-bit bins (mode 3 needs d <= 258)"); }
    b->dev_W = w->dev_W;
    const size_t Qd = (size_t)Q_max * c.d;
    BCUDA_B(cudaMalloc((void **)&b->u_a, Qd));
    BCUDA_B(cudaMalloc((void **)&b->u_b, Qd));
    BCUDA_B(cudaMalloc((void **)&b->ub, Qd * 4));
    BCUDA_B(cudaMalloc((void **)&b->av, Qd * 4));
    BCUDA_B(cudaMalloc((void **)&b->sv, Qd * 4));
    BCUDA_B(cudaMalloc((void **)&b->u9, Qd * 4));
    BCUDA_B(cudaMalloc((void **)&b->quirk, (size_t)Q_max * 4));
    BCUDA_B(cudaMalloc((void **)&b->ub8, Qd));
    BCUDA_B(cudaMalloc((void **)&b->umax, (size_t)Q_max * 4));
    {
        // tensor-core scorer: d a multiple of 64 whose query planes (d * 256 bytes per block of 64 queries) fit shared memory
        const char *env_mma = getenv("QMANN_BIGMEM_MMA");
        const size_t frag_bytes = (size_t)c.d * 256;
        b->mma_ok = c.mode == 2 && c.frac_bin == 2 && c.d % 64 == 0 && frag_bytes + 256 + (size_t)2 * 2 * 16 * c.d <= (size_t)b->smem_optin &&
                    !(env_mma && atoi(env_mma) == 0);
        if (b->mma_ok) BCUDA_B(cudaMalloc((void **)&b->bfrag, (size_t)((Q_max + MMA_QB - 1) / MMA_QB) * frag_bytes));
    }
    b->bin8 = (b->NB <= 256) ? 1 : 0;                                  // mode 2: 2*127+1 bins fit a byte
    BCUDA_B(cudaMalloc((void **)&b->bins, std::max<size_t>(16, (size_t)Q_max * S_local * (b->bin8 ? 1 : 2))));
    BCUDA_B(cudaMalloc((void **)&b->pq, (size_t)Q_max * b->NB));
    BCUDA_B(cudaMalloc((void **)&b->thr, (size_t)Q_max * 4));
    BCUDA_B(cudaMalloc((void **)&b->nsel, (size_t)Q_max * 4));
    if (c.V) BCUDA_B(cudaMalloc((void **)&b->zbuf, (size_t)Q_max * c.V * 4));
    for (unsigned h = 0; h < c.H; h++) {
        if (!c.lin_map) continue;
        BCUDA_B(cudaMalloc((void **)&b->dev_H[h], (size_t)c.d * c.d));
        k_big_quant_H<<<64, 256>>>(w->dev_Hm[h], b->dev_H[h], c.d * c.d, c.iwl_w[h], c.frac_w[h]);
        BCUDA_B(cudaMalloc((void **)&b->dev_hmax[h], c.d));
        k_big_rowmax_H<<<(c.d + 7) / 8, 256>>>(b->dev_H[h], c.d, b->dev_hmax[h]);
        count_launch();
    }
    // k_big_scores_fast (mode 2, two fractional bits in the query format, d/16 a power of two): one pass over each
    // hop's M for the row maxima, and a re-quantised copy only where Q_att(M) differs from M.  The caller's memory
    // must not change while this object lives.  QMANN_BIGMEM_FAST=0 keeps the per-product kernel (A/B tests).
    const unsigned c16 = c.d / 16;
    const char *env_fast = getenv("QMANN_BIGMEM_FAST");
    const bool want_fast = c.mode == 2 && c.frac_bin == 2 && (c16 & (c16 - 1)) == 0 && c16 <= 32 && S_local > 0 &&
                           !(env_fast && atoi(env_fast) == 0);
    b->nine_on = !(env_fast && atoi(env_fast) == 0);
    { const char *env_ham = getenv("QMANN_BIGMEM_HAM"); b->ham_off = env_ham && atoi(env_ham) == 0; }
    b->fast_update = c.mode == 2 && c.frac_bin == 2 && c.d % 16 == 0 && c.lin_map && !(env_fast && atoi(env_fast) == 0);
    if (want_fast) {
        unsigned *dev_flag = nullptr;
        BCUDA_B(cudaMalloc((void **)&dev_flag, 4));
        const unsigned gx = (unsigned)std::min<unsigned long long>((S_local + 7) / 8, (unsigned long long)b->sm_count * 16);
        for (unsigned h = 0; h < c.H; h++) {
            BCUDA_B(cudaMalloc((void **)&b->rowmax[h], S_local));
            BCUDA_B(cudaMemset(dev_flag, 0, 4));
            k_big_prep_mem<<<std::max(1u, gx), 256>>>(b->M[h], S_local, c.d, b->f[h], nullptr, b->rowmax[h], dev_flag);
            count_launch();
            unsigned flag = 0;
            BCUDA_B(cudaMemcpy(&flag, dev_flag, 4, cudaMemcpyDeviceToHost));
            b->Y[h] = b->M[h];
            if (flag) {
                if (cudaMalloc((void **)&b->Y_own[h], (size_t)S_local * c.d) != cudaSuccess) {
                    cudaGetLastError();                      // no room for the copy: this hop keeps the per-product kernel
                    b->Y_own[h] = nullptr;
                    continue;
                }
                k_big_prep_mem<<<std::max(1u, gx), 256>>>(b->M[h], S_local, c.d, b->f[h], b->Y_own[h], b->rowmax[h], dev_flag);
                count_launch();
                b->Y[h] = b->Y_own[h];
            }
            b->fast[h] = true;
        }
        BCUDA_B(cudaDeviceSynchronize());
        cudaFree(dev_flag);
    }
    {
        // tcgen05 scorer: every fast hop's Y as a TMA tensor, the B planes' image; QMANN_BIGMEM_TC=0 keeps the mma.sync kernel
        const char *env_tc = getenv("QMANN_BIGMEM_TC");
        const size_t need = (size_t)(c.d / 128) * 4 * 8192 + (size_t)(TCS_NY + 3 * TCS_NP) * TCS_TILE + 1024 + 512 + 2 * 8192;
        bool ok = b->mma_ok && c.d % 128 == 0 && c.d >= 128 && need <= (size_t)b->smem_optin && !(env_tc && atoi(env_tc) == 0);
        for (unsigned h = 0; h < c.H && ok; h++)
            if (b->fast[h]) ok = tmap_bytes_rows(&b->tmY[h], b->Y[h], S_local, c.d);
        if (ok) BCUDA_B(cudaMalloc((void **)&b->bplanes, (size_t)((Q_max + TCS_QB - 1) / TCS_QB) * (c.d / 128) * 4 * 8192));
        b->tc_ok = ok;
        const char *env_tq = getenv("QMANN_BIGMEM_TQ");
        const size_t need_tq = (size_t)(TCS_NY + 3 * 3) * TCS_TILE + 2048;
        { const char *env_w = getenv("QMANN_BIGMEM_TQ_WIDE"); b->tq_wide = env_w && atoi(env_w) != 0; }
        b->tq_ok = ok && c.d <= 256 && b->bin8 && c.mode == 2 && b->NB <= 256 && need_tq <= (size_t)b->smem_optin && !(env_tq && atoi(env_tq) == 0);
    }
    {
        const char *env_ov = getenv("QMANN_BIGMEM_OVERLAP");
        b->overlap = env_ov ? atoi(env_ov) : 0;          // opt-in: measured slower (see qmann_bigmem_hop_scores)
        if (b->overlap && b->tq_ok) {
            BCUDA_B(cudaStreamCreateWithFlags(&b->s2, cudaStreamNonBlocking));
            BCUDA_B(cudaEventCreateWithFlags(&b->ev_blk, cudaEventDisableTiming));
            BCUDA_B(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
        }
    }
    BCUDA_B(cudaDeviceSynchronize());
    *out = b;
    return QMANN_OK;
}

void qmann_bigmem_destroy(qmann_bigmem *b)
{
    if (!b) return;
    cudaFree(b->u_a); cudaFree(b->u_b); cudaFree(b->ub); cudaFree(b->av); cudaFree(b->sv); cudaFree(b->u9); cudaFree(b->quirk); cudaFree(b->bins);
    cudaFree(b->ub8); cudaFree(b->umax); cudaFree(b->bfrag); cudaFree(b->bplanes);
    cudaFree(b->xhist); cudaFree(b->xpartial);
    if (b->gexec) cudaGraphExecDestroy(b->gexec);
    if (b->ev_blk) cudaEventDestroy(b->ev_blk);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    if (b->s2) cudaStreamDestroy(b->s2);
    for (int h = 0; h < MAXH; h++) { cudaFree(b->Y_own[h]); cudaFree(b->rowmax[h]); }
    cudaFree(b->pq); cudaFree(b->thr); cudaFree(b->nsel); cudaFree(b->zbuf);
    for (int h = 0; h < MAXH; h++) { cudaFree(b->dev_H[h]); cudaFree(b->dev_hmax[h]); }
    if (b->pev[0]) for (int i = 0; i < 2 * MAXH; i++) cudaEventDestroy(b->pev[i]);
    delete b;
}

uint32_t qmann_bigmem_num_bins(const qmann_bigmem *b) { return b ? b->NB : 0; }

int qmann_bigmem_begin(qmann_bigmem *b, const int8_t *dev_u0, uint32_t Q, void *stream)
{
    if (!b || !dev_u0) return bfail(QMANN_E_ARG, "null argument");
    if (Q == 0 || Q > b->Q_max) return bfail(QMANN_E_ARG, "Q must be in 1..Q_max");
    b->Q = Q;
    b->fu = (int)b->cfg.frac_w[0];        // emb_q outputs in the hop-0 weight format (MemN2N.c:826)
    BCUDA(cudaMemcpyAsync(b->u_a, dev_u0, (size_t)Q * b->cfg.d, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return QMANN_OK;
}

int qmann_bigmem_hop_scores(qmann_bigmem *b, uint32_t h, uint32_t *dev_hist, void *stream)
{
    if (!b || !dev_hist) return bfail(QMANN_E_ARG, "null argument");
    if (h >= b->cfg.H || b->Q == 0) return bfail(QMANN_E_ARG, "bad hop or no query batch (call qmann_bigmem_begin)");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned Q = b->Q, d = b->cfg.d;
    const HopFmt &f = b->f[h];
    // mode 3, nine-bit form: A = sat9(code << (8 - iwl_att - frac)); usable when both shifts are in 0..8 (QMANN_BIGMEM_FAST=0: literal)
    const int sh_m = 8 - f.ia - f.fw, sh_u = 8 - f.ia - b->fu;
    const bool nine = b->cfg.mode == 3 && b->nine_on && sh_m >= 0 && sh_m <= 8 && sh_u >= 0 && sh_u <= 8;
    k_big_prep_query<<<Q, 256, 0, st>>>(b->u_a, b->fu, Q, d, (int)b->cfg.mode, f, b->ub, b->av, b->sv, b->ub8, b->umax, b->u9, b->quirk, nine ? sh_u : -1);
    count_launch();
    BCUDA(cudaMemsetAsync(dev_hist, 0, (size_t)Q * b->NB * 4, st));
    if (b->S_local) {
        ScoreParams sp;
        sp.M = b->M[h]; sp.S_local = b->S_local; sp.d = d; sp.Q = Q; sp.f = f; sp.const_scale = b->cfg.const_scale;
        sp.ub = b->ub; sp.av = b->av; sp.sv = b->sv; sp.bins = b->bins; sp.bin8 = b->bin8;
        sp.u9 = b->u9; sp.quirk = b->quirk; sp.nine = nine ? 1 : 0; sp.sh_m = sh_m; sp.only_quirk = 0;
        sp.bias = (b->cfg.mode == 3) ? b->bias : (unsigned)f.la;
        int rc;
        const bool prof = b->profile && b->pused + 2 <= 2 * MAXH;
        if (prof) BCUDA(cudaEventRecord(b->pev[b->pused], st));
        bool hist_done = false;
        if (b->fast[h] && b->tq_ok && Q >= 4) {
            TqScoreParams tp;
            tp.tmY = b->tmY[h];
            tp.Y = b->Y[h]; tp.rowmax = b->rowmax[h]; tp.S_local = b->S_local; tp.d = d; tp.Q = Q; tp.la = f.la; tp.fb = f.fb;
            tp.ub8 = b->ub8; tp.umax = b->umax; tp.bins = reinterpret_cast<unsigned char *>(b->bins); tp.bias = (unsigned)f.la; tp.qblk0 = 0;
            tp.clk = nullptr;
#ifdef QMANN_TC_TRACE
            static unsigned long long *clk_host_q = nullptr;
            if (!clk_host_q) BCUDA(cudaHostAlloc((void **)&clk_host_q, 16 * 8, cudaHostAllocMapped));
            { unsigned long long *dp = nullptr; BCUDA(cudaHostGetDevicePointer((void **)&dp, clk_host_q, 0)); tp.clk = dp; }
            g_tcs_clk = clk_host_q;
#endif
            static bool attr_done = false;
            if (!attr_done) {
                BCUDA(cudaFuncSetAttribute(k_big_scores_tq<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, b->smem_optin));
                BCUDA(cudaFuncSetAttribute(k_big_scores_tq<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, b->smem_optin));
                BCUDA(cudaFuncSetAttribute(k_big_hist_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)HL_WARPS * 256 * 32 * 2)));
                attr_done = true;
            }
            const unsigned long long tiles = (b->S_local + 127) / 128;
            const unsigned nblk = (Q + TQ_QB - 1) / TQ_QB;
            // One CTA per SM at a time.  With several query blocks over a small shard (multi-GPU) a CTA per SM and block would see only
            // a handful of tiles, and every CTA pays for writing its query planes into tensor memory and for filling the pipeline:
            // use fewer, longer CTAs per block -- gx * nblk = waves * SMs with at least ~24 tiles per CTA where the work allows it.
            unsigned gx = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>(tiles, (unsigned long long)b->sm_count));
            if (nblk >= 2 && !b->tq_wide) {
                const unsigned long long waves = std::max<unsigned long long>(1, std::min<unsigned long long>(nblk, tiles * nblk / ((unsigned long long)b->sm_count * 24)));
                gx = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>(tiles, (unsigned long long)b->sm_count * waves / nblk));
            }
            // Opt-in (QMANN_BIGMEM_OVERLAP=1: shards of >= 2^18 slots, 2: always): one scorer launch per query block, the histogram of
            // block b on a second stream under the scorer of block b + 1; the scorer gives up one plane stage (145 KB) so that a 4-warp
            // histogram CTA (64 KB) fits on the same SM.  Measured at Q = 1024 on one GPU: 1.77 ms per hop against 1.63 ms for the
            // two kernels back to back -- both live on the shared-memory pipe (plane construction + operand reads, counter
            // load/stores), so running them together only slows the scorer.  Kept for the record and covered by the tests.
            const bool ovl = b->s2 && nblk >= 2 && b->bin8 && b->NB <= 256 && (b->overlap >= 2 || (b->overlap == 1 && b->S_local >= (1ull << 18)));
            if (ovl) {
                const size_t smem = (size_t)(TCS_NY + 3 * 2) * TCS_TILE + 2048, hl_smem = (size_t)4 * 256 * 32 * 2;
                for (unsigned blk = 0; blk < nblk; blk++) {
                    tp.qblk0 = blk;
                    k_big_scores_tq<2><<<dim3(gx, 1), TCS_WARPS * 32, smem, st>>>(tp);
                    count_launch();
                    BCUDA(cudaEventRecord(b->ev_blk, st));
                    BCUDA(cudaStreamWaitEvent(b->s2, b->ev_blk, 0));
                    const unsigned qb0 = blk * TQ_QB, qn = std::min(TQ_QB, Q - qb0);
                    k_big_hist_lanes<<<3 * b->sm_count, 4 * 32, hl_smem, b->s2>>>(reinterpret_cast<const unsigned char *>(b->bins) + (size_t)qb0 * b->S_local, b->S_local, qn,
                                                                                 b->NB, 16384u, dev_hist + (size_t)qb0 * b->NB);
                    count_launch();
                }
                BCUDA(cudaEventRecord(b->ev_join, b->s2));
                BCUDA(cudaStreamWaitEvent(st, b->ev_join, 0));
                hist_done = true;
            } else {
                const size_t smem = (size_t)(TCS_NY + 3 * 3) * TCS_TILE + 2048;
                k_big_scores_tq<3><<<dim3(gx, nblk), TCS_WARPS * 32, smem, st>>>(tp);
                count_launch();
            }
            BCUDA(cudaPeekAtLastError());
            rc = QMANN_OK;
        }
        else if (b->fast[h] && b->tc_ok && Q >= 4) {
            const unsigned qblocks = (Q + TCS_QB - 1) / TCS_QB, KC = d / 128;
            k_big_prep_bplanes<<<std::min(512u, (qblocks * KC * 2048u + 255u) / 256u), 256, 0, st>>>(b->ub8, Q, d, b->bplanes);
            count_launch();
            TcScoreParams tp;
            tp.tmY = b->tmY[h];
            tp.Y = b->Y[h]; tp.rowmax = b->rowmax[h]; tp.S_local = b->S_local; tp.d = d; tp.Q = Q; tp.la = f.la; tp.fb = f.fb;
            tp.ub8 = b->ub8; tp.umax = b->umax; tp.bplanes = b->bplanes; tp.bins = b->bins; tp.bin8 = b->bin8; tp.bias = (unsigned)f.la;
            tp.clk = nullptr;
#ifdef QMANN_TC_TRACE
            static unsigned long long *clk_host = nullptr;
            if (!clk_host) BCUDA(cudaHostAlloc((void **)&clk_host, 16 * 8, cudaHostAllocMapped));
            { unsigned long long *dp = nullptr; BCUDA(cudaHostGetDevicePointer((void **)&dp, clk_host, 0)); tp.clk = dp; }
            g_tcs_clk = clk_host;
#endif
            const size_t smem = (size_t)KC * 4 * 8192 + (size_t)(TCS_NY + 3 * TCS_NP) * TCS_TILE + 1024 + 512 + 2 * 8192;
            static bool attr_done = false;
            if (!attr_done) { BCUDA(cudaFuncSetAttribute(k_big_scores_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, b->smem_optin)); attr_done = true; }
            const unsigned long long tiles = (b->S_local + 127) / 128;
            const unsigned gx = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>(tiles, (unsigned long long)b->sm_count));
            k_big_scores_tc<<<dim3(gx, qblocks), TCS_WARPS * 32, smem, st>>>(tp);
            count_launch();
            BCUDA(cudaPeekAtLastError());
            rc = QMANN_OK;
        }
        else if (b->fast[h] && b->mma_ok && Q >= 4) {
            const unsigned qblocks = (Q + MMA_QB - 1) / MMA_QB;
            const unsigned frag_vec = (d / 32) * 8 * 2 * 32;
            k_big_prep_bfrag<<<std::min(256u, (qblocks * frag_vec + 255) / 256), 256, 0, st>>>(b->ub8, Q, d, b->bfrag);
            count_launch();
            MmaScoreParams mp;
            mp.Y = b->Y[h]; mp.rowmax = b->rowmax[h]; mp.S_local = b->S_local; mp.d = d; mp.Q = Q; mp.la = f.la; mp.fb = f.fb;
            mp.ub8 = b->ub8; mp.umax = b->umax; mp.bfrag = b->bfrag; mp.bins = b->bins; mp.bin8 = b->bin8; mp.bias = (unsigned)f.la;
            // shared memory: query planes + umax + two tile buffers per warp; as many warps (<= 16) as fit
            const size_t fixed = (size_t)frag_vec * 16 + 256, per_warp = (size_t)2 * 16 * d;
            unsigned warps = (unsigned)std::min<size_t>(16, ((size_t)b->smem_optin - fixed) / per_warp);
            if (warps == 0) return bfail(QMANN_E_NOMEM, "tensor-core scorer does not fit shared memory");
            const size_t smem = fixed + (size_t)warps * per_warp;
            BCUDA(cudaFuncSetAttribute(k_big_scores_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned long long tiles = (b->S_local + 15) / 16;
            const unsigned gx = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>((tiles + warps - 1) / warps, (unsigned long long)b->sm_count));
            k_big_scores_mma<<<dim3(gx, qblocks), warps * 32, smem, st>>>(mp);
            count_launch();
            BCUDA(cudaPeekAtLastError());
            rc = QMANN_OK;
        }
        else if (b->fast[h]) {
            FastScoreParams fp;
            fp.Y = b->Y[h]; fp.rowmax = b->rowmax[h]; fp.S_local = b->S_local; fp.d = d; fp.Q = Q; fp.la = f.la; fp.fb = f.fb;
            fp.ub8 = b->ub8; fp.umax = b->umax; fp.bins = b->bins; fp.bin8 = b->bin8; fp.bias = (unsigned)f.la;
            rc = dispatch_scores_fast(b, fp, st);
        }
        else if (b->cfg.mode == 3 && nine && d % 16 == 0 && !b->ham_off) {
            // packed-byte Hamming scorer; a block of queries that holds the -2^iwl value is declined there and served by the
            // second launch on the literal path (same block size, so the two partitions agree)
            HamParams hp;
            hp.M = b->M[h]; hp.S_local = b->S_local; hp.d = d; hp.Q = Q; hp.sh_m = sh_m; hp.sh_u = sh_u; hp.u8 = b->u_a; hp.quirk = b->quirk;
            hp.bins = b->bins; hp.bin8 = b->bin8; hp.bias = b->bias;
            sp.only_quirk = 1;
            rc = (Q >= 8) ? launch_ham<16>(b, hp, st) : launch_ham<4>(b, hp, st);
            if (!rc) rc = (Q >= 8) ? launch_scores<3, 16>(b, sp, st) : launch_scores<3, 4>(b, sp, st);
        }
        else if (b->cfg.mode == 3) rc = (Q >= 8) ? launch_scores<3, 8>(b, sp, st) : (Q >= 4 ? launch_scores<3, 4>(b, sp, st) : launch_scores<3, 1>(b, sp, st));
        else rc = (Q >= 16) ? launch_scores<2, 16>(b, sp, st) : (Q >= 4 ? launch_scores<2, 4>(b, sp, st) : launch_scores<2, 1>(b, sp, st));
        if (rc) return rc;
        if (prof) { BCUDA(cudaEventRecord(b->pev[b->pused + 1], st)); b->pused += 2; }
        if (hist_done) return QMANN_OK;
        // mode 2 bins are code + la of THIS hop; the histogram is indexed with the common bias
        const int use_smem = b->NB <= 8192;
        const unsigned gx = (unsigned)std::min<unsigned long long>((b->S_local + 255) / 256, (unsigned long long)b->sm_count * 4);
        const char *env_hl = getenv("QMANN_BIGMEM_HIST_LANES");          // 0: the shared-memory-atomic kernel (A/B tests)
        const bool hist_lanes_on = !(env_hl && atoi(env_hl) == 0);
        const size_t hl_smem = (size_t)HL_WARPS * 256 * 32 * 2;
        // (small jobs keep k_big_hist: a task's end-of-run reduction over 255 bins x 32 lanes needs ~16 K slots to amortise)
        if (b->bin8 && b->NB <= 256 && hist_lanes_on && hl_smem <= (size_t)b->smem_optin && (unsigned long long)Q * b->S_local / 16384 >= 2ull * b->sm_count * HL_WARPS) {
            static bool hl_attr = false;
            if (!hl_attr) { BCUDA(cudaFuncSetAttribute(k_big_hist_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hl_smem)); hl_attr = true; }
            // tasks of 2^16 slots (2048 values per lane) when that leaves at least four per warp, else 2^14
            const unsigned long long nw = (unsigned long long)b->sm_count * HL_WARPS, work = (unsigned long long)Q * b->S_local;
            const unsigned chunk = (work / 65536 >= 4 * nw) ? 65536u : 16384u;
            const unsigned long long ntask = ((b->S_local + chunk - 1) / chunk) * Q;
            const unsigned hgx = (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>((ntask + HL_WARPS - 1) / HL_WARPS, (unsigned long long)b->sm_count));
            k_big_hist_lanes<<<hgx, HL_WARPS * 32, hl_smem, st>>>(reinterpret_cast<const unsigned char *>(b->bins), b->S_local, Q, b->NB, chunk, dev_hist);
        }
        else if (b->bin8) k_big_hist<unsigned char><<<dim3(std::max(1u, gx), Q), 256, use_smem ? b->NB * 4 : 0, st>>>(reinterpret_cast<const unsigned char *>(b->bins), b->S_local, b->NB, dev_hist, use_smem);
        else         k_big_hist<unsigned short><<<dim3(std::max(1u, gx), Q), 256, use_smem ? b->NB * 4 : 0, st>>>(reinterpret_cast<const unsigned short *>(b->bins), b->S_local, b->NB, dev_hist, use_smem);
        count_launch();
        BCUDA(cudaPeekAtLastError());
    }
    return QMANN_OK;
}

int qmann_bigmem_hop_read(qmann_bigmem *b, uint32_t h, const uint32_t *dev_hist, int32_t *dev_partial, float *dev_pbin, void *stream)
{
    if (!b || !dev_hist || !dev_partial) return bfail(QMANN_E_ARG, "null argument");
    if (h >= b->cfg.H || b->Q == 0) return bfail(QMANN_E_ARG, "bad hop or no query batch");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned Q = b->Q, d = b->cfg.d;
    const HopFmt &f = b->f[h];
    SoftmaxParams sp;
    sp.hist = dev_hist; sp.NB = b->NB; sp.bias = (b->cfg.mode == 3) ? b->bias : (unsigned)f.la; sp.mode = (int)b->cfg.mode;
    sp.fa = f.fa; sp.ia = f.ia; sp.const_scale = b->cfg.const_scale; sp.iff = f.iff; sp.ff = f.ff;
    sp.pq = b->pq; sp.thr = b->thr; sp.pbin = dev_pbin; sp.total_out = nullptr;
    k_big_softmax<<<Q, SOFTMAX_RANGES, 0, st>>>(sp);
    count_launch();
    BCUDA(cudaMemsetAsync(dev_partial, 0, (size_t)Q * d * 4, st));
    BCUDA(cudaMemsetAsync(b->nsel, 0, (size_t)Q * 4, st));
    if (b->S_local) {
        ReadParams rp;
        rp.bins = b->bins; rp.pq = b->pq; rp.thr = b->thr; rp.C = b->C[h]; rp.S_local = b->S_local; rp.NB = b->NB; rp.d = d; rp.f = f;
        rp.partial = dev_partial; rp.nsel = b->nsel;
        const unsigned long long items = b->S_local / (b->bin8 ? 16 : 8) + 1;                    // 128-bit vectors of bins per query
        const unsigned gx = (unsigned)std::min<unsigned long long>((items + 1023) / 1024, (unsigned long long)b->sm_count * 4);      // four per thread
        if (b->bin8) k_big_read<unsigned char><<<dim3(std::max(1u, gx), Q), 256, 0, st>>>(rp);
        else         k_big_read<unsigned short><<<dim3(std::max(1u, gx), Q), 256, 0, st>>>(rp);
        count_launch();
    }
    BCUDA(cudaPeekAtLastError());
    return QMANN_OK;
}

int qmann_bigmem_hop_update(qmann_bigmem *b, uint32_t h, const int32_t *dev_partial, int8_t *dev_o, int8_t *dev_g, void *stream)
{
    if (!b || !dev_partial) return bfail(QMANN_E_ARG, "null argument");
    if (h >= b->cfg.H || b->Q == 0) return bfail(QMANN_E_ARG, "bad hop or no query batch");
    UpdateParams up;
    up.partial = dev_partial; up.u_in = b->u_a; up.u_out = b->u_b; up.ub = b->ub; up.Hq = b->cfg.lin_map ? b->dev_H[h] : nullptr;
    up.d = b->cfg.d; up.fu = b->fu; up.f = b->f[h]; up.dbg_o = dev_o; up.dbg_g = dev_g;
    if (b->fast_update && up.Hq)       // ub8 / umax are those of this hop's k_big_prep_query (mode 2: the same Q_bin(u) feeds scorer and map)
        k_big_update_fast<<<dim3(b->Q, (b->cfg.d + 255) / 256), 256, (size_t)b->cfg.d + (size_t)(b->cfg.d / 16) * 4, (cudaStream_t)stream>>>(up, b->ub8, b->dev_hmax[h]);
    else
        k_big_update<<<dim3(b->Q, (b->cfg.d + 7) / 8), 256, (size_t)b->cfg.d * 4, (cudaStream_t)stream>>>(up);
    count_launch();
    BCUDA(cudaPeekAtLastError());
    std::swap(b->u_a, b->u_b);
    b->fu = b->f[h].ff;
    return QMANN_OK;
}

int qmann_bigmem_state(qmann_bigmem *b, int8_t *dev_u, int32_t *frac_bits, void *stream)
{
    if (!b) return bfail(QMANN_E_ARG, "null argument");
    if (dev_u && b->Q) BCUDA(cudaMemcpyAsync(dev_u, b->u_a, (size_t)b->Q * b->cfg.d, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (frac_bits) *frac_bits = b->fu;
    return QMANN_OK;
}

int qmann_bigmem_profile_enable(qmann_bigmem *b, int enable)
{
    if (!b) return bfail(QMANN_E_ARG, "null argument");
    if (enable && !b->pev[0])
        for (int i = 0; i < 2 * MAXH; i++) BCUDA(cudaEventCreate(&b->pev[i]));
    b->profile = enable != 0;
    b->pused = 0; b->pms = 0.0; b->pcount = 0;
    return QMANN_OK;
}

// Folds the pending event pairs (call after the forward has been enqueued; synchronises on them) and returns
// the accumulated k_big_scores time and launch count since the last reset.
int qmann_bigmem_profile_read(qmann_bigmem *b, float *ms_scores, uint32_t *n_launches, int reset)
{
    if (!b) return bfail(QMANN_E_ARG, "null argument");
    for (unsigned i = 0; i + 2 <= b->pused; i += 2) {
        float ms = 0.f;
        BCUDA(cudaEventSynchronize(b->pev[i + 1]));
        BCUDA(cudaEventElapsedTime(&ms, b->pev[i], b->pev[i + 1]));
        b->pms += ms; b->pcount++;
    }
    b->pused = 0;
    if (ms_scores) *ms_scores = (float)b->pms;
    if (n_launches) *n_launches = b->pcount;
    if (reset) { b->pms = 0.0; b->pcount = 0; }
    return QMANN_OK;
}

int qmann_bigmem_finish(qmann_bigmem *b, uint32_t *dev_pred, float *dev_z, float *dev_h, void *stream)
{
    if (!b) return bfail(QMANN_E_ARG, "null argument");
    if (b->Q == 0) return bfail(QMANN_E_ARG, "no query batch");
    if (!b->dev_W || !b->cfg.V) return bfail(QMANN_E_ARG, "the memory was created without an answer projection (dev_W, V)");
    cudaStream_t st = (cudaStream_t)stream;
    float *z = dev_z ? dev_z : b->zbuf;
    k_big_answer<<<(b->Q * 32 + 127) / 128, 128, (size_t)4 * b->cfg.d * 4, st>>>(b->dev_W, b->u_a, b->fu, b->Q, b->cfg.V, b->cfg.d, z, dev_pred, dev_h);
    count_launch();
    BCUDA(cudaPeekAtLastError());
    return QMANN_OK;
}

// ---- the whole slot-sharded forward as ONE call: phases + the two exchanges per hop on the caller's NCCL communicator ----
namespace {
// ncclAllReduce resolved at run time from the NCCL the process already uses (the library itself does not link NCCL):
//   ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
nccl_allreduce_fn nccl_allreduce()
{
    static nccl_allreduce_fn fn = []() -> nccl_allreduce_fn {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        return h ? (nccl_allreduce_fn)dlsym(h, "ncclAllReduce") : nullptr;
    }();
    return fn;
}
constexpr int NCCL_INT32 = 2, NCCL_UINT32 = 3, NCCL_SUM = 0;      // nccl.h: ncclInt32, ncclUint32, ncclSum

int forward_phases(qmann_bigmem *b, void *comm, const int8_t *dev_u0, uint32_t Q, uint32_t *dev_pred, cudaStream_t st)
{
    nccl_allreduce_fn ar = comm ? nccl_allreduce() : nullptr;
    if (comm && !ar) return bfail(QMANN_E_ARG, "libnccl.so.2 not found: cannot run the exchanges of a sharded forward");
    int rc;
    if ((rc = qmann_bigmem_begin(b, dev_u0, Q, st))) return rc;
    for (uint32_t h = 0; h < b->cfg.H; h++) {
        if ((rc = qmann_bigmem_hop_scores(b, h, b->xhist, st))) return rc;
        if (ar && ar(b->xhist, b->xhist, (size_t)Q * b->NB, NCCL_UINT32, NCCL_SUM, comm, st) != 0) return bfail(QMANN_E_CUDA, "ncclAllReduce (score histograms) failed");
        if ((rc = qmann_bigmem_hop_read(b, h, b->xhist, b->xpartial, nullptr, st))) return rc;
        if (ar && ar(b->xpartial, b->xpartial, (size_t)Q * b->cfg.d, NCCL_INT32, NCCL_SUM, comm, st) != 0) return bfail(QMANN_E_CUDA, "ncclAllReduce (partial reads) failed");
        if ((rc = qmann_bigmem_hop_update(b, h, b->xpartial, nullptr, nullptr, st))) return rc;
    }
    if (dev_pred && b->dev_W && b->cfg.V) return qmann_bigmem_finish(b, dev_pred, nullptr, nullptr, st);
    return QMANN_OK;
}
}  // namespace

int qmann_bigmem_forward_sharded(qmann_bigmem *b, void *nccl_comm, const int8_t *dev_u0, uint32_t Q, uint32_t *dev_pred, void *stream)
{
    if (!b || !dev_u0) return bfail(QMANN_E_ARG, "null argument");
    if (Q == 0 || Q > b->Q_max) return bfail(QMANN_E_ARG, "Q must be in 1..Q_max");
    cudaStream_t st = (cudaStream_t)stream;
    if (!b->xhist) BCUDA(cudaMalloc((void **)&b->xhist, (size_t)b->Q_max * b->NB * 4));
    if (!b->xpartial) BCUDA(cudaMalloc((void **)&b->xpartial, (size_t)b->Q_max * b->cfg.d * 4));
    const char *env_graph = getenv("QMANN_BIGMEM_GRAPH");
    const bool use_graph = !(env_graph && atoi(env_graph) == 0) && !b->profile && st != nullptr;      // the legacy default stream cannot be captured
    if (!use_graph) return forward_phases(b, nccl_comm, dev_u0, Q, dev_pred, st);
    if (!b->warm || b->warm_comm != nccl_comm) {
        // the first forward on a communicator runs un-captured: NCCL sets its channels up inside the first collective (allocations,
        // synchronisation), which a stream capture must not see; it is captured from the second call on
        b->warm = true; b->warm_comm = nccl_comm;
        return forward_phases(b, nccl_comm, dev_u0, Q, dev_pred, st);
    }
    if (!b->gexec || b->g_u0 != dev_u0 || b->g_pred != dev_pred || b->g_comm != nccl_comm || b->g_Q != Q) {
        // (re)capture: about twenty launches, two memsets and six collectives become one graph launch per forward
        if (b->gexec) { cudaGraphExecDestroy(b->gexec); b->gexec = nullptr; }
        cudaGraph_t graph = nullptr;
        BCUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        const int rc = forward_phases(b, nccl_comm, dev_u0, Q, dev_pred, st);
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess || !graph) return bfail(QMANN_E_CUDA, std::string("stream capture of the forward failed: ") + cudaGetErrorString(ce));
        const cudaError_t ie = cudaGraphInstantiate(&b->gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { b->gexec = nullptr; return bfail(QMANN_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
        b->g_u0 = dev_u0; b->g_pred = dev_pred; b->g_comm = nccl_comm; b->g_Q = Q;
    }
    BCUDA(cudaGraphLaunch(b->gexec, st));
    return QMANN_OK;
}

}  // extern "C"

// qmann_fast.cuh -- the production forward kernel k_story (included by qmann_forward.cu after
// qmann_kernels.cuh).
//
// One warp per story, persistent CTAs, dynamic work claiming.  A warp takes a story through the WHOLE path without
// leaving the SM:
//
//   DENSE source (the reference's boundary format, MemN2N.c:2294-2350): the warp streams the story's dense fp32
//     bag-of-words rows from HBM into its private shared-memory staging buffers with bulk asynchronous copies
//     (cp.async.bulk + mbarrier transaction counts, SASS UBLKCP), NB chunks of R rows in flight, and compacts them
//     on the fly into a list of 16-bit table-row offsets in its own scratch -- no global compact record, no second
//     kernel.  The first chunks of the warp's NEXT story are issued before the forward of the current one, and a bulk
//     L2 prefetch runs pf_dist claims ahead, so HBM streams underneath the (issue-bound) forward arithmetic.
//   RECORD source (word-id input, qmann_forward_ids): the compact records k_ids_compact wrote.
//
// then per hop: gather-and-sum embedding from the int8 tables in shared memory, scorer (packed SWAR dot product, plain
// fixed-point dot product or nine-bit Hamming form), softmax, weighted read over the slots with a non-zero quantised
// weight (C_h rows gathered from L2), linear map (product-table gather), saturating update; finally the answer
// projection with the int8 prefilter and the argmax.
//
// It handles the stories that are regular: every bag-of-words value is 1.0 after count splitting, the entry list fits
// the warp's slot, every weight format has an integer bit.  Anything else is appended to p.slow_list and processed by
// the next tier (unpacked k_story, then k_compact + the general k_forward: same arithmetic, every input) -- further CUDA
// launches, not a CPU fallback.
//
// DUMP instantiations additionally write what the arithmetic produced (scores, attention codes = selected slots, read,
// linear map, controller state per hop, exact logits and the prefilter's candidate set, and which tier finished the
// story) so that the parity tests observe the production kernels themselves.
#pragma once
#include "qmann_kernels.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// bulk-copy / mbarrier primitives (sm_90+ PTX; SASS: UBLKCP, SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void line_prefetch_l2(const void *src) { asm volatile("prefetch.global.L2 [%0];" ::"l"(src)); }

// ---------------------------------------------------------------------------------------------
// in-kernel compaction of one staged dense row (shared memory) into 16-bit entries col * DP
// ---------------------------------------------------------------------------------------------
// Appends one entry per set bit of mm.  ca/cb: float4 index of the lane's two loads inside the row's span, sh: position of
// column 0 inside the first float4.
__device__ __forceinline__ unsigned emit_units_s(unsigned mm, unsigned ca, unsigned cb, unsigned sh, unsigned DP, unsigned short *__restrict__ ent,
                                                 unsigned cap, unsigned base, unsigned lt)
{
    unsigned any;
    while ((any = __ballot_sync(0xffffffffu, mm != 0u)) != 0u) {
        const unsigned k = (unsigned)(__ffs((int)mm) - 1);               // garbage when mm == 0, unused
        const unsigned col = ((k >= 4u) ? cb : ca) * 4u + (k & 3u) - sh;
        const unsigned pos = base + __popc(any & lt);
        if (mm != 0u && pos < cap) ent[pos] = (unsigned short)(col * DP);
        mm &= mm - 1u;
        base += __popc(any);
    }
    return base;
}

// A chunk holding values other than 0.0 / 1.0: counts n in 2..nmax with n * colmax <= split_lim become n unit entries
// (then Q_w(Q_w(n) * Q_w(T)) = n * T for every table, hop and dimension), anything else makes the story irregular.
__device__ __forceinline__ unsigned emit_general_s(const FwdParams &p, const float (&v)[8], unsigned ca, unsigned cb, unsigned sh, unsigned short *__restrict__ ent,
                                                   unsigned cap, unsigned base, unsigned lt, bool &irregular)
{
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) m |= nzbit(v[j]) << j;
    unsigned rep = 0, col = 0, any;
    bool bad = false;
    while ((any = __ballot_sync(0xffffffffu, (m | rep) != 0u)) != 0u) {
        const bool had = ((m | rep) != 0u);
        if (rep) {
            rep--;
        } else if (had) {
            const unsigned k = (unsigned)(__ffs((int)m) - 1);
            m &= m - 1u;
            float x = v[0];
#pragma unroll
            for (int j = 1; j < 8; j++) x = (k == (unsigned)j) ? v[j] : x;
            col = ((k >= 4u) ? cb : ca) * 4u + (k & 3u) - sh;
            if (__float_as_uint(x) != 0x3F800000u) {
                const float n = truncf(x);
                if (n == x && x >= 2.0f && x <= (float)p.nmax && (unsigned)n * (unsigned)p.colmax[col] <= p.split_lim) rep = (unsigned)n - 1u;
                else bad = true;
            }
        }
        const unsigned pos = base + __popc(any & lt);
        if (had && pos < cap) ent[pos] = (unsigned short)(col * p.DP);
        base += __popc(any);
    }
    irregular |= __any_sync(0xffffffffu, bad);
    return base;
}

// One row of V floats staged at float index fr of `buf` (16-byte aligned buffer).  ALIGNED: fr % 4 == 0 and V % 4 == 0.
template <bool ALIGNED>
__device__ __forceinline__ unsigned scan_row_s(const FwdParams &p, const float *__restrict__ buf, unsigned fr, unsigned short *__restrict__ ent, unsigned cap,
                                               unsigned base, unsigned lane, bool &irregular)
{
    const unsigned lt = (1u << lane) - 1u;
    const unsigned sh = ALIGNED ? 0u : (fr & 3u);
    const float4 *b4 = reinterpret_cast<const float4 *>(buf) + (fr >> 2);
    const unsigned V = p.V;
    const unsigned n4 = (sh + V + 3u) >> 2;
    for (unsigned c0 = 0; c0 < n4; c0 += 64) {
        const unsigned ca = c0 + lane, cb = ca + 32u;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = 0.0f;
        if (ca < n4) { const float4 t = b4[ca]; v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        if (cb < n4) { const float4 t = b4[cb]; v[4] = t.x; v[5] = t.y; v[6] = t.z; v[7] = t.w; }
        if (!ALIGNED) {
            // elements in front of column 0 (first float4) or behind column V-1 (last float4) belong to the neighbours
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned fa = 4u * ca + (unsigned)j, fb = 4u * cb + (unsigned)j;
                if (fa < sh || fa >= sh + V) v[j] = 0.0f;
                if (fb < sh || fb >= sh + V) v[4 + j] = 0.0f;
            }
        }
        if (!chunk_irregular<4>(v)) base = emit_units_s(unit_mask<4>(v), ca, cb, sh, p.DP, ent, cap, base, lt);
        else base = emit_general_s(p, v, ca, cb, sh, ent, cap, base, lt, irregular);
    }
    return base;
}

// acc[j] += table codes of dims 16q..16q+15 over the entries of `row`; the table lives in GLOBAL memory (B and C_h:
// few rows per story, L2-resident) -- tabq points at dims 16q.. of table row 0.  Lanes past their row's end gather
// the all-zero row V through the pseudo entry at `zaddr`.
template <int LPR>
__device__ __forceinline__ void embed_glob(const FwdParams &p, unsigned ws, unsigned lane, const unsigned char *__restrict__ tabq, int row, int acc[16],
                                           const int sel[4])
{
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0;
    unsigned beg = 0, len = 0;
    if (row >= 0) {
        const unsigned short *rend = reinterpret_cast<const unsigned short *>(smem + ws + p.fl.o_rend);
        beg = row ? rend[row - 1] : 0u;
        len = rend[row] - beg;
    }
    const unsigned maxlen = __reduce_max_sync(0xffffffffu, len);
    unsigned ea = ws + 2u * beg;
    const unsigned zaddr = ws + p.fl.o_zent;
#pragma unroll 1
    for (unsigned k0 = 0; k0 < maxlen; k0 += 4) {
        uint4 t[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned off = *reinterpret_cast<const unsigned short *>(smem + ((k0 + i < len) ? ea + 2u * i : zaddr));
            t[i] = __ldg(reinterpret_cast<const uint4 *>(tabq + off));
        }
        ea += 8u;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            acc[0] = __dp4a((int)t[i].x, sel[0], acc[0]);   acc[1] = __dp4a((int)t[i].x, sel[1], acc[1]);
            acc[2] = __dp4a((int)t[i].x, sel[2], acc[2]);   acc[3] = __dp4a((int)t[i].x, sel[3], acc[3]);
            acc[4] = __dp4a((int)t[i].y, sel[0], acc[4]);   acc[5] = __dp4a((int)t[i].y, sel[1], acc[5]);
            acc[6] = __dp4a((int)t[i].y, sel[2], acc[6]);   acc[7] = __dp4a((int)t[i].y, sel[3], acc[7]);
            acc[8] = __dp4a((int)t[i].z, sel[0], acc[8]);   acc[9] = __dp4a((int)t[i].z, sel[1], acc[9]);
            acc[10] = __dp4a((int)t[i].z, sel[2], acc[10]); acc[11] = __dp4a((int)t[i].z, sel[3], acc[11]);
            acc[12] = __dp4a((int)t[i].w, sel[0], acc[12]); acc[13] = __dp4a((int)t[i].w, sel[1], acc[13]);
            acc[14] = __dp4a((int)t[i].w, sel[2], acc[14]); acc[15] = __dp4a((int)t[i].w, sel[3], acc[15]);
        }
    }
}

// Same gather from a table in SHARED memory (A_h, unpacked kernels): tabq = shared byte offset of dims 16q.. of row 0.
template <int LPR>
__device__ __forceinline__ void embed_smem(const FwdParams &p, unsigned ws, unsigned lane, unsigned tabq, int row, int acc[16], const int sel[4])
{
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0;
    unsigned beg = 0, len = 0;
    if (row >= 0) {
        const unsigned short *rend = reinterpret_cast<const unsigned short *>(smem + ws + p.fl.o_rend);
        beg = row ? rend[row - 1] : 0u;
        len = rend[row] - beg;
    }
    const unsigned maxlen = __reduce_max_sync(0xffffffffu, len);
    unsigned ea = ws + 2u * beg;
    const unsigned zaddr = ws + p.fl.o_zent;
#pragma unroll 4
    for (unsigned k = 0; k < maxlen; k++, ea += 2u) {
        const unsigned off = *reinterpret_cast<const unsigned short *>(smem + ((k < len) ? ea : zaddr));
        const uint4 t = *reinterpret_cast<const uint4 *>(smem + tabq + off);
        acc[0] = __dp4a((int)t.x, sel[0], acc[0]);   acc[1] = __dp4a((int)t.x, sel[1], acc[1]);
        acc[2] = __dp4a((int)t.x, sel[2], acc[2]);   acc[3] = __dp4a((int)t.x, sel[3], acc[3]);
        acc[4] = __dp4a((int)t.y, sel[0], acc[4]);   acc[5] = __dp4a((int)t.y, sel[1], acc[5]);
        acc[6] = __dp4a((int)t.y, sel[2], acc[6]);   acc[7] = __dp4a((int)t.y, sel[3], acc[7]);
        acc[8] = __dp4a((int)t.z, sel[0], acc[8]);   acc[9] = __dp4a((int)t.z, sel[1], acc[9]);
        acc[10] = __dp4a((int)t.z, sel[2], acc[10]); acc[11] = __dp4a((int)t.z, sel[3], acc[11]);
        acc[12] = __dp4a((int)t.w, sel[0], acc[12]); acc[13] = __dp4a((int)t.w, sel[1], acc[13]);
        acc[14] = __dp4a((int)t.w, sel[2], acc[14]); acc[15] = __dp4a((int)t.w, sel[3], acc[15]);
    }
}

// One lane's share (16 dims) of  sum_t ( Q_att( Q_att(M[r][t]) * Q_bin(u[t]) ) + la ):
// y = clamp(m) + L in one VIADDMNMX.RELU, x = y*u - L*u in one IMAD (cub = -L*u), truncating shift
// in three ops, clamp + la in one more, accumulate.                        layer_cuda.cu:105-141
// KA = sign of (frac_att - frac_w), the re-quantisation of M from the weight to the addressing format:
//   KA == 0: same grid, one clamp (la == lw)
//   KA  > 0: clamp_att(clamp_w(a) << k) == clamp_att(a << k) because la <= lw << k
//   KA  < 0: trunc0(clamp_w(a) / 2^k) == clamp(trunc0(a / 2^k), lw >> k) (both maps are odd and monotone)
template <int KA>
__device__ __forceinline__ int score_fast(const int acc[16], const int ub[16], const int cub[16], int L, int ksh, int la, int fb, int mb)
{
    const int L2 = 2 * L, la2 = 2 * la, mk = (1 << ksh) - 1;
    int part = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        int a = acc[j];
        if (KA > 0) a <<= ksh;
        if (KA < 0) a = shr0m(a, ksh, mk);
        const int y = clamp_biased(a, L, L2);
        const int x = y * ub[j] + cub[j];
        part += clamp_biased(shr0m(x, fb, mb), la, la2);
    }
    return part;
}

// ---------------------------------------------------------------------------------------------
// Packed (SWAR) memory embedding + dot-product scorer of k_story<.., 2, true, ..>.
//
// The A_h tables of this kernel's image hold BIASED bytes code + cm_h[column] (cm_h = max |code| of the column), so a
// row sum is formed on four dims per 32-bit add with no carries between bytes as long as the row's bias
// B = sum of cm_h over its entries is <= 127 (then the true sum s - B is in [-B, B] and Q_w is the identity).
// Stories with a row above that bound are left to the unpacked kernel.  From the packed sums:
//   a = s - B per byte, y = Q_att(a) per byte (ka = 0: a, +1: 2a, -1: trunc0(a/2)), and the score
//   4 * sum_t trunc0(y_t u_t / 4) = sum y_t u_t - sum (x_t mod 4) + 4 #{x_t < 0, x_t mod 4 != 0}
// exactly as in k_big_scores_fast (qmann_bigmem.cu), valid when no product saturates: |y_t| < tau(|u_t|) =
// ceil(512 / |u_t|), tested exactly per byte; a row that fails is recomputed product by product.
// All byte identities are checked exhaustively on the host (tests/test_identities.py).
// ---------------------------------------------------------------------------------------------
constexpr unsigned SW_H = 0x80808080u, SW_L = 0x7F7F7F7Fu, SW_1 = 0x01010101u;

__device__ __forceinline__ unsigned swar_unbias(unsigned s, unsigned Bw) { return ((s | SW_H) - Bw) ^ (~s & SW_H); }
__device__ __forceinline__ unsigned swar_half0(unsigned a)          // trunc0(a / 2) per signed byte
{
    const unsigned neg = (a >> 7) & SW_1;
    const unsigned t = ((a & SW_L) + neg) ^ (a & SW_H);
    return ((t >> 1) & SW_L) | (t & SW_H);
}

struct SwarQuery {                     // per lane: its 16 dims of Q_bin(u), packed
    unsigned Uw[4], U0[4], U0s[4], U1[4], Us4[4], Tw[4];
};

// One pass: the packed row sums acc4 of row r (bias Bw replicated per byte) -> 4 * partial score of this lane's 16 dims
// (+48) and the saturation flag.  y4 returns Q_att(M) for the exact path.
template <int KA>
__device__ __forceinline__ int swar_score(const unsigned (&acc4)[4], unsigned Bw, const SwarQuery &sq, unsigned (&y4)[4], unsigned &flag)
{
    int D = 0;
    unsigned cs = 0, f = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const unsigned a = swar_unbias(acc4[w], Bw);
        unsigned y;
        if (KA == 0) y = a;
        else if (KA > 0) y = (a << 1) & 0xFEFEFEFEu;
        else y = swar_half0(a);
        y4[w] = y;
        unsigned fill;
        asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(fill) : "r"(y));
        f |= ((y ^ fill) + (fill & SW_1)) + sq.Tw[w];                 // bit 7 of a byte: |y| >= tau
        D = __dp4a((int)y, (int)sq.Uw[w], D);
        const unsigned t0 = (y << 1) & sq.U1[w];
        const unsigned t1 = (y & sq.U0s[w]) ^ t0;
        const unsigned bm = (y & sq.U0[w]) | t1;                       // x mod 4 per byte
        const unsigned wv = bm + 0x03030303u;
        cs += wv & ~(((y >> 5) & 0x04040404u) ^ sq.Us4[w]);
    }
    flag = f & SW_H;
    return D - (int)__dp4a(cs, SW_1, 0u) + 48;
}

// DP == 64 (LPR == 4): lane l owns dims 2l, 2l+1 of every d-vector.  (a0, a1) and (b0, b1) += the table codes of those dims over the
// entries [ea, ea + 2 na) and [eb, eb + 2 nb) of the warp's entry list (shared byte addresses; 16-bit byte offsets column * DP): one
// coalesced 64-byte table row per entry and warp, two rows' gathers interleaved so that their L2 round trips overlap.
__device__ __forceinline__ unsigned ldg_na_u16(const void *ptr)
{
    unsigned short r;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(ptr));
    return (unsigned)r;
}
__device__ __forceinline__ void gather2x2(unsigned ea, unsigned na, unsigned eb, unsigned nb, const unsigned char *__restrict__ tab, unsigned lane, int &a0, int &a1,
                                          int &b0, int &b1)
{
    const unsigned char *t16 = tab + 2u * lane;
    const unsigned n = max(na, nb);
#pragma unroll 4
    for (unsigned e = 0; e < n; e++) {
        unsigned short offa = 0, offb = 0;
        if (e < na) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(offa) : "r"(ea + 2u * e));
        if (e < nb) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(offb) : "r"(eb + 2u * e));
        unsigned wa = 0, wb = 0;
        if (e < na) wa = ldg_na_u16(t16 + offa);
        if (e < nb) wb = ldg_na_u16(t16 + offb);
        a0 += (int)(signed char)(wa & 0xFFu);
        a1 += (int)(signed char)(wa >> 8);
        b0 += (int)(signed char)(wb & 0xFFu);
        b1 += (int)(signed char)(wb >> 8);
    }
}

// Tier tag written to dbg.dev_path / counted in p.path_count
constexpr unsigned PATH_PACKED = 1u, PATH_UNPACKED = 2u, PATH_GENERAL = 3u;

template <int LPR, int MODE, bool SWAR, bool DENSE, bool DUMP, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_story(const __grid_constant__ FwdParams p)
{
    constexpr int G = 32 / LPR;
    const FastLayout &fl = p.fl;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned g = lane / LPR, q = lane % LPR;
    const unsigned n_work = p.work_list ? *p.work_count : p.n_stories;
    if (n_work == 0u) return;                                  // nothing was left to this kernel
    if (p.path_count && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.path_count + (SWAR ? 0 : 1), (unsigned long long)n_work);     // stories entering this tier
    const unsigned d = p.d, DP = p.DP, V = p.V;
    {
        // this kernel's shared-memory image: the A_h tables, their packed column maxima, tau and the int8 image of W;
        // B, C_h (a few rows per story and hop) and the fp32 rows of W (candidates only) are read from L2
        for (unsigned h = 0; h < p.H; h++) {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.img + p.offA[h]);
            uint4 *dst = reinterpret_cast<uint4 *>(smem + fl.sA[h]);
            for (unsigned i = threadIdx.x; i < (V + 1) * DP / 16; i += blockDim.x) dst[i] = src[i];
        }
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.img + p.offCM[0]);
            uint4 *dst = reinterpret_cast<uint4 *>(smem + fl.sCM);
            for (unsigned i = threadIdx.x; i < ((V + 1) * 4 + 15) / 16; i += blockDim.x) dst[i] = src[i];
            if (threadIdx.x < 8) reinterpret_cast<uint4 *>(smem + fl.sTAU)[threadIdx.x] = reinterpret_cast<const uint4 *>(p.img + p.offTAU)[threadIdx.x];
        }
        const uint4 *src8 = reinterpret_cast<const uint4 *>(p.img + p.offW8);
        uint4 *dst8 = reinterpret_cast<uint4 *>(smem + fl.sW8);
        for (unsigned i = threadIdx.x; i < p.w8_bytes / 16; i += blockDim.x) dst8[i] = src8[i];
    }
    const unsigned wso = fl.tables_bytes + wid * fl.warp_bytes;
    unsigned char *ws = smem + wso;
    unsigned short *ent_s = reinterpret_cast<unsigned short *>(ws);
    unsigned short *rend_s = reinterpret_cast<unsigned short *>(ws + fl.o_rend);
    int *sc = reinterpret_cast<int *>(ws + fl.o_sc);
    float *ex = reinterpret_cast<float *>(ws + fl.o_ex);
    unsigned char *pq = ws + fl.o_pq;
    signed char *uvec = reinterpret_cast<signed char *>(ws + fl.o_uvec);
    int *ub32 = reinterpret_cast<int *>(ws + fl.o_ub32);
    signed char *ovec = reinterpret_cast<signed char *>(ws + fl.o_ovec);
    float *ufl = reinterpret_cast<float *>(ws + fl.o_ufl);
    float *zbuf = reinterpret_cast<float *>(ws);      // aliases the entry list (dead by the answer phase)
    unsigned char *brow = ws + fl.o_brow;              // SWAR: row biases [H][S_pad]
    unsigned short *perm = reinterpret_cast<unsigned short *>(ws + fl.o_perm);     // rows ordered by entry count
    unsigned *cnt_s = reinterpret_cast<unsigned *>(ws + fl.o_cnt);
    const unsigned bar0 = smem_u32(ws + fl.o_bar);
    const unsigned stage0 = smem_u32(ws + fl.o_stage);
    if (lane == 0) {
        *reinterpret_cast<unsigned short *>(ws + fl.o_zent) = (unsigned short)(V * DP);
        if (DENSE) {
            for (unsigned b = 0; b < fl.NB; b++) mbar_init(bar0 + 8u * b, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    __syncthreads();

    const unsigned lgDP = 31u - (unsigned)__clz((int)DP);
    int sel[4];
    asm volatile("mov.u32 %0, 0x00000001;" : "=r"(sel[0]));
    asm volatile("mov.u32 %0, 0x00000100;" : "=r"(sel[1]));
    asm volatile("mov.u32 %0, 0x00010000;" : "=r"(sel[2]));
    asm volatile("mov.u32 %0, 0x01000000;" : "=r"(sel[3]));
    const bool aligned = DENSE && (V % 4u == 0u);
    const unsigned row_bytes = V * 4u;

    // ---- work claiming and the dense stream ----
    auto claim = [&]() -> unsigned {
        unsigned i = 0;
        if (lane == 0) i = atomicAdd(p.counter, 1u);
        return __shfl_sync(0xffffffffu, i, 0);
    };
    // chunk c of a story: c == 0 is the question row, c >= 1 the sentence rows (c-1)*R .. ; returns the byte range of the
    // arena it covers and the 16-byte aligned range that is copied
    struct Chunk { const char *base; unsigned long long a0; unsigned bytes, skew; bool short_tail; };
    auto chunk_of = [&](unsigned story, unsigned long long soff, unsigned S, unsigned c) -> Chunk {
        Chunk ck;
        unsigned long long b0, b1, arena;
        if (c == 0) { ck.base = reinterpret_cast<const char *>(p.dq); b0 = (unsigned long long)story * row_bytes; b1 = b0 + row_bytes; arena = p.q_bytes; }
        else {
            const unsigned r0 = (c - 1u) * fl.R, n = min(fl.R, S - r0);
            ck.base = reinterpret_cast<const char *>(p.dm); b0 = (soff + r0) * row_bytes; b1 = b0 + (unsigned long long)n * row_bytes; arena = p.m_bytes;
        }
        ck.a0 = b0 & ~15ull;
        unsigned long long a1 = (b1 + 15ull) & ~15ull;
        ck.short_tail = false;
        if (a1 > arena) { a1 = arena & ~15ull; ck.short_tail = a1 < b1; }       // the arena's last bytes when its size is not a multiple of 16
        ck.bytes = (a1 > ck.a0) ? (unsigned)(a1 - ck.a0) : 0u;
        ck.skew = (unsigned)(b0 - ck.a0);
        return ck;
    };
    auto issue = [&](unsigned story, unsigned long long soff, unsigned S, unsigned c) {
        if (lane == 0) {
            const Chunk ck = chunk_of(story, soff, S, c);
            const unsigned b = c % fl.NB;
            if (ck.bytes) {
                mbar_expect_tx(bar0 + 8u * b, ck.bytes);
                bulk_g2s(stage0 + b * fl.buf_bytes, ck.base + ck.a0, ck.bytes, bar0 + 8u * b);
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * b) : "memory");
            }
        }
    };
    auto prefetch_ahead = [&](unsigned idx) {
        // the dense rows of the story that will be claimed pf_dist claims from now: into L2, no shared memory involved
        if (!DENSE || p.pf_mode == 0u || p.work_list) return;
        const unsigned j = idx + p.pf_dist;
        if (j >= n_work) return;
        const unsigned story = p.story0 + j;
        const unsigned long long soff = p.sen_off[story];
        const unsigned S = (unsigned)(p.sen_off[story + 1] - soff);
        const unsigned long long b0 = soff * row_bytes & ~15ull;
        unsigned long long b1 = ((soff + S) * row_bytes + 15ull) & ~15ull;
        if (b1 > p.m_bytes) b1 = p.m_bytes & ~15ull;
        const char *base = reinterpret_cast<const char *>(p.dm);
        if (p.pf_mode == 1u) {
            if (lane == 0)
                for (unsigned long long o = b0; o < b1; o += 16384ull) bulk_prefetch_l2(base + o, (unsigned)min(16384ull, b1 - o));
        } else {
            for (unsigned long long o = b0 + 128ull * lane; o < b1; o += 4096ull) line_prefetch_l2(base + o);
        }
        if (lane == 1) {
            const unsigned long long qb = (unsigned long long)story * row_bytes;
            line_prefetch_l2(reinterpret_cast<const char *>(p.dq) + qb);
            if (row_bytes > 128u) line_prefetch_l2(reinterpret_cast<const char *>(p.dq) + qb + row_bytes - 4u);
        }
    };

    unsigned phase_bits = 0;                         // parity to wait for, per staging buffer
    unsigned idx_next = claim();
    unsigned long long soff_next = 0;
    unsigned S_next = 0, w_next = 0;
    auto open_next = [&]() {
        // resolve the next work item and, for the dense source, put its first chunks in flight
        if (idx_next >= n_work) return;
        w_next = p.work_list ? p.work_list[idx_next] : idx_next;
        const unsigned story = p.story0 + w_next;
        soff_next = p.sen_off[story];
        S_next = (unsigned)(p.sen_off[story + 1] - soff_next);
        if (DENSE) {
            const unsigned nch = 1u + (S_next + fl.R - 1u) / fl.R;
            for (unsigned c = 0; c < min(nch, fl.NB); c++) issue(story, soff_next, S_next, c);
            prefetch_ahead(idx_next);
        }
    };
    open_next();

#pragma unroll 1
    for (;;) {
        if (idx_next >= n_work) break;
        const unsigned w = w_next;
        const unsigned story = p.story0 + w;
        const unsigned long long soff = soff_next;
        const unsigned S = S_next;
        unsigned ans_idx = ANS_NONE;
        bool decline = false;

        if (DENSE) {
            // ---- stream the story's rows through the staging buffers and compact them ----
            const unsigned nch = 1u + (S + fl.R - 1u) / fl.R;
            unsigned base = 0;
            bool irregular = false;
#pragma unroll 1
            for (unsigned c = 0; c < nch; c++) {
                const unsigned b = c % fl.NB;
                mbar_wait(bar0 + 8u * b, (phase_bits >> b) & 1u);
                phase_bits ^= 1u << b;
                const Chunk ck = chunk_of(story, soff, S, c);
                irregular |= ck.short_tail;
                const float *buf = reinterpret_cast<const float *>(ws + fl.o_stage + b * fl.buf_bytes);
                const unsigned rows = (c == 0) ? 1u : min(fl.R, S - (c - 1u) * fl.R);
                const unsigned r_first = (c == 0) ? 0u : 1u + (c - 1u) * fl.R;
                unsigned fr = ck.skew >> 2;
                for (unsigned rr = 0; rr < rows; rr++, fr += V) {
                    base = aligned ? scan_row_s<true>(p, buf, fr, ent_s, fl.LW, base, lane, irregular)
                                   : scan_row_s<false>(p, buf, fr, ent_s, fl.LW, base, lane, irregular);
                    if (lane == 0) rend_s[r_first + rr] = (unsigned short)min(base, 0xFFFFu);
                }
                __syncwarp();
                if (c + fl.NB < nch) issue(story, soff, S, c + fl.NB);
            }
            decline = irregular || base > fl.LW;
            if (p.da) {
                // answer: index of the (last) 1.0 in the one-hot row (lib/layer_cuda.cu:2196)
                const float *arow = p.da + (size_t)story * V;
                for (unsigned c0 = 0; c0 < V; c0 += 32) {
                    const unsigned c = c0 + lane;
                    const bool hot = (c < V) && (ldg_stream1(arow + c) == 1.0f);
                    const unsigned bb = __ballot_sync(0xffffffffu, hot);
                    if (bb) ans_idx = c0 + 31 - __clz(bb);
                }
            }
            // the buffers are free again: claim the next story and put its first chunks in flight under this forward
            idx_next = claim();
            open_next();
        } else {
            const unsigned char *rec = p.rec + (size_t)w * p.rec_stride;
            const unsigned *hdr = reinterpret_cast<const unsigned *>(rec);
            const unsigned n_ent = hdr[0], flags = hdr[1], n_exc = hdr[4];
            ans_idx = hdr[2];
            decline = (flags != 0u || n_exc != 0u || n_ent > fl.LW);
            if (!decline) {
                const unsigned short *rend_g = reinterpret_cast<const unsigned short *>(rec + p.off_rend);
                for (unsigned r = lane; r < S + 1; r += 32) rend_s[r] = rend_g[r];
                const unsigned *ent_g = reinterpret_cast<const unsigned *>(rec + p.off_ent);
                for (unsigned k = lane; k < n_ent; k += 32) ent_s[k] = (unsigned short)(ent_g[k] * DP);
            }
            idx_next = claim();
            open_next();
        }
        if (lane < 17) cnt_s[lane] = 0;
        __syncwarp();
        if (decline) {
            // not a regular story: leave it to the next tier
            if (lane == 0) p.slow_list[atomicAdd(p.slow_count, 1u)] = w;
            continue;
        }
        if (SWAR) {
            // per-row biases of the hops and the narrow test (every row: B <= 127, B << ka <= 127).  cm10[column] packs
            // the column maxima of up to three hops in 10-bit fields, so one add per entry serves all hops (rows of
            // more than 8 entries could overflow a field and are not narrow)
            const unsigned *cm10 = reinterpret_cast<const unsigned *>(smem + fl.sCM);
            bool ok = true;
            for (unsigned r = lane; r < S; r += 32) {
                const unsigned e0 = rend_s[r], e1 = rend_s[r + 1];
                unsigned B10 = 0;
                for (unsigned e = e0; e < e1; e++) B10 += cm10[ent_s[e] >> lgDP];
                ok = ok && (e1 - e0 <= 8u);
                for (unsigned h = 0; h < p.H; h++) {
                    const unsigned B = (B10 >> (10u * h)) & 0x3FFu;
                    const int ka = p.fa[h] - p.fw[h];
                    ok = ok && ((ka > 0 ? (B << ka) : B) <= 127u);
                    brow[h * p.S_pad + r] = (unsigned char)min(B, 255u);
                }
            }
            if (!__all_sync(0xffffffffu, ok)) {
                if (lane == 0) p.slow_list[atomicAdd(p.slow_count, 1u)] = w;
                continue;
            }
        }
        {
            // Rows ordered by entry count (counting sort, counts capped at 15): the G rows of a pass then have nearly
            // the same length and the gather loop does not idle on the longest one.  Only the order in which rows are
            // embedded changes; every score is stored at its own row index.
            for (unsigned r0 = 0; r0 < S; r0 += 32) {
                const unsigned r = r0 + lane;
                const unsigned key = (r < S) ? min((unsigned)(rend_s[r + 1] - rend_s[r]), 15u) : 16u;
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                if (lane == (unsigned)(__ffs((int)peers) - 1)) cnt_s[key] += (unsigned)__popc(peers);
                __syncwarp();
            }
            {
                const unsigned c = (lane < 17) ? cnt_s[lane] : 0u;
                unsigned incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((int)lane >= o) incl += t;
                }
                __syncwarp();
                if (lane < 17) cnt_s[lane] = incl - c;
            }
            __syncwarp();
            for (unsigned r0 = 0; r0 < S; r0 += 32) {
                const unsigned r = r0 + lane;
                const unsigned key = (r < S) ? min((unsigned)(rend_s[r + 1] - rend_s[r]), 15u) : 16u;
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                const unsigned base = cnt_s[key];
                const unsigned rank = (unsigned)__popc(peers & ((1u << lane) - 1u));
                if (r < S) perm[base + rank] = (unsigned short)r;
                __syncwarp();
                if (rank == 0) cnt_s[key] = base + (unsigned)__popc(peers);
                __syncwarp();
            }
        }

        int acc[16];
        // ---- question embedding u0 = Q_w0(sum)                                 MemN2N.c:826, layer_cuda.cu:49 ----
        embed_glob<LPR>(p, wso, lane, p.img + p.offB + 16u * q, (g == 0) ? 0 : -1, acc, sel);
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
        int fu = p.fw[0];
        if (DUMP && p.dbg.dev_u0)
            for (unsigned j = lane; j < d; j += 32) p.dbg.dev_u0[(size_t)story * d + j] = (float)uvec[j] / (float)(1 << fu);

#pragma unroll 1
        for (unsigned h = 0; h < p.H; h++) {
            const int fw = p.fw[h], lw = p.lw[h];
            const int fa = p.fa[h], la = p.la[h];
            const int ff = p.ff[h], lf = p.lf[h];
            const int fb = p.fb, lb = p.lb;
            const int mb = (1 << fb) - 1;

            // u operand Q_bin(u) as int32                                        MemN2N.c:847,873
            for (unsigned j = lane; j < DP; j += 32) ub32[j] = (j < d) ? qi_requant((int)uvec[j], fu, lb, fb) : 0;
            __syncwarp();
            const int ka = fa - fw;
            // scorer constants: clamp limit of the re-quantised memory value and -L*u per dim
            const int Ls = (ka < 0) ? min(la, lw >> (-ka)) : la;
            const bool simple = (ka == 0) ? (la == lw) : (ka > 0 ? (la <= (lw << ka)) : true);
            int ub[16], cub[16];
            unsigned au[16];
            unsigned su_bits = 0;
            // Mode 3 in nine bits.  The reference compares 31-bit magnitudes |x| * 2^(31-iwl) (layer_cuda.cu:384-428) and
            // keeps bits 30..24 of their difference (same signs) or sum (opposite signs, whose carry into bit 31 flips the
            // sign).  Every magnitude is a multiple of 2^23 (codes have at most 7 bits, k = 31-iwl-frac >= 23) or the
            // saturated 0x7FFFFFFF, whose low 23 one-bits never borrow or carry against multiples of 2^23: so the
            // element is exact on A = sat9(code << (k-23)) in [-255, 255]:  w = |A_m - A_u|,  e = 127 - ((w >> 1) & 127),
            // negative iff the signs differ and w < 256.  The -2^iwl -> 0 encode quirk keeps its sign on the
            // memory side (the sign test uses the code); a query holding it takes the literal path below.
            // Checked on every pair of 8-bit codes against tests/golden/kat_appx_element.npz.
            const int sh_m = 8 - p.ia[h] - fw, sh_u = 8 - p.ia[h] - fu;
            bool fast3 = (MODE == 3) && sh_m >= 0 && sh_u >= 0 && sh_m <= 8 && sh_u <= 8;
            const bool sat_m = fast3 && ((127 << sh_m) >= 256);
            int U9[16];
            if (MODE == 3) {
                const uint4 t = *reinterpret_cast<const uint4 *>(uvec + 16 * q);
                const unsigned tw[4] = {t.x, t.y, t.z, t.w};
                bool quirk = false;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int uc = sbyte(tw[j >> 2], j & 3);
                    unsigned s_, m_;
                    appx_encode(uc, fu, p.ia[h], s_, m_);     // layer.c:215-233, layer_cuda.cu:2515
                    au[j] = m_;
                    su_bits |= (s_ >> 31) << j;
                    const int tu = fast3 ? (uc << sh_u) : 0;
                    quirk |= (tu == -256);
                    U9[j] = (16u * q + j < d) ? max(-255, min(tu, 255)) : 255;     // padding dims: w = 255 -> e = 0
                }
                fast3 = fast3 && !__any_sync(0xffffffffu, quirk);
            } else {
#pragma unroll
                for (int w4 = 0; w4 < 4; w4++) {
                    const int4 t = *reinterpret_cast<const int4 *>(ub32 + 16 * q + 4 * w4);
                    ub[4 * w4 + 0] = t.x; ub[4 * w4 + 1] = t.y; ub[4 * w4 + 2] = t.z; ub[4 * w4 + 3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < 16; j++) cub[j] = -Ls * ub[j];
            }

            // ---- memory embedding + addressing, G rows per pass ----
            if (SWAR) {
                SwarQuery sq;
                {
                    // Q_bin(u) and the saturation thresholds as bytes, built once per hop by the whole warp in the (idle until
                    // the answer phase) ufl region, then 16 bytes per lane
                    const unsigned char *tau = smem + fl.sTAU;
                    unsigned char *ub8_s = reinterpret_cast<unsigned char *>(ufl), *tw_s = ub8_s + DP;
                    for (unsigned j = lane; j < DP; j += 32) {
                        const int u = ub32[j];
                        ub8_s[j] = (unsigned char)(u & 0xFF);
                        tw_s[j] = tau[abs(u)];
                    }
                    __syncwarp();
                    const uint4 u4 = *reinterpret_cast<const uint4 *>(ub8_s + 16u * q), t4 = *reinterpret_cast<const uint4 *>(tw_s + 16u * q);
                    const unsigned uw4[4] = {u4.x, u4.y, u4.z, u4.w}, tw4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                    for (int w4 = 0; w4 < 4; w4++) {
                        const unsigned uw = uw4[w4];
                        sq.Uw[w4] = uw; sq.Tw[w4] = tw4[w4];
                        sq.U0[w4] = uw & SW_1; sq.U0s[w4] = (uw & SW_1) << 1; sq.U1[w4] = uw & 0x02020202u;
                        sq.Us4[w4] = (uw >> 5) & 0x04040404u;
                    }
                    __syncwarp();
                }
                const unsigned tabq = fl.sA[h] + 16u * q;
                const unsigned zaddr = wso + fl.o_zent;
#pragma unroll 1
                for (unsigned r0 = 0; r0 < S; r0 += G) {
                    const unsigned r = (r0 + g < S) ? (unsigned)perm[r0 + g] : S;
                    unsigned beg = 0, len = 0, Bw = 0;
                    if (r < S) {
                        beg = rend_s[r];
                        len = rend_s[r + 1] - beg;
                        Bw = (unsigned)brow[h * p.S_pad + r] * SW_1;
                    }
                    const unsigned maxlen = __reduce_max_sync(0xffffffffu, len);
                    unsigned acc4[4] = {0u, 0u, 0u, 0u};
                    unsigned ea = wso + 2u * beg;
#pragma unroll 2
                    for (unsigned k = 0; k < maxlen; k++, ea += 2u) {
                        const unsigned coff = *reinterpret_cast<const unsigned short *>(smem + ((k < len) ? ea : zaddr));
                        const uint4 t = *reinterpret_cast<const uint4 *>(smem + tabq + coff);
                        acc4[0] += t.x; acc4[1] += t.y; acc4[2] += t.z; acc4[3] += t.w;
                    }
                    unsigned y4[4], flag;
                    int part;
                    if (ka == 0) part = swar_score<0>(acc4, Bw, sq, y4, flag);
                    else if (ka > 0) part = swar_score<1>(acc4, Bw, sq, y4, flag);
                    else part = swar_score<-1>(acc4, Bw, sq, y4, flag);
                    int tot = group_sum<LPR>(part) >> 2;               // exact: a multiple of 4
#pragma unroll
                    for (int o = 1; o < LPR; o <<= 1) flag |= __shfl_xor_sync(0xffffffffu, flag, o);
                    if (__any_sync(0xffffffffu, flag != 0u)) {
                        // some product of this row saturates: the reference order, product by product
                        int sp = 0;
#pragma unroll
                        for (int j = 0; j < 16; j++) sp += qi_mul(sbyte(y4[j >> 2], j & 3), ub[j], la, fb);
                        sp = group_sum<LPR>(sp);
                        if (flag) tot = sp;
                    }
                    if (q == 0 && r < S) sc[r] = qi_clamp(tot, la);
                }
            } else
#pragma unroll 1
            for (unsigned r0 = 0; r0 < S; r0 += G) {
                const unsigned r = (r0 + g < S) ? (unsigned)perm[r0 + g] : S;
                embed_smem<LPR>(p, wso, lane, fl.sA[h] + 16u * q, (r < S) ? (int)(r + 1) : -1, acc, sel);
                int part = 0;
                if (MODE == 3 && fast3) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int n = qi_clamp(acc[j], lw);
                        int t = n << sh_m;
                        if (sat_m) t = (t == -256) ? 0 : max(-255, min(t, 255));
                        const int w9 = (int)__sad(t, U9[j], 0u);
                        const int e = ~(w9 >> 1) & 0x7F;
                        part += (((n ^ U9[j]) < 0) && (w9 < 256)) ? -e : e;
                    }
                } else if (MODE == 3) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        unsigned sm, am;
                        appx_encode(qi_clamp(acc[j], lw), fw, p.ia[h], sm, am);
                        const unsigned sv = ((su_bits >> j) & 1u) << 31;
                        part += (16u * q + j < d) ? appx_element_x128(sm, am, sv, au[j]) : 0;
                    }
                } else if (!simple) {
#pragma unroll
                    for (int j = 0; j < 16; j++) part += qi_mul(qi_requant(qi_clamp(acc[j], lw), fw, la, fa), ub[j], la, fb);
                } else {
                    if (ka == 0) part = score_fast<0>(acc, ub, cub, Ls, 0, la, fb, mb);
                    else if (ka > 0) part = score_fast<1>(acc, ub, cub, Ls, ka, la, fb, mb);
                    else part = score_fast<-1>(acc, ub, cub, Ls, -ka, la, fb, mb);
                    part -= 16 * la;
                }
                const int tot = group_sum<LPR>(part);
                if (q == 0 && r < S) sc[r] = (MODE == 3) ? tot : qi_clamp(tot, la);
            }
            __syncwarp();

            // ---- attention normalisation (layer_cuda.cu:1895-1916, 1969-2060) ----
            float mx = -INFINITY;
            for (unsigned r = lane; r < S; r += 32) {
                float sv;
                if (MODE == 3) {
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
            if (DUMP && p.dbg.dev_s)
                for (unsigned r = lane; r < S; r += 32) p.dbg.dev_s[(size_t)h * p.sum_sen + soff + r] = ex[r];
            float tsum = 0.0f;
            for (unsigned r = lane; r < S; r += 32) { const float e = __expf(ex[r] - mx); ex[r] = e; tsum += e; }
            __syncwarp();
            // The reference's weight is p = fl32(fl64(e / total)) with the double total accumulated in slot order, and only
            // trunc(p * 2^ff) enters the read.  A float tree total differs from the double one by < 2^-19 relative
            // (S <= 4096), so  v = e / tsum * 2^ff  fixes the code whenever it is not within 4e-6 relative of an integer
            // >= 1; otherwise (exact ties at the top are the usual cause: p = 1/2, 1/4) the warp forms the exact total.
            bool exact_total = !p.fast_softmax;
            unsigned nnz = 0;
            if (!exact_total) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
                const float sc2 = (float)(1 << ff) / tsum;
                const float tol = (float)(S / 32u + 8u) * 1.1920929e-7f;       // float total: S/32 sequential + 5 tree additions, the scaling
                bool amb = false;
                for (unsigned r = lane; r < S; r += 32) {
                    const float v = ex[r] * sc2;
                    const float n = rintf(v);
                    amb |= (n >= 1.0f) && (fabsf(v - n) <= tol * v);
                }
                exact_total = __any_sync(0xffffffffu, amb);
                if (!exact_total) {
#pragma unroll 1
                    for (unsigned r0 = 0; r0 < S; r0 += 32) {
                        const unsigned r = r0 + lane;
                        unsigned code = 0;
                        if (r < S) code = (unsigned)min((int)(ex[r] * sc2), lf);
                        if (DUMP && p.dbg.dev_pcode && r < S) p.dbg.dev_pcode[(size_t)h * p.sum_sen + soff + r] = (unsigned char)code;
                        const unsigned b = __ballot_sync(0xffffffffu, code != 0u);
                        if (code) {
                            const unsigned k = nnz + __popc(b & ((1u << lane) - 1u));
                            sc[k] = (int)r;
                            pq[k] = (unsigned char)code;
                        }
                        nnz += __popc(b);
                    }
                }
            }
            if (exact_total) {
                double total = 0.0;
#pragma unroll 4
                for (unsigned r = 0; r < S; r++) total += (double)ex[r];
#pragma unroll 1
                for (unsigned r0 = 0; r0 < S; r0 += 32) {
                    const unsigned r = r0 + lane;
                    unsigned code = 0;
                    if (r < S) code = (unsigned)qi_encode((float)((double)ex[r] / total), p.iff[h], ff);      // layer_cuda.cu:561
                    if (DUMP && p.dbg.dev_pcode && r < S) p.dbg.dev_pcode[(size_t)h * p.sum_sen + soff + r] = (unsigned char)code;
                    const unsigned b = __ballot_sync(0xffffffffu, code != 0u);
                    if (code) {
                        const unsigned k = nnz + __popc(b & ((1u << lane) - 1u));
                        sc[k] = (int)r;
                        pq[k] = (unsigned char)code;
                    }
                    nnz += __popc(b);
                }
            }
            __syncwarp();

            if (LPR == 4 && (!p.lin_map || p.lut)) {
                // ---- DP == 64: lane l owns dims 2l, 2l+1 of o, g and u.  Weighted read over the selected slots (layer_cuda.cu:547-579):
                //      their C_h rows are gathered with one coalesced 64-byte table row per entry, two slots at a time ----
                const unsigned c0 = 2u * lane;
                int o0 = 0, o1 = 0;
                const unsigned char *ctab = p.img + p.offC[h];
#pragma unroll 1
                for (unsigned k = 0; k < nnz; k += 2) {
                    const bool two = (k + 1 < nnz);
                    const unsigned ra = (unsigned)sc[k] + 1u, rb = two ? (unsigned)sc[k + 1] + 1u : ra;      // record rows (0 = question)
                    const unsigned ba = rend_s[ra - 1], bb = rend_s[rb - 1];
                    const unsigned na = rend_s[ra] - ba, nb = two ? rend_s[rb] - bb : 0u;
                    int a0 = 0, a1 = 0, b0 = 0, b1 = 0;
                    gather2x2(smem_u32(ws) + 2u * ba, na, smem_u32(ws) + 2u * bb, nb, ctab, lane, a0, a1, b0, b1);
                    const int pa = (int)pq[k], pb = two ? (int)pq[k + 1] : 0;
                    o0 += qi_mul(pa, qi_requant(qi_clamp(a0, lw), fw, lf, ff), lf, ff) + qi_mul(pb, qi_requant(qi_clamp(b0, lw), fw, lf, ff), lf, ff);
                    o1 += qi_mul(pa, qi_requant(qi_clamp(a1, lw), fw, lf, ff), lf, ff) + qi_mul(pb, qi_requant(qi_clamp(b1, lw), fw, lf, ff), lf, ff);
                }
                o0 = qi_clamp(o0, lf); o1 = qi_clamp(o1, lf);
                if (DUMP && p.dbg.dev_o) {
                    if (c0 < d) p.dbg.dev_o[((size_t)h * p.n_total + story) * d + c0] = (float)o0 / (float)(1 << ff);
                    if (c0 + 1 < d) p.dbg.dev_o[((size_t)h * p.n_total + story) * d + c0 + 1] = (float)o1 / (float)(1 << ff);
                }
                // ---- linear map (MemN2N.c:873, layer_cuda.cu:49-68): g[i] = Q_w(sum_j T[j][Q_bin(u[j])][i]), one byte of a product-table
                //      row per (j, i); update (MemN2N.c:889, layer_cuda.cu:1535) ----
                int g0, g1, gfrac;
                if (p.lin_map) {
                    g0 = 0; g1 = 0;
                    const unsigned short *lut16 = reinterpret_cast<const unsigned short *>(p.lut + p.offL[h]) + lane;
#pragma unroll 25
                    for (unsigned j = 0; j < d; j++) {
                        const unsigned row = j * 255u + (unsigned)(ub32[j] + 127);
                        const unsigned w_ = ldg_na_u16(lut16 + (size_t)row * (DP / 2));
                        g0 += (int)(signed char)(w_ & 0xFFu);
                        g1 += (int)(signed char)(w_ >> 8);
                    }
                    g0 = qi_clamp(g0, lw); g1 = qi_clamp(g1, lw);
                    gfrac = fw;
                } else {
                    g0 = (c0 < d) ? (int)uvec[c0] : 0; g1 = (c0 + 1 < d) ? (int)uvec[c0 + 1] : 0; gfrac = fu;
                }
                if (DUMP && p.dbg.dev_g) {
                    if (c0 < d) p.dbg.dev_g[((size_t)h * p.n_total + story) * d + c0] = (float)g0 / (float)(1 << gfrac);
                    if (c0 + 1 < d) p.dbg.dev_g[((size_t)h * p.n_total + story) * d + c0 + 1] = (float)g1 / (float)(1 << gfrac);
                }
                const int n0 = (c0 < d) ? qi_clamp(qi_requant(g0, gfrac, lf, ff) + o0, lf) : 0;
                const int n1 = (c0 + 1 < d) ? qi_clamp(qi_requant(g1, gfrac, lf, ff) + o1, lf) : 0;
                __syncwarp();                                        // every lane has read uvec / ub32 of this hop
                *reinterpret_cast<unsigned short *>(uvec + c0) = (unsigned short)((n0 & 0xFF) | ((n1 & 0xFF) << 8));
            } else {
            // ---- weighted read over the slots with a non-zero quantised weight (layer_cuda.cu:547-579) ----
                int oacc[16];
    #pragma unroll
                for (int j = 0; j < 16; j++) oacc[j] = 0;
                const unsigned char *ctab = p.img + p.offC[h] + 16u * q;
    #pragma unroll 1
                for (unsigned k0 = 0; k0 < nnz; k0 += G) {
                    const unsigned k = k0 + g;
                    const int r = (k < nnz) ? sc[k] : -1;
                    const int pc = (k < nnz) ? (int)pq[k] : 0;
                    embed_glob<LPR>(p, wso, lane, ctab, (r >= 0) ? r + 1 : -1, acc, sel);
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
                if (DUMP && p.dbg.dev_o)
                    for (unsigned j = lane; j < d; j += 32) p.dbg.dev_o[((size_t)h * p.n_total + story) * d + j] = (float)ovec[j] / (float)(1 << ff);
    
                // ---- linear map (MemN2N.c:873, layer_cuda.cu:49-68) and update (MemN2N.c:889, layer_cuda.cu:1535) ----
                if (p.lin_map && p.lut) {
                    // g[i] = Q_w(sum_j T[j][Q_bin(u[j])][i]): every product Q_w(Q_w(Hm[i][j]) * Q_bin(u[j])) is a function of
                    // one 8-bit activation, so the d*d quantised products become a gather-and-sum of d table rows
                    int gacc[16];
    #pragma unroll
                    for (int k = 0; k < 16; k++) gacc[k] = 0;
                    const signed char *lut = p.lut + p.offL[h] + 16u * q;
    #pragma unroll 1
                    for (unsigned jb = 0; jb < d; jb += 8 * G) {
                        // eight independent 128-bit gathers (L2-resident table) in flight per lane before the first use
                        uint4 t[8];
    #pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const unsigned j = jb + (unsigned)i * G + g;
                            t[i] = make_uint4(0u, 0u, 0u, 0u);
                            if (j < d) t[i] = __ldg(reinterpret_cast<const uint4 *>(lut + (size_t)(j * 255u + (unsigned)(ub32[j] + 127)) * DP));
                        }
    #pragma unroll
                        for (int i = 0; i < 8; i++) {
                            gacc[0] = __dp4a((int)t[i].x, sel[0], gacc[0]);   gacc[1] = __dp4a((int)t[i].x, sel[1], gacc[1]);
                            gacc[2] = __dp4a((int)t[i].x, sel[2], gacc[2]);   gacc[3] = __dp4a((int)t[i].x, sel[3], gacc[3]);
                            gacc[4] = __dp4a((int)t[i].y, sel[0], gacc[4]);   gacc[5] = __dp4a((int)t[i].y, sel[1], gacc[5]);
                            gacc[6] = __dp4a((int)t[i].y, sel[2], gacc[6]);   gacc[7] = __dp4a((int)t[i].y, sel[3], gacc[7]);
                            gacc[8] = __dp4a((int)t[i].z, sel[0], gacc[8]);   gacc[9] = __dp4a((int)t[i].z, sel[1], gacc[9]);
                            gacc[10] = __dp4a((int)t[i].z, sel[2], gacc[10]); gacc[11] = __dp4a((int)t[i].z, sel[3], gacc[11]);
                            gacc[12] = __dp4a((int)t[i].w, sel[0], gacc[12]); gacc[13] = __dp4a((int)t[i].w, sel[1], gacc[13]);
                            gacc[14] = __dp4a((int)t[i].w, sel[2], gacc[14]); gacc[15] = __dp4a((int)t[i].w, sel[3], gacc[15]);
                        }
                    }
    #pragma unroll
                    for (int o = LPR; o < 32; o <<= 1)
    #pragma unroll
                        for (int k = 0; k < 16; k++) gacc[k] += __shfl_xor_sync(0xffffffffu, gacc[k], o);
                    if (g == 0) {
                        const uint4 o4 = *reinterpret_cast<const uint4 *>(ovec + 16 * q);
                        const unsigned ow[4] = {o4.x, o4.y, o4.z, o4.w};
                        unsigned packed[4];
    #pragma unroll
                        for (int w4 = 0; w4 < 4; w4++) {
                            unsigned v = 0;
    #pragma unroll
                            for (int b = 0; b < 4; b++) {
                                const int g_w = qi_clamp(gacc[4 * w4 + b], lw);
                                if (DUMP && p.dbg.dev_g && 16u * q + 4 * w4 + b < d)
                                    p.dbg.dev_g[((size_t)h * p.n_total + story) * d + 16u * q + 4 * w4 + b] = (float)g_w / (float)(1 << fw);
                                const int a_f = qi_requant(g_w, fw, lf, ff);
                                v |= ((unsigned)(qi_clamp(a_f + sbyte(ow[w4], b), lf) & 0xFF)) << (8 * b);
                            }
                            packed[w4] = v;
                        }
                        *reinterpret_cast<uint4 *>(uvec + 16 * q) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    }
                } else {
                    const int lw2 = 2 * lw;
                    const signed char *Hg = reinterpret_cast<const signed char *>(p.img + p.offH[h]);      // int8 [d][HS], L2-resident
    #pragma unroll 1
                    for (unsigned i0 = 0; i0 < d; i0 += 32) {
                        const unsigned i = i0 + lane;
                        int a_f = 0, g_w = 0;
                        if (p.lin_map) {
                            const unsigned *hrow = reinterpret_cast<const unsigned *>(Hg + (size_t)min(i, d - 1) * p.HS);
                            int s_ = 0;
                            const unsigned d4 = (d + 3) / 4;
    #pragma unroll 2
                            for (unsigned j4 = 0; j4 < d4; j4++) {
                                const unsigned hw = __ldg(hrow + j4);
                                const int4 uu = *reinterpret_cast<const int4 *>(ub32 + 4 * j4);
                                s_ += clamp_biased(shr0m(sbyte_prmt<0>(hw) * uu.x, fb, mb), lw, lw2);
                                s_ += clamp_biased(shr0m(sbyte_prmt<1>(hw) * uu.y, fb, mb), lw, lw2);
                                s_ += clamp_biased(shr0m(sbyte_prmt<2>(hw) * uu.z, fb, mb), lw, lw2);
                                s_ += clamp_biased(shr0m(((int)hw >> 24) * uu.w, fb, mb), lw, lw2);
                            }
                            g_w = qi_clamp(s_ - (int)(4u * d4) * lw, lw);
                            a_f = qi_requant(g_w, fw, lf, ff);
                        } else if (i < d) {
                            g_w = (int)uvec[i];
                            a_f = qi_requant(g_w, fu, lf, ff);
                        }
                        __syncwarp();
                        if (i < d) {
                            if (DUMP && p.dbg.dev_g) p.dbg.dev_g[((size_t)h * p.n_total + story) * d + i] = (float)g_w / (float)(1 << (p.lin_map ? fw : fu));
                            uvec[i] = (signed char)qi_clamp(a_f + (int)ovec[i], lf);
                        }
                    }
                }
            }
            fu = ff;
            __syncwarp();
            if (DUMP && p.dbg.dev_u)
                for (unsigned j = lane; j < d; j += 32) p.dbg.dev_u[((size_t)h * p.n_total + story) * d + j] = (float)uvec[j] / (float)(1 << fu);
        }

        // ---- answer projection, sequential fp32 (MemN2N.c:902-906, layer_cuda.cu:69-82), argmax on the
        //      probabilities (layer_cuda.cu:1918-1939) ----
        // Only the rows that can hold the largest probability need their fp32 logit: z_i = sum_j fl(W_ij u_j) in index
        // order.  With W = s (W8 + eps), |eps| <= 1/2, and u = n / 2^fu, the integer dot D_i = sum_j W8_ij n_j (IDP.4A)
        // satisfies |z_i 2^fu / s - D_i| <= E = |n|_1 / 2 + gamma_{d+1} 127 |n|_1 (quantisation of W + fp32 rounding of
        // the chain).  Every row whose probability can tie with the maximum has z_i >= z_max - 1e-5, hence
        // D_i >= D_max - (2 E + 1e-5 2^fu / s): those candidates (1.4 rows on average) are computed exactly, from the
        // fp32 rows in L2; all other rows get -inf, i.e. e = 0.  A tie among candidates, or h[y] requested, computes
        // every row exactly.
        for (unsigned j = lane; j < DP; j += 32) ufl[j] = (j < d) ? (float)uvec[j] / (float)(1 << fu) : 0.0f;
        __syncwarp();
        const unsigned d4 = (d + 3) / 4;
        const float *Wg = reinterpret_cast<const float *>(p.img + p.offW);
        auto exact_z = [&](unsigned i) {
            const float4 *wr = reinterpret_cast<const float4 *>(Wg + (size_t)i * p.WS);
            float z = 0.0f;
#pragma unroll 4
            for (unsigned j4 = 0; j4 < d4; j4++) {
                const float4 ww = __ldg(wr + j4);
                const float4 uu = *reinterpret_cast<const float4 *>(ufl + 4 * j4);
                z = __fadd_rn(z, __fmul_rn(ww.x, uu.x));
                z = __fadd_rn(z, __fmul_rn(ww.y, uu.y));
                z = __fadd_rn(z, __fmul_rn(ww.z, uu.z));
                z = __fadd_rn(z, __fmul_rn(ww.w, uu.w));
            }
            return z;
        };
        float zmax = -INFINITY;
        unsigned n_cand = 0, cand_idx = 0;
        bool need_full = !p.w8_ok || p.want_h;
        if (!need_full) {
            int n1 = 0;
            for (unsigned j = lane; j < DP; j += 32) n1 += abs((int)uvec[j]);
            n1 = __reduce_add_sync(0xffffffffu, n1);
            const int T = n1 + (n1 >> 6) + p.ans_margin + 2;
            int *zi = reinterpret_cast<int *>(zbuf);
            const unsigned nw16 = (d + 15) / 16;
            int Dmax = INT_MIN;
#pragma unroll 1
            for (unsigned i0 = 0; i0 < V; i0 += 128) {
                int D[4] = {0, 0, 0, 0};
                unsigned wrow[4];
#pragma unroll
                for (int k = 0; k < 4; k++) wrow[k] = fl.sW8 + min(i0 + 32 * k + lane, V - 1) * p.W8S;
#pragma unroll 1
                for (unsigned w16 = 0; w16 < nw16; w16++) {
                    const uint4 uu = *reinterpret_cast<const uint4 *>(uvec + 16 * w16);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint4 ww = *reinterpret_cast<const uint4 *>(smem + wrow[k] + 16u * w16);
                        D[k] = __dp4a((int)ww.x, (int)uu.x, D[k]);
                        D[k] = __dp4a((int)ww.y, (int)uu.y, D[k]);
                        D[k] = __dp4a((int)ww.z, (int)uu.z, D[k]);
                        D[k] = __dp4a((int)ww.w, (int)uu.w, D[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned i = i0 + 32 * k + lane;
                    if (i < V) { zi[i] = D[k]; Dmax = max(Dmax, D[k]); }
                }
            }
            Dmax = __reduce_max_sync(0xffffffffu, Dmax);
            const int thr = Dmax - T;
#pragma unroll 1
            for (unsigned i0 = 0; i0 < V; i0 += 32) {
                const unsigned i = i0 + lane;
                if (i < V) {
                    float z = -INFINITY;
                    const bool cnd = zi[i] >= thr;
                    if (cnd) z = exact_z(i);
                    zbuf[i] = z;
                    zmax = fmaxf(zmax, z);
                    if (DUMP && p.dbg.dev_cand) p.dbg.dev_cand[(size_t)story * V + i] = cnd ? 1 : 0;
                    if (DUMP && p.dbg.dev_z) p.dbg.dev_z[(size_t)story * V + i] = z;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
            __syncwarp();
#pragma unroll 1
            for (unsigned i0 = 0; i0 < V; i0 += 32) {
                const unsigned i = i0 + lane;
                const bool cand = (i < V) && (__expf(zbuf[i] - zmax) >= 0.99999905f);
                const unsigned b = __ballot_sync(0xffffffffu, cand);
                if (b) { n_cand += __popc(b); cand_idx = i0 + 31 - __clz(b); }
            }
            need_full = n_cand > 1;                          // near-tie: the double total decides, every row exactly
        }
        if (need_full) {
            zmax = -INFINITY;
#pragma unroll 1
            for (unsigned i0 = 0; i0 < V; i0 += 32) {
                const unsigned i = i0 + lane;
                if (i < V) {
                    const float z = exact_z(i); zbuf[i] = z; zmax = fmaxf(zmax, z);
                    if (DUMP && p.dbg.dev_cand) p.dbg.dev_cand[(size_t)story * V + i] = 2;
                    if (DUMP && p.dbg.dev_z) p.dbg.dev_z[(size_t)story * V + i] = z;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
            __syncwarp();
            // h_i = fl(e_i / total) is monotone in e_i = __expf(z_i - max): only slots whose e is within 2^-20 of the
            // largest can share the maximal probability
            n_cand = 0; cand_idx = 0;
#pragma unroll 1
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
        }
        __syncwarp();
        unsigned pred_i = cand_idx;
        float h_true_v = 0.0f;
        if (need_full && ((n_cand > 1) || p.want_h)) {
            double total = 0.0;
#pragma unroll 2
            for (unsigned i = 0; i < V; i++) total += (double)zbuf[i];
            float best = -INFINITY;
            unsigned best_i = 0;
#pragma unroll 1
            for (unsigned i0 = 0; i0 < V; i0 += 32) {
                const unsigned i = i0 + lane;
                if (i < V) {
                    const float hv = (float)((double)zbuf[i] / total);
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
            h_true_v = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(h_true_v)));
        }
        if (lane == 0) {
            if (p.pred) p.pred[story] = pred_i;
            if (p.h_true) p.h_true[story] = h_true_v;
            if (p.match && ans_idx != ANS_NONE && pred_i == ans_idx) atomicAdd(p.match, 1u);
            if (DUMP && p.dbg.dev_path) p.dbg.dev_path[story] = (unsigned char)(SWAR ? PATH_PACKED : PATH_UNPACKED);
        }
        __syncwarp();
    }
}

}  // namespace

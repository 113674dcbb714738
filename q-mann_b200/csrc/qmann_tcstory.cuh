// qmann_tcstory.cuh -- k_story_tc: the dense-input production kernel with the sentence embedding on the tensor cores.
//
// The reference forms the memory embedding as the dense product M_h = X * A_h^T of the fp32 bag-of-words rows and the
// quantised table (dense_mat_fwd, lib/layer.c:2646; _cuda_mat_mat_trans_product, lib/layer_cuda.cu:105-141).  Here the
// dense rows go from HBM into shared memory with TMA tensor loads (128-byte swizzle, no thread touches them on the way) and
// are multiplied as they lie -- fp32 read as tf32 -- with all three hops' tables by tcgen05.mma kind::tf32; counts and codes
// are small integers, so the fp32 accumulators in tensor memory hold the exact row sums (qmann_tc.cuh).  No compaction, no
// gather.  What the CUDA cores still do per story: check that every value is a count the integer forms cover, pack the row
// sums to bytes, and the forward proper (packed scorer, softmax, weighted read, linear map, update, answer prefilter).
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-15    forward warps, four per team; warp w owns quadrant w % 4: it packs its story's row sums into its private
//                 tensor-memory slot (tcgen05.st), releases the accumulator tile, and runs the story to its prediction
//   warp 16       producer: claims groups of four stories, issues the TMA boxes (32 rows x 32 columns) of every K chunk
//   warp 17       issues the MMAs: M = 128 rows (rows 32t..32t+31 of the four stories, one story per 32-lane quadrant of tensor
//                 memory), N = 160 (3 hops x 52 dims + the three row-bias columns), K = 8 per instruction
//   warps 18-23   validate the staged rows (every value 0/1, or a count whose n copies of a unit entry are exact); every
//                 stage is split among the six warps so that it is released quickly
// Stories this tier does not cover (a row sum beyond a byte, irregular values, very many selected slots) are appended to
// p.slow_list for the unpacked k_story tier / the general kernel.
#pragma once
#include "qmann_fast.cuh"
#include "qmann_tc.cuh"

namespace {

constexpr unsigned TC_NT = 160;                 // MMA N: table rows
constexpr unsigned TC_HCOLS = 52;               // accumulator columns per hop (13 packed words)
constexpr unsigned TC_DW = 13;                  // packed words per hop and row
constexpr unsigned TC_BIAS0 = 156;              // columns 156.. : sum of column maxima per hop
constexpr unsigned TC_NSTAGE = 3, TC_STAGE_BYTES = 128 * 128, TC_TABCH_BYTES = TC_NT * 128;
constexpr unsigned TC_NG = 16;                  // group descriptor ring
constexpr unsigned TC_NVAL = 6;                 // validator warps: each checks a sixth of EVERY stage, so each sees every phase of every
                                                // full barrier and is one of the arrivals that release it (a barrier can then never lap a waiter)
constexpr unsigned TC_MAX_TEAMS = 4;
constexpr unsigned TC_W_PRODUCER = 4 * TC_MAX_TEAMS, TC_W_MMA = TC_W_PRODUCER + 1, TC_W_VAL0 = TC_W_MMA + 1, TC_WARPS = TC_W_VAL0 + TC_NVAL;
static_assert(TC_WARPS <= 24, "trace layout");
constexpr unsigned TC_PK_COLS = 40;             // packed slot: 3 hops x 13 words (+1) per row set
constexpr unsigned TC_NNZ_CAP = 16, TC_ENT_CAP = 160;
// per-warp scratch (bytes)
constexpr unsigned TW_SELR = 0, TW_PQ = 16, TW_UVEC = 32, TW_OVEC = 96, TW_SQ = 160, TW_UB8 = 576, TW_TW8 = 640, TW_ENT = 704, TW_REND = 1024,
                   TW_ZENT = 1064, TW_BYTES = 1072;
// control block (bytes from its base)
// (one accumulator-ready barrier per team: a waiter may then never be more than one phase behind its barrier)
constexpr unsigned TCB_TAB = 0, TCB_FULL = 8, TCB_EMPTY = 32, TCB_DFREE = 56, TCB_TMEM = 64, TCB_DFULL = 96, TCB_GBAR = 128, TCB_GDESC = 256,
                   TCB_BAD = TCB_GDESC + TC_NG * 48, TCB_TAU = TCB_BAD + TC_NG * 4, TCB_BYTES = TCB_TAU + 128;

struct TcGroup {
    unsigned story[4];        // chunk index of the story in each quadrant (0xFFFFFFFF: empty)
    unsigned soff[4];         // first arena row
    unsigned char S[4];       // sentences
    unsigned n_tiles;         // 1 or 2 accumulator tiles (rows 0-31, rows 32-63); 0 terminates
    unsigned tile_base;       // sequence number of the group's first tile among the tiles of its team
    unsigned pad;
};
static_assert(sizeof(TcGroup) == 48, "TcGroup layout");

struct alignas(64) TcParams {
    CUtensorMap tmX;          // dense sentence arena [sum_sen][V] fp32, box 32 x 32, 128-byte swizzle
    CUtensorMap tmT;          // tables [160][V] fp32 (tf32-exact integers), box 32 x 160
    FwdParams f;
    unsigned n_groups, total_rows, kch, n_teams;
    volatile unsigned *trace;  // QMANN_TC_TRACE builds: mapped host memory, 4 progress words per warp of CTA 0
};
#ifdef QMANN_TC_TRACE
#if QMANN_TC_TRACE == 2
#define TCT(slot, val) do { } while (0)          /* clock accounting only */
#else
#define TCT(slot, val) do { if (tp.trace && (threadIdx.x & 31) == 0) { tp.trace[(blockIdx.x * 24 + (threadIdx.x >> 5)) * 4 + (slot)] = (val); } } while (0)
#endif
#define TCK_DECL unsigned long long tck_acc[4] = {0ull, 0ull, 0ull, 0ull}; long long tck_t = clock64();
#define TCK(slot) do { const long long n_ = clock64(); tck_acc[slot] += (unsigned long long)(n_ - tck_t); tck_t = n_; } while (0)
#define TCK_FLUSH do { if (tp.trace && (threadIdx.x & 31) == 0) for (int i_ = 0; i_ < 4; i_++) tp.trace[148 * 24 * 4 + (blockIdx.x * 24 + (threadIdx.x >> 5)) * 4 + i_] = (unsigned)(tck_acc[i_] >> 4); } while (0)
#else
#define TCT(slot, val) do { } while (0)
#define TCK_DECL
#define TCK(slot) do { } while (0)
#define TCK_FLUSH do { } while (0)
#endif

// One staged K chunk (128 rows x 32 columns): every value must be 0.0 or 1.0, or an integer count n in 2..nmax whose n copies
// of the unit entry are exact in every table (n * colmax <= split_lim); otherwise the story is marked for the next tier.
__device__ __forceinline__ void tc_validate_stage(const FwdParams &p, const unsigned char *stage, unsigned t, unsigned k, const TcGroup *gd,
                                                  unsigned char *bad, unsigned lane, unsigned v)
{
    // the stage is 32 groups of four rows (one 128-bit load per lane covers a group); validator v takes groups v, v + NVAL, ...
    const unsigned pc = lane & 7u;
#pragma unroll 1
    for (unsigned i0 = 2u * v; i0 < 32; i0 += 2u * TC_NVAL) {
        const unsigned r0 = 4u * i0 + (lane >> 3), r1 = r0 + 4u;
        const float4 a = *reinterpret_cast<const float4 *>(stage + r0 * 128u + pc * 16u);
        const float4 b = *reinterpret_cast<const float4 *>(stage + r1 * 128u + pc * 16u);
        float acc0, acc1;
        {
            const float t0 = __fmaf_rn(a.x, a.x, -a.x), t1 = __fmaf_rn(a.y, a.y, -a.y), t2 = __fmaf_rn(a.z, a.z, -a.z), t3 = __fmaf_rn(a.w, a.w, -a.w);
            acc0 = __fmaf_rn(t0, t0, t1 * t1);
            acc0 = __fmaf_rn(t2, t2, acc0);
            acc0 = __fmaf_rn(t3, t3, acc0);
            const float s0 = __fmaf_rn(b.x, b.x, -b.x), s1 = __fmaf_rn(b.y, b.y, -b.y), s2 = __fmaf_rn(b.z, b.z, -b.z), s3 = __fmaf_rn(b.w, b.w, -b.w);
            acc1 = __fmaf_rn(s0, s0, s1 * s1);
            acc1 = __fmaf_rn(s2, s2, acc1);
            acc1 = __fmaf_rn(s3, s3, acc1);
        }
        if (!(acc0 + acc1 == 0.0f)) {
            // some value of these eight is not 0/1 (or is not finite)
#pragma unroll 1
            for (unsigned e = 0; e < 8; e++) {
                const float4 &v4 = (e < 4) ? a : b;
                const unsigned ee = e & 3u;
                const float x = (ee == 0) ? v4.x : (ee == 1) ? v4.y : (ee == 2) ? v4.z : v4.w;
                if (x == 0.0f || x == 1.0f) continue;
                const unsigned row = (e < 4) ? r0 : r1;
                const unsigned j = row >> 5, r = 32u * t + (row & 31u);
                if (r >= gd->S[j]) continue;                           // a row of the next story (read ahead), not ours
                const unsigned col = 32u * k + 4u * (pc ^ (row & 7u)) + ee;
                bool ok = false;
                if (col < p.V) {
                    const float n = truncf(x);
                    ok = (n == x) && x >= 2.0f && x <= (float)p.nmax && (unsigned)n * (unsigned)p.colmax[col] <= p.split_lim;
                }
                if (!ok) bad[j] = 1;
            }
        }
    }
}

// One dense row in global memory (V <= 256, V % 4 == 0, 16-byte aligned): lane l loads float4 l and l + 32 ...
__device__ __forceinline__ void tc_row_load(const FwdParams &p, const float *__restrict__ rowp, unsigned lane, float (&v)[8])
{
    const float4 *b4 = reinterpret_cast<const float4 *>(rowp);
    const unsigned n4 = p.V >> 2;
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = 0.0f;
    if (lane < n4) { const float4 t = ldg_stream4(b4 + lane); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    if (lane + 32u < n4) { const float4 t = ldg_stream4(b4 + lane + 32u); v[4] = t.x; v[5] = t.y; v[6] = t.z; v[7] = t.w; }
}
// ... and its entries (column * DP) are appended to ent[] from position `base`
__device__ __noinline__ unsigned tc_row_emit(const FwdParams &p, const float (&v)[8], unsigned short *__restrict__ ent, unsigned cap, unsigned base, unsigned lane,
                                             bool &irregular)
{
    const unsigned lt = (1u << lane) - 1u;
    if (!chunk_irregular<4>(v)) return emit_units_s(unit_mask<4>(v), lane, lane + 32u, 0u, p.DP, ent, cap, base, lt);
    return emit_general_s(p, v, lane, lane + 32u, 0u, ent, cap, base, lt, irregular);
}

// (a0, a1) += table codes of this lane's dims 2*lane, 2*lane+1 over entries [beg, end) of the warp's entry list (byte offsets
// column * DP, DP = 64); table rows in global memory (L2-resident), one coalesced 64-byte row read per entry and warp.
// Two rows at once (entries [0, na) and [na, nb)) so that their L2 round trips overlap.
__device__ __forceinline__ void tc_gather2(unsigned ent_sa, unsigned na, unsigned nb, const unsigned char *__restrict__ tab, unsigned lane, int &a0, int &a1,
                                           int &b0, int &b1)
{
    const unsigned char *t16 = tab + 2u * lane;
    const unsigned lb_ = nb - na, n = max(na, lb_);
#pragma unroll 4
    for (unsigned e = 0; e < n; e++) {
        unsigned short offa = 0, offb = 0;
        if (e < na) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(offa) : "r"(ent_sa + 2u * e));
        if (e < lb_) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(offb) : "r"(ent_sa + 2u * (na + e)));
        unsigned wa = 0, wb = 0;
        if (e < na) wa = ldg_na_u16(t16 + offa);
        if (e < lb_) wb = ldg_na_u16(t16 + offb);
        a0 += (int)(signed char)(wa & 0xFFu);
        a1 += (int)(signed char)(wa >> 8);
        b0 += (int)(signed char)(wb & 0xFFu);
        b1 += (int)(signed char)(wb >> 8);
    }
}

// Packed scorer of one row held by this lane: y[w] = four dims of the row sum (bytes, weight format), query constants from
// the warp's scratch (uniform addresses).  Returns 4 * score (+ 3 per dim) and the saturation flag; see swar_score.
__device__ __forceinline__ int tc_score_row(unsigned sq_sa, const unsigned (&y)[TC_DW], unsigned &flag)
{
    int D = 0;
    unsigned cs = 0, f = 0;
#pragma unroll
    for (unsigned w = 0; w < TC_DW; w++) {
        unsigned Uw, Tw, U1, Us4;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(Uw), "=r"(Tw), "=r"(U1), "=r"(Us4) : "r"(sq_sa + 16u * w));
        const unsigned U0 = Uw & SW_1, U0s = U0 << 1;
        const unsigned yy = y[w];
        unsigned fill;
        asm("prmt.b32 %0, %1, %1, 0xBA98;" : "=r"(fill) : "r"(yy));
        f |= ((yy ^ fill) + (fill & SW_1)) + Tw;                          // bit 7 of a byte: |y| >= tau(|u|)
        D = __dp4a((int)yy, (int)Uw, D);
        const unsigned t0 = (yy << 1) & U1;
        const unsigned t1 = (yy & U0s) ^ t0;
        const unsigned bm = (yy & U0) | t1;                               // x mod 4 per byte
        const unsigned wv = bm + 0x03030303u;
        cs += wv & ~(((yy >> 5) & 0x04040404u) ^ Us4);
    }
    flag = f & SW_H;
    return D - (int)__dp4a(cs, SW_1, 0u) + 12 * (int)TC_DW;
}

__device__ __forceinline__ unsigned tc_pack4(unsigned b0, unsigned b1, unsigned b2, unsigned b3)
{
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}
__device__ __forceinline__ unsigned tc_ibits(unsigned fbits) { return __float_as_uint(__uint_as_float(fbits) + 12582912.0f); }      // low byte = the integer

// z_i = sum_j fl(W_ij u_j), sequential fp32 without contraction (layer_cuda.cu:69-82); few rows per story need it
__device__ __noinline__ float tc_exact_z(const float *__restrict__ wrow, const float *ufl, unsigned d4)
{
    const float4 *wr = reinterpret_cast<const float4 *>(wrow);
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
}
// a row with a saturating product: the reference order, product by product (rare)
__device__ __noinline__ int tc_exact_row(const unsigned *y, const signed char *ub8, int la, int fb)
{
    int sp = 0;
    for (unsigned j = 0; j < 4u * TC_DW; j++) sp += qi_mul(sbyte(y[j >> 2], j & 3), (int)ub8[j], la, fb);
    return sp;
}

template <bool DUMP>
__global__ void __launch_bounds__(TC_WARPS * 32, 1) k_story_tc(const __grid_constant__ TcParams tp)
{
    using namespace qtc;
    const FwdParams &p = tp.f;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n_work = p.n_stories;
    if (n_work == 0u) return;
    if (p.path_count && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.path_count + 0, (unsigned long long)n_work);
    const unsigned sraw = smem_u32(smem);
    const unsigned sbase = (sraw + 1023u) & ~1023u;
    unsigned char *gbase = smem + (sbase - sraw);
    const unsigned tabs = sbase, ring = tabs + tp.kch * TC_TABCH_BYTES, cb = ring + TC_NSTAGE * TC_STAGE_BYTES;
    unsigned char *cbg = gbase + (cb - sbase);
    TcGroup *gdesc = reinterpret_cast<TcGroup *>(cbg + TCB_GDESC);
    unsigned char *badf = cbg + TCB_BAD;
    const unsigned KCH = tp.kch;

    if (threadIdx.x == 0) {
        mbar_init(cb + TCB_TAB, 1);
        for (unsigned s = 0; s < TC_NSTAGE; s++) { mbar_init(cb + TCB_FULL + 8 * s, 1); mbar_init(cb + TCB_EMPTY + 8 * s, 1 + TC_NVAL); }
        for (unsigned i = 0; i < TC_MAX_TEAMS; i++) mbar_init(cb + TCB_DFULL + 8 * i, 1 + TC_NVAL);
        mbar_init(cb + TCB_DFREE, 4);
        for (unsigned i = 0; i < TC_NG; i++) mbar_init(cb + TCB_GBAR + 8 * i, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tp.tmX);
        tma_prefetch_desc(&tp.tmT);
    }
    if (threadIdx.x < 128) cbg[TCB_TAU + threadIdx.x] = p.img[p.offTAU + threadIdx.x];     // saturation thresholds tau[|u|]
    if (warp == TC_W_MMA) tmem_alloc(cb + TCB_TMEM, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *reinterpret_cast<const unsigned *>(cbg + TCB_TMEM);

    if (warp == TC_W_PRODUCER) {
        // ================= producer =================
        if (lane == 0) {
            mbar_expect_tx(cb + TCB_TAB, KCH * TC_TABCH_BYTES);
            for (unsigned k = 0; k < KCH; k++) tma_load_2d(tabs + k * TC_TABCH_BYTES, &tp.tmT, (int)(32 * k), 0, cb + TCB_TAB);
            unsigned it = 0;
            unsigned team_tiles[TC_MAX_TEAMS] = {0u, 0u, 0u, 0u};
            TCK_DECL
            for (unsigned gl = 0;; gl++) {
                const unsigned g = atomicAdd(p.counter, 1u);
                TCT(0, 0x100u + gl); TCT(1, g);
                if (g >= tp.n_groups) {
                    for (unsigned i = 0; i < tp.n_teams; i++) {
                        TcGroup *gd = &gdesc[(gl + i) % TC_NG];
                        gd->n_tiles = 0;
                        mbar_arrive(cb + TCB_GBAR + 8 * ((gl + i) % TC_NG));
                    }
                    TCK(1); TCK_FLUSH;
                    break;
                }
                TcGroup *gd = &gdesc[gl % TC_NG];
                unsigned soffs[4], Ss[4], smax = 0;
                for (unsigned j = 0; j < 4; j++) {
                    const unsigned idx = 4u * g + j;
                    unsigned so = 0, S = 0, st = 0xFFFFFFFFu;
                    if (idx < n_work) {
                        const unsigned long long a = p.sen_off[p.story0 + idx];
                        so = (unsigned)a; S = (unsigned)(p.sen_off[p.story0 + idx + 1] - a); st = idx;
                    }
                    soffs[j] = so; Ss[j] = S; smax = max(smax, S);
                    gd->story[j] = st; gd->soff[j] = so; gd->S[j] = (unsigned char)S;
                    badf[(gl % TC_NG) * 4 + j] = 0;
                }
                const unsigned n_tiles = smax > 32u ? 2u : 1u;
                gd->n_tiles = n_tiles; gd->tile_base = team_tiles[gl % tp.n_teams];
                team_tiles[gl % tp.n_teams] += n_tiles;
                mbar_arrive(cb + TCB_GBAR + 8 * (gl % TC_NG));
                for (unsigned t = 0; t < n_tiles; t++)
                    for (unsigned k = 0; k < KCH; k++, it++) {
                        const unsigned s = it % TC_NSTAGE, ph = (it / TC_NSTAGE) & 1u;
                        TCT(2, it);
                        TCK(1);
                        mbar_wait(cb + TCB_EMPTY + 8 * s, ph ^ 1u);
                        TCK(0);
                        TCT(3, it);
                        mbar_expect_tx(cb + TCB_FULL + 8 * s, TC_STAGE_BYTES);
                        for (unsigned j = 0; j < 4; j++) {
                            const unsigned row = (Ss[j] > 32u * t) ? soffs[j] + 32u * t : tp.total_rows;      // past the arena: zero fill
                            tma_load_2d(ring + s * TC_STAGE_BYTES + j * 4096u, &tp.tmX, (int)(32 * k), (int)row, cb + TCB_FULL + 8 * s);
                        }
                    }
            }
        }
    } else if (warp == TC_W_MMA) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const unsigned idesc = umma_idesc_tf32(128, TC_NT);
            TCT(0, 1u);
            mbar_wait(cb + TCB_TAB, 0);
            TCT(0, 2u);
            unsigned it = 0, tcnt = 0;
            TCK_DECL
            for (unsigned gl = 0;; gl++) {
                mbar_wait(cb + TCB_GBAR + 8 * (gl % TC_NG), (gl / TC_NG) & 1u);
                const unsigned n_tiles = gdesc[gl % TC_NG].n_tiles;
                TCK(3);
                if (n_tiles == 0) { TCK_FLUSH; break; }
                for (unsigned t = 0; t < n_tiles; t++, tcnt++) {
                    TCT(0, 0x1000u + tcnt);
                    TCK(2);
                    mbar_wait(cb + TCB_DFREE, (tcnt & 1u) ^ 1u);
                    TCK(0);
                    TCT(0, 0x2000u + tcnt);
                    tc_fence_after();
                    for (unsigned k = 0; k < KCH; k++, it++) {
                        const unsigned s = it % TC_NSTAGE, ph = (it / TC_NSTAGE) & 1u;
                        TCT(1, it);
                        TCK(2);
                        mbar_wait(cb + TCB_FULL + 8 * s, ph);
                        TCK(1);
                        TCT(2, it);
                        tc_fence_after();
#pragma unroll
                        for (unsigned j = 0; j < 4; j++)
                            umma_tf32(tmem, umma_desc_sw128(ring + s * TC_STAGE_BYTES + 32u * j), umma_desc_sw128(tabs + k * TC_TABCH_BYTES + 32u * j), idesc,
                                      (k | j) ? 1u : 0u);
                        umma_commit(cb + TCB_EMPTY + 8 * s);
                    }
                    umma_commit(cb + TCB_DFULL + 8 * (gl % tp.n_teams));
                }
            }
        }
    } else if (warp >= TC_W_VAL0) {
        // ================= validators =================
        const unsigned v = warp - TC_W_VAL0;
        unsigned it = 0;
        TCK_DECL
        for (unsigned gl = 0;; gl++) {
            mbar_wait(cb + TCB_GBAR + 8 * (gl % TC_NG), (gl / TC_NG) & 1u);
            const TcGroup *gd = &gdesc[gl % TC_NG];
            const unsigned n_tiles = gd->n_tiles;
            TCK(3);
            if (n_tiles == 0) { TCK_FLUSH; break; }
            for (unsigned t = 0; t < n_tiles; t++) {
                for (unsigned k = 0; k < KCH; k++, it++) {
                    const unsigned s = it % TC_NSTAGE, ph = (it / TC_NSTAGE) & 1u;
                    TCT(0, it);
                    TCK(1);
                    mbar_wait(cb + TCB_FULL + 8 * s, ph);
                    TCK(0);
                    TCT(1, it);
                    tc_validate_stage(p, gbase + (ring - sbase) + s * TC_STAGE_BYTES, t, k, gd, badf + (gl % TC_NG) * 4, lane, v);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(cb + TCB_EMPTY + 8 * s);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(cb + TCB_DFULL + 8 * (gl % tp.n_teams));
            }
        }
    } else if (warp < 4 * tp.n_teams) {
        // ================= forward warps =================
        const unsigned fwi = warp, team = fwi >> 2, q = warp & 3u;
        const unsigned d = p.d, DP = p.DP, V = p.V;
        const unsigned wso = (cb - sraw) + TCB_BYTES + fwi * TW_BYTES;        // offset of the warp's scratch inside smem[]
        unsigned char *ws = smem + wso;
        const unsigned ws_sa = sraw + wso;
        unsigned char *selr = ws + TW_SELR, *pq = ws + TW_PQ;
        signed char *uvec = reinterpret_cast<signed char *>(ws + TW_UVEC), *ovec = reinterpret_cast<signed char *>(ws + TW_OVEC);
        signed char *ub8 = reinterpret_cast<signed char *>(ws + TW_UB8);
        unsigned char *tw8 = ws + TW_TW8;
        unsigned short *ent = reinterpret_cast<unsigned short *>(ws + TW_ENT), *rend = reinterpret_cast<unsigned short *>(ws + TW_REND);
        float *ufl = reinterpret_cast<float *>(ws + TW_SQ);                   // answer phase: aliases the query constants
        const unsigned zaddr = ws_sa + TW_ZENT, ent_sa = ws_sa + TW_ENT, sq_sa = ws_sa + TW_SQ;
        if (lane == 0) *reinterpret_cast<unsigned short *>(ws + TW_ZENT) = (unsigned short)(V * DP);
        __syncwarp();
        const unsigned tq = tmem + ((32u * q) << 16);
        const unsigned tpk = tq + TC_NT + team * (2u * TC_PK_COLS);
        const unsigned char *tau = cbg + TCB_TAU;
        const unsigned row_floats = V;

        TCK_DECL
#pragma unroll 1
        for (unsigned gl = team;; gl += tp.n_teams) {
            TCT(0, 0x100u + gl);
            TCK(3);
            mbar_wait(cb + TCB_GBAR + 8 * (gl % TC_NG), (gl / TC_NG) & 1u);
            TCK(0);
            const TcGroup *gd = &gdesc[gl % TC_NG];
            const unsigned n_tiles = gd->n_tiles;
            TCT(0, 0x200u + gl); TCT(1, n_tiles);
            if (n_tiles == 0) { TCK_FLUSH; break; }
            const unsigned w = gd->story[q], S = gd->S[q], tile_base = gd->tile_base;
            const unsigned long long soff = gd->soff[q];
            bool decline = false;
            if (w != 0xFFFFFFFFu && lane < (V * 4u + 127u) / 128u)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p.dq + (size_t)(p.story0 + w) * V) + 128u * lane));
            // ---- row sums of my story: accumulator tile -> bytes -> my packed slot ----
            for (unsigned t = 0; t < n_tiles; t++) {
                TCT(2, 0x100u + tile_base + t);
                TCK(2);
                mbar_wait(cb + TCB_DFULL + 8 * team, (tile_base + t) & 1u);
                TCK(1);
                TCT(2, 0x200u + tile_base + t);
                tc_fence_after();
                const bool rvalid = (32u * t + lane) < S;
                bool wide = false;
                {
                    unsigned bv[4];
                    tmem_ld4(tq + TC_BIAS0, bv);
                    tmem_wait_ld();
                    for (unsigned h = 0; h < p.H; h++) {
                        const unsigned B = tc_ibits(bv[h]) & 0xFFFFu;
                        const int ka = p.fa[h] - p.fw[h];
                        wide |= ((ka > 0 ? (B << ka) : B) > 127u);
                    }
                }
                decline |= __any_sync(0xffffffffu, wide && rvalid);
                for (unsigned h = 0; h < p.H; h++) {
                    unsigned pk[13];
                    {
                        unsigned v0[16], v1[16], v2[16], v3[4];
                        tmem_ld16(tq + TC_HCOLS * h, v0);
                        tmem_ld16(tq + TC_HCOLS * h + 16u, v1);
                        tmem_ld16(tq + TC_HCOLS * h + 32u, v2);
                        tmem_ld4(tq + TC_HCOLS * h + 48u, v3);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            pk[i] = tc_pack4(tc_ibits(v0[4 * i]), tc_ibits(v0[4 * i + 1]), tc_ibits(v0[4 * i + 2]), tc_ibits(v0[4 * i + 3]));
                            pk[4 + i] = tc_pack4(tc_ibits(v1[4 * i]), tc_ibits(v1[4 * i + 1]), tc_ibits(v1[4 * i + 2]), tc_ibits(v1[4 * i + 3]));
                            pk[8 + i] = tc_pack4(tc_ibits(v2[4 * i]), tc_ibits(v2[4 * i + 1]), tc_ibits(v2[4 * i + 2]), tc_ibits(v2[4 * i + 3]));
                        }
                        pk[12] = tc_pack4(tc_ibits(v3[0]), tc_ibits(v3[1]), tc_ibits(v3[2]), tc_ibits(v3[3]));
                    }
                    const unsigned dst = tpk + TC_PK_COLS * t + TC_DW * h;
                    {
                        const unsigned a0[4] = {pk[0], pk[1], pk[2], pk[3]}, a1[4] = {pk[4], pk[5], pk[6], pk[7]}, a2[4] = {pk[8], pk[9], pk[10], pk[11]};
                        tmem_st4(dst, a0); tmem_st4(dst + 4, a1); tmem_st4(dst + 8, a2);
                        tmem_st1(dst + 12, pk[12]);
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(cb + TCB_DFREE);
                TCK(2);
                TCT(2, 0x300u + tile_base + t);
            }
            // The four warps of a team run their stories in step (a named barrier per phase): they then fetch the same
            // instruction lines at the same time, and the 32 KB instruction cache holds the four teams' positions.
            const bool empty = (w == 0xFFFFFFFFu);                            // empty quadrant of the last group
            decline |= (badf[(gl % TC_NG) * 4 + q] != 0) || (S == 0u);
            bool dead = empty;
            const unsigned story = p.story0 + w;
            const unsigned nrs = S > 32u ? 2u : 1u;
            unsigned ans_idx = ANS_NONE;
            if (p.da && !dead) {
                const float *arow = p.da + (size_t)story * V;
                for (unsigned c0 = 0; c0 < V; c0 += 32) {
                    const unsigned c = c0 + lane;
                    const bool hot = (c < V) && (ldg_stream1(arow + c) == 1.0f);
                    const unsigned bb = __ballot_sync(0xffffffffu, hot);
                    if (bb) ans_idx = c0 + 31 - __clz(bb);
                }
            }
            // ---- question embedding u0 = Q_w0(sum)                          MemN2N.c:826, layer_cuda.cu:49 ----
            // From here on lane l owns dims 2l, 2l+1 of every d-vector (u, o, g): gathers read one coalesced 64-byte table row
            // per entry, sums and updates stay in the lane's registers.
            const unsigned c0 = 2u * lane;
            int u0c = 0, u1c = 0;                                             // controller state, codes with fu fractional bits
            if (!decline && !dead) {
                bool irregular = false;
                float qv[8];
                tc_row_load(p, p.dq + (size_t)story * row_floats, lane, qv);
                const unsigned n = tc_row_emit(p, qv, ent, TC_ENT_CAP, 0u, lane, irregular);
                decline = irregular || n > TC_ENT_CAP;
                __syncwarp();
                if (!decline) {
                    int x0 = 0, x1 = 0;
                    tc_gather2(ent_sa, n, n, p.img + p.offB, lane, u0c, u1c, x0, x1);
                    u0c = qi_clamp(u0c, p.lw[0]); u1c = qi_clamp(u1c, p.lw[0]);
                }
            }
            if (decline && !dead) {
                if (lane == 0) p.slow_list[atomicAdd(p.slow_count, 1u)] = w;
                dead = true;
            }
            int fu = p.fw[0];
            if (DUMP && p.dbg.dev_u0 && !dead) {
                if (c0 < d) p.dbg.dev_u0[(size_t)story * d + c0] = (float)u0c / (float)(1 << fu);
                if (c0 + 1 < d) p.dbg.dev_u0[(size_t)story * d + c0 + 1] = (float)u1c / (float)(1 << fu);
            }

#pragma unroll 1
            for (unsigned h = 0; h < p.H; h++) {
                asm volatile("bar.sync %0, 128;" ::"r"(1u + team) : "memory");
                if (!dead) do {
                const int fw = p.fw[h], lw = p.lw[h];
                const int fa = p.fa[h], la = p.la[h];
                const int ff = p.ff[h], lf = p.lf[h];
                const int fb = p.fb, lb = p.lb;
                const int ka = fa - fw;
                TCT(3, 0x10u + h);
                // Q_bin(u) and the saturation thresholds as bytes, then the per-word query constants  MemN2N.c:847,873
                {
                    const int b0 = (c0 < d) ? qi_requant(u0c, fu, lb, fb) : 0, b1 = (c0 + 1 < d) ? qi_requant(u1c, fu, lb, fb) : 0;
                    *reinterpret_cast<unsigned short *>(ub8 + c0) = (unsigned short)((b0 & 0xFF) | ((b1 & 0xFF) << 8));
                    *reinterpret_cast<unsigned short *>(tw8 + c0) = (unsigned short)((unsigned)tau[abs(b0)] | ((unsigned)tau[abs(b1)] << 8));
                }
                __syncwarp();
                if (lane < TC_DW) {
                    const unsigned uw = reinterpret_cast<const unsigned *>(ub8)[lane], tw = reinterpret_cast<const unsigned *>(tw8)[lane];
                    *reinterpret_cast<uint4 *>(ws + TW_SQ + 16u * lane) = make_uint4(uw, tw, uw & 0x02020202u, (uw >> 5) & 0x04040404u);
                }
                __syncwarp();
                // ---- addressing: this lane's rows lane, lane + 32 ----
                int scode[2] = {0, 0};
#pragma unroll 1
                for (unsigned rs = 0; rs < nrs; rs++) {
                    unsigned a[16];
                    tmem_ld16(tpk + TC_PK_COLS * rs + TC_DW * h, a);
                    tmem_wait_ld();
                    unsigned y[TC_DW], flag;
                    // Q_att of the row sums: same grid, one more fractional bit (2a, exact: 2B <= 127), or one less (trunc0(a / 2))
                    if (ka == 0) {
#pragma unroll
                        for (unsigned w_ = 0; w_ < TC_DW; w_++) y[w_] = a[w_];
                    } else if (ka > 0) {
#pragma unroll
                        for (unsigned w_ = 0; w_ < TC_DW; w_++) y[w_] = (a[w_] << 1) & 0xFEFEFEFEu;
                    } else {
#pragma unroll
                        for (unsigned w_ = 0; w_ < TC_DW; w_++) y[w_] = swar_half0(a[w_]);
                    }
                    const int part = tc_score_row(sq_sa, y, flag);
                    int tot = part >> 2;                                       // exact: a multiple of 4
                    if (flag != 0u && (32u * rs + lane) < S) tot = tc_exact_row(y, ub8, la, fb);
                    if (rs == 0) scode[0] = qi_clamp(tot, la); else scode[1] = qi_clamp(tot, la);
                }
                // ---- attention normalisation (layer_cuda.cu:1895-1916, 1969-2060) ----
                float sv[2], ev[2];
                float mx = -INFINITY;
#pragma unroll
                for (int rs = 0; rs < 2; rs++) {
                    const bool val = (32u * rs + lane) < S;
                    sv[rs] = (float)scode[rs] / (float)(1 << fa);
                    if (val) mx = fmaxf(mx, sv[rs]);
                    if (DUMP && p.dbg.dev_s && val) p.dbg.dev_s[(size_t)h * p.sum_sen + soff + 32u * rs + lane] = sv[rs];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float tsum = 0.0f;
#pragma unroll
                for (int rs = 0; rs < 2; rs++) {
                    const bool val = (32u * rs + lane) < S;
                    ev[rs] = val ? __expf(sv[rs] - mx) : 0.0f;
                    tsum += ev[rs];
                }
                // see k_story: codes from a float total unless some weight sits within the float error of a truncation boundary
                bool exact_total = !p.fast_softmax;
                unsigned code[2] = {0u, 0u};
                if (!exact_total) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
                    const float sc2 = (float)(1 << ff) / tsum;
                    const float tol = (float)(S / 32u + 8u) * 1.1920929e-7f;
                    bool amb = false;
#pragma unroll
                    for (int rs = 0; rs < 2; rs++) {
                        const float vv = ev[rs] * sc2;
                        const float n = rintf(vv);
                        amb |= (n >= 1.0f) && (fabsf(vv - n) <= tol * vv);
                        code[rs] = (unsigned)min((int)vv, lf);
                    }
                    exact_total = __any_sync(0xffffffffu, amb);
                }
                if (exact_total) {
                    double total = 0.0;
#pragma unroll 1
                    for (unsigned r = 0; r < S; r++) total += (double)__shfl_sync(0xffffffffu, (r < 32u) ? ev[0] : ev[1], (int)(r & 31u));
#pragma unroll
                    for (int rs = 0; rs < 2; rs++)
                        code[rs] = ((32u * rs + lane) < S) ? (unsigned)qi_encode((float)((double)ev[rs] / total), p.iff[h], ff) : 0u;      // layer_cuda.cu:561
                }
                unsigned nnz = 0;
#pragma unroll
                for (int rs = 0; rs < 2; rs++) {
                    const unsigned r = 32u * rs + lane;
                    if (r >= S) code[rs] = 0u;
                    if (DUMP && p.dbg.dev_pcode && r < S) p.dbg.dev_pcode[(size_t)h * p.sum_sen + soff + r] = (unsigned char)code[rs];
                    const unsigned b = __ballot_sync(0xffffffffu, code[rs] != 0u);
                    if (code[rs]) {
                        const unsigned k = nnz + __popc(b & ((1u << lane) - 1u));
                        if (k < TC_NNZ_CAP) { selr[k] = (unsigned char)r; pq[k] = (unsigned char)code[rs]; }
                    }
                    nnz += __popc(b);
                }
                if (nnz > TC_NNZ_CAP) { decline = true; break; }
                __syncwarp();
                // ---- weighted read over the selected slots (layer_cuda.cu:547-579): their C_h rows from the dense rows ----
                int o0 = 0, o1 = 0;
                {
                    bool irregular = false;
                    const unsigned char *ctab = p.img + p.offC[h];
#pragma unroll 1
                    for (unsigned k = 0; k < nnz; k += 2) {
                        // two selected slots per step: their dense rows are loaded together, then their C_h rows gathered together
                        const bool two = (k + 1 < nnz);
                        float va[8], vb[8];
                        tc_row_load(p, p.dm + (soff + selr[k]) * (size_t)row_floats, lane, va);
                        tc_row_load(p, p.dm + (soff + selr[two ? k + 1 : k]) * (size_t)row_floats, lane, vb);
                        const unsigned na = tc_row_emit(p, va, ent, TC_ENT_CAP, 0u, lane, irregular);
                        const unsigned nb = two ? tc_row_emit(p, vb, ent, TC_ENT_CAP, na, lane, irregular) : na;
                        if (irregular || nb > TC_ENT_CAP) { irregular = true; break; }
                        __syncwarp();
                        int a0 = 0, a1 = 0, b0 = 0, b1 = 0;
                        tc_gather2(ent_sa, na, nb, ctab, lane, a0, a1, b0, b1);
                        const int pa = (int)pq[k], pb = two ? (int)pq[k + 1] : 0;
                        o0 += qi_mul(pa, qi_requant(qi_clamp(a0, lw), fw, lf, ff), lf, ff) + qi_mul(pb, qi_requant(qi_clamp(b0, lw), fw, lf, ff), lf, ff);
                        o1 += qi_mul(pa, qi_requant(qi_clamp(a1, lw), fw, lf, ff), lf, ff) + qi_mul(pb, qi_requant(qi_clamp(b1, lw), fw, lf, ff), lf, ff);
                        __syncwarp();
                    }
                    if (irregular) { decline = true; break; }
                }
                o0 = qi_clamp(o0, lf); o1 = qi_clamp(o1, lf);
                if (DUMP && p.dbg.dev_o) {
                    if (c0 < d) p.dbg.dev_o[((size_t)h * p.n_total + story) * d + c0] = (float)o0 / (float)(1 << ff);
                    if (c0 + 1 < d) p.dbg.dev_o[((size_t)h * p.n_total + story) * d + c0 + 1] = (float)o1 / (float)(1 << ff);
                }
                // ---- linear map (MemN2N.c:873, layer_cuda.cu:49-68) and update (MemN2N.c:889, layer_cuda.cu:1535) ----
                int g0, g1, gfrac;
                if (p.lin_map) {
                    // g[i] = Q_w(sum_j T[j][Q_bin(u[j])][i]): every product Q_w(Q_w(Hm[i][j]) * Q_bin(u[j])) is one byte of a table row
                    g0 = 0; g1 = 0;
                    const unsigned short *lut16 = reinterpret_cast<const unsigned short *>(p.lut + p.offL[h]) + lane;
#pragma unroll 25
                    for (unsigned j = 0; j < d; j++) {
                        const unsigned row = j * 255u + (unsigned)((int)ub8[j] + 127);
                        const unsigned w_ = ldg_na_u16(lut16 + (size_t)row * (DP / 2));
                        g0 += (int)(signed char)(w_ & 0xFFu);
                        g1 += (int)(signed char)(w_ >> 8);
                    }
                    g0 = qi_clamp(g0, lw); g1 = qi_clamp(g1, lw);
                    gfrac = fw;
                } else {
                    g0 = u0c; g1 = u1c; gfrac = fu;
                }
                if (DUMP && p.dbg.dev_g) {
                    if (c0 < d) p.dbg.dev_g[((size_t)h * p.n_total + story) * d + c0] = (float)g0 / (float)(1 << gfrac);
                    if (c0 + 1 < d) p.dbg.dev_g[((size_t)h * p.n_total + story) * d + c0 + 1] = (float)g1 / (float)(1 << gfrac);
                }
                u0c = (c0 < d) ? qi_clamp(qi_requant(g0, gfrac, lf, ff) + o0, lf) : 0;
                u1c = (c0 + 1 < d) ? qi_clamp(qi_requant(g1, gfrac, lf, ff) + o1, lf) : 0;
                fu = ff;
                if (DUMP && p.dbg.dev_u) {
                    if (c0 < d) p.dbg.dev_u[((size_t)h * p.n_total + story) * d + c0] = (float)u0c / (float)(1 << fu);
                    if (c0 + 1 < d) p.dbg.dev_u[((size_t)h * p.n_total + story) * d + c0 + 1] = (float)u1c / (float)(1 << fu);
                }
                __syncwarp();
                } while (0);
                if (decline && !dead) {
                    if (lane == 0) p.slow_list[atomicAdd(p.slow_count, 1u)] = w;
                    dead = true;
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1u + team) : "memory");
            if (dead) continue;

            TCT(3, 0x50u);
            // ---- answer projection with the int8 prefilter (see k_story): W8 and the fp32 rows from L2 ----
            // z_i = sum_j fl(W_ij u_j) in index order is needed exactly only for the rows that can hold the largest probability:
            // the integer dot D_i = sum_j W8_ij n_j bounds z_i, every row within 1e-5 of the best logit has D_i >= D_max - T.
            *reinterpret_cast<unsigned short *>(uvec + c0) = (unsigned short)((u0c & 0xFF) | ((u1c & 0xFF) << 8));
            ufl[c0] = (float)u0c / (float)(1 << fu);
            ufl[c0 + 1] = (float)u1c / (float)(1 << fu);
            __syncwarp();
            const unsigned d4 = (d + 3) / 4;
            const float *Wg = reinterpret_cast<const float *>(p.img + p.offW);
            const unsigned nw16 = (d + 15) / 16;
            const unsigned char *W8g = p.img + p.offW8;
            auto w8_dot = [&](unsigned i) {
                const unsigned char *wrow = W8g + (size_t)min(i, V - 1) * p.W8S;
                int D = 0;
#pragma unroll 4
                for (unsigned w16 = 0; w16 < nw16; w16++) {
                    const uint4 ww = __ldg(reinterpret_cast<const uint4 *>(wrow + 16u * w16));
                    const uint4 uu = *reinterpret_cast<const uint4 *>(uvec + 16 * w16);
                    D = __dp4a((int)ww.x, (int)uu.x, D);
                    D = __dp4a((int)ww.y, (int)uu.y, D);
                    D = __dp4a((int)ww.z, (int)uu.z, D);
                    D = __dp4a((int)ww.w, (int)uu.w, D);
                }
                return D;
            };
            unsigned n_cand = 0, cand_idx = 0;
            bool need_full = !p.w8_ok || p.want_h;
            if (!need_full) {
                int n1 = abs(u0c) + abs(u1c);
                n1 = __reduce_add_sync(0xffffffffu, n1);
                const int T = n1 + (n1 >> 6) + p.ans_margin + 2;
                int Dmax = INT_MIN;
#pragma unroll 2
                for (unsigned i = lane; i < V; i += 32) Dmax = max(Dmax, w8_dot(i));
                Dmax = __reduce_max_sync(0xffffffffu, Dmax);
                const int thr = Dmax - T;
                // second pass: the candidates' exact logits; per lane the two largest are enough to see a near-tie
                float z1 = -INFINITY, z2 = -INFINITY;
                unsigned i1 = 0;
#pragma unroll 1
                for (unsigned i = lane; i < V; i += 32) {
                    const bool cnd = w8_dot(i) >= thr;
                    float z = -INFINITY;
                    if (cnd) z = tc_exact_z(Wg + (size_t)i * p.WS, ufl, d4);
                    if (DUMP && p.dbg.dev_cand) p.dbg.dev_cand[(size_t)story * V + i] = cnd ? 1 : 0;
                    if (DUMP && p.dbg.dev_z) p.dbg.dev_z[(size_t)story * V + i] = z;
                    if (cnd) {
                        if (!(z1 > z)) { z2 = z1; z1 = z; i1 = i; }            // ties keep the larger index in (z1, i1)
                        else if (z > z2) z2 = z;
                    }
                }
                float zmax = z1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
                const bool c1 = __expf(z1 - zmax) >= 0.99999905f, c2 = __expf(z2 - zmax) >= 0.99999905f;
                const unsigned b1 = __ballot_sync(0xffffffffu, c1), b2 = __ballot_sync(0xffffffffu, c2);
                n_cand = __popc(b1) + __popc(b2);
                cand_idx = __shfl_sync(0xffffffffu, i1, b1 ? (31 - __clz(b1)) : 0);
                need_full = n_cand != 1;                         // near-tie: the double total decides, every row exactly
            }
            unsigned pred_i = cand_idx;
            float h_true_v = 0.0f;
            if (need_full) {
                // every row exactly; h_i = fl(e_i / total) with the double total in index order (layer_cuda.cu:1969-2060),
                // argmax with the last index among equals (layer_cuda.cu:1918-1939)
                float er[8];
                float zmax = -INFINITY;
#pragma unroll 1
                for (unsigned k = 0; k < 8; k++) {
                    const unsigned i = 32u * k + lane;
                    float z = -INFINITY;
                    if (i < V) {
                        z = tc_exact_z(Wg + (size_t)i * p.WS, ufl, d4);
                        if (DUMP && p.dbg.dev_cand) p.dbg.dev_cand[(size_t)story * V + i] = 2;
                        if (DUMP && p.dbg.dev_z) p.dbg.dev_z[(size_t)story * V + i] = z;
                    }
                    er[k] = z;
                    zmax = fmaxf(zmax, z);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
                double total = 0.0;
#pragma unroll 1
                for (unsigned k = 0; k < 8; k++) {
                    const unsigned i = 32u * k + lane;
                    er[k] = (i < V) ? __expf(er[k] - zmax) : 0.0f;
                    const float mine = er[k];
#pragma unroll 1
                    for (unsigned l = 0; l < 32 && 32u * k + l < V; l++) total += (double)__shfl_sync(0xffffffffu, mine, (int)l);
                }
                float best = -INFINITY;
                unsigned best_i = 0;
#pragma unroll 1
                for (unsigned k = 0; k < 8; k++) {
                    const unsigned i = 32u * k + lane;
                    if (i < V) {
                        const float hv = (float)((double)er[k] / total);
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
                if (DUMP && p.dbg.dev_path) p.dbg.dev_path[story] = (unsigned char)PATH_PACKED;
            }
            __syncwarp();
        }
    }
    TCT(0, 0xE0Du);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == TC_W_MMA) qtc::tmem_dealloc(tmem, 512);
}

// tables of the MMA: tab[n][v] fp32, n = 52 h + c -> code of A_h[v][c]; n = 156 + h -> max_c |code| of column v (row bias)
__global__ void k_prep_tc(const unsigned char *__restrict__ img, float *__restrict__ tab, unsigned V, unsigned d, unsigned DP, unsigned H, unsigned offA0,
                          unsigned offA1, unsigned offA2)
{
    const unsigned offs[3] = {offA0, offA1, offA2};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < TC_NT * V; i += gridDim.x * blockDim.x) {
        const unsigned n = i / V, v = i % V;
        float x = 0.0f;
        if (n < TC_BIAS0) {
            const unsigned h = n / TC_HCOLS, c = n % TC_HCOLS;
            if (h < H && c < d) x = (float)reinterpret_cast<const signed char *>(img + offs[h])[(size_t)v * DP + c];
        } else if (n - TC_BIAS0 < H) {
            const signed char *row = reinterpret_cast<const signed char *>(img + offs[n - TC_BIAS0]) + (size_t)v * DP;
            int mx = 0;
            for (unsigned c = 0; c < d; c++) mx = max(mx, abs((int)row[c]));
            x = (float)mx;
        }
        tab[i] = x;
    }
}

}  // namespace

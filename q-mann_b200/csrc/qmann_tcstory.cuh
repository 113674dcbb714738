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
//   warps 18-20   validate the staged rows (every value 0/1, or a count whose n copies of a unit entry are exact), one warp
//                 per ring stage
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
constexpr unsigned TC_NVAL = 3;                 // validator warps: one per ring stage (a warp waits only on the barrier it releases)
constexpr unsigned TC_MAX_TEAMS = 4;
constexpr unsigned TC_W_PRODUCER = 4 * TC_MAX_TEAMS, TC_W_MMA = TC_W_PRODUCER + 1, TC_W_VAL0 = TC_W_MMA + 1, TC_WARPS = TC_W_VAL0 + TC_NVAL;
static_assert(TC_NVAL == TC_NSTAGE, "one validator per stage");
constexpr unsigned TC_PK_COLS = 40;             // packed slot: 3 hops x 13 words (+1) per row set
constexpr unsigned TC_NNZ_CAP = 16, TC_ENT_CAP = 160;
// per-warp scratch (bytes)
constexpr unsigned TW_SELR = 0, TW_PQ = 16, TW_UVEC = 32, TW_OVEC = 96, TW_SQ = 160, TW_UB8 = 576, TW_TW8 = 640, TW_ENT = 704, TW_REND = 1024,
                   TW_ZENT = 1064, TW_BYTES = 1072;
// control block (bytes from its base)
// (one accumulator-ready barrier per team: a waiter may then never be more than one phase behind its barrier)
constexpr unsigned TCB_TAB = 0, TCB_FULL = 8, TCB_EMPTY = 32, TCB_DFREE = 56, TCB_TMEM = 64, TCB_DFULL = 96, TCB_GBAR = 128, TCB_GDESC = 256,
                   TCB_BAD = TCB_GDESC + TC_NG * 48, TCB_BYTES = TCB_BAD + TC_NG * 4;

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
#define TCT(slot, val) do { if (tp.trace && (threadIdx.x & 31) == 0) { tp.trace[(blockIdx.x * 24 + (threadIdx.x >> 5)) * 4 + (slot)] = (val); } } while (0)
#else
#define TCT(slot, val) do { } while (0)
#endif

// One staged K chunk (128 rows x 32 columns): every value must be 0.0 or 1.0, or an integer count n in 2..nmax whose n copies
// of the unit entry are exact in every table (n * colmax <= split_lim); otherwise the story is marked for the next tier.
__device__ __forceinline__ void tc_validate_stage(const FwdParams &p, const unsigned char *stage, unsigned t, unsigned k, const TcGroup *gd,
                                                  unsigned char *bad, unsigned lane)
{
    const unsigned pc = lane & 7u;
#pragma unroll 2
    for (unsigned i0 = 0; i0 < 32; i0 += 2) {
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

// Scan one dense row in global memory (V % 4 == 0, 16-byte aligned) and append its entries (column * DP) to ent[].
__device__ __forceinline__ unsigned tc_scan_row_g(const FwdParams &p, const float *__restrict__ rowp, unsigned short *__restrict__ ent, unsigned cap,
                                                  unsigned base, unsigned lane, bool &irregular)
{
    const unsigned lt = (1u << lane) - 1u;
    const float4 *b4 = reinterpret_cast<const float4 *>(rowp);
    const unsigned n4 = p.V >> 2;
    for (unsigned c0 = 0; c0 < n4; c0 += 64) {
        const unsigned ca = c0 + lane, cb = ca + 32u;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = 0.0f;
        if (ca < n4) { const float4 t = ldg_stream4(b4 + ca); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        if (cb < n4) { const float4 t = ldg_stream4(b4 + cb); v[4] = t.x; v[5] = t.y; v[6] = t.z; v[7] = t.w; }
        if (!chunk_irregular<4>(v)) base = emit_units_s(unit_mask<4>(v), ca, cb, 0u, p.DP, ent, cap, base, lt);
        else base = emit_general_s(p, v, ca, cb, 0u, ent, cap, base, lt, irregular);
    }
    return base;
}

// acc[j] += table codes of dims 16q.. over the entries of mini-record row `row` (entries in the warp's scratch at shared
// address ent_sa, row ends in rend[]); table in global memory (L2).  Lanes without a row gather the all-zero row.
__device__ __forceinline__ void tc_embed_glob(unsigned ent_sa, const unsigned short *rend, unsigned zaddr, const unsigned char *__restrict__ tabq, int row,
                                              int acc[16], const int sel[4])
{
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = 0;
    unsigned beg = 0, len = 0;
    if (row >= 0) {
        beg = row ? rend[row - 1] : 0u;
        len = rend[row] - beg;
    }
    const unsigned maxlen = __reduce_max_sync(0xffffffffu, len);
    unsigned ea = ent_sa + 2u * beg;
#pragma unroll 1
    for (unsigned k0 = 0; k0 < maxlen; k0 += 4) {
        uint4 t[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            unsigned short off;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(off) : "r"((k0 + i < len) ? ea + 2u * i : zaddr));
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

// Packed scorer of one row held by this lane: y[w] = four dims of the row sum (bytes, weight format), query constants from
// the warp's scratch (uniform addresses).  Returns 4 * score (+ 3 per dim) and the saturation flag; see swar_score.
template <int KA>
__device__ __forceinline__ int tc_score_row(const unsigned (&a)[16], unsigned sq_sa, unsigned (&y)[TC_DW], unsigned &flag)
{
    int D = 0;
    unsigned cs = 0, f = 0;
#pragma unroll
    for (unsigned w = 0; w < TC_DW; w++) {
        unsigned Uw, Tw, U1, Us4;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(Uw), "=r"(Tw), "=r"(U1), "=r"(Us4) : "r"(sq_sa + 16u * w));
        const unsigned U0 = Uw & SW_1, U0s = U0 << 1;
        unsigned yy;
        if (KA == 0) yy = a[w];
        else if (KA > 0) yy = (a[w] << 1) & 0xFEFEFEFEu;
        else yy = swar_half0(a[w]);
        y[w] = yy;
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
        for (unsigned s = 0; s < TC_NSTAGE; s++) { mbar_init(cb + TCB_FULL + 8 * s, 1); mbar_init(cb + TCB_EMPTY + 8 * s, 2); }
        for (unsigned i = 0; i < TC_MAX_TEAMS; i++) mbar_init(cb + TCB_DFULL + 8 * i, 1 + TC_NVAL);
        mbar_init(cb + TCB_DFREE, 4);
        for (unsigned i = 0; i < TC_NG; i++) mbar_init(cb + TCB_GBAR + 8 * i, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tp.tmX);
        tma_prefetch_desc(&tp.tmT);
    }
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
            for (unsigned gl = 0;; gl++) {
                const unsigned g = atomicAdd(p.counter, 1u);
                TCT(0, 0x100u + gl); TCT(1, g);
                if (g >= tp.n_groups) {
                    for (unsigned i = 0; i < tp.n_teams; i++) {
                        TcGroup *gd = &gdesc[(gl + i) % TC_NG];
                        gd->n_tiles = 0;
                        mbar_arrive(cb + TCB_GBAR + 8 * ((gl + i) % TC_NG));
                    }
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
                        mbar_wait(cb + TCB_EMPTY + 8 * s, ph ^ 1u);
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
            for (unsigned gl = 0;; gl++) {
                mbar_wait(cb + TCB_GBAR + 8 * (gl % TC_NG), (gl / TC_NG) & 1u);
                const unsigned n_tiles = gdesc[gl % TC_NG].n_tiles;
                if (n_tiles == 0) break;
                for (unsigned t = 0; t < n_tiles; t++, tcnt++) {
                    TCT(0, 0x1000u + tcnt);
                    mbar_wait(cb + TCB_DFREE, (tcnt & 1u) ^ 1u);
                    TCT(0, 0x2000u + tcnt);
                    tc_fence_after();
                    for (unsigned k = 0; k < KCH; k++, it++) {
                        const unsigned s = it % TC_NSTAGE, ph = (it / TC_NSTAGE) & 1u;
                        TCT(1, it);
                        mbar_wait(cb + TCB_FULL + 8 * s, ph);
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
        for (unsigned gl = 0;; gl++) {
            mbar_wait(cb + TCB_GBAR + 8 * (gl % TC_NG), (gl / TC_NG) & 1u);
            const TcGroup *gd = &gdesc[gl % TC_NG];
            const unsigned n_tiles = gd->n_tiles;
            if (n_tiles == 0) break;
            for (unsigned t = 0; t < n_tiles; t++) {
                for (unsigned k = 0; k < KCH; k++, it++) {
                    // this warp owns ring stage v: it sees every phase of full[v] and is one of the two arrivals that release it,
                    // so the barrier can never run two phases ahead of its wait
                    const unsigned s = it % TC_NSTAGE, ph = (it / TC_NSTAGE) & 1u;
                    if (s != v) continue;
                    TCT(0, it);
                    mbar_wait(cb + TCB_FULL + 8 * s, ph);
                    TCT(1, it);
                    tc_validate_stage(p, gbase + (ring - sbase) + s * TC_STAGE_BYTES, t, k, gd, badf + (gl % TC_NG) * 4, lane);
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
        constexpr int LPR = 4, G = 8;
        const unsigned g8 = lane / LPR, ql = lane % LPR;
        int sel[4];
        asm volatile("mov.u32 %0, 0x00000001;" : "=r"(sel[0]));
        asm volatile("mov.u32 %0, 0x00000100;" : "=r"(sel[1]));
        asm volatile("mov.u32 %0, 0x00010000;" : "=r"(sel[2]));
        asm volatile("mov.u32 %0, 0x01000000;" : "=r"(sel[3]));
        const unsigned char *tau = p.img + p.offTAU;
        const unsigned row_floats = V;

#pragma unroll 1
        for (unsigned gl = team;; gl += tp.n_teams) {
            TCT(0, 0x100u + gl);
            mbar_wait(cb + TCB_GBAR + 8 * (gl % TC_NG), (gl / TC_NG) & 1u);
            const TcGroup *gd = &gdesc[gl % TC_NG];
            const unsigned n_tiles = gd->n_tiles;
            TCT(0, 0x200u + gl); TCT(1, n_tiles);
            if (n_tiles == 0) break;
            const unsigned w = gd->story[q], S = gd->S[q], tile_base = gd->tile_base;
            const unsigned long long soff = gd->soff[q];
            bool decline = false;
            // ---- row sums of my story: accumulator tile -> bytes -> my packed slot ----
            for (unsigned t = 0; t < n_tiles; t++) {
                TCT(2, 0x100u + tile_base + t);
                mbar_wait(cb + TCB_DFULL + 8 * team, (tile_base + t) & 1u);
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
                    unsigned pk[16];
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        unsigned v[16];
                        tmem_ld16(tq + TC_HCOLS * h + 16u * c, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 4; i++) pk[4 * c + i] = tc_pack4(tc_ibits(v[4 * i]), tc_ibits(v[4 * i + 1]), tc_ibits(v[4 * i + 2]), tc_ibits(v[4 * i + 3]));
                    }
                    {
                        unsigned v[4];
                        tmem_ld4(tq + TC_HCOLS * h + 48u, v);
                        tmem_wait_ld();
                        pk[12] = tc_pack4(tc_ibits(v[0]), tc_ibits(v[1]), tc_ibits(v[2]), tc_ibits(v[3]));
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
                TCT(2, 0x300u + tile_base + t);
            }
            if (w == 0xFFFFFFFFu) continue;                                   // empty quadrant of the last group
            decline |= (badf[(gl % TC_NG) * 4 + q] != 0) || (S == 0u);
            const unsigned story = p.story0 + w;
            const unsigned nrs = S > 32u ? 2u : 1u;
            unsigned ans_idx = ANS_NONE;
            if (p.da) {
                const float *arow = p.da + (size_t)story * V;
                for (unsigned c0 = 0; c0 < V; c0 += 32) {
                    const unsigned c = c0 + lane;
                    const bool hot = (c < V) && (ldg_stream1(arow + c) == 1.0f);
                    const unsigned bb = __ballot_sync(0xffffffffu, hot);
                    if (bb) ans_idx = c0 + 31 - __clz(bb);
                }
            }
            int acc[16];
            // ---- question embedding u0 = Q_w0(sum)                          MemN2N.c:826, layer_cuda.cu:49 ----
            if (!decline) {
                bool irregular = false;
                const unsigned n = tc_scan_row_g(p, p.dq + (size_t)story * row_floats, ent, TC_ENT_CAP, 0u, lane, irregular);
                if (lane == 0) rend[0] = (unsigned short)min(n, 0xFFFFu);
                decline = irregular || n > TC_ENT_CAP;
                __syncwarp();
            }
            if (decline) {
                if (lane == 0) p.slow_list[atomicAdd(p.slow_count, 1u)] = w;
                continue;
            }
            TCT(3, 1u);
            tc_embed_glob(ent_sa, rend, zaddr, p.img + p.offB + 16u * ql, (g8 == 0) ? 0 : -1, acc, sel);
            if (g8 == 0) {
                unsigned packed[4];
#pragma unroll
                for (int w4 = 0; w4 < 4; w4++) {
                    unsigned v = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) v |= ((unsigned)(qi_clamp(acc[4 * w4 + b], p.lw[0]) & 0xFF)) << (8 * b);
                    packed[w4] = v;
                }
                *reinterpret_cast<uint4 *>(uvec + 16 * ql) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
            __syncwarp();
            int fu = p.fw[0];
            if (DUMP && p.dbg.dev_u0)
                for (unsigned j = lane; j < d; j += 32) p.dbg.dev_u0[(size_t)story * d + j] = (float)uvec[j] / (float)(1 << fu);

#pragma unroll 1
            for (unsigned h = 0; h < p.H && !decline; h++) {
                const int fw = p.fw[h], lw = p.lw[h];
                const int fa = p.fa[h], la = p.la[h];
                const int ff = p.ff[h], lf = p.lf[h];
                const int fb = p.fb, lb = p.lb;
                const int ka = fa - fw;
                TCT(3, 0x10u + h);
                // Q_bin(u) and the saturation thresholds as bytes, then the per-word query constants  MemN2N.c:847,873
                for (unsigned j = lane; j < 64; j += 32) {
                    const int u = (j < d) ? qi_requant((int)uvec[j], fu, lb, fb) : 0;
                    ub8[j] = (signed char)u;
                    tw8[j] = __ldg(tau + abs(u));
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
                    int part;
                    if (ka == 0) part = tc_score_row<0>(a, sq_sa, y, flag);
                    else if (ka > 0) part = tc_score_row<1>(a, sq_sa, y, flag);
                    else part = tc_score_row<-1>(a, sq_sa, y, flag);
                    int tot = part >> 2;                                       // exact: a multiple of 4
                    if (flag != 0u && (32u * rs + lane) < S) {
                        // some product of this row saturates: the reference order, product by product
                        int sp = 0;
                        for (unsigned j = 0; j < 4u * TC_DW; j++) {
                            const unsigned yw = (j < 4) ? y[0] : (j < 8) ? y[1] : (j < 12) ? y[2] : (j < 16) ? y[3] : (j < 20) ? y[4] : (j < 24) ? y[5] : (j < 28) ? y[6]
                                                : (j < 32) ? y[7] : (j < 36) ? y[8] : (j < 40) ? y[9] : (j < 44) ? y[10] : (j < 48) ? y[11] : y[12];
                            sp += qi_mul(sbyte(yw, j & 3), (int)ub8[j], la, fb);
                        }
                        tot = sp;
                    }
                    scode[rs] = qi_clamp(tot, la);
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
                {
                    unsigned base = 0;
                    bool irregular = false;
                    for (unsigned k = 0; k < nnz; k++) {
                        base = tc_scan_row_g(p, p.dm + (soff + selr[k]) * (size_t)row_floats, ent, TC_ENT_CAP, base, lane, irregular);
                        if (lane == 0) rend[k] = (unsigned short)min(base, 0xFFFFu);
                    }
                    if (irregular || base > TC_ENT_CAP) { decline = true; break; }
                    __syncwarp();
                }
                int oacc[16];
#pragma unroll
                for (int j = 0; j < 16; j++) oacc[j] = 0;
                const unsigned char *ctab = p.img + p.offC[h] + 16u * ql;
#pragma unroll 1
                for (unsigned k0 = 0; k0 < nnz; k0 += G) {
                    const unsigned k = k0 + g8;
                    const int pc = (k < nnz) ? (int)pq[k] : 0;
                    tc_embed_glob(ent_sa, rend, zaddr, ctab, (k < nnz) ? (int)k : -1, acc, sel);
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
                if (g8 == 0) {
                    unsigned packed[4];
#pragma unroll
                    for (int w4 = 0; w4 < 4; w4++) {
                        unsigned v = 0;
#pragma unroll
                        for (int b = 0; b < 4; b++) v |= ((unsigned)(qi_clamp(oacc[4 * w4 + b], lf) & 0xFF)) << (8 * b);
                        packed[w4] = v;
                    }
                    *reinterpret_cast<uint4 *>(ovec + 16 * ql) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                }
                __syncwarp();
                if (DUMP && p.dbg.dev_o)
                    for (unsigned j = lane; j < d; j += 32) p.dbg.dev_o[((size_t)h * p.n_total + story) * d + j] = (float)ovec[j] / (float)(1 << ff);

                // ---- linear map (MemN2N.c:873, layer_cuda.cu:49-68) and update (MemN2N.c:889, layer_cuda.cu:1535) ----
                if (p.lin_map) {
                    int gacc[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) gacc[k] = 0;
                    const signed char *lut = p.lut + p.offL[h] + 16u * ql;
#pragma unroll 1
                    for (unsigned jb = 0; jb < d; jb += 8 * G) {
                        uint4 t[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const unsigned j = jb + (unsigned)i * G + g8;
                            t[i] = make_uint4(0u, 0u, 0u, 0u);
                            if (j < d) t[i] = __ldg(reinterpret_cast<const uint4 *>(lut + (size_t)(j * 255u + (unsigned)((int)ub8[j] + 127)) * DP));
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
                    if (g8 == 0) {
                        const uint4 o4 = *reinterpret_cast<const uint4 *>(ovec + 16 * ql);
                        const unsigned ow[4] = {o4.x, o4.y, o4.z, o4.w};
                        unsigned packed[4];
#pragma unroll
                        for (int w4 = 0; w4 < 4; w4++) {
                            unsigned v = 0;
#pragma unroll
                            for (int b = 0; b < 4; b++) {
                                const int g_w = qi_clamp(gacc[4 * w4 + b], lw);
                                if (DUMP && p.dbg.dev_g && 16u * ql + 4 * w4 + b < d)
                                    p.dbg.dev_g[((size_t)h * p.n_total + story) * d + 16u * ql + 4 * w4 + b] = (float)g_w / (float)(1 << fw);
                                const int a_f = qi_requant(g_w, fw, lf, ff);
                                v |= ((unsigned)(qi_clamp(a_f + sbyte(ow[w4], b), lf) & 0xFF)) << (8 * b);
                            }
                            packed[w4] = v;
                        }
                        *reinterpret_cast<uint4 *>(uvec + 16 * ql) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    }
                } else {
                    for (unsigned i = lane; i < 64; i += 32) {
                        const int g_w = (i < d) ? (int)uvec[i] : 0;
                        const int a_f = qi_requant(g_w, fu, lf, ff);
                        if (DUMP && p.dbg.dev_g && i < d) p.dbg.dev_g[((size_t)h * p.n_total + story) * d + i] = (float)g_w / (float)(1 << fu);
                        __syncwarp();
                        uvec[i] = (i < d) ? (signed char)qi_clamp(a_f + (int)ovec[i], lf) : (signed char)0;
                    }
                }
                fu = ff;
                __syncwarp();
                if (DUMP && p.dbg.dev_u)
                    for (unsigned j = lane; j < d; j += 32) p.dbg.dev_u[((size_t)h * p.n_total + story) * d + j] = (float)uvec[j] / (float)(1 << fu);
            }
            if (decline) {
                if (lane == 0) p.slow_list[atomicAdd(p.slow_count, 1u)] = w;
                continue;
            }

            TCT(3, 0x50u);
            // ---- answer projection with the int8 prefilter (see k_story), logits in registers, W8 and W rows from L2 ----
            for (unsigned j = lane; j < 64; j += 32) ufl[j] = (j < d) ? (float)uvec[j] / (float)(1 << fu) : 0.0f;
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
            float zr[8];                      // logits / exponentials of rows i = 32 k + lane
#pragma unroll
            for (int k = 0; k < 8; k++) zr[k] = -INFINITY;
            float zmax = -INFINITY;
            unsigned n_cand = 0, cand_idx = 0;
            bool need_full = !p.w8_ok || p.want_h;
            if (!need_full) {
                int n1 = 0;
                for (unsigned j = lane; j < 64; j += 32) n1 += abs((int)uvec[j]);
                n1 = __reduce_add_sync(0xffffffffu, n1);
                const int T = n1 + (n1 >> 6) + p.ans_margin + 2;
                const unsigned nw16 = (d + 15) / 16;
                const unsigned char *W8g = p.img + p.offW8;
                int Dv[8];
                int Dmax = INT_MIN;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const unsigned i = 32u * k + lane;
                    int D = 0;
                    if (32u * k < V) {
                        const unsigned char *wrow = W8g + (size_t)min(i, V - 1) * p.W8S;
#pragma unroll 4
                        for (unsigned w16 = 0; w16 < nw16; w16++) {
                            const uint4 ww = __ldg(reinterpret_cast<const uint4 *>(wrow + 16u * w16));
                            const uint4 uu = *reinterpret_cast<const uint4 *>(uvec + 16 * w16);
                            D = __dp4a((int)ww.x, (int)uu.x, D);
                            D = __dp4a((int)ww.y, (int)uu.y, D);
                            D = __dp4a((int)ww.z, (int)uu.z, D);
                            D = __dp4a((int)ww.w, (int)uu.w, D);
                        }
                    }
                    Dv[k] = (i < V) ? D : INT_MIN;
                    Dmax = max(Dmax, Dv[k]);
                }
                Dmax = __reduce_max_sync(0xffffffffu, Dmax);
                const int thr = Dmax - T;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const unsigned i = 32u * k + lane;
                    if (i < V) {
                        const bool cnd = Dv[k] >= thr;
                        if (cnd) zr[k] = exact_z(i);
                        zmax = fmaxf(zmax, zr[k]);
                        if (DUMP && p.dbg.dev_cand) p.dbg.dev_cand[(size_t)story * V + i] = cnd ? 1 : 0;
                        if (DUMP && p.dbg.dev_z) p.dbg.dev_z[(size_t)story * V + i] = zr[k];
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const unsigned i = 32u * k + lane;
                    const bool cand = (i < V) && (__expf(zr[k] - zmax) >= 0.99999905f);
                    const unsigned b = __ballot_sync(0xffffffffu, cand);
                    if (b) { n_cand += __popc(b); cand_idx = 32u * k + 31 - __clz(b); }
                }
                need_full = n_cand > 1;                          // near-tie: the double total decides, every row exactly
            }
            if (need_full) {
                zmax = -INFINITY;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const unsigned i = 32u * k + lane;
                    if (i < V) {
                        zr[k] = exact_z(i);
                        zmax = fmaxf(zmax, zr[k]);
                        if (DUMP && p.dbg.dev_cand) p.dbg.dev_cand[(size_t)story * V + i] = 2;
                        if (DUMP && p.dbg.dev_z) p.dbg.dev_z[(size_t)story * V + i] = zr[k];
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
                n_cand = 0; cand_idx = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const unsigned i = 32u * k + lane;
                    bool cand = false;
                    if (i < V) {
                        zr[k] = __expf(zr[k] - zmax);
                        cand = (zr[k] >= 0.99999905f);
                    } else zr[k] = 0.0f;
                    const unsigned b = __ballot_sync(0xffffffffu, cand);
                    if (b) { n_cand += __popc(b); cand_idx = 32u * k + 31 - __clz(b); }
                }
            }
            unsigned pred_i = cand_idx;
            float h_true_v = 0.0f;
            if (need_full && ((n_cand > 1) || p.want_h)) {
                double total = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (32u * k < V)
                        for (unsigned l = 0; l < 32; l++) {
                            const float e = __shfl_sync(0xffffffffu, zr[k], (int)l);
                            if (32u * k + l < V) total += (double)e;
                        }
                float best = -INFINITY;
                unsigned best_i = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const unsigned i = 32u * k + lane;
                    if (i < V) {
                        const float hv = (float)((double)zr[k] / total);
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

// qmann_forward.cu -- batched quantized MemN2N inference forward for sm_100a (include/qmann_abi.h
// part 2).  Replaces the reference's one-story-at-a-time host loop (MemN2N/MemN2N.c:2377-2703,
// 31 kernel launches per story) by tiers of kernels per chunk of stories:
//
//   k_story    : (qmann_fast.cuh) the production kernel: one warp per story streams the story's dense fp32
//                bag-of-words rows (the reference's boundary format, MemN2N.c:2294-2350) from HBM with bulk
//                asynchronous copies, compacts them in shared memory and runs the whole forward; packed
//                (SWAR) tier first, unpacked tier for what that declines.
//   k_compact  : dense arenas -> compact records in global memory, for the stories k_story declines
//                (fractional values, large counts, very dense stories) and for the instrumented pass.
//   k_forward  : the general kernel (every input, optional dumps of every intermediate): one warp per story,
//                persistent CTAs, all quantised weight tables resident in shared memory as int8 codes.
//
// Arithmetic follows SURVEY.md Appendix A (integer forms proven against the reference by
// tests/golden/kat_*.npz); every formula cites the reference line it reproduces.
#include "qmann_fixed.cuh"
#include "qmann_common.h"
#include "../../include/qmann_abi.h"

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <time.h>
#include <unistd.h>

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

#include "qmann_kernels.cuh"
#include "qmann_fast.cuh"
#include "qmann_tcstory.cuh"

namespace {

int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define QCUDA(expr)                                                                          \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) return fail(QMANN_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

}  // namespace

// =============================================================================================
// host objects
// =============================================================================================
struct qmann_model {
    qmann_config cfg;
    int device;
    int sm_count;
    unsigned compact_ctas_per_sm = 8;   // grid cap of k_compact in CTAs per SM (QMANN_COMPACT_CTAS_PER_SM)
    int max_smem = 0;
    unsigned char *dev_img = nullptr;
    signed char *dev_lut = nullptr;          // linear-map product tables (k_prep_lut), NULL when too large
    FwdParams base;                // everything but the per-call fields
    unsigned LPR, NW, smem_bytes;
    unsigned NW_fast = 0;          // warps per CTA of k_story
    unsigned long long *dev_path_count = nullptr;   // [4] stories that entered each tier (qmann_path_counts)
    unsigned story_chunk_cap = 0;  // stories per launch of the dense production path (records are needed only for declined stories)
    double prof_acc_c = 0.0, prof_acc_f = 0.0;      // folded event pairs
    size_t prof_pairs = 0;
    unsigned rec_stride, off_rend, off_exc, off_ent, lcap;
    // chunk scratch
    unsigned chunk_cap;
    unsigned char *dev_rec = nullptr;
    uint2 *dev_heap = nullptr;
    unsigned long long heap_cap;
    // control block (one 32-byte memset per chunk): heap_used u64 | counter | slow_count | counter2
    unsigned long long *dev_heap_used = nullptr;
    unsigned *dev_counter = nullptr, *dev_err = nullptr;
    unsigned *dev_slow_list = nullptr;       // [chunk_cap] chunk indices the fast kernel left to the general one
    unsigned *dev_slow_list2 = nullptr;      // [chunk_cap] chunk indices the packed kernel left to the unpacked fast kernel
    unsigned char *dev_img_swar = nullptr;   // image with biased A_h tables (packed path of k_forward_fast)
    bool swar_ok = false;
    bool fast_ok = false;                    // every weight format has an integer bit (Q_w(1.0) = 2^frac_w)
    unsigned char *dev_colmax = nullptr;     // [V] max |code| per column over all embedding tables (count splitting)
    unsigned nmax = 0, split_lim = 127;
    // tensor-core tier (k_story_tc): fp32 image of the A_h tables [160][V] for the tf32 MMA, its tensor map, geometry
    float *dev_tc_tab = nullptr;
    bool tc_ok = false;
    bool dense_stream = false;               // QMANN_DENSE_STREAM=1 at create: k_story streams the dense rows itself (staging buffers laid out)
    CUtensorMap tc_tmT;
    unsigned tc_kch = 0, tc_teams = 0, tc_smem = 0;
    // qmann_infer_host staging (grow-only device arenas, two streams)
    float *e2e_m = nullptr, *e2e_q = nullptr, *e2e_a = nullptr, *e2e_h = nullptr;
    uint32_t *e2e_pred = nullptr, *e2e_match = nullptr;
    size_t e2e_m_cap = 0, e2e_n_cap = 0;
    cudaStream_t e2e_compute = nullptr, e2e_copy = nullptr;
    std::vector<cudaEvent_t> e2e_events;
    // sentence offsets of the host entries (grow-only device buffer, no allocation per call)
    qmann_batch *host_batch = nullptr;
    size_t host_batch_cap = 0;
    // qmann_infer_ids_host staging (grow-only)
    uint16_t *ids_dev = nullptr;
    uint32_t *rowoff_dev = nullptr, *ans_dev = nullptr, *e2e_pred2 = nullptr;
    float *e2e_h2 = nullptr;
    size_t ids_cap = 0, rows_cap = 0, e2e_n_cap2 = 0;
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
constexpr unsigned CTRL_BYTES = 128;          // control block of a chunk: heap_used u64 | 30 x u32 counters
// counter slots: 0 packed tier claims, 1 unpacked tier claims, 2 packed->unpacked list length, 3 ->general list length,
// 8.. general-kernel claims per slice of the declined list
constexpr unsigned CT_PACKED = 0, CT_UNPACKED = 1, CT_LIST2 = 2, CT_LIST1 = 3, CT_GENERAL0 = 8;

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

// k_story: up to 24 warps per SM where shared memory and the 80-register budget allow (MAXT 768), 16 otherwise; the DUMP
// instantiations exist for MAXT 512 only.
template <int LPR, int MODE, bool SWAR, bool DENSE>
int launch_story_t(const qmann_model *m, const FwdParams &p, bool dump, cudaStream_t st)
{
    // small launches do not need all the warps: about one story per warp and SM at least, 8 warps minimum
    const unsigned per_sm = (p.n_stories + (unsigned)m->sm_count - 1) / (unsigned)m->sm_count;
    unsigned nw = std::min(m->NW_fast, std::max(std::min(8u, m->NW_fast), (per_sm + 3) / 4 * 4));
    if (dump) nw = std::min(nw, 16u);
    // 28 / 32 warps (64 registers each) exist for the packed record tier of the d <= 64 / dot-attention shape only (QMANN_FAST_WARPS=28 / 32)
    constexpr bool HAS896 = SWAR && !DENSE && LPR == 4 && MODE == 2;
    if (!HAS896) nw = std::min(nw, 24u);
    const unsigned smem = p.fl.tables_bytes + nw * p.fl.warp_bytes;
    const unsigned grid = (unsigned)m->sm_count;
#define QM_LAUNCH(DUMP_, MAXT_)                                                                                              \
    do {                                                                                                                      \
        static bool attr_done = false;                                                                                        \
        if (!attr_done) {                                                                                                     \
            QCUDA(cudaFuncSetAttribute(k_story<LPR, MODE, SWAR, DENSE, DUMP_, MAXT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, m->max_smem)); \
            attr_done = true;                                                                                                 \
        }                                                                                                                     \
        k_story<LPR, MODE, SWAR, DENSE, DUMP_, MAXT_><<<grid, nw * 32, smem, st>>>(p);                                        \
    } while (0)
    if (dump) QM_LAUNCH(true, 512);
    else if (nw > 28) { if constexpr (HAS896) QM_LAUNCH(false, 1024); }
    else if (nw > 24) { if constexpr (HAS896) QM_LAUNCH(false, 896); }
    else if (nw > 16) QM_LAUNCH(false, 768);
    else QM_LAUNCH(false, 512);
#undef QM_LAUNCH
    count_launch();
    QCUDA(cudaPeekAtLastError());
    return QMANN_OK;
}
template <int LPR, bool DENSE>
int launch_story_l(const qmann_model *m, const FwdParams &p, bool swar, bool dump, cudaStream_t st)
{
    if (swar) return launch_story_t<LPR, 2, true, DENSE>(m, p, dump, st);
    return m->cfg.mode == 3 ? launch_story_t<LPR, 3, false, DENSE>(m, p, dump, st) : launch_story_t<LPR, 2, false, DENSE>(m, p, dump, st);
}
template <bool DENSE>
int launch_story_d(const qmann_model *m, const FwdParams &p, bool swar, bool dump, cudaStream_t st)
{
    switch (m->LPR) {
        case 4: return launch_story_l<4, DENSE>(m, p, swar, dump, st);
        case 8: return launch_story_l<8, DENSE>(m, p, swar, dump, st);
        case 16: return launch_story_l<16, DENSE>(m, p, swar, dump, st);
        default: return launch_story_l<32, DENSE>(m, p, swar, dump, st);
    }
}
int launch_story(const qmann_model *m, const FwdParams &p, bool swar, bool dense, bool dump, cudaStream_t st)
{
    return dense ? launch_story_d<true>(m, p, swar, dump, st) : launch_story_d<false>(m, p, swar, dump, st);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tmap_encode_fn tmap_encoder()
{
    static tmap_encode_fn fn = []() -> tmap_encode_fn {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        return (tmap_encode_fn)f;
    }();
    return fn;
}
// fp32 matrix [rows][cols] (row pitch cols * 4, a multiple of 16 bytes), boxes of 32 columns x box_rows rows, 128-byte swizzle
bool tmap_rows_f32(CUtensorMap *out, const void *base, unsigned long long rows, unsigned cols, unsigned box_rows)
{
    tmap_encode_fn enc = tmap_encoder();
    if (!enc || rows == 0) return false;
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, box_rows}, es[2] = {1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef QMANN_TC_TRACE
unsigned *g_tc_trace = nullptr;
unsigned g_tc_trace_blocks = 0, g_tc_trace_warps = 0;
#endif
// k_story_tc over the whole chunk (dense arenas, every story <= 64 sentences); declined stories -> p.slow_list
int launch_story_tc(const qmann_model *m, const FwdParams &p, const float *dev_m, unsigned long long total_rows, bool dump, cudaStream_t st)
{
    TcParams tp;
    memset(&tp, 0, sizeof(tp));
    if (!tmap_rows_f32(&tp.tmX, dev_m, total_rows, m->cfg.V, 32)) return fail(QMANN_E_CUDA, "cuTensorMapEncodeTiled failed for the sentence arena");
    tp.tmT = m->tc_tmT;
    tp.f = p;
    tp.n_groups = (p.n_stories + 3) / 4;
    tp.total_rows = (unsigned)total_rows;
    tp.kch = m->tc_kch;
    tp.n_teams = m->tc_teams;
    const unsigned grid = (unsigned)std::min<unsigned>((unsigned)m->sm_count, tp.n_groups);
    const unsigned block = TC_WARPS * 32;
#ifdef QMANN_TC_TRACE
    if (!g_tc_trace) QCUDA(cudaHostAlloc((void **)&g_tc_trace, 2 * 148 * 24 * 16, cudaHostAllocMapped));
    memset(g_tc_trace, 0, 2 * 148 * 24 * 16);
    { unsigned *dp = nullptr; QCUDA(cudaHostGetDevicePointer((void **)&dp, g_tc_trace, 0)); tp.trace = dp; }
    g_tc_trace_blocks = grid; g_tc_trace_warps = block / 32;
#endif
    static bool attr_done[2] = {false, false};
    if (dump) {
        if (!attr_done[1]) { QCUDA(cudaFuncSetAttribute(k_story_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, m->max_smem)); attr_done[1] = true; }
        k_story_tc<true><<<grid, block, m->tc_smem, st>>>(tp);
    } else {
        if (!attr_done[0]) { QCUDA(cudaFuncSetAttribute(k_story_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, m->max_smem)); attr_done[0] = true; }
        k_story_tc<false><<<grid, block, m->tc_smem, st>>>(tp);
    }
    count_launch();
    QCUDA(cudaPeekAtLastError());
    return QMANN_OK;
}
}  // namespace

extern "C" {

#ifdef QMANN_TC_TRACE
// debug builds only: where every warp of the last k_story_tc launch stands (CTAs that have not reached the end)
void qmann_tc_trace_dump(void)
{
    if (!g_tc_trace) return;
    // clock accumulators (units of 16 cycles) of CTAs 0 and 1, four per warp (meaning per role, see qmann_tcstory.cuh)
    for (unsigned b = 0; b < 2 && b < g_tc_trace_blocks; b++) {
        const unsigned *t = g_tc_trace + 148 * 24 * 4 + (size_t)b * 24 * 4;
        fprintf(stderr, "CTA %u clocks/16\n", b);
        for (unsigned w = 0; w < g_tc_trace_warps; w++) fprintf(stderr, "  warp %2u: %9u %9u %9u %9u\n", w, t[4 * w], t[4 * w + 1], t[4 * w + 2], t[4 * w + 3]);
    }
    for (unsigned b = 0; b < g_tc_trace_blocks; b++) {
        const unsigned *t = g_tc_trace + (size_t)b * 24 * 4;
        bool done = true;
        for (unsigned w = 0; w < g_tc_trace_warps; w++) done = done && (t[4 * w] == 0xE0Du);
        if (done) continue;
        fprintf(stderr, "CTA %u\n", b);
        for (unsigned w = 0; w < g_tc_trace_warps; w++) fprintf(stderr, "  warp %2u: %08x %08x %08x %08x\n", w, t[4 * w], t[4 * w + 1], t[4 * w + 2], t[4 * w + 3]);
    }
    fflush(stderr);
}
#endif
const char *qmann_last_error(void) { return g_err.c_str(); }
const char *qmann_version(void) { return "qmann_b200 0.1 (sm_100a)"; }
uint64_t qmann_launch_count(void) { return g_launches.load(); }

static int model_build(qmann_model *m, const qmann_config *cfg, const qmann_weights *w);

int qmann_model_create(qmann_model **out, const qmann_config *cfg, const qmann_weights *w)
{
    if (!out || !cfg || !w) return fail(QMANN_E_ARG, "null argument");
    *out = nullptr;
    qmann_model *m = new qmann_model();
    const int rc = model_build(m, cfg, w);
    if (rc != QMANN_OK) { qmann_model_destroy(m); return rc; }      // frees whatever the failed build had allocated
    *out = m;
    return QMANN_OK;
}

static int model_build(qmann_model *m, const qmann_config *cfg, const qmann_weights *w)
{
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

    m->cfg = c;
    QCUDA(cudaGetDevice(&m->device));
    cudaDeviceProp prop;
    QCUDA(cudaGetDeviceProperties(&prop, m->device));
    m->sm_count = prop.multiProcessorCount;
    if (const char *e = getenv("QMANN_COMPACT_CTAS_PER_SM")) m->compact_ctas_per_sm = (unsigned)std::min(64, std::max(1, atoi(e)));

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
    p.offB = take((c.V + 1) * DP);                     // +1: the all-zero row idle gather lanes read
    for (unsigned h = 0; h < c.H; h++) { p.offA[h] = take((c.V + 1) * DP); p.offC[h] = take((c.V + 1) * DP); }
    for (unsigned h = 0; h < c.H; h++) p.offH[h] = c.lin_map ? take(c.d * HS) : 0;
    p.offCM[0] = take((c.V + 1) * 4);                  // cm10[V+1]: column maxima of A_h, three hops per word
    p.offTAU = take(128);
    p.offW = take(c.V * WS * 4);
    p.tables_bytes = off;                              // shared-memory image of the general kernel: everything up to here
    unsigned W8S = round_up(c.d, 16);
    if (((W8S / 16) & 1u) == 0) W8S += 16;             // odd number of 16-byte units: conflict-free lane-per-row reads
    p.W8S = W8S;
    p.w8_bytes = round_up(c.V * W8S, 16);
    p.offW8 = take(p.w8_bytes);                        // global only; k_forward_fast places it at offW in its shared memory
    p.img_bytes = off;
    p.V = c.V; p.d = c.d; p.S_max = c.S_max; p.H = c.H; p.lin_map = c.lin_map; p.const_scale = c.const_scale;
    p.DP = DP; p.HS = HS; p.WS = WS;
    for (unsigned h = 0; h < c.H; h++) {
        p.fw[h] = c.frac_w[h]; p.iw[h] = c.iwl_w[h]; p.lw[h] = fixed_max(c.iwl_w[h], c.frac_w[h]);
        p.fa[h] = c.frac_att[h]; p.ia[h] = c.iwl_att[h]; p.la[h] = fixed_max(c.iwl_att[h], c.frac_att[h]);
        p.ff[h] = c.frac[h]; p.iff[h] = c.iwl[h]; p.lf[h] = fixed_max(c.iwl[h], c.frac[h]);
    }
    p.fb = c.frac_bin; p.lb = fixed_max(c.iwl_bin, c.frac_bin);
    p.en_sc_att = c.en_sc_att ? 1 : 0; p.en_non_lin = c.en_non_lin ? 1 : 0;
    for (unsigned h = 0; h < c.H; h++) p.sc_w[h] = c.sc_att_w[h];

    // ---- per-warp scratch layout ----
    const unsigned S_pad = round_up(c.S_max, 32);
    unsigned o = 0;
    auto wtake = [&](unsigned bytes) { unsigned r = o; o += round_up(bytes, 16); return r; };
    const unsigned ent_fixed = round_up(c.V, 32) * 4;            // the region is reused as zbuf[V]
    p.o_rend = 0;  // placeholder, set below
    int max_smem = 0;
    QCUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
    m->max_smem = max_smem;
    unsigned NW = 0, LW = 0;
    unsigned fixed_warp = 0;
    {
        // everything except the entry list
        o = 0;
        wtake(0);
        unsigned o_rend = wtake((c.S_max + 2) * 2), o_sc = wtake(S_pad * 4), o_ex = wtake(S_pad * 4), o_pq = wtake(S_pad);
        unsigned o_uvec = wtake(DP), o_ub32 = wtake(DP * 4), o_ovec = wtake(DP), o_ufl = wtake(DP * 4), o_exc = wtake(MAX_EXC * 8);
        unsigned o_zent = wtake(16);
        unsigned o_brow = wtake(c.H * S_pad);                 // packed path: row biases per hop
        unsigned o_perm = wtake(S_pad * 2), o_cnt = wtake(20 * 4);        // packed path: rows ordered by length
        fixed_warp = o;
        p.o_brow = o_brow; p.o_perm = o_perm; p.o_cnt = o_cnt;
        p.o_rend = o_rend; p.o_sc = o_sc; p.o_ex = o_ex; p.o_pq = o_pq; p.o_uvec = o_uvec; p.o_ub32 = o_ub32;
        p.o_ovec = o_ovec; p.o_ufl = o_ufl; p.o_exc = o_exc; p.o_zent = o_zent;
    }
    const unsigned want_LW = std::min(65535u, round_up(12 * (c.S_max + 1), 32));
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
        return fail(QMANN_E_NOMEM, "model tables (" + std::to_string(p.tables_bytes) + " B) plus per-warp scratch do not fit the SM's " +
                                       std::to_string(max_smem) + " B of shared memory");
    }
    const unsigned ent_region = round_up(std::max(LW * 4, ent_fixed), 16);
    // entry list sits first in the warp scratch; shift the other offsets behind it
    p.o_rend += ent_region; p.o_sc += ent_region; p.o_ex += ent_region; p.o_pq += ent_region; p.o_uvec += ent_region;
    p.o_ub32 += ent_region; p.o_ovec += ent_region; p.o_ufl += ent_region; p.o_exc += ent_region; p.o_zent += ent_region; p.o_brow += ent_region; p.o_perm += ent_region; p.o_cnt += ent_region;
    p.warp_bytes = ent_region + fixed_warp;
    p.LW = LW; p.S_pad = S_pad;
    m->NW = NW;
    m->NW_fast = 0;            // sized below, once count splitting and the packed path are known
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
    // linear-map product tables: H * d * 255 rows of DP bytes (2.4 MB at d = 50); skipped beyond 64 MB
    if (c.lin_map && (size_t)c.H * c.d * 255 * DP <= (64u << 20)) {
        QCUDA(cudaMalloc((void **)&m->dev_lut, (size_t)c.H * c.d * 255 * DP));
        for (unsigned h = 0; h < c.H; h++) {
            p.offL[h] = (unsigned)((size_t)h * c.d * 255 * DP);
            k_prep_lut<<<256, 256>>>(w->dev_Hm[h], m->dev_lut + p.offL[h], c.d, DP, c.iwl_w[h], c.frac_w[h], c.frac_bin);
            count_launch();
        }
        p.lut = m->dev_lut;
    }
    k_prep_ans<<<64, 256>>>(w->dev_W, reinterpret_cast<float *>(m->dev_img + p.offW), c.V, c.d, WS);
    count_launch();
    {
        // int8 image of W for the answer prefilter: W8 = rint(W / s), s = max|W| / 127 (host side, once)
        std::vector<float> hw((size_t)c.V * c.d);
        QCUDA(cudaMemcpy(hw.data(), w->dev_W, hw.size() * sizeof(float), cudaMemcpyDeviceToHost));
        double wmax = 0.0;
        bool finite = true;
        for (float x : hw) { if (!std::isfinite(x)) finite = false; wmax = std::max(wmax, (double)std::fabs(x)); }
        const char *env_pf = getenv("QMANN_ANS_PREFILTER");
        p.w8_ok = (finite && wmax > 0.0 && !(env_pf && atoi(env_pf) == 0)) ? 1 : 0;
        if (p.w8_ok) {
            const double sc = wmax / 127.0;
            std::vector<signed char> h8((size_t)p.w8_bytes, 0);
            for (unsigned i = 0; i < c.V; i++)
                for (unsigned j = 0; j < c.d; j++) h8[(size_t)i * W8S + j] = (signed char)std::lrint((double)hw[(size_t)i * c.d + j] / sc);
            QCUDA(cudaMemcpy(m->dev_img + p.offW8, h8.data(), h8.size(), cudaMemcpyHostToDevice));
            // 1e-5 (the window in which two logits can share the maximal probability) in units of the integer dot product;
            // the final u has frac[H-1] fractional bits (frac_w[0] when there is no hop)
            const unsigned fu_last = c.H ? c.frac[c.H - 1] : c.frac_w[0];
            const double mu = 1e-5 * std::ldexp(1.0, (int)fu_last) / sc;
            p.ans_margin = (mu < 1e9) ? (int)std::ceil(mu) + 1 : 0;
            if (!(mu < 1e9)) p.w8_ok = 0;
        }
    }
    // count splitting (k_compact): per-column max |code| and the largest count every weight format represents
    QCUDA(cudaMalloc((void **)&m->dev_colmax, c.V));
    QCUDA(cudaMemset(m->dev_colmax, 0, c.V));
    {
        std::vector<unsigned> offs = {p.offB};
        for (unsigned h = 0; h < c.H; h++) { offs.push_back(p.offA[h]); offs.push_back(p.offC[h]); }
        for (unsigned o_ : offs) {
            k_colmax<<<(c.V + 127) / 128, 128>>>(reinterpret_cast<const signed char *>(m->dev_img + o_), c.V, DP, m->dev_colmax);
            count_launch();
        }
        unsigned nmax = 127;
        bool unit_ok = true;
        for (unsigned h = 0; h < c.H; h++) {
            nmax = std::min(nmax, (unsigned)fixed_max(c.iwl_w[h], c.frac_w[h]) >> c.frac_w[h]);
            if (c.iwl_w[h] < 1) unit_ok = false;
        }
        m->nmax = unit_ok ? nmax : 0;
        // the optional layers (scale before the softmax, RELU after the update) exist in the general kernel only
        m->fast_ok = unit_ok && !c.en_sc_att && !c.en_non_lin;
        // n unit entries stand for a count n only while n * max|code| stays inside EVERY hop's weight format
        // (the reference clamps each product Q_w(Q_w(n) * Q_w(T)) to that hop's limit, lib/layer_cuda.cu:120)
        unsigned sl = 127;
        for (unsigned h = 0; h < c.H; h++) sl = std::min(sl, (unsigned)p.lw[h]);
        m->split_lim = sl;
    }
    // packed path (mode 2): every hop with an 8-bit memory/addressing format, two fractional bits in the query operand and
    // |frac_att - frac_w| <= 1.  QMANN_SWAR=0 keeps the unpacked kernel only (A/B tests).
    {
        const char *env_swar = getenv("QMANN_SWAR");
        bool ok = m->fast_ok && c.mode == 2 && c.frac_bin == 2 && c.H <= 3 && (DP & (DP - 1)) == 0 && !(env_swar && atoi(env_swar) == 0);
        for (unsigned h = 0; h < c.H && ok; h++) {
            const int ka = (int)c.frac_att[h] - (int)c.frac_w[h];
            ok = p.lw[h] == 127 && p.la[h] == 127 && ka >= -1 && ka <= 1;
        }
        if (ok) {
            k_prep_tau<<<1, 128>>>(m->dev_img + p.offTAU);
            count_launch();
            QCUDA(cudaMalloc((void **)&m->dev_img_swar, p.img_bytes));
            QCUDA(cudaMemcpy(m->dev_img_swar, m->dev_img, p.img_bytes, cudaMemcpyDeviceToDevice));
            for (unsigned h = 0; h < c.H; h++) {
                k_prep_bias<<<(c.V + 128) / 128, 128>>>(reinterpret_cast<signed char *>(m->dev_img_swar + p.offA[h]),
                                                        reinterpret_cast<unsigned *>(m->dev_img_swar + p.offCM[0]), h, c.V, DP);
                count_launch();
            }
            m->swar_ok = true;
        }
    }
    // ---- production kernel k_story: shared-memory image (A_h, column maxima, tau, int8 W) and per-warp scratch ----
    {
        FastLayout &fl = p.fl;
        unsigned t = 0;
        auto ttake = [&](unsigned bytes) { unsigned o_ = t; t += round_up(bytes, 16); return o_; };
        for (unsigned h = 0; h < c.H; h++) fl.sA[h] = ttake((c.V + 1) * DP);
        fl.sCM = ttake((c.V + 1) * 4);
        fl.sTAU = ttake(128);
        fl.sW8 = ttake(p.w8_bytes);
        fl.tables_bytes = round_up(t, 128);
        const unsigned row_bytes = c.V * 4;
        unsigned R = std::max(1u, 2048u / row_bytes), NB = 2;
        if (const char *e = getenv("QMANN_STAGE_ROWS")) R = (unsigned)std::max(1, atoi(e));
        if (const char *e = getenv("QMANN_STAGE_BUFS")) NB = (unsigned)std::min(8, std::max(1, atoi(e)));
        fl.R = R; fl.NB = NB;
        fl.buf_bytes = round_up(R * row_bytes + 32, 128);
        unsigned o2 = 0;
        auto w2 = [&](unsigned bytes, unsigned al = 16) { o2 = round_up(o2, al); unsigned r_ = o2; o2 += bytes; return r_; };
        // entry list first (16-bit table-row offsets); the region doubles as zbuf[V] in the answer phase
        const unsigned want_ent = std::min(65535u, round_up(8 * (c.S_max + 1), 32));
        const unsigned ent_bytes = round_up(std::max(want_ent * 2, c.V * 4), 16);
        w2(ent_bytes);
        fl.LW = ent_bytes / 2;
        fl.o_rend = w2((c.S_max + 2) * 2);
        fl.o_sc = w2(S_pad * 4); fl.o_ex = w2(S_pad * 4); fl.o_pq = w2(S_pad);
        fl.o_uvec = w2(DP); fl.o_ub32 = w2(DP * 4); fl.o_ovec = w2(DP); fl.o_ufl = w2(DP * 4);
        fl.o_zent = w2(16);
        fl.o_brow = w2(c.H * S_pad); fl.o_perm = w2(S_pad * 2); fl.o_cnt = w2(20 * 4);
        // the staging buffers of the dense stream (4.3 KB per warp at C2) are only laid out when that opt-in mode is on: without
        // them 24 warps fit beside the tables instead of 20
        const char *env_ds = getenv("QMANN_DENSE_STREAM"), *env_tc0 = getenv("QMANN_TC");
        m->dense_stream = env_ds && atoi(env_ds) == 1;
        // (the tensor-core tier hands what it declines to k_story with the DENSE source, which needs them too)
        const bool need_stage = m->dense_stream || (env_tc0 && atoi(env_tc0) == 1);
        fl.o_bar = w2(8 * NB, 8);
        fl.o_stage = w2(need_stage ? NB * fl.buf_bytes : 0u, 128);
        fl.warp_bytes = round_up(o2, 128);
        // entries are 16-bit offsets column * DP
        if ((size_t)(c.V + 1) * DP > 65535u) m->fast_ok = false;
        unsigned cap = 28;      // 28 measured best for the packed record tier (24: 0.302 ms, 28: 0.292, 32: 0.307 on C2; profiles/r02_story_warps_sweep.txt)
        if (const char *e = getenv("QMANN_FAST_WARPS")) cap = (unsigned)std::max(1, atoi(e));
        m->NW_fast = 0;
        for (unsigned want : {32u, 28u, 24u, 22u, 20u, 18u, 16u, 14u, 12u, 10u, 8u, 6u, 4u, 2u, 1u}) {
            if (want > cap) continue;
            if ((size_t)fl.tables_bytes + (size_t)want * fl.warp_bytes + 1024 <= (size_t)max_smem) { m->NW_fast = want; break; }
        }
        if (m->NW_fast == 0) m->fast_ok = false;
        p.colmax = m->dev_colmax; p.nmax = m->nmax; p.split_lim = m->split_lim;
        const char *efs = getenv("QMANN_FAST_SOFTMAX");
        p.fast_softmax = (efs && atoi(efs) == 0) ? 0 : 1;
        p.pf_dist = 384; p.pf_mode = 1;
        if (const char *e = getenv("QMANN_PF_DIST")) p.pf_dist = (unsigned)std::max(0, atoi(e));
        if (const char *e = getenv("QMANN_PF_MODE")) p.pf_mode = (unsigned)std::max(0, atoi(e));
        if (p.pf_dist == 0) p.pf_mode = 0;
    }
    // ---- tensor-core tier: dense rows x A_h tables on tcgen05 (k_story_tc) ----
    {
        const char *env_tc = getenv("QMANN_TC");
        const unsigned kch = (c.V + 31) / 32;
        bool ok = m->swar_ok && c.H <= 3 && c.d <= TC_HCOLS && DP == 64 && c.V % 4 == 0 && c.V <= 256 && (!c.lin_map || p.lut) && c.S_max >= 1 &&
                  (env_tc && atoi(env_tc) == 1) && tmap_encoder() != nullptr;      // opt-in until it beats the two-kernel path
        unsigned teams = TC_MAX_TEAMS;
        if (const char *e = getenv("QMANN_TC_TEAMS")) teams = (unsigned)std::min<int>(TC_MAX_TEAMS, std::max(1, atoi(e)));
        unsigned need = 0;
        for (; ok && teams >= 1; teams--) {
            need = kch * TC_TABCH_BYTES + TC_NSTAGE * TC_STAGE_BYTES + TCB_BYTES + 4 * teams * TW_BYTES + 1024;
            if (need <= (unsigned)max_smem) break;
        }
        if (ok && teams >= 1) {
            QCUDA(cudaMalloc((void **)&m->dev_tc_tab, (size_t)TC_NT * c.V * sizeof(float)));
            k_prep_tc<<<64, 256>>>(m->dev_img, m->dev_tc_tab, c.V, c.d, DP, c.H, p.offA[0], c.H > 1 ? p.offA[1] : 0, c.H > 2 ? p.offA[2] : 0);
            count_launch();
            ok = tmap_rows_f32(&m->tc_tmT, m->dev_tc_tab, TC_NT, c.V, TC_NT);
            m->tc_ok = ok; m->tc_kch = kch; m->tc_teams = teams; m->tc_smem = need;
        }
    }
    QCUDA(cudaPeekAtLastError());
    QCUDA(cudaDeviceSynchronize());
    p.img = m->dev_img;

    // ---- compact-record geometry and chunk scratch ----
    m->lcap = LW;                 // a story that does not fit a warp's shared-memory list goes to the heap
    m->off_rend = REC_HDR_BYTES;
    m->off_exc = round_up(REC_HDR_BYTES + (c.S_max + 2) * 2, 16);
    m->off_ent = m->off_exc + MAX_EXC * 8;
    m->rec_stride = round_up(m->off_ent + m->lcap * 4, 16);
    m->chunk_cap = 32768;                    // records held at a time (ids input, instrumented pass, declined stories)
    m->story_chunk_cap = 8 * m->chunk_cap;   // stories per launch of the dense production path
    QCUDA(cudaMalloc((void **)&m->dev_rec, (size_t)m->chunk_cap * m->rec_stride));
    m->heap_cap = 8ull << 20;                                        // 8 Mi entries = 64 MiB
    QCUDA(cudaMalloc((void **)&m->dev_heap, m->heap_cap * sizeof(uint2)));
    // control block (one memset per chunk): heap_used u64 | 30 x u32 counters
    QCUDA(cudaMalloc((void **)&m->dev_heap_used, CTRL_BYTES));
    m->dev_counter = reinterpret_cast<unsigned *>(m->dev_heap_used + 1);
    QCUDA(cudaMalloc((void **)&m->dev_slow_list, (size_t)m->story_chunk_cap * sizeof(unsigned)));
    QCUDA(cudaMalloc((void **)&m->dev_slow_list2, (size_t)m->story_chunk_cap * sizeof(unsigned)));
    QCUDA(cudaMalloc((void **)&m->dev_err, sizeof(unsigned)));
    QCUDA(cudaMemset(m->dev_err, 0, sizeof(unsigned)));
    QCUDA(cudaMalloc((void **)&m->dev_path_count, 4 * sizeof(unsigned long long)));
    QCUDA(cudaMemset(m->dev_path_count, 0, 4 * sizeof(unsigned long long)));
    p.rec = m->dev_rec; p.rec_stride = m->rec_stride; p.off_rend = m->off_rend; p.off_exc = m->off_exc; p.off_ent = m->off_ent;
    p.heap = m->dev_heap; p.counter = m->dev_counter; p.err_flag = m->dev_err; p.path_count = m->dev_path_count;
    return QMANN_OK;
}

void qmann_model_destroy(qmann_model *m)
{
    if (!m) return;
    cudaFree(m->dev_tc_tab);
    cudaFree(m->dev_img); cudaFree(m->dev_lut); cudaFree(m->dev_rec); cudaFree(m->dev_heap); cudaFree(m->dev_heap_used);
    cudaFree(m->dev_path_count); cudaFree(m->dev_slow_list); cudaFree(m->dev_slow_list2); cudaFree(m->dev_img_swar); cudaFree(m->dev_err); cudaFree(m->dev_colmax);
    if (m->host_batch) qmann_batch_destroy(m->host_batch);
    cudaFree(m->ids_dev); cudaFree(m->rowoff_dev); cudaFree(m->ans_dev); cudaFree(m->e2e_pred2); cudaFree(m->e2e_h2);
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

// Batch descriptor of the host entries: owned by the model, its device buffer only grows, the offsets are uploaded on
// `st` (the copy stream, ahead of the first chunk).
static int host_batch_prepare(qmann_model *m, const uint32_t *n_sen, uint32_t N, cudaStream_t st, qmann_batch **out)
{
    if (!m->host_batch) { m->host_batch = new qmann_batch(); m->host_batch->dev_sen_off = nullptr; }
    qmann_batch *b = m->host_batch;
    b->N = N;
    b->sen_off.resize((size_t)N + 1);
    b->sen_off[0] = 0;
    b->max_sen = 0;
    for (uint32_t i = 0; i < N; i++) {
        b->sen_off[i + 1] = b->sen_off[i] + n_sen[i];
        b->max_sen = std::max(b->max_sen, n_sen[i]);
    }
    b->sum_sen = b->sen_off[N];
    if ((size_t)N + 1 > m->host_batch_cap) {
        cudaFree(b->dev_sen_off); b->dev_sen_off = nullptr; m->host_batch_cap = 0;
        QCUDA(cudaMalloc((void **)&b->dev_sen_off, ((size_t)N + 1) * sizeof(unsigned long long)));
        m->host_batch_cap = (size_t)N + 1;
    }
    QCUDA(cudaMemcpyAsync(b->dev_sen_off, b->sen_off.data(), ((size_t)N + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    *out = b;
    return QMANN_OK;
}

// The two input formats of a batch: the dense fp32 arenas (the reference's boundary, compacted by k_compact) or the
// word-id lists they are built from (k_ids_compact).  Both produce the same compact records.
struct FwdInput {
    const float *dev_m = nullptr, *dev_q = nullptr, *dev_a = nullptr;
    const uint16_t *dev_ids = nullptr;
    const uint32_t *dev_row_off = nullptr, *dev_ans = nullptr;
};

// Event triples of the optional profile are folded into running sums before the vector can grow without bound.
static int profile_fold(qmann_model *m)
{
    for (size_t i = 0; i + 3 <= m->prof_used; i += 3) {
        float a = 0.f, b2 = 0.f;
        QCUDA(cudaEventSynchronize(m->prof_events[i + 2]));
        QCUDA(cudaEventElapsedTime(&a, m->prof_events[i], m->prof_events[i + 1]));
        QCUDA(cudaEventElapsedTime(&b2, m->prof_events[i + 1], m->prof_events[i + 2]));
        m->prof_acc_c += a; m->prof_acc_f += b2;
    }
    m->prof_pairs += m->prof_used / 3;
    m->prof_used = 0;
    return QMANN_OK;
}

// Forward of stories [first, first+count) of the batch; data pointers are the FULL arenas.
//
// Production path (dbg == NULL or dbg->production):
//   dense input, 16-byte aligned arenas:  k_story<packed, DENSE> -> k_story<unpacked, DENSE> on what it declines ->
//                                         k_compact + k_forward (general) on what that declines, in slices of chunk_cap records
//   word-id input / unaligned arenas:     k_ids_compact | k_compact over the chunk, then k_story<.., RECORD> tiers, k_forward
// Instrumented path (dbg given, production == 0): k_compact | k_ids_compact, then k_forward<DEBUG> for every story.
static int forward_range(qmann_model *m, const qmann_batch *b, uint32_t first, uint32_t count, const FwdInput &in,
                         uint32_t *dev_pred, float *dev_h_true, uint32_t *dev_match, const qmann_debug *dbg, cudaStream_t st)
{
    const bool production = (dbg == nullptr) || dbg->production != 0;
    const bool dump = dbg != nullptr;
    const float *dev_m = in.dev_m, *dev_q = in.dev_q, *dev_a = in.dev_a;
    const bool vec4 = (m->cfg.V % 4 == 0) && (((uintptr_t)dev_m | (uintptr_t)dev_q) % 16 == 0);
    const bool vec2 = (m->cfg.V % 2 == 0) && (((uintptr_t)dev_m | (uintptr_t)dev_q) % 8 == 0);
    const bool fast = production && m->fast_ok && m->NW_fast > 0;
    const char *env_stream = getenv("QMANN_DENSE_STREAM");
    // k_story's own bulk-copy stream + in-kernel compaction measured slower than k_compact followed by the record tiers
    // (0.66 vs 0.53 ms on C2, profiles/r02_k_story_dense_stream.txt): opt-in only
    const bool aligned16 = !in.dev_ids && (((uintptr_t)dev_m | (uintptr_t)dev_q) % 16 == 0);
    // first tier on the tensor cores (k_story_tc): dense arenas, stories of at most 64 sentences
    const bool use_tc = fast && m->tc_ok && aligned16 && b->max_sen <= 64 && b->sum_sen > 0 && b->sum_sen < 0x7FFFFFFFull;
    (void)env_stream;
    const bool stream_dense = use_tc || (fast && aligned16 && m->dense_stream);
    const uint32_t cap = stream_dense ? m->story_chunk_cap : m->chunk_cap;
    unsigned *ctr = m->dev_counter;
    for (uint32_t s0 = first; s0 < first + count; s0 += cap) {
        const uint32_t n = std::min<uint32_t>(cap, first + count - s0);
        QCUDA(cudaMemsetAsync(m->dev_heap_used, 0, CTRL_BYTES, st));
        cudaEvent_t *pe = nullptr;
        if (m->profile) {
            if (m->prof_used + 3 > 3 * 1024) { const int rcf = profile_fold(m); if (rcf) return rcf; }
            if (m->prof_used + 3 > m->prof_events.size()) {
                for (int k = 0; k < 3; k++) { cudaEvent_t e; QCUDA(cudaEventCreate(&e)); m->prof_events.push_back(e); }
            }
            pe = &m->prof_events[m->prof_used];
            m->prof_used += 3;
            QCUDA(cudaEventRecord(pe[0], st));
        }
        auto compact_params = [&]() {
            CompactParams cp;
            cp.m = dev_m; cp.q = dev_q; cp.a = dev_a; cp.sen_off = b->dev_sen_off; cp.V = m->cfg.V; cp.S_max = m->cfg.S_max;
            cp.story0 = s0; cp.n_stories = n; cp.rec = m->dev_rec; cp.rec_stride = m->rec_stride; cp.off_rend = m->off_rend;
            cp.off_exc = m->off_exc; cp.off_ent = m->off_ent; cp.lcap = m->lcap; cp.heap = m->dev_heap; cp.heap_cap = m->heap_cap;
            cp.heap_used = m->dev_heap_used; cp.colmax = m->dev_colmax; cp.nmax = m->nmax; cp.split_lim = m->split_lim;
            cp.work_list = nullptr; cp.work_count = nullptr; cp.work_off = 0; cp.work_cap = 0;
            return cp;
        };
        auto launch_compact = [&](const CompactParams &cp, unsigned n_est) -> int {
            const unsigned cblocks = std::max(1u, std::min<unsigned>((n_est + 7) / 8, (unsigned)m->sm_count * m->compact_ctas_per_sm));
            if (vec4) k_compact<4><<<cblocks, 256, 0, st>>>(cp);
            else if (vec2) k_compact<2><<<cblocks, 256, 0, st>>>(cp);
            else           k_compact<1><<<cblocks, 256, 0, st>>>(cp);
            count_launch();
            QCUDA(cudaPeekAtLastError());
            return QMANN_OK;
        };
        if (!stream_dense) {
            // records for the whole chunk
            if (in.dev_ids) {
                IdsParams ip;
                ip.ids = in.dev_ids; ip.row_off = in.dev_row_off; ip.ans = in.dev_ans; ip.sen_off = b->dev_sen_off; ip.V = m->cfg.V;
                ip.story0 = s0; ip.n_stories = n; ip.rec = m->dev_rec; ip.rec_stride = m->rec_stride; ip.off_rend = m->off_rend;
                ip.off_exc = m->off_exc; ip.off_ent = m->off_ent; ip.lcap = m->lcap; ip.heap = m->dev_heap; ip.heap_cap = m->heap_cap;
                ip.heap_used = m->dev_heap_used; ip.colmax = m->dev_colmax; ip.nmax = m->nmax; ip.split_lim = m->split_lim;
                const unsigned cblocks = std::min<unsigned>((n + 7) / 8, (unsigned)m->sm_count * 8);
                k_ids_compact<<<cblocks, 256, 0, st>>>(ip);
                count_launch();
                QCUDA(cudaPeekAtLastError());
            } else {
                const int rcc = launch_compact(compact_params(), n);
                if (rcc) return rcc;
            }
        }
        if (pe) QCUDA(cudaEventRecord(pe[1], st));

        FwdParams p = m->base;
        p.sen_off = b->dev_sen_off; p.story0 = s0; p.n_stories = n; p.n_total = b->N; p.sum_sen = b->sum_sen;
        p.pred = dev_pred; p.h_true = dev_h_true; p.match = dev_match; p.want_h = (dev_h_true != nullptr);
        p.dm = dev_m; p.dq = dev_q; p.da = dev_a;
        p.m_bytes = b->sum_sen * (unsigned long long)m->cfg.V * 4ull; p.q_bytes = (unsigned long long)b->N * m->cfg.V * 4ull;
        if (dbg) p.dbg = *dbg;
        int rc;
        if (fast) {
            // regular stories in the production kernel(s); whatever they decline goes through the general one
            if (use_tc) {
                FwdParams ps = p;
                ps.counter = ctr + CT_PACKED;
                ps.slow_list = m->dev_slow_list2; ps.slow_count = ctr + CT_LIST2;
                rc = launch_story_tc(m, ps, dev_m, b->sum_sen, dump, st);
                if (rc) return rc;
                p.work_list = m->dev_slow_list2; p.work_count = ctr + CT_LIST2;
            } else if (m->swar_ok) {
                // packed embedding + scorer first; stories with a row whose column maxima add up above 127 go to the unpacked kernel
                FwdParams ps = p;
                ps.img = m->dev_img_swar;
                ps.counter = ctr + CT_PACKED;
                ps.slow_list = m->dev_slow_list2; ps.slow_count = ctr + CT_LIST2;
                rc = launch_story(m, ps, true, stream_dense, dump, st);
                if (rc) return rc;
                p.work_list = m->dev_slow_list2; p.work_count = ctr + CT_LIST2;
            }
            {
                FwdParams pf = p;
                pf.counter = ctr + CT_UNPACKED;
                pf.slow_list = m->dev_slow_list; pf.slow_count = ctr + CT_LIST1;
                rc = launch_story(m, pf, false, stream_dense, dump, st);
                if (rc) return rc;
            }
            p.work_list = m->dev_slow_list; p.work_count = ctr + CT_LIST1;
        }
        // general kernel: everything (instrumented pass) or the declined stories, whose records exist already (record source)
        // or are made now, chunk_cap at a time (dense stream)
        const unsigned slices = (fast && stream_dense) ? (n + m->chunk_cap - 1) / m->chunk_cap : 1u;
        for (unsigned sl = 0; sl < slices; sl++) {
            FwdParams pg = p;
            pg.counter = ctr + CT_GENERAL0 + sl;
            pg.work_off = 0; pg.work_cap = 0xFFFFFFFFu; pg.rec_by_pos = 0;
            if (fast && stream_dense) {
                CompactParams cp = compact_params();
                cp.work_list = m->dev_slow_list; cp.work_count = ctr + CT_LIST1; cp.work_off = sl * m->chunk_cap; cp.work_cap = m->chunk_cap;
                rc = launch_compact(cp, std::min<unsigned>(n, 4096u));
                if (rc) return rc;
                pg.work_off = sl * m->chunk_cap; pg.work_cap = m->chunk_cap; pg.rec_by_pos = 1;
            }
            rc = launch_forward(m, pg, dump, st);
            if (rc) return rc;
        }
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
    FwdInput in;
    in.dev_m = dev_m; in.dev_q = dev_q; in.dev_a = dev_a;
    return forward_range(m, b, 0, b->N, in, dev_pred, dev_h_true, dev_match, dbg, (cudaStream_t)stream);
}

int qmann_forward_ids(qmann_model *m, const qmann_batch *b, const uint16_t *dev_ids, const uint32_t *dev_row_off, const uint32_t *dev_ans,
                      uint32_t *dev_pred, float *dev_h_true, uint32_t *dev_match, const qmann_debug *dbg, void *stream)
{
    if (!m || !b || !dev_ids || !dev_row_off) return fail(QMANN_E_ARG, "null argument");
    if (b->max_sen > m->cfg.S_max) return fail(QMANN_E_ARG, "a story has more sentences than S_max");
    if (m->cfg.V > 65535u) return fail(QMANN_E_ARG, "ids are 16-bit");
    if (b->N == 0) return QMANN_OK;
    FwdInput in;
    in.dev_ids = dev_ids; in.dev_row_off = dev_row_off; in.dev_ans = dev_ans;
    return forward_range(m, b, 0, b->N, in, dev_pred, dev_h_true, dev_match, dbg, (cudaStream_t)stream);
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
#define QC2(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail(QMANN_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } while (0)
    if (!m->e2e_compute) QC2(cudaStreamCreateWithFlags(&m->e2e_compute, cudaStreamNonBlocking));
    if (!m->e2e_copy) QC2(cudaStreamCreateWithFlags(&m->e2e_copy, cudaStreamNonBlocking));
    qmann_batch *b = nullptr;
    int rc = host_batch_prepare(m, n_sen, N, m->e2e_copy, &b);
    if (rc) return rc;
    if (b->max_sen > m->cfg.S_max) return fail(QMANN_E_ARG, "a story has more sentences than S_max");
    const size_t V = m->cfg.V;
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
    QC2(cudaMemsetAsync(m->dev_err, 0, sizeof(unsigned), sc));          // every call starts clean (a failed batch does not poison the next)
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
        FwdInput in;
        in.dev_m = dm; in.dev_q = dq; in.dev_a = da;
        rc = forward_range(m, b, s0, n, in, m->e2e_pred, dh, da ? m->e2e_match : nullptr, nullptr, sc);
        if (rc) { cudaStreamSynchronize(sx); cudaStreamSynchronize(sc); return rc; }      // nothing of this call stays in flight
    }
    QC2(cudaMemcpyAsync(pred_host, m->e2e_pred, (size_t)N * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
    uint32_t mt = 0;
    QC2(cudaMemcpyAsync(&mt, m->e2e_match, sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
    std::vector<float> ht;
    if (dh) { ht.resize(N); QC2(cudaMemcpyAsync(ht.data(), dh, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, sc)); }
    unsigned err = 0;
    QC2(cudaMemcpyAsync(&err, m->dev_err, sizeof(unsigned), cudaMemcpyDeviceToHost, sc));
    QC2(cudaStreamSynchronize(sc));
    if (err) QC2(cudaMemset(m->dev_err, 0, sizeof(unsigned)));
#undef QC2
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

int qmann_infer_ids_host(qmann_model *m, const uint16_t *ids_host, const uint32_t *row_off_host, const uint32_t *ans_host,
                         const uint32_t *n_sen, uint32_t N, uint32_t *pred_host, uint32_t *match, float *cost)
{
    if (!m || !ids_host || !row_off_host || !n_sen || !pred_host) return fail(QMANN_E_ARG, "null argument");
    if (m->cfg.V > 65535u) return fail(QMANN_E_ARG, "ids are 16-bit");
#define QC2(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail(QMANN_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } while (0)
    if (!m->e2e_compute) QC2(cudaStreamCreateWithFlags(&m->e2e_compute, cudaStreamNonBlocking));
    if (!m->e2e_copy) QC2(cudaStreamCreateWithFlags(&m->e2e_copy, cudaStreamNonBlocking));
    qmann_batch *b = nullptr;
    int rc = host_batch_prepare(m, n_sen, N, m->e2e_copy, &b);
    if (rc) return rc;
    if (b->max_sen > m->cfg.S_max) return fail(QMANN_E_ARG, "a story has more sentences than S_max");
    const size_t R = (size_t)N + b->sum_sen;                 // rows: one question and n_sen sentences per story
    const size_t n_ids = row_off_host[R];
    if (n_ids > m->ids_cap) {
        cudaFree(m->ids_dev); m->ids_dev = nullptr; m->ids_cap = 0;
        QC2(cudaMalloc((void **)&m->ids_dev, std::max<size_t>(1, n_ids) * sizeof(uint16_t)));
        m->ids_cap = n_ids;
    }
    if (R + 1 > m->rows_cap) {
        cudaFree(m->rowoff_dev); m->rowoff_dev = nullptr; m->rows_cap = 0;
        QC2(cudaMalloc((void **)&m->rowoff_dev, (R + 1) * sizeof(uint32_t)));
        m->rows_cap = R + 1;
    }
    if (N > m->e2e_n_cap2) {
        cudaFree(m->ans_dev); cudaFree(m->e2e_h2); cudaFree(m->e2e_pred2);
        m->ans_dev = nullptr; m->e2e_h2 = nullptr; m->e2e_pred2 = nullptr; m->e2e_n_cap2 = 0;
        QC2(cudaMalloc((void **)&m->ans_dev, (size_t)N * sizeof(uint32_t)));
        QC2(cudaMalloc((void **)&m->e2e_h2, (size_t)N * sizeof(float)));
        QC2(cudaMalloc((void **)&m->e2e_pred2, (size_t)N * sizeof(uint32_t)));
        m->e2e_n_cap2 = N;
    }
    if (!m->e2e_match) QC2(cudaMalloc((void **)&m->e2e_match, sizeof(uint32_t)));
    if (!m->e2e_compute) QC2(cudaStreamCreateWithFlags(&m->e2e_compute, cudaStreamNonBlocking));
    if (!m->e2e_copy) QC2(cudaStreamCreateWithFlags(&m->e2e_copy, cudaStreamNonBlocking));
    cudaStream_t sc = m->e2e_compute, sx = m->e2e_copy;
    float *dh = (ans_host && cost) ? m->e2e_h2 : nullptr;
    QC2(cudaMemsetAsync(m->e2e_match, 0, sizeof(uint32_t), sc));
    QC2(cudaMemsetAsync(m->dev_err, 0, sizeof(unsigned), sc));
    // same pipeline as qmann_infer_host: the copy stream runs ahead chunk by chunk
    const uint32_t CH = 8192;
    size_t ev_i = 0;
    for (uint32_t s0 = 0; s0 < N; s0 += CH) {
        const uint32_t n = std::min<uint32_t>(CH, N - s0);
        const size_t r0 = b->sen_off[s0] + s0, r1 = b->sen_off[s0 + n] + s0 + n;          // rows of this chunk
        const size_t i0 = row_off_host[r0], i1 = row_off_host[r1];
        if (i1 > i0) QC2(cudaMemcpyAsync(m->ids_dev + i0, ids_host + i0, (i1 - i0) * sizeof(uint16_t), cudaMemcpyHostToDevice, sx));
        QC2(cudaMemcpyAsync(m->rowoff_dev + r0, row_off_host + r0, (r1 - r0 + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, sx));
        if (ans_host) QC2(cudaMemcpyAsync(m->ans_dev + s0, ans_host + s0, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, sx));
        if (ev_i >= m->e2e_events.size()) {
            cudaEvent_t ev;
            QC2(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            m->e2e_events.push_back(ev);
        }
        cudaEvent_t ev = m->e2e_events[ev_i++];
        QC2(cudaEventRecord(ev, sx));
        QC2(cudaStreamWaitEvent(sc, ev, 0));
        FwdInput in;
        in.dev_ids = m->ids_dev; in.dev_row_off = m->rowoff_dev; in.dev_ans = ans_host ? m->ans_dev : nullptr;
        rc = forward_range(m, b, s0, n, in, m->e2e_pred2, dh, ans_host ? m->e2e_match : nullptr, nullptr, sc);
        if (rc) { cudaStreamSynchronize(sx); cudaStreamSynchronize(sc); return rc; }
    }
    QC2(cudaMemcpyAsync(pred_host, m->e2e_pred2, (size_t)N * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
    uint32_t mt = 0;
    QC2(cudaMemcpyAsync(&mt, m->e2e_match, sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
    std::vector<float> ht;
    if (dh) { ht.resize(N); QC2(cudaMemcpyAsync(ht.data(), dh, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, sc)); }
    unsigned err = 0;
    QC2(cudaMemcpyAsync(&err, m->dev_err, sizeof(unsigned), cudaMemcpyDeviceToHost, sc));
    QC2(cudaStreamSynchronize(sc));
    if (err) QC2(cudaMemset(m->dev_err, 0, sizeof(unsigned)));
#undef QC2
    if (match) *match = mt;
    if (cost && dh) {
        float cacc = *cost;
        for (uint32_t i = 0; i < N; i++) cacc = (float)((double)cacc + -1.0 * (double)ht[i]);
        *cost = cacc;
    }
    if (err) return fail(QMANN_E_ARG, "a story holds an id >= V or overflowed the compaction heap");
    return QMANN_OK;
}

int qmann_profile_enable(qmann_model *m, int enable)
{
    if (!m) return fail(QMANN_E_ARG, "null model");
    m->profile = enable != 0;
    m->prof_used = 0;
    m->prof_acc_c = m->prof_acc_f = 0.0;
    m->prof_pairs = 0;
    return QMANN_OK;
}

int qmann_profile_read(qmann_model *m, float *ms_compact, float *ms_forward, uint32_t *n_pairs)
{
    if (!m) return fail(QMANN_E_ARG, "null model");
    const int rc = profile_fold(m);
    if (rc) return rc;
    if (ms_compact) *ms_compact = (float)m->prof_acc_c;
    if (ms_forward) *ms_forward = (float)m->prof_acc_f;
    if (n_pairs) *n_pairs = (uint32_t)m->prof_pairs;
    m->prof_acc_c = m->prof_acc_f = 0.0;
    m->prof_pairs = 0;
    return QMANN_OK;
}

int qmann_check_errors(qmann_model *m, void *stream, uint32_t *flags)
{
    if (!m) return fail(QMANN_E_ARG, "null model");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned err = 0;
    QCUDA(cudaMemcpyAsync(&err, m->dev_err, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    QCUDA(cudaStreamSynchronize(st));
    if (err) QCUDA(cudaMemset(m->dev_err, 0, sizeof(unsigned)));
    if (flags) *flags = err;
    return QMANN_OK;
}

int qmann_path_counts(qmann_model *m, void *stream, uint64_t tiers[3])
{
    if (!m || !tiers) return fail(QMANN_E_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long h[4] = {0, 0, 0, 0};
    QCUDA(cudaMemcpyAsync(h, m->dev_path_count, sizeof(h), cudaMemcpyDeviceToHost, st));
    QCUDA(cudaMemsetAsync(m->dev_path_count, 0, sizeof(h), st));
    QCUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 3; i++) tiers[i] = h[i];
    return QMANN_OK;
}

}  // extern "C"

// umma_probe.cu -- hardware check of the primitives in q-mann_b200/csrc/qmann_tc.cuh before they carry the production
// kernel: TMA 2-D boxes with 128-byte swizzle -> tcgen05.mma kind::tf32 (M=128, N=160, both operands K-major in shared
// memory) -> accumulators in tensor memory -> tcgen05.ld, plus a tcgen05.st/ld round trip.  Exactness of the tf32 product
// on bag-of-words counts x int8 codes is checked against an integer reference on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I q-mann_b200/csrc -o umma_probe profiles/tools/umma_probe.cu
#include "qmann_tc.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace qtc;

constexpr unsigned V = 256, NT = 160, S = 50, NSTORY = 4, KCH = V / 32, NSTAGE = 3;
constexpr unsigned STAGE_BYTES = 128 * 128, TAB_CH_BYTES = NT * 128;

struct Params {
    alignas(64) CUtensorMap tmX;
    alignas(64) CUtensorMap tmT;
    float *out;        // [NSTORY][64][NT]
    unsigned *out2;    // [128][8] st/ld round trip
};

extern __shared__ unsigned char smem_raw[];

__global__ void __launch_bounds__(192, 1) k_probe(const __grid_constant__ Params p)
{
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const unsigned tabs = base, stages = tabs + KCH * TAB_CH_BYTES, bars = stages + NSTAGE * STAGE_BYTES;
    const unsigned bar_tab = bars, bar_full = bars + 8, bar_empty = bar_full + 8 * NSTAGE, bar_done = bar_empty + 8 * NSTAGE, tmem_slot = bar_done + 8;
    if (threadIdx.x == 0) {
        mbar_init(bar_tab, 1);
        for (unsigned s = 0; s < NSTAGE; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    unsigned tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

    if (warp == 4 && lane == 0) {
        // producer: tables once, then 2 tiles x 8 K-chunks of 4 boxes (32 rows of each story)
        mbar_expect_tx(bar_tab, KCH * TAB_CH_BYTES);
        for (unsigned k = 0; k < KCH; k++) tma_load_2d(tabs + k * TAB_CH_BYTES, &p.tmT, (int)(32 * k), 0, bar_tab);
        unsigned it = 0;
        for (unsigned t = 0; t < 2; t++)
            for (unsigned k = 0; k < KCH; k++, it++) {
                const unsigned s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
                for (unsigned j = 0; j < NSTORY; j++)
                    tma_load_2d(stages + s * STAGE_BYTES + j * 4096u, &p.tmX, (int)(32 * k), (int)(j * S + 32 * t), bar_full + 8 * s);
            }
    } else if (warp == 5 && lane == 0) {
        const unsigned idesc = umma_idesc_tf32(128, NT);
        mbar_wait(bar_tab, 0);
        unsigned it = 0;
        for (unsigned t = 0; t < 2; t++) {
            for (unsigned k = 0; k < KCH; k++, it++) {
                const unsigned s = it % NSTAGE, ph = (it / NSTAGE) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                for (unsigned j = 0; j < 4; j++)
                    umma_tf32(tmem + t * NT, umma_desc_sw128(stages + s * STAGE_BYTES + 32u * j), umma_desc_sw128(tabs + k * TAB_CH_BYTES + 32u * j), idesc,
                              (k | j) ? 1u : 0u);
                umma_commit(bar_empty + 8 * s);
            }
        }
        umma_commit(bar_done);
    } else if (warp < 4) {
        mbar_wait(bar_done, 0);
        tc_fence_after();
        const unsigned tq = tmem + ((32u * warp) << 16);
        for (unsigned t = 0; t < 2; t++)
            for (unsigned c0 = 0; c0 < NT; c0 += 16) {
                unsigned v[16];
                tmem_ld16(tq + t * NT + c0, v);
                tmem_wait_ld();
                for (int i = 0; i < 16; i++) p.out[((size_t)warp * 64 + 32 * t + lane) * NT + c0 + i] = __uint_as_float(v[i]);
            }
        // st / ld round trip in the columns behind the accumulators
        unsigned w4[4] = {0xA0000000u + threadIdx.x, 0xB0000000u + threadIdx.x, 0xC0000000u + threadIdx.x, 0xD0000000u + threadIdx.x};
        tmem_st4(tq + 2 * NT, w4);
        tmem_st4(tq + 2 * NT + 4, w4);
        tmem_wait_st();
        unsigned r8[8];
        tmem_ld8(tq + 2 * NT, r8);
        tmem_wait_ld();
        for (int i = 0; i < 8; i++) p.out2[threadIdx.x * 8 + i] = r8[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main()
{
    const unsigned R = NSTORY * S;
    std::vector<float> X((size_t)R * V, 0.f), T((size_t)NT * V, 0.f);
    unsigned long long st = 12345;
    auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(st >> 33); };
    for (unsigned r = 0; r < R; r++) {
        for (int e = 0; e < 5; e++) X[(size_t)r * V + rnd() % 192] += 1.0f;
        X[(size_t)r * V + 192 + (r % 64)] = 1.0f;
    }
    X[7 * V + 3] = 3.0f;
    for (unsigned n = 0; n < 153; n++)
        for (unsigned k = 0; k < V; k++) T[(size_t)n * V + k] = (float)((int)(rnd() % 255) - 127);
    float *dX, *dT, *dout; unsigned *dout2;
    CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dT, T.size() * 4)); CK(cudaMalloc(&dout, (size_t)NSTORY * 64 * NT * 4)); CK(cudaMalloc(&dout2, 128 * 8 * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dT, T.data(), T.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xFF, (size_t)NSTORY * 64 * NT * 4));
    void *fn = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    EncodeFn enc = (EncodeFn)fn;
    Params p;
    {
        cuuint64_t dims[2] = {V, R}, strides[1] = {V * 4};
        cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
        CUresult r = enc(&p.tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode X failed %d\n", (int)r); return 2; }
        cuuint64_t dimsT[2] = {V, NT};
        cuuint32_t boxT[2] = {32, NT};
        r = enc(&p.tmT, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dT, dimsT, strides, boxT, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode T failed %d\n", (int)r); return 2; }
    }
    p.out = dout; p.out2 = dout2;
    const unsigned smem = KCH * TAB_CH_BYTES + NSTAGE * STAGE_BYTES + 1024 + 256;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<<<1, 192, smem>>>(p);
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)NSTORY * 64 * NT);
    std::vector<unsigned> out2(128 * 8);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out2.data(), dout2, out2.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (unsigned q = 0; q < NSTORY; q++)
        for (unsigned r = 0; r < 64; r++)
            for (unsigned n = 0; n < NT; n++) {
                const unsigned row = q * S + r;
                long long ref = 0;
                if (row < R)
                    for (unsigned k = 0; k < V; k++) ref += (long long)X[(size_t)row * V + k] * (long long)T[(size_t)n * V + k];
                const float got = out[((size_t)q * 64 + r) * NT + n];
                if (got != (float)ref) { if (bad < 10) printf("mismatch story %u row %u n %u: got %f want %lld\n", q, r, n, got, ref); bad++; }
            }
    size_t bad2 = 0;
    for (unsigned t = 0; t < 128; t++)
        for (int i = 0; i < 8; i++) {
            const unsigned want = (0xA0000000u + 0x10000000u * (i & 3)) + t;
            if (out2[t * 8 + i] != want) { if (bad2 < 5) printf("st/ld mismatch thread %u word %d: %08x want %08x\n", t, i, out2[t * 8 + i], want); bad2++; }
        }
    printf("UMMA_PROBE mma mismatches %zu of %zu, st/ld mismatches %zu\n", bad, out.size(), bad2);
    return (bad || bad2) ? 1 : 0;
}

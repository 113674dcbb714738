// umma_ts_probe.cu -- hardware check of tcgen05.mma kind::i8 with the A operand in TENSOR MEMORY (written with tcgen05.st) and
// the B operand K-major in shared memory (128-byte swizzle), as k_big_scores_tq (q-mann_b200/csrc/qmann_bigmem.cu) uses it:
// D[128 queries][128 slots] (int32, TMEM) = A[128 queries][K] (TMEM, lane = row, four 8-bit K elements per 32-bit column)
//                                           * B[128 slots][K]^T (shared memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I q-mann_b200/csrc -o umma_ts_probe profiles/tools/umma_ts_probe.cu
#include "qmann_tc.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace qtc;

constexpr unsigned KB = 128;          // K bytes per row (one swizzle atom), 4 MMAs of K = 32

__host__ __device__ constexpr unsigned idesc_i8(unsigned M, unsigned N) { return (2u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }

__device__ __forceinline__ void umma_i8_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long b_desc, unsigned idesc, unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

extern __shared__ unsigned char smem_raw[];

__global__ void __launch_bounds__(160, 1) k_probe(const signed char *A, const signed char *B, int *out)
{
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned sraw = smem_u32(smem_raw), base = (sraw + 1023u) & ~1023u;
    unsigned char *g = smem_raw + (base - sraw);
    const unsigned btile = base, bar = base + 128 * KB, slot = bar + 8;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 4) tmem_alloc(slot, 512);
    // B tile: row r (slot) at r * 128, its 16-byte chunk c at ((c ^ (r & 7)) * 16)
    for (unsigned i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const unsigned r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4 *>(g + r * 128 + ((c ^ (r & 7)) * 16)) = *reinterpret_cast<const uint4 *>(B + r * KB + c * 16);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    unsigned tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (warp < 4) {
        // A: row = 32 warp + lane, 32 columns of four K bytes each at TMEM columns 256 ..
        const unsigned row = 32 * warp + lane;
        const unsigned tq = tmem + ((32u * warp) << 16);
        for (unsigned c = 0; c < KB / 4; c += 4) {
            const uint4 v = *reinterpret_cast<const uint4 *>(A + row * KB + 4 * c);
            const unsigned w4[4] = {v.x, v.y, v.z, v.w};
            tmem_st4(tq + 256 + c, w4);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4 && lane == 0) {
        const unsigned idesc = idesc_i8(128, 128);
        for (unsigned j = 0; j < 4; j++) umma_i8_ts(tmem, tmem + 256 + 8 * j, umma_desc_sw128(btile) + 2ull * j, idesc, j ? 1u : 0u);
        umma_commit(bar);
    }
    if (warp < 4) {
        mbar_wait(bar, 0);
        tc_fence_after();
        const unsigned tq = tmem + ((32u * warp) << 16);
        for (unsigned c0 = 0; c0 < 128; c0 += 16) {
            unsigned v[16];
            tmem_ld16(tq + c0, v);
            tmem_wait_ld();
            for (int i = 0; i < 16; i++) out[(32 * warp + lane) * 128 + c0 + i] = (int)v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4) tmem_dealloc(tmem, 512);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main()
{
    std::vector<signed char> A(128 * KB), B(128 * KB);
    unsigned long long st = 4242;
    auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(st >> 33); };
    for (auto &x : A) x = (signed char)((int)(rnd() % 255) - 127);
    for (auto &x : B) x = (signed char)((int)(rnd() % 255) - 127);
    signed char *dA, *dB; int *dout;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dout, 128 * 128 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0x7F, 128 * 128 * 4));
    const unsigned smem = 128 * KB + 1024 + 64;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_probe<<<1, 160, smem>>>(dA, dB, dout);
    CK(cudaDeviceSynchronize());
    std::vector<int> out(128 * 128);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (unsigned q = 0; q < 128; q++)
        for (unsigned s = 0; s < 128; s++) {
            int ref = 0;
            for (unsigned k = 0; k < KB; k++) ref += (int)A[q * KB + k] * (int)B[s * KB + k];
            if (out[q * 128 + s] != ref) { if (bad < 10) printf("mismatch q %u s %u: got %d want %d\n", q, s, out[q * 128 + s], ref); bad++; }
        }
    printf("UMMA_TS_PROBE mismatches %zu of %zu\n", bad, out.size());
    return bad ? 1 : 0;
}

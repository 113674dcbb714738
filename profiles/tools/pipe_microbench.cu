// pipe_microbench.cu -- issue rate (warp instructions / clock / SM) of the integer / packed ops the
// forward kernels are built from, measured on the GPU it runs on.  Development tool, not product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench pipe_microbench.cu && ./pipe_microbench
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define OPS(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
enum { IMAD, IMADHI, IDP4A, VIADD2, VIADDMNMX, VIADDMNMX2, LOP3, SHFR, PRMT, HFMA2, HMNMX2, HADD2F32, FFMARZ, I2FP, IADD3, FMNMX, FADD, VIMNMX2, MIX_ALU_FMA, LDS32, LDS128, NOPS };
static const char *names[] = {"IMAD", "IMAD.HI", "IDP.4A", "VIADD.16x2", "VIADDMNMX.RELU", "VIADDMNMX.S16x2.RELU", "LOP3", "SHF.R.S32", "PRMT", "HFMA2",
                              "HMNMX2", "HADD2.F32", "FFMA.RZ", "I2FP.F32.S32", "IADD3", "FMNMX", "FADD", "VIMNMX.S16x2", "IMAD+LOP3 alternating", "LDS.32", "LDS.128"};

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(unsigned *out, unsigned long long *cyc, int iters, unsigned seed)
{
    __shared__ uint4 sm[1024];
    sm[threadIdx.x] = make_uint4(threadIdx.x, seed, 3, 4);
    __syncthreads();
    unsigned r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = seed * (i + 1) + threadIdx.x;
    unsigned b = seed | 1u, c = seed ^ 0x1234u;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#define ONE(i)                                                                                                     \
    if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));                      \
    if (OP == IMADHI) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));                                \
    if (OP == IDP4A) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(c));                   \
    if (OP == VIADD2) r[i] = __vadd2(r[i], b);                                                                     \
    if (OP == VIADDMNMX) r[i] = __viaddmin_s32_relu((int)r[i], (int)b, (int)c);                                    \
    if (OP == VIADDMNMX2) r[i] = __viaddmin_s16x2_relu(r[i], b, c);                                                \
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(c));                  \
    if (OP == SHFR) asm volatile("shr.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(b & 1u));                                \
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c & 0x7777u));              \
    if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));                   \
    if (OP == HMNMX2) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b));                                 \
    if (OP == HADD2F32) { float f; asm volatile("{.reg .f16 lo, hi; mov.b32 {lo,hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f) : "r"(r[i])); r[i] = __float_as_uint(f) + 1; } \
    if (OP == FFMARZ) asm volatile("fma.rz.f32 %0, %0, %1, %2;" : "+f"(*(float *)&r[i]) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c))); \
    if (OP == I2FP) { float f; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(r[i])); r[i] = __float_as_uint(f); } \
    if (OP == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));                                    \
    if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(*(float *)&r[i]) : "f"(__uint_as_float(b)));       \
    if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float *)&r[i]) : "f"(__uint_as_float(b)));     \
    if (OP == VIMNMX2) r[i] = __vmaxs2(r[i], b);                                                                   \
    if (OP == MIX_ALU_FMA) { if (i & 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(c)); } \
    if (OP == LDS32) r[i] += ((unsigned *)sm)[(r[i] * 4 + i) & 4095];                                              \
    if (OP == LDS128) { uint4 v = sm[(r[i] + i) & 1023]; r[i] += v.x + v.w; }
        OPS(ONE)
#undef ONE
    }
    const unsigned long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

#define EMIT(OPX, i)                                                                                            \
    if (OPX == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));                      \
    if (OPX == IDP4A) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(c));                   \
    if (OPX == VIADD2) r[i] = __vadd2(r[i], b);                                                                     \
    if (OPX == VIADDMNMX) r[i] = __viaddmin_s32_relu((int)r[i], (int)b, (int)c);                                    \
    if (OPX == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(c));                  \
    if (OPX == SHFR) asm volatile("shr.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(sh));                                    \
    if (OPX == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c & 0x7777u));              \
    if (OPX == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));                   \
    if (OPX == HMNMX2) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b));                                 \
    if (OPX == FFMARZ) asm volatile("fma.rz.f32 %0, %0, %1, %2;" : "+f"(*(float *)&r[i]) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c))); \
    if (OPX == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));                                    \
    if (OPX == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(*(float *)&r[i]) : "f"(__uint_as_float(b)));       \
    if (OPX == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float *)&r[i]) : "f"(__uint_as_float(b)));     \
    if (OPX == VIMNMX2) r[i] = __vmaxs2(r[i], b);                                                                   \
    if (OPX == I2FP) { float f; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(r[i])); r[i] = __float_as_uint(f); }

template <int A, int B>
__global__ void __launch_bounds__(1024, 1) kpair(unsigned *out, unsigned long long *cyc, int iters, unsigned seed, unsigned sh)
{
    unsigned r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = seed * (i + 1) + threadIdx.x;
    unsigned b = seed | 1u, c = seed ^ 0x1234u;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        EMIT(A, 0) EMIT(B, 1) EMIT(A, 2) EMIT(B, 3) EMIT(A, 4) EMIT(B, 5) EMIT(A, 6) EMIT(B, 7)
    }
    const unsigned long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int A, int B>
void runpair(unsigned *out, unsigned long long *cyc, int sms)
{
    const int iters = 4096;
    kpair<A, B><<<sms, 1024>>>(out, cyc, iters, 12345u, 1u);
    kpair<A, B><<<sms, 1024>>>(out, cyc, iters, 12345u, 1u);
    cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)h[i];
    avg /= sms;
    const double winst = 32.0 * iters * 8;
    printf("%-22s + %-22s %7.3f warp-inst/clk/SM\n", names[A], names[B], winst / avg);
}

template <int OP>
void run(unsigned *out, unsigned long long *cyc, int sms)
{
    const int iters = 4096;
    k<OP><<<sms, 1024>>>(out, cyc, iters, 12345u);
    k<OP><<<sms, 1024>>>(out, cyc, iters, 12345u);
    cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)h[i];
    avg /= sms;
    const double winst = 32.0 * iters * 8;   // warp-level target ops per SM (32 warps)
    printf("%-24s %7.3f warp-inst/clk/SM  (%.2f clk per warp-inst per SMSP)\n", names[OP], winst / avg, avg / (winst / 4));
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    unsigned *out;
    unsigned long long *cyc;
    cudaMalloc(&out, sms * 1024 * 4);
    cudaMalloc(&cyc, sms * 8);
    printf("%s, %d SMs; 32 warps/SM, 8 independent chains per thread (loop overhead included)\n", p.name, sms);
    run<IMAD>(out, cyc, sms); run<IMADHI>(out, cyc, sms); run<IDP4A>(out, cyc, sms); run<VIADD2>(out, cyc, sms);
    run<VIADDMNMX>(out, cyc, sms); run<VIADDMNMX2>(out, cyc, sms); run<LOP3>(out, cyc, sms); run<SHFR>(out, cyc, sms);
    run<PRMT>(out, cyc, sms); run<HFMA2>(out, cyc, sms); run<HMNMX2>(out, cyc, sms); run<HADD2F32>(out, cyc, sms);
    run<FFMARZ>(out, cyc, sms); run<I2FP>(out, cyc, sms); run<IADD3>(out, cyc, sms); run<FMNMX>(out, cyc, sms);
    run<FADD>(out, cyc, sms); run<VIMNMX2>(out, cyc, sms); run<MIX_ALU_FMA>(out, cyc, sms); run<LDS32>(out, cyc, sms); run<LDS128>(out, cyc, sms);
    printf("pairs (4+4 alternating independent chains): ~3.7 = different pipes, ~2 = same half-rate pipe\n");
    runpair<SHFR, SHFR>(out, cyc, sms); runpair<IMAD, SHFR>(out, cyc, sms); runpair<LOP3, SHFR>(out, cyc, sms); runpair<IMAD, VIADDMNMX>(out, cyc, sms);
    runpair<LOP3, VIADDMNMX>(out, cyc, sms); runpair<IMAD, IDP4A>(out, cyc, sms); runpair<LOP3, IDP4A>(out, cyc, sms); runpair<IADD3, IMAD>(out, cyc, sms);
    runpair<IADD3, LOP3>(out, cyc, sms); runpair<FMNMX, LOP3>(out, cyc, sms); runpair<FMNMX, IMAD>(out, cyc, sms); runpair<FADD, IMAD>(out, cyc, sms);
    runpair<FADD, LOP3>(out, cyc, sms); runpair<VIMNMX2, LOP3>(out, cyc, sms); runpair<VIMNMX2, IMAD>(out, cyc, sms); runpair<HFMA2, IMAD>(out, cyc, sms);
    runpair<HFMA2, LOP3>(out, cyc, sms); runpair<PRMT, IMAD>(out, cyc, sms); runpair<VIADD2, IMAD>(out, cyc, sms); runpair<I2FP, IMAD>(out, cyc, sms);
    runpair<I2FP, LOP3>(out, cyc, sms); runpair<FFMARZ, IMAD>(out, cyc, sms); runpair<FFMARZ, FADD>(out, cyc, sms); runpair<IADD3, FADD>(out, cyc, sms);
    runpair<IADD3, VIMNMX2>(out, cyc, sms); runpair<HMNMX2, LOP3>(out, cyc, sms); runpair<HMNMX2, IMAD>(out, cyc, sms);
    return 0;
}

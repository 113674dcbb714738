// qmann_fixed.cuh -- fixed-point arithmetic of the Q-MANN forward as sm_100a device code.
//
// Two families:
//  * lit_*  : the reference macro semantics on fp32 tensors, value for value
//             (CUDA_FLOAT2FIXED / FIXED2FLOAT / FLOAT_QUANT / FIXED_MUL / FIXED_ADD,
//             reference lib/layer_cuda.h:207-259).  Used by the per-layer cuda_* shim, whose
//             inputs are arbitrary fp32 device tensors.
//  * qi_*   : the same arithmetic on integer codes (proven equal to the
//             literal forms on every 8-bit operand pair by tests/test_oracle_golden.py against
//             tests/golden/kat_fixed_mul.npz).  Used by the fused batched kernels, which keep
//             every activation as an 8-bit code.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace qmann {

struct Fmt {
    int iwl, frac;
};

// ----------------------------------------------------------------------------------------------
// literal (fp32 tensor) forms
// ----------------------------------------------------------------------------------------------

// CUDA_FIXED_MAX_FIXED(iwl,frac) = (int)((unsigned)(1<<(iwl+frac))-1)           layer_cuda.h:207
__host__ __device__ __forceinline__ int fixed_max(int iwl, int frac) { return (int)((1u << ((iwl + frac) & 31)) - 1u); }

// CUDA_FLOAT2FIXED: sign-magnitude code word, truncation toward zero, saturation, and the
// "negative zero" (sign bit with zero magnitude) the reference produces for -2^-frac < x < 0.
// `x` is a double because FLOAT_QUANT of a FLOAT_QUANT product/sum is formed in double
// (CUDA_FLOAT_QUANT is a ternary with a double literal).                        layer_cuda.h:233,246
__device__ __forceinline__ unsigned lit_float2fixed(double x, int iwl, int frac)
{
    const int lim = fixed_max(iwl, frac);
    const float one = (float)(int)(1u << (frac & 31));
    const float maxf = (float)(unsigned)lim / one;       // the reference converts MAX_FIXED as unsigned (I2FP.F32.U32)
    const float minf = -maxf;
    int n;
    if (x > (double)maxf) n = lim;
    else if (x < (double)minf) n = -lim;
    else n = __double2int_rz(x * (double)(int)(1u << (frac & 31)));   // cvt.rzi: saturating, NaN -> 0
    if (x >= 0.0) return (unsigned)n;
    return ((unsigned)(~n) + 1u) | 0x80000000u;
}

// CUDA_FIXED2FLOAT                                                               layer_cuda.h:247
__device__ __forceinline__ float lit_fixed2float(unsigned code, int frac)
{
    const float one = (float)(int)(1u << (frac & 31));
    if ((code & 0x80000000u) == 0u) return (float)code / one;
    const int v = (int)(~(code & 0x7FFFFFFFu) + 1u);
    return (float)v / one;
}

// CUDA_FLOAT_QUANT, including the binary (iwl+frac==0) branch                   layer_cuda.h:253
__device__ __forceinline__ double lit_quant(double x, int iwl, int frac)
{
    if (iwl + frac == 0) return (x >= 0.0) ? 1.0 : -1.0;
    return (double)lit_fixed2float(lit_float2fixed(x, iwl, frac), frac);
}

// CUDA_FIXED_MUL / CUDA_FIXED_ADD                                                layer_cuda.h:257-258
__device__ __forceinline__ float lit_fixed_mul(float a, float b, int iwl_a, int frac_a, int iwl_b, int frac_b)
{
    return (float)lit_quant(lit_quant((double)a, iwl_a, frac_a) * lit_quant((double)b, iwl_b, frac_b), iwl_a, frac_a);
}
__device__ __forceinline__ float lit_fixed_add(float a, float b, int iwl_a, int frac_a, int iwl_b, int frac_b)
{
    return (float)lit_quant(lit_quant((double)a, iwl_a, frac_a) + lit_quant((double)b, iwl_b, frac_b), iwl_a, frac_a);
}

// ----------------------------------------------------------------------------------------------
// integer-code forms (8-bit word length: iwl + frac <= 7, |code| <= 127)
// ----------------------------------------------------------------------------------------------

__device__ __forceinline__ int qi_clamp(int v, int lim) { return max(-lim, min(v, lim)); }

// v / 2^sh rounded toward zero (the reference truncates through (int)(float))
__device__ __forceinline__ int qi_shr0(int v, int sh)
{
    return (v + ((v >> 31) & ((1 << sh) - 1))) >> sh;
}

// fp32 value -> signed code of format (iwl,frac)
__device__ __forceinline__ int qi_encode(float x, int iwl, int frac)
{
    const int lim = fixed_max(iwl, frac);
    const float scale = (float)(1 << frac);
    const float maxf = (float)lim / scale;
    if (x > maxf) return lim;
    if (x < -maxf) return -lim;
    return __float2int_rz(x * scale);
}

// code on a 2^-frac_from grid -> code of format (iwl_to, frac_to): shift (toward zero) then saturate
__device__ __forceinline__ int qi_requant(int n, int frac_from, int lim_to, int frac_to)
{
    const int sh = frac_to - frac_from;
    const int v = (sh >= 0) ? (n << sh) : qi_shr0(n, -sh);
    return qi_clamp(v, lim_to);
}

// FIXED_MUL on codes: a in format A (limit lim_a), b with frac_b fractional bits; result in format A
__device__ __forceinline__ int qi_mul(int a, int b, int lim_a, int frac_b)
{
    return qi_clamp(qi_shr0(a * b, frac_b), lim_a);
}

// One element of the approximate (Hamming) attention, reference lib/layer_cuda.cu:384-428 and
// :218-326 with num_bit = 8: returns e*128 where e = +-(127 - D7)/128 is the weighted bit-match
// similarity.  am/av are the 31-bit magnitudes (|x| * 2^(31-iwl), saturated to 0x7FFFFFFF, with
// the -2^iwl -> 0 quirk applied by the caller), sm/sv the sign bits (0 or 0x80000000).
// The opposite-sign branch uses the three-input ADD nvcc emits for `sign|(abs+min)`
// (signed-overflow UB in the source; see oracle/qmann_oracle.c and tests/golden/kat_appx_element.npz).
__device__ __forceinline__ int appx_element_x128(unsigned sm, unsigned am, unsigned sv, unsigned av)
{
    unsigned fm, fv;
    const unsigned amin = min(am, av);
    if (sm == sv) {
        fm = sm | (am - amin);
        fv = sv | (av - amin);
    } else if (am >= av) {
        fm = sm + am + amin;
        fv = sv;
    } else {
        fm = sm;
        fv = sv + av + amin;
    }
    const unsigned x = fm ^ fv;
    const int d7 = (int)((x >> 24) & 0x7Fu);
    const int e = 127 - d7;
    return (x & 0x80000000u) ? -e : e;
}

// 31-bit sign-magnitude encode of a code n on a 2^-frac_n grid at format (iwl, 31-iwl), as
// CUDA_FLOAT2FIXED(x, iwl, 31-iwl) does for x = n / 2^frac_n:
//   |x| >  2^iwl            -> magnitude 0x7FFFFFFF
//   x  == +2^iwl            -> (int)(2^31) saturates to 0x7FFFFFFF
//   x  == -2^iwl            -> (int) gives INT_MIN, two's complement negate leaves magnitude 0
__device__ __forceinline__ void appx_encode(int n, int frac_n, int iwl, unsigned &sign, unsigned &mag)
{
    const int k = 31 - iwl - frac_n;                 // x * 2^(31-iwl) = n * 2^k
    const unsigned a = (unsigned)abs(n);
    sign = (n < 0) ? 0x80000000u : 0u;
    // a * 2^k compared with 2^31 without overflow: a < 2^(31-k)
    const unsigned long long v = (k >= 0) ? ((unsigned long long)a << k) : ((unsigned long long)a >> (-k));
    if (v > 0x80000000ull) mag = 0x7FFFFFFFu;
    else if (v == 0x80000000ull) mag = (n < 0) ? 0u : 0x7FFFFFFFu;
    else mag = (unsigned)v;
}

}  // namespace qmann

// ccgp_math.h -- small FP64 math helpers shared by every kernel (host+device).
//
// dexp_neg(s) = exp(-s) for s >= 0: the Gaussian correlation exp(-Dist) of
// corr.matrix ([A]:351-360 `return(exp(-Dist))`).  Cody-Waite reduction by ln2
// in two pieces, then exp(r) = 1 + r + r^2 q(r) with a degree-9 near-minimax q
// on |r| <= ln2/2 (approximation error 1.6e-17, fitted in 60-digit mpmath at
// Chebyshev nodes), scaled by 2^k through the exponent field.  No special-case
// paths: the argument is always finite and <= 0; results below 2^-1000 flush
// to zero (irrelevant next to the unit diagonal).  14 FMA-pipe ops + 3 integer.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define CCGP_HD __host__ __device__ __forceinline__
#else
#define CCGP_HD static inline
#endif

CCGP_HD double ccgp_scale2(double v, int k) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    u += ((uint64_t)(int64_t)k) << 52;
    memcpy(&v, &u, 8);
    return v;
#endif
}

#if defined(__CUDACC__)
// Coefficients live in constant memory so each Horner step is one DFMA with a constant-bank
// operand (64-bit immediates would cost two extra uniform moves per step).
static __constant__ double c_dexp[16] = {
    6755399441055744.0, 1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    2.5100424157005067e-08, 2.7620138719733994e-07, 2.7557268378684192e-06, 2.480152119021773e-05,
    0.00019841269863105968, 0.0013888888917281794, 0.008333333333330051, 0.04166666666662399,
    0.16666666666666669, 0.5000000000000001, 700.0, 0.0};

// device version, branch-free.  CLAMP=false: the caller guarantees s < 1e8 (checked once per
// candidate from the parameter row and the design's bounding box), so only the power-of-two
// exponent is clamped (one integer max); CLAMP=true clamps the argument itself at 700.
template <bool CLAMP>
__device__ __forceinline__ double dexp_neg_dev(double s) {
    if (CLAMP) s = fmin(s, c_dexp[14]);
    double t = fma(-s, c_dexp[1], c_dexp[0]);
    double kf = t - c_dexp[0];
    double r = fma(kf, c_dexp[2], -s);
    r = fma(kf, c_dexp[3], r);
    double q = c_dexp[4];
    q = fma(q, r, c_dexp[5]);
    q = fma(q, r, c_dexp[6]);
    q = fma(q, r, c_dexp[7]);
    q = fma(q, r, c_dexp[8]);
    q = fma(q, r, c_dexp[9]);
    q = fma(q, r, c_dexp[10]);
    q = fma(q, r, c_dexp[11]);
    q = fma(q, r, c_dexp[12]);
    q = fma(q, r, c_dexp[13]);
    double r2 = r * r;
    double e = fma(r2, q, r) + 1.0;
    // t = MAGIC + k exactly, so k sits in the low word of t (two's complement)
    int k = __double2loint(t);
    if (!CLAMP) k = max(k, -1010);
    return __hiloint2double(__double2hiint(e) + (k << 20), __double2loint(e));
}
#endif

CCGP_HD double dexp_neg(double s) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rint() by add/sub
    double t = fma(-s, 1.4426950408889634074, MAGIC);
    double kf = t - MAGIC;
    double r = fma(kf, -6.93147180369123816490e-01, -s);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double q = 2.5100424157005067e-08;
    q = fma(q, r, 2.7620138719733994e-07);
    q = fma(q, r, 2.7557268378684192e-06);
    q = fma(q, r, 2.480152119021773e-05);
    q = fma(q, r, 0.00019841269863105968);
    q = fma(q, r, 0.0013888888917281794);
    q = fma(q, r, 0.008333333333330051);
    q = fma(q, r, 0.04166666666662399);
    q = fma(q, r, 0.16666666666666669);
    q = fma(q, r, 0.5000000000000001);
    double r2 = r * r;
    double e = fma(r2, q, r) + 1.0;
    int k = (int)kf;
    if (s > 690.0) return 0.0;
    return ccgp_scale2(e, k);
}

// Double-precision tanh(x/2) and 2*atanh(y) of the sum-product check node (qkd_ldpc_algorithm.cpp:55-71) for the float64
// streaming kernels: branch-free, no slow paths, no table -- about half the instructions of the CUDA libm versions, whose
// range branches diverge inside a warp (the float64 SPA check-node kernel spent 6x the time its HBM traffic needs).
// Accuracy against glibc (itself 1-2 ulp): <= 4 ulp for tanh(x/2), <= 5 ulp for 2*atanh over (-1, 1), <= 1 ulp next to the
// pole and identical saturation (tanh -> exactly 1.0 from |x| = 38.25 on, atanh(+-1) = +-inf, NaN in -> NaN out) --
// tools/spa_f64/spa_f64_math_check.cpp, 2.5 * 10^6 arguments per range. The parity bar
// against the reference (glibc) is the one the libm versions were held to: >= 99.5 % equal iteration counts
// (tests/test_gpu_large.py). Plain C++ so that the same text runs on the host checker.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define QK_HD __host__ __device__ __forceinline__
#else
#define QK_HD inline
#endif

namespace qk {

// Polynomial coefficients: Taylor of (exp(r) - 1 - r) / r^2 and of atanh(s) / s. On the device they live in constant
// memory, so that the DFMA reads them as a constant-bank operand instead of a 64-bit immediate rebuilt in a uniform
// register before every use (24 UMOV per tanh in the first version).
#define QK_SPA_EXP_COEFS                                                                                          \
    {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,     \
     1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5}
#define QK_SPA_ATANH_COEFS                                                                                        \
    {1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0}
#if defined(__CUDACC__)
static __constant__ double kSpaExpC_dev[12] = QK_SPA_EXP_COEFS;
static __constant__ double kSpaAtanhC_dev[10] = QK_SPA_ATANH_COEFS;
#endif
static const double kSpaExpC_host[12] = QK_SPA_EXP_COEFS;
static const double kSpaAtanhC_host[10] = QK_SPA_ATANH_COEFS;
#if defined(__CUDA_ARCH__)
#define QK_SPA_EXPC kSpaExpC_dev
#define QK_SPA_ATANHC kSpaAtanhC_dev
#else
#define QK_SPA_EXPC kSpaExpC_host
#define QK_SPA_ATANHC kSpaAtanhC_host
#endif

QK_HD uint64_t f64_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    std::memcpy(&u, &x, 8);
    return u;
#endif
}
QK_HD double f64_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    std::memcpy(&x, &u, 8);
    return x;
#endif
}

// tanh(x / 2) = em / (em + 2), em = exp(|x|) - 1. exp: |x| = k ln2 + r, |r| <= ln2 / 2, exp(r) = 1 + r + r^2 q(r) (Taylor to
// r^13: 4e-18), em = 2^k (1 + p) - 1 with p = r + r^2 q -- for k == 0 simply p, so small arguments keep full relative
// accuracy. |x| is capped at 40: the quotient is 1 - 8.5e-18 there and rounds to exactly 1.0, as tanh does from 38.2 on.
QK_HD double spa_tanh_half_f64(double x) {
    const double z = fmin(fabs(x), 40.0);
    const double kf = rint(z * 1.4426950408889634074);
    const double r = fma(-kf, 1.90821492927058770002e-10, fma(-kf, 6.93147180369123816490e-01, z));
    double q = QK_SPA_EXPC[0];
    for (int i = 1; i < 12; ++i) q = fma(q, r, QK_SPA_EXPC[i]);
    const double p = fma(r * r, q, r);                                   // exp(r) - 1
    const double scale = f64_from_bits((uint64_t)((int64_t)kf + 1023) << 52);   // 2^k, 0 <= k <= 58
    const double em = (kf == 0.0) ? p : fma(scale, p, scale - 1.0);
    const double t = em / (em + 2.0);
    const double s = copysign(t, x);
    return (x != x) ? x : s;
}

// 2 atanh(y) = ln((1 + |y|) / (1 - |y|)) with the sign of y, |y| <= 1. |y| <= 0.17: the series 2 (y + y^3/3 + ...) in y
// itself. Otherwise a = 1 + |y|, b = 1 - |y| = 2^eb * mb (b is exact from 0.5 on): the mantissas are brought within a
// factor sqrt(2) of each other and ln(a / mb') = 2 atanh(s), s = (a - mb') / (a + mb'), |s| <= 0.1716 -- one division for
// quotient and logarithm together. y = +-1 gives +-inf, NaN gives NaN (0/0 of quirk Q3 upstream).
QK_HD double spa_two_atanh_f64(double y) {
    const double ay = fabs(y);
    const bool small = ay <= 0.17;
    const double a = 1.0 + ay, b = 1.0 - ay;
    const uint64_t bb = f64_bits(b);
    const int eb = (int)((bb >> 52) & 0x7ffu) - 1023;
    double mb = f64_from_bits((bb & 0x000fffffffffffffull) | 0x3ff0000000000000ull);   // [1, 2)
    int e = -eb;
    const bool up = a > mb * 1.4142135623730951, down = a * 1.4142135623730951 < mb;
    mb = up ? mb + mb : (down ? 0.5 * mb : mb);
    e += up ? 1 : (down ? -1 : 0);
    const double num = small ? ay : a - mb, den = small ? 1.0 : a + mb;
    const double s = num / den;
    const double ef = small ? 0.0 : (double)e;
    const double w = s * s;
    double p = QK_SPA_ATANHC[0];
    for (int i = 1; i < 10; ++i) p = fma(p, w, QK_SPA_ATANHC[i]);
    const double s2 = s + s;
    const double series = fma(s2 * w, p, s2);                            // 2 atanh(s)
    const double res = fma(ef, 6.93147180369123816490e-01, fma(ef, 1.90821492927058770002e-10, series));
    const double inf = f64_from_bits(0x7ff0000000000000ull);
    const double out = copysign((ay == 1.0) ? inf : res, y);
    return (ay <= 1.0) ? out : f64_from_bits(0x7ff8000000000000ull);   // NaN in, or |y| > 1: NaN
}

}  // namespace qk

// pmath.h -- portable, bit-reproducible single-precision elementary functions for the B200 wave_tracer hot path.
//
// Why this exists.  The reference evaluates sin / cos / exp / log / pow / atan2 / acos / tan / hypot through the host libm
// (m::sin ... in /root/reference/include/wt/math/common.hpp -> std::sin -> glibc).  CUDA's libm and glibc agree only to 1-2 ulp, and the
// wave-optical path is ill-conditioned in exactly those ulps: a propagation phase is k*L ~ 1e5..1e8 rad, so one ulp in a sampled
// direction moves an interference term by 1e-3..1e-2 (DESIGN.md "Arithmetic contract").  A GPU result could therefore never be
// compared with a CPU result tighter than ~1e-3 however faithful the code.  Here every function is computed in IEEE binary64 from
// +, -, *, /, sqrt, fma, rint and bit manipulation only -- operations that are correctly rounded on the host AND on the device -- in a
// fixed order, and rounded to binary32 once at the end.  Host and device therefore return the SAME BITS for the same argument
// (tests/test_pmath.py: host vs. device bit-exact), and because the internal error is ~1e-15 the result is the correctly rounded float in
// all but ~1e-7 of the cases (measured against binary64 references: > 99.99 %; never more than 1 ulp from glibc's faithful sinf / cosf /
// expf, identical to them for ~99 % of the arguments).
//
// No dependency on the rest of the code base: included by the device code (dmath.cuh); the CPU checker of the test suite includes it too.
// Must be compiled without implicit contraction (nvcc --fmad=false, gcc -ffp-contract=off); every fused operation is an explicit fma().
#ifndef WT_PMATH_H
#define WT_PMATH_H

#include <stdint.h>
#include <math.h>
#ifndef __CUDACC__
#include <string.h>
#endif

#if defined(__CUDACC__)
#define PM_HD __host__ __device__ inline
// the f32 entry points (sinf, expf, atan2f, ...): -DPM_API_NOINLINE keeps ONE copy of each per kernel instead of one per call site (code size:
// the elementary functions were a quarter of k_shade's 700 KB of SASS).  Same arithmetic either way.
#ifdef PM_API_NOINLINE
#define PM_API __host__ __device__ __noinline__
#else
#define PM_API __host__ __device__ inline
#endif
#else
#define PM_HD inline
#define PM_API inline
#endif

namespace pm {

PM_HD uint64_t d2u(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
PM_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
PM_HD uint32_t f2u(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
PM_HD double pm_inf() { return u2d(0x7ff0000000000000ull); }
PM_HD double pm_nan() { return u2d(0x7ff8000000000000ull); }
PM_HD double pow2i(int n) { return u2d((uint64_t)(n + 1023) << 52); }     // 2^n, -1022 <= n <= 1023

constexpr double kPio2Hi = 1.5707963267948966, kPio2Lo = 6.123233995736766e-17;     // pi/2 = hi + lo (tools/gen_pmath_tables.py)
constexpr double kTwoOverPi = 0.6366197723675814;
constexpr double kLn2Hi = 0.6931471805599453, kLn2Lo = 2.3190468138462996e-17, kLog2e = 1.4426950408889634;
constexpr double kPiD = 3.141592653589793, kPiLoD = 1.2246467991473532e-16;

// ---- argument reduction: x = n * pi/2 + r, |r| <= ~pi/4; returns n mod 4.  x is a float value held in a double.
// |x| < 2^40: two-constant Cody-Waite in binary64 with fma (n < 2^41: n*kPio2Hi is exact inside the fma, the neglected tail of pi/2
// contributes n * 2^-107).  |x| >= 2^40 (the float is m * 2^e, m < 2^24, 17 <= e <= 104): x mod pi/2 = (m * (2^e mod pi/2)) mod pi/2 from
// a table of residues (absolute error ~3e-9: arguments whose float spacing is >= 2^17 rad carry no phase information; the value is
// still deterministic and a true sine / cosine of a point within 3e-9 of the reduced argument).
PM_HD int rem_pio2(double x, double& r) {
    static const double kR[88] = {
        0.04210325344140148, 0.08420650688280296, 0.1684130137656059, 0.3368260275312118,
        0.6736520550624236, 1.3473041101248473, 1.1238118934547983, 0.6768274601146997,
        1.3536549202293995, 1.1365135136639022, 0.7022307005329076, 1.4044614010658152,
        1.238126475336734, 0.9054566238785713, 0.240116920962246, 0.480233841924492,
        0.960467683848984, 0.35013904090307124, 0.7002780818061425, 1.400556163612285,
        1.2303160004296732, 0.8898356740644496, 0.20887502133400254, 0.4177500426680051,
        0.8355000853360102, 0.10020384387712368, 0.20040768775424736, 0.4008153755084947,
        0.8016307510169894, 0.03246517523908234, 0.06493035047816469, 0.12986070095632937,
        0.25972140191265874, 0.5194428038253175, 1.038885607650635, 0.5069748885063734,
        1.0139497770127468, 0.45710322723059693, 0.9142064544611939, 0.2576165821274912,
        0.5152331642549824, 1.030466328509965, 0.49013633022503306, 0.9802726604500661,
        0.38974899410523567, 0.7794979882104713, 1.5589959764209427, 1.5471956260469888,
        1.523594925299081, 1.4763935238032653, 1.381990720811634, 1.1931851148283712,
        0.8155739028618457, 0.060351478928794756, 0.12070295785758951, 0.24140591571517903,
        0.48281183143035805, 0.9656236628607161, 0.36045099892653565, 0.7209019978530713,
        1.4418039957061426, 1.3128116646173884, 1.0548270024398803, 0.5388576780848638,
        1.0777153561697277, 0.5846343855445587, 1.1692687710891174, 0.7677412153833382,
        1.5354824307666763, 1.500168534738456, 1.4295407426820153, 1.288285158569134,
        1.0057739903433711, 0.4407516538918458, 0.8815033077836916, 0.19221028877248647,
        0.38442057754497294, 0.7688411550899459, 1.5376823101798918, 1.5045682935648867,
        1.438340260334877, 1.3058841938748573, 1.0409720609548179, 0.5111477951147392,
        1.0222955902294784, 0.47379485366406027, 0.9475897073281205, 0.3243830878613445 };
    static const unsigned char kQ[88] = {
        3, 2, 0, 0, 0, 0, 1, 3, 2, 1, 3, 2, 1, 3, 3, 2, 0, 1, 2, 0, 1, 3, 3, 2, 0, 1, 2, 0, 0, 1, 2, 0, 0, 0, 0, 1, 2, 1, 2, 1, 2, 0, 1, 2, 1, 2, 0, 1, 3, 3, 3, 3, 3, 3,
        2, 0, 0, 0, 1, 2, 0, 1, 3, 3, 2, 1, 2, 1, 2, 1, 3, 3, 3, 3, 2, 1, 2, 0, 0, 1, 3, 3, 3, 3, 2, 1, 2, 1 };
    int q0 = 0;
    bool neg = false;
    if (fabs(x) >= 1099511627776.0) {           // 2^40
        neg = x < 0.0;
        const uint64_t u = d2u(fabs(x));
        const int ex = (int)(u >> 52) - 1023;                              // x = 1.f * 2^ex
        const uint64_t m = ((u & 0xfffffffffffffull) | 0x10000000000000ull) >> 29;   // 24-bit integer significand of the float
        const int e = ex - 23;                                             // x = m * 2^e
        const int i = e - 17;
        q0 = (int)((m * (uint64_t)kQ[i]) & 3u);
        x = (double)m * kR[i];                                             // < 2^25
    }
    const double n = rint(x * kTwoOverPi);
    r = fma(-n, kPio2Hi, x);
    r = fma(-n, kPio2Lo, r);
    int q = (int)(((long long)n + (long long)q0) & 3ll);
    if (neg) { r = -r; q = (4 - q) & 3; }
    return q;
}
// Taylor kernels on |r| <= pi/4 (+ slack): truncation < 3e-14 (sin, r^15/15!), 2e-15 (cos, r^16/16!)
PM_HD double sin_k(double r) {
    const double z = r * r;
    double p = -7.647163731819816e-13;                  // -1/15!
    p = fma(p, z, 1.6059043836821613e-10);              //  1/13!
    p = fma(p, z, -2.505210838544172e-08);              // -1/11!
    p = fma(p, z, 2.7557319223985893e-06);              //  1/9!
    p = fma(p, z, -0.0001984126984126984);              // -1/7!
    p = fma(p, z, 0.008333333333333333);                //  1/5!
    p = fma(p, z, -0.16666666666666666);                // -1/3!
    return fma(r * z, p, r);
}
PM_HD double cos_k(double r) {
    const double z = r * r;
    double p = 4.779477332387385e-14;                   //  1/16!
    p = fma(p, z, -1.1470745597729725e-11);             // -1/14!
    p = fma(p, z, 2.08767569878681e-09);                //  1/12!
    p = fma(p, z, -2.755731922398589e-07);              // -1/10!
    p = fma(p, z, 2.48015873015873e-05);                //  1/8!
    p = fma(p, z, -0.001388888888888889);               // -1/6!
    p = fma(p, z, 0.041666666666666664);                //  1/4!
    p = fma(p, z, -0.5);
    return fma(p, z, 1.0);
}
PM_HD void sincos_d(double x, double& s, double& c) {
    if (!(fabs(x) < pm_inf())) { s = c = pm_nan(); return; }
    double r; const int q = rem_pio2(x, r);
    const double sr = sin_k(r), cr = cos_k(r);
    s = (q & 1) ? cr : sr; c = (q & 1) ? sr : cr;
    if (q & 2) s = -s;
    if ((q + 1) & 2) c = -c;
}
PM_API float sinf(float x) { if (x == 0.f) return x; double s, c; sincos_d((double)x, s, c); return (float)s; }
PM_API float cosf(float x) { double s, c; sincos_d((double)x, s, c); return (float)c; }
PM_API void sincosf(float x, float* s, float* c) { double sd, cd; sincos_d((double)x, sd, cd); *s = x == 0.f ? x : (float)sd; *c = (float)cd; }
PM_API float tanf(float x) { if (x == 0.f) return x; double s, c; sincos_d((double)x, s, c); return (float)(s / c); }

// ---- exp / log in binary64
PM_HD double exp_d(double x) {          // |x| < ~700 ; relative error ~2e-16
    if (x != x) return x;
    if (x > 709.0) return pm_inf();
    if (x < -745.0) return 0.0;
    const double n = rint(x * kLog2e);
    double r = fma(-n, kLn2Hi, x);
    r = fma(-n, kLn2Lo, r);             // |r| <= ln2/2
    double p = 1.6059043836821613e-10;  // 1/13!
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 0.0001984126984126984);
    p = fma(p, r, 0.001388888888888889);
    p = fma(p, r, 0.008333333333333333);
    p = fma(p, r, 0.041666666666666664);
    p = fma(p, r, 0.16666666666666666);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int ni = (int)n;
    // two-step scaling keeps 2^n representable when the result is subnormal / near overflow
    const int n1 = ni / 2, n2 = ni - n1;
    return p * pow2i(n1) * pow2i(n2);
}
PM_HD double log_d(double x) {          // x > 0 finite normal (every positive float is a normal double); relative error ~2e-16
    const uint64_t u = d2u(x);
    int e = (int)(u >> 52) - 1023;
    double m = u2d((u & 0xfffffffffffffull) | 0x3ff0000000000000ull);      // [1, 2)
    if (m > 1.4142135623730951) { m *= 0.5; ++e; }
    const double s = (m - 1.0) / (m + 1.0), z = s * s;                      // |s| <= 0.1716
    double p = 0.05263157894736842;     // 1/19
    p = fma(p, z, 0.058823529411764705);
    p = fma(p, z, 0.06666666666666667);
    p = fma(p, z, 0.07692307692307693);
    p = fma(p, z, 0.09090909090909091);
    p = fma(p, z, 0.1111111111111111);
    p = fma(p, z, 0.14285714285714285);
    p = fma(p, z, 0.2);
    p = fma(p, z, 0.3333333333333333);
    const double l = 2.0 * fma(s * z, p, s);
    const double ed = (double)e;
    return fma(ed, kLn2Hi, fma(ed, kLn2Lo, l));
}
PM_API float expf(float x) {
    if (x != x) return x;
    if (x > 88.8f) return (float)pm_inf();
    if (x < -104.f) return 0.f;
    return (float)exp_d((double)x);
}
PM_API float logf(float x) {
    if (x != x || x < 0.f) return (float)pm_nan();
    if (x == 0.f) return -(float)pm_inf();
    if (!(x < (float)pm_inf())) return x;
    if (x == 1.f) return 0.f;
    return (float)log_d((double)x);
}
PM_API float powf(float x, float y) {
    if (y == 0.f || x == 1.f) return 1.f;
    if (x != x || y != y) return (float)pm_nan();
    const bool yint = floorf(y) == y, yodd = yint && fabsf(y) < 16777216.f && (((long long)y) & 1ll);
    if (x == 0.f) { const float r = y > 0.f ? 0.f : (float)pm_inf(); return (yodd && f2u(x) >> 31) ? -r : r; }
    const float ax = fabsf(x);
    if (x < 0.f && !yint) return (float)pm_nan();
    float r;
    if (!(ax < (float)pm_inf())) r = y > 0.f ? (float)pm_inf() : 0.f;
    else if (!(fabsf(y) < (float)pm_inf())) r = (ax > 1.f) == (y > 0.f) ? (float)pm_inf() : 0.f;
    else {
        const double t = (double)y * log_d((double)ax);
        r = t > 100.0 ? (float)pm_inf() : t < -120.0 ? 0.f : (float)exp_d(t);
    }
    return (x < 0.f && yodd) ? -r : r;
}

// ---- inverse trigonometric functions in binary64
PM_HD double atan_d(double t) {         // t >= 0 finite or +inf; absolute error ~1e-15
    // atan(t) = pi/2 - atan(1/t) for t > 1; two tangent-subtraction steps bring the argument below tan(pi/16); odd series up to u^19
    const bool inv = t > 1.0;
    double u = inv ? 1.0 / t : t, off = 0.0;
    if (u > 0.41421356237309503) { u = (u - 1.0) / (u + 1.0); off = 0.7853981633974483; }                 // atan(u) = pi/4 + atan((u-1)/(u+1))
    if (u > 0.198912367379658) { u = (u - 0.41421356237309503) / fma(u, 0.41421356237309503, 1.0); off += 0.39269908169872414; }
    else if (u < -0.198912367379658) { u = (u + 0.41421356237309503) / fma(-u, 0.41421356237309503, 1.0); off -= 0.39269908169872414; }
    const double z = u * u;
    double p = 0.05263157894736842;     // 1/19
    p = fma(p, z, -0.058823529411764705);
    p = fma(p, z, 0.06666666666666667);
    p = fma(p, z, -0.07692307692307693);
    p = fma(p, z, 0.09090909090909091);
    p = fma(p, z, -0.1111111111111111);
    p = fma(p, z, 0.14285714285714285);
    p = fma(p, z, -0.2);
    p = fma(p, z, 0.3333333333333333);
    const double a = off + fma(-(u * z), p, u);
    return inv ? (kPio2Hi - a) + kPio2Lo : a;
}
PM_HD double atan2_d(double y, double x) {
    if (x != x || y != y) return pm_nan();
    const bool ny = (d2u(y) >> 63) != 0, nx = (d2u(x) >> 63) != 0;
    const double ay = fabs(y), ax = fabs(x);
    double a;
    if (ay == 0.0) a = 0.0;
    else if (ax == 0.0) a = kPio2Hi;
    else if (!(ax < pm_inf()) && !(ay < pm_inf())) a = 0.7853981633974483;
    else if (!(ay < pm_inf())) a = kPio2Hi;
    else if (!(ax < pm_inf())) a = 0.0;
    else a = atan_d(ay / ax);
    if (nx) a = (kPiD - a) + kPiLoD;
    return ny ? -a : a;
}
PM_API float atan2f(float y, float x) {
    if (y == 0.f && !(f2u(x) >> 31) && x == x) return y;            // +-0 for x >= +0
    return (float)atan2_d((double)y, (double)x);
}
PM_API float acosf(float x) {
    if (!(fabsf(x) <= 1.f)) return (float)pm_nan();
    const double xd = (double)x;
    return (float)atan2_d(sqrt((1.0 - xd) * (1.0 + xd)), xd);
}
PM_API float hypotf(float a, float b) {
    if (!(fabsf(a) < (float)pm_inf()) || !(fabsf(b) < (float)pm_inf())) return (a != a && fabsf(b) < (float)pm_inf()) || (b != b && fabsf(a) < (float)pm_inf()) || (a != a && b != b) ? (float)pm_nan() : (float)pm_inf();
    const double ad = (double)a, bd = (double)b;
    return (float)sqrt(fma(ad, ad, bd * bd));
}

} // namespace pm
#endif

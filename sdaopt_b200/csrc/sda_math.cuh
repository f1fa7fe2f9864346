// sda_math.cuh -- FP64 elementary functions of the two hot loops with their coefficients in
// constant memory.
//
// Why: ncu on the annealing launch showed 11.4 % of all issued instructions to be UMOV -- libdevice's
// log/exp/cos carry their polynomial coefficients as 64-bit literals, and sm_100 materialises every
// literal with two UMOVs right before the DFMA that uses it.  The same polynomials with the
// coefficients in a __constant__ table compile to one LDCU.128 per two coefficients.  The range
// reductions are also specialised to what the callers need:
//   cos_2pi(v)  = cos(2 pi v), period reduction v - rint(v) is exact (no pi/2 Cody-Waite, no quadrant
//                 select): the Rastrigin-family term cos(2*pi*x) (go_funcs_R.py:91, README.md:65-66);
//   log_pos(x)  = log(x) for positive normal x (anything else goes to libdevice);
//   exp_any(a)  = exp(a) for -708 <= a < 709; +inf from 709 on, 0 below -708.
// Accuracy (tests/test_gpu_parity.py::test_device_math_accuracy, against x87 long double):
// log_pos and exp_any <= 1 ulp-class like libdevice; cos_2pi <= 4e-16 ABSOLUTE, which is what a term
// `x^2 - 10 cos(2 pi x)` needs.  cos_2pi is the mathematically exact-period value; the reference's
// cos(fl(2*pi*x)) differs from it by up to |2 pi x| 2^-53 |sin| <= 3.6e-15 for |x| <= 5.12 -- four
// orders of magnitude inside the 1e-12 relative parity criterion.  The tape-replay path keeps
// libdevice in visit_from_polar (sda_device.cuh).
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace sda {

// (-1)^n (2 pi)^(2n) / (2n)!, n = 1 .. 12
__constant__ double kCos2PiC[12] = {
    -1.97392088021787159846e+01, 6.49393940226682957473e+01,  -8.54568172066937279396e+01,
    6.02446413718766606848e+01,  -2.64262567833743986512e+01, 7.90353637131846920028e+00,
    -1.71439071108867202575e+00, 2.82005968455791233840e-01,  -3.63828411425456688111e-02,
    3.77983420068003957148e-03,  -3.22991067207097808708e-04, 2.30999569450704433926e-05};
// 1/n!, n = 2 .. 13
__constant__ double kExpC[12] = {
    5.00000000000000000000e-01, 1.66666666666666657415e-01, 4.16666666666666643537e-02,
    8.33333333333333321769e-03, 1.38888888888888894189e-03, 1.98412698412698412526e-04,
    2.48015873015873015658e-05, 2.75573192239858925110e-06, 2.75573192239858882758e-07,
    2.50521083854417202239e-08, 2.08767569878681001866e-09, 1.60590438368216133409e-10};
// log(1+f) = f - s (f - R(s^2)), s = f / (2 + f): the classic 14th-degree Remez fit of
// R(z) = Lg1 z + ... + Lg7 z^7 on [0, 0.1716] (error < 2^-58.45), then ln2 split in two
__constant__ double kLogC[10] = {
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
    2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
    1.479819860511658591e-01, 6.93147180369123816490e-01, 1.90821492927058770002e-10, 0.0};

constexpr double kMagicRint = 6755399441055744.0;   // 1.5 * 2^52

// The special-case paths of the functions below go to libdevice.  They are (almost) never taken, but
// libdevice is bitcode: left to itself the compiler may inline cos / log / fmod -- ~1.1 KB of literals in
// the kernel's constant bank, ~100 bytes more stack, a longer hot loop -- and whether it does depended on
// which -split-compile partition the caller landed in: two builds of IDENTICAL sources from different
// directories ran the annealing launch at 1,306 and 1,221 M evaluations/s
// (profiles/r2_ab_split_compile_bisect.log).  Out of line, the fast path is the same in every build.
__device__ __noinline__ double cold_cos(double x) { return cos(x); }
__device__ __noinline__ double cold_log(double x) { return log(x); }
__device__ __noinline__ double cold_fmod(double a, double b) { return fmod(a, b); }

// a / b and sqrt(x) for NORMAL positive b, x (and a normal quotient): MUFU seed (2^-23), two Newton
// steps, one residual correction -- 8 branch-free instructions instead of libdevice's ~14 plus the
// BSSY/CALL scaffolding of its special-case path.  Faithfully rounded: never more than 1 ulp off
// (measured: the correctly rounded value in ~95 % of the square roots, ~99.9 % of the quotients).
__device__ __forceinline__ double div_normal(double a, double b)
{
#ifdef SDA_LIBDEVICE_DIV
    return __ddiv_rn(a, b);
#endif
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    const double q = __dmul_rn(a, r);
    return fma(fma(-q, b, a), r, q);
}
__device__ __forceinline__ double sqrt_normal(double x)
{
#ifdef SDA_LIBDEVICE_DIV
    return sqrt(x);
#endif
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));      // ~2^-22
    const double hx = __dmul_rn(0.5, x);
    y = fma(y, fma(-hx, __dmul_rn(y, y), 0.5), y);                 // y (1.5 - x/2 y^2)
    y = fma(y, fma(-hx, __dmul_rn(y, y), 0.5), y);
    const double g = __dmul_rn(x, y);                              // ~sqrt(x)
    return fma(fma(-g, g, x), __dmul_rn(0.5, y), g);               // g + (x - g^2) y / 2
}

__device__ __forceinline__ double cos_2pi(double v)
{
#ifdef SDA_LIBDEVICE_MATH   // A/B switch: the libdevice calls this file replaces
    return cos(6.283185307179586476925286766559 * v);
#endif
    if (!(fabs(v) < 2251799813685248.0)) return cold_cos(6.283185307179586476925286766559 * v);  // 2^51, inf, nan
    const double q = __dadd_rn(__dadd_rn(v, kMagicRint), -kMagicRint);   // rint(v)
    double a = fabs(__dsub_rn(v, q));                                    // exact, in [0, 0.5]
    const bool flip = a > 0.25;
    if (flip) a = __dsub_rn(0.5, a);                                     // exact; cos(2 pi t) = -cos(2 pi (1/2 - t))
    const double z = __dmul_rn(a, a);
    // (Horner on purpose: Estrin's scheme shortens the dependency chain from 11 to 5 FMAs but costs
    // three more multiplies and four more live registers -- measured 5 % SLOWER on the headline shape,
    // profiles/r2_ab_occupancy_estrin.log: the launch is bound by issue slots, not by latency)
    double p = kCos2PiC[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = fma(p, z, kCos2PiC[i]);
    p = fma(p, z, 1.0);
    return flip ? -p : p;
}

__device__ __forceinline__ double exp_any(double a)
{
#if defined(SDA_LIBDEVICE_MATH) || defined(SDA_LIBDEVICE_LOGEXP)
    return exp(a);
#endif
    // a >= 709: +inf (the true value is >= 8.2e307; every caller clips far below that), nan -> nan
    if (!(a < 709.0)) return (a != a) ? a : __longlong_as_double(0x7ff0000000000000LL);
    if (a < -708.0) return 0.0;
    const double t = fma(a, 1.4426950408889634074, kMagicRint);
    const int ki = __double2loint(t);
    const double k = __dadd_rn(t, -kMagicRint);
    double r = fma(k, -kLogC[7], a);
    r = fma(k, -kLogC[8], r);
    double q = kExpC[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) q = fma(q, r, kExpC[i]);
    double p = fma(__dmul_rn(r, r), q, r);
    p = __dadd_rn(p, 1.0);                                  // in [0.70, 1.42]
    return __hiloint2double(__double2hiint(p) + (ki << 20), __double2loint(p));
}

#ifdef SDA_NOINLINE_LOG
__device__ __noinline__ double log_pos(double x)
#else
__device__ __forceinline__ double log_pos(double x)
#endif
{
#if defined(SDA_LIBDEVICE_MATH) || defined(SDA_LIBDEVICE_LOGEXP)
    return log(x);
#endif
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    if ((unsigned)(hx - 0x00100000) >= 0x7fe00000u) return cold_log(x);   // 0, denormal, negative, inf, nan
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;                  // mantissa above sqrt(2): halve it
    const double m = __hiloint2double(hx | (i ^ 0x3ff00000), lx);
    k += i >> 20;
    const double dk = __dadd_rn(__hiloint2double(0x43300000, k ^ 0x80000000), -4503601774854144.0);  // (double)k
    const double f = __dadd_rn(m, -1.0);
    const double s = div_normal(f, __dadd_rn(2.0, f));
    const double z = __dmul_rn(s, s);
    const double w = __dmul_rn(z, z);
    const double t1 = __dmul_rn(w, fma(w, fma(w, kLogC[5], kLogC[3]), kLogC[1]));
    const double t2 = __dmul_rn(z, fma(w, fma(w, fma(w, kLogC[6], kLogC[4]), kLogC[2]), kLogC[0]));
    const double R = __dadd_rn(t2, t1);
    const double hfsq = __dmul_rn(0.5, __dmul_rn(f, f));
    // dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f)
    const double inner = fma(s, __dadd_rn(hfsq, R), __dmul_rn(dk, kLogC[8]));
    return __dsub_rn(__dmul_rn(dk, kLogC[7]), __dsub_rn(__dsub_rn(hfsq, inner), f));
}

}  // namespace sda

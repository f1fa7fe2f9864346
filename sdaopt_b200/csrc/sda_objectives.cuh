// sda_objectives.cuh -- K3: device-side ports of the benchmark objectives.
//
// One chain group (G lanes, sda_device.cuh) evaluates one point: lane gl handles terms gl, gl+G,
// ...; partial sums meet in an xor-butterfly so all lanes of the group return the same double.  x points to shared memory (or any memory
// the whole warp can read).  Reference bodies: sdaopt/benchmark/go_benchmark_functions/*.py and
// sdaopt/benchmark/suttonchen.py; formulas keep the reference's operation order per term, the
// summation order differs (allowed by the 1e-12 relative criterion).
#pragma once

#include "sda_device.cuh"

#include "sda_objectives_cat.cuh"

#ifdef SDA_USER_FUNCTOR_HEADER
#include SDA_USER_FUNCTOR_HEADER
#endif

namespace sda {

// suttonchen.py:23-40  (A=1, C=39.432, M=6, K=9).  Lane L owns atoms L, L+G, ...: for each it
// accumulates sum_j r_ij^-9 and sum_j r_ij^-6 over all other atoms from shared memory.
// One pair term: r^-1 by rsqrt (1 ulp) instead of an IEEE division plus a square root, r^-6 and r^-9
// by multiplication (relative error ~1e-15, inside the 1e-11 parity bound); the j == i term is
// masked to zero instead of branched around, so that the four terms of a trip are independent
// instruction streams.
template <class XV>
__device__ __forceinline__ void sutton_chen_term(const XV &x, int i, int j, double xi, double yi,
                                                 double zi, double &rep, double &rho)
{
    const double dx = xi - x[3 * j], dy = yi - x[3 * j + 1], dz = zi - x[3 * j + 2];
    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    const double inv1 = (j == i) ? 0.0 : rsqrt(r2);
    const double inv2 = inv1 * inv1;
    const double inv6 = inv2 * inv2 * inv2;
    rho += inv6;
    rep = fma(inv6 * inv2, inv1, rep);
}

// Four atoms j per trip with their own accumulators: at D = 3N >= 768 only 6-8 warps fit an SM
// (the chain state lives in shared memory), and one dependent rsqrt/multiply chain per warp left the
// FP64 pipe idle (ncu-less evidence: 67 G pair terms/s at N = 512 against 371 G at N = 128,
// profiles/r1_config3_config4_shapes.log).
template <class XV>
__device__ __forceinline__ double eval_sutton_chen(const XV &x, int dim, const Group &g)
{
    const int n = dim / 3;
    double part = 0.0;
    for (int i = g.gl; i < n; i += g.G) {
        const double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
        double rep0 = 0.0, rho0 = 0.0, rep1 = 0.0, rho1 = 0.0;
        double rep2 = 0.0, rho2 = 0.0, rep3 = 0.0, rho3 = 0.0;
        int j = 0;
        for (; j + 3 < n; j += 4) {
            sutton_chen_term(x, i, j, xi, yi, zi, rep0, rho0);
            sutton_chen_term(x, i, j + 1, xi, yi, zi, rep1, rho1);
            sutton_chen_term(x, i, j + 2, xi, yi, zi, rep2, rho2);
            sutton_chen_term(x, i, j + 3, xi, yi, zi, rep3, rho3);
        }
        for (; j < n; ++j) sutton_chen_term(x, i, j, xi, yi, zi, rep0, rho0);
        const double rep = (rep0 + rep1) + (rep2 + rep3);
        const double rho = (rho0 + rho1) + (rho2 + rho3);
        part += 0.5 * rep - 39.432 * sqrt(rho);
    }
    return group_sum(part, g);
}

// catalogue part 2 (ids 11..59), out of line
template <class XV>
__device__ __noinline__ double eval_catalogue2(int fid, const XV &x, int dim, const double *params,
                                              const Group &g, unsigned long long salt)
{
    const int lane = g.gl, SDA_STRIDE = g.G;
    double a = 0.0, b = 0.0;
    (void)params; (void)salt;
    switch (fid) {
    // ---- catalogue widening; formulas: go_benchmark_functions/go_funcs_<initial>.py
    // n-D forms: lane `lane` of the group takes terms lane, lane+G, ...
    case SDA_FN_ALPINE01:  // sum |x sin x + 0.1 x|
        for (int k = lane; k < dim; k += SDA_STRIDE) { const double v = x[k]; a += fabs(v * sin(v) + 0.1 * v); }
        return group_sum(a, g);
    case SDA_FN_BROWN:  // sum (x_i^2)^(x_{i+1}^2+1) + (x_{i+1}^2)^(x_i^2+1)
        for (int k = lane; k + 1 < dim; k += SDA_STRIDE) {
            const double p0 = x[k] * x[k], p1 = x[k + 1] * x[k + 1];
            a += pow(p0, p1 + 1.0) + pow(p1, p0 + 1.0);
        }
        return group_sum(a, g);
    case SDA_FN_CIGAR:  // x_0^2 + 1e6 sum_{i>=1} x_i^2
        for (int k = lane; k < dim; k += SDA_STRIDE) { if (k > 0) a += x[k] * x[k]; }
        return x[0] * x[0] + 1e6 * group_sum(a, g);
    case SDA_FN_COSINE_MIXTURE:  // -0.1 sum cos(5 pi x) - sum x^2
        for (int k = lane; k < dim; k += SDA_STRIDE) { const double v = x[k]; a += cos(5.0 * kPi * v); b += v * v; }
        return -0.1 * group_sum(a, g) - group_sum(b, g);
    case SDA_FN_DEB01:  // -(1/N) sum sin(5 pi x)^6
        for (int k = lane; k < dim; k += SDA_STRIDE) { const double t = sin(5 * kPi * x[k]); const double t2 = t * t; a += t2 * t2 * t2; }
        return -(1.0 / dim) * group_sum(a, g);
    case SDA_FN_DIXON_PRICE:  // sum_{i=2..N} i (2 x_i^2 - x_{i-1})^2 + (x_1 - 1)^2
        for (int k = lane; k + 1 < dim; k += SDA_STRIDE) { const double t = 2.0 * x[k + 1] * x[k + 1] - x[k]; a += (double)(k + 2) * (t * t); }
        return group_sum(a, g) + (x[0] - 1.0) * (x[0] - 1.0);
    case SDA_FN_QING:  // sum (x_i^2 - i)^2
        for (int k = lane; k < dim; k += SDA_STRIDE) { const double t = x[k] * x[k] - (double)(k + 1); a += t * t; }
        return group_sum(a, g);
    case SDA_FN_QUINTIC:  // sum |x^5 - 3x^4 + 4x^3 + 2x^2 - 10x - 4|
        for (int k = lane; k < dim; k += SDA_STRIDE) {
            const double v = x[k], v2 = v * v, v3 = v2 * v, v4 = v2 * v2, v5 = v4 * v;
            a += fabs(v5 - 3 * v4 + 4 * v3 + 2 * v2 - 10 * v - 4);
        }
        return group_sum(a, g);
    case SDA_FN_SALOMON: {  // 1 - cos(2 pi u) + 0.1 u, u = |x|
        for (int k = lane; k < dim; k += SDA_STRIDE) a += x[k] * x[k];
        const double u = sqrt(group_sum(a, g));
        return 1 - cos(kTwoPi * u) + 0.1 * u;
    }
    case SDA_FN_SCHWEFEL20:
        for (int k = lane; k < dim; k += SDA_STRIDE) a += fabs(x[k]);
        return group_sum(a, g);
    case SDA_FN_SCHWEFEL21:
        for (int k = lane; k < dim; k += SDA_STRIDE) a = fmax(a, fabs(x[k]));
        for (int o = g.G >> 1; o > 0; o >>= 1) a = fmax(a, __shfl_xor_sync(g.mask, a, o));
        return a;
    case SDA_FN_SCHWEFEL22:
        b = 1.0;
        for (int k = lane; k < dim; k += SDA_STRIDE) { const double v = fabs(x[k]); a += v; b *= v; }
        return group_sum(a, g) + group_prod(b, g);
    case SDA_FN_STEP:
        for (int k = lane; k < dim; k += SDA_STRIDE) a += floor(fabs(x[k]));
        return group_sum(a, g);
    case SDA_FN_STYBLINSKI_TANG:  // sum(x^4 - 16 x^2 + 5 x) / 2
        for (int k = lane; k < dim; k += SDA_STRIDE) { const double v = x[k], v2 = v * v; a += v2 * v2 - 16 * v2 + 5 * v; }
        return group_sum(a, g) / 2;
    case SDA_FN_TRID:  // sum (x-1)^2 - sum x_i x_{i-1}
        for (int k = lane; k < dim; k += SDA_STRIDE) {
            const double t = x[k] - 1.0;
            a += t * t;
            if (k > 0) b += x[k] * x[k - 1];
        }
        return group_sum(a, g) - group_sum(b, g);
    case SDA_FN_ZACHAROV: {  // u + (v/2)^2 + (v/2)^4
        for (int k = lane; k < dim; k += SDA_STRIDE) { a += x[k] * x[k]; b += (double)(k + 1) * x[k]; }
        const double u = group_sum(a, g), h = 0.5 * group_sum(b, g), h2 = h * h;
        return u + h2 + h2 * h2;
    }
    case SDA_FN_WAVY:  // 1 - (1/N) sum cos(10x) exp(-x^2/2)
        for (int k = lane; k < dim; k += SDA_STRIDE) { const double v = x[k]; a += cos(10 * v) * exp(-(v * v) / 2.0); }
        return 1.0 - (1.0 / dim) * group_sum(a, g);
    case SDA_FN_STEP2:  // sum (floor(x) + 0.5)^2
        for (int k = lane; k < dim; k += SDA_STRIDE) { const double t = floor(x[k]) + 0.5; a += t * t; }
        return group_sum(a, g);
    case SDA_FN_PLATEAU:
        for (int k = lane; k < dim; k += SDA_STRIDE) a += floor(fabs(x[k]));
        return 30.0 + group_sum(a, g);
    case SDA_FN_SODP:  // sum |x_i|^(i+1)
        for (int k = lane; k < dim; k += SDA_STRIDE) a += pow(fabs(x[k]), (double)(k + 2));
        return group_sum(a, g);
    case SDA_FN_TRIGONOMETRIC02:  // 1 + sum 8 sin^2(7 t) + 6 sin^2(14 t) + t, t = (x-0.9)^2
        for (int k = lane; k < dim; k += SDA_STRIDE) {
            const double d = x[k] - 0.9, t = d * d, s7 = sin(7 * t), s14 = sin(14 * t);
            a += 8 * (s7 * s7) + 6 * (s14 * s14) + t;
        }
        return 1.0 + group_sum(a, g);
    case SDA_FN_XIN_SHE_YANG02:  // sum|x| exp(-sum sin(x^2))
        for (int k = lane; k < dim; k += SDA_STRIDE) { const double v = x[k]; a += fabs(v); b += sin(v * v); }
        return group_sum(a, g) * exp(-group_sum(b, g));
    case SDA_FN_SINE_ENVELOPE:  // sum (sin^2(sqrt(q)) - 0.5) / (1 + 0.001 q)^2 + 0.5, q = x_i^2 + x_{i+1}^2
        for (int k = lane; k + 1 < dim; k += SDA_STRIDE) {
            const double q = x[k] * x[k] + x[k + 1] * x[k + 1], sn = sin(sqrt(q)), dn = 1 + 0.001 * q;
            a += (sn * sn - 0.5) / (dn * dn) + 0.5;
        }
        return group_sum(a, g);
    case SDA_FN_PATHOLOGICAL:
        for (int k = lane; k + 1 < dim; k += SDA_STRIDE) {
            const double p0 = x[k], p1 = x[k + 1], sn = sin(sqrt(100 * (p0 * p0) + p1 * p1));
            const double q = p0 * p0 - 2 * p0 * p1 + p1 * p1;
            a += 0.5 + (sn * sn - 0.5) / (1. + 0.001 * (q * q));
        }
        return group_sum(a, g);
    case SDA_FN_RANA:
        for (int k = lane; k + 1 < dim; k += SDA_STRIDE) {
            const double p0 = x[k], p1 = x[k + 1];
            const double t1 = sqrt(fabs(p1 + p0 + 1)), t2 = sqrt(fabs(p1 - p0 + 1));
            a += (p1 + 1) * cos(t2) * sin(t1) + p0 * cos(t1) * sin(t2);
        }
        return group_sum(a, g);
    // fixed 2-D forms: every lane evaluates the expression (no reduction)
    case SDA_FN_EASOM: { const double p = x[0], q = x[1]; const double e = (p - kPi) * (p - kPi) + (q - kPi) * (q - kPi); return -cos(p) * cos(q) * exp(-e); }
    case SDA_FN_SIX_HUMP_CAMEL: { const double p = x[0], q = x[1], p2 = p * p, q2 = q * q; return (4 - 2.1 * p2 + p2 * p2 / 3) * p2 + p * q + (4 * q2 - 4) * q2; }
    case SDA_FN_BEALE: { const double p = x[0], q = x[1]; const double t1 = 1.5 - p + p * q, t2 = 2.25 - p + p * (q * q), t3 = 2.625 - p + p * (q * q * q); return t1 * t1 + t2 * t2 + t3 * t3; }
    case SDA_FN_HIMMEL_BLAU: { const double p = x[0], q = x[1]; const double t1 = p * p + q - 11, t2 = p + q * q - 7; return t1 * t1 + t2 * t2; }
    case SDA_FN_MATYAS: { const double p = x[0], q = x[1]; return 0.26 * (p * p + q * q) - 0.48 * p * q; }
    case SDA_FN_MC_CORMICK: { const double p = x[0], q = x[1]; return sin(p + q) + (p - q) * (p - q) - 1.5 * p + 2.5 * q + 1; }
    case SDA_FN_GOLDSTEIN_PRICE: {
        const double p = x[0], q = x[1];
        const double s1 = p + q + 1, s2 = 2 * p - 3 * q;
        const double t1 = 1 + s1 * s1 * (19 - 14 * p + 3 * (p * p) - 14 * q + 6 * p * q + 3 * (q * q));
        const double t2 = 30 + s2 * s2 * (18 - 32 * p + 12 * (p * p) + 48 * q - 36 * p * q + 27 * (q * q));
        return t1 * t2;
    }
    case SDA_FN_LEVY13: {
        const double p = x[0], q = x[1];
        const double s3p = sin(3 * kPi * p), s3q = sin(3 * kPi * q), s2q = sin(kTwoPi * q);
        return s3p * s3p + (p - 1) * (p - 1) * (1 + s3q * s3q) + (q - 1) * (q - 1) * (1 + s2q * s2q);
    }
    case SDA_FN_BOHACHEVSKY1: { const double p = x[0], q = x[1]; return p * p + 2 * (q * q) - 0.3 * cos(3 * kPi * p) - 0.4 * cos(4 * kPi * q) + 0.7; }
    case SDA_FN_BUKIN06: { const double p = x[0], q = x[1]; return 100 * sqrt(fabs(q - 0.01 * (p * p))) + 0.01 * fabs(p + 10); }
    case SDA_FN_EGG_CRATE: { const double p = x[0], q = x[1], sp = sin(p), sq = sin(q); return p * p + q * q + 25 * (sp * sp + sq * sq); }
    case SDA_FN_SCHAFFER01: { const double u = x[0] * x[0] + x[1] * x[1], sn = sin(u), dn = 1 + 0.001 * u; return 0.5 + (sn * sn - 0.5) / (dn * dn); }
    case SDA_FN_DROP_WAVE: { const double nx = x[0] * x[0] + x[1] * x[1]; return -(1 + cos(12 * sqrt(nx))) / (0.5 * nx + 2); }
    case SDA_FN_THREE_HUMP_CAMEL: { const double p = x[0], q = x[1], p2 = p * p; return 2.0 * p2 - 1.05 * (p2 * p2) + (p2 * p2 * p2) / 6.0 + p * q + q * q; }
    case SDA_FN_ZETTL: { const double p = x[0], q = x[1]; const double t = p * p + q * q - 2 * p; return t * t + 0.25 * p; }
    case SDA_FN_LEON: { const double p = x[0], q = x[1]; const double t = q - p * p; return 100. * (t * t) + (1 - p) * (1 - p); }
    case SDA_FN_CUBE: { const double p = x[0], q = x[1]; const double t = q - p * p * p; return 100.0 * (t * t) + (1.0 - p) * (1.0 - p); }
    case SDA_FN_BRENT: { const double p = x[0], q = x[1]; return (p + 10.0) * (p + 10.0) + (q + 10.0) * (q + 10.0) + exp(-(p * p) - q * q); }
    case SDA_FN_PRICE01: { const double t1 = fabs(x[0]) - 5.0, t2 = fabs(x[1]) - 5.0; return t1 * t1 + t2 * t2; }
    case SDA_FN_BIRD: { const double p = x[0], q = x[1], c1 = 1 - cos(q), c2 = 1 - sin(p); return sin(p) * exp(c1 * c1) + cos(q) * exp(c2 * c2) + (p - q) * (p - q); }
    case SDA_FN_ACKLEY02: return -200 * exp(-0.02 * sqrt(x[0] * x[0] + x[1] * x[1]));
    case SDA_FN_ACKLEY03: { const double p = x[0], q = x[1]; return -200 * exp(-0.02 * sqrt(p * p + q * q)) + 5 * exp(cos(3 * p) + sin(3 * q)); }
    case SDA_FN_ZIRILLI: { const double p = x[0], q = x[1], p2 = p * p; return 0.25 * (p2 * p2) - 0.5 * p2 + 0.1 * p + 0.5 * (q * q); }
    case SDA_FN_ADJIMAN: { const double p = x[0], q = x[1]; return cos(p) * sin(q) - p / (q * q + 1); }
    default:
        return nan_value();
    }
}

// Evaluate objective `fid` at x[0..dim).  Warp-uniform arguments; warp-uniform result.
// x is anything indexable: a pointer into shared memory, or a PerturbedView of it.
template <class XV>
__device__ __forceinline__ double eval_objective(int fid, const XV &x, int dim,
                                                 const double *params, const Group &g,
                                                 unsigned long long salt = 0)
{
    const int lane = g.gl, SDA_STRIDE = g.G;
    double a = 0.0, b = 0.0;
    switch (fid) {
    case SDA_FN_RASTRIGIN: {  // go_funcs_R.py:91   10*N + sum(x^2 - 10 cos(2 pi x))
        for (int k = lane; k < dim; k += SDA_STRIDE) {
            const double v = x[k];
            a += v * v - 10.0 * cos_2pi(v);
        }
        return 10.0 * dim + group_sum(a, g);
    }
    case SDA_FN_SHIFTED_RASTRIGIN: {  // README.md:65-66
        const double shift = params[0];
        for (int k = lane; k < dim; k += SDA_STRIDE) {
            const double d = x[k] - shift;
            a += d * d - 10.0 * cos_2pi(d);
        }
        return group_sum(a, g) + 10.0 * dim;
    }
    case SDA_FN_ROSENBROCK:  // scipy.optimize.rosen: sum(100 (x[i+1]-x[i]^2)^2 + (1-x[i])^2)
        for (int k = lane; k + 1 < dim; k += SDA_STRIDE) {
            const double v = x[k];
            const double t = x[k + 1] - v * v;
            const double u = 1.0 - v;
            a += 100.0 * (t * t) + u * u;
        }
        return group_sum(a, g);
    case SDA_FN_ACKLEY01:  // go_funcs_A.py:45-48
        for (int k = lane; k < dim; k += SDA_STRIDE) {
            const double v = x[k];
            a += v * v;
            b += cos_2pi(v);
        }
        a = group_sum(a, g);
        b = group_sum(b, g);
        return -20. * exp(-0.2 * sqrt(a / dim)) - exp(b / dim) + 20. + 2.718281828459045;
    case SDA_FN_SCHWEFEL01:  // go_funcs_S.py:335-336   (sum x^2) ** sqrt(pi)
        for (int k = lane; k < dim; k += SDA_STRIDE) a += x[k] * x[k];
        return pow(group_sum(a, g), 1.7724538509055159);
    case SDA_FN_SCHWEFEL26:  // go_funcs_S.py:616
        for (int k = lane; k < dim; k += SDA_STRIDE) {
            const double v = x[k];
            a += v * sin(sqrt(fabs(v)));
        }
        return 418.982887 * dim - group_sum(a, g);
    case SDA_FN_EXPONENTIAL:  // go_funcs_E.py:306
        for (int k = lane; k < dim; k += SDA_STRIDE) a += x[k] * x[k];
        return -exp(-0.5 * group_sum(a, g));
    case SDA_FN_SUTTON_CHEN:
        return eval_sutton_chen(x, dim, g);
    case SDA_FN_CONSTANT:
        return params[0];
    case SDA_FN_SPHERE:  // go_funcs_S.py:1159
        for (int k = lane; k < dim; k += SDA_STRIDE) a += x[k] * x[k];
        return group_sum(a, g);
    case SDA_FN_GRIEWANK:  // go_funcs_G.py:173-175
        b = 1.0;
        for (int k = lane; k < dim; k += SDA_STRIDE) {
            const double v = x[k];
            a += v * v / 4000.0;
            b *= cos(v / sqrt((double)(k + 1)));
        }
        return group_sum(a, g) - group_prod(b, g) + 1.0;
#ifdef SDA_USER_FUNCTOR_HEADER
    // user-supplied device functor: every lane of the group evaluates it (group-uniform result)
    case SDA_FN_USER:
        return ::sda_user_objective(x, dim, params);
#endif

    default:
        // ids 11..59 and 60..: the catalogue, out of line.  Only the eleven named objectives above are
        // inlined into the annealing loop: with sixty inlined cases the kernel was ~650 KB of SASS, and
        // how the compiler dispatched that switch (one 287-entry jump table or not) flipped between
        // builds of identical sources with -split-compile -- a 7 % swing on the headline shape
        // (profiles/r2_ab_split_compile_bisect.log).  A call per evaluation is noise for these.
        return fid < SDA_FN_AMGM ? eval_catalogue2(fid, x, dim, params, g, salt) : eval_catalogue(fid, x, dim, g, salt);
    }
}

// x with one coordinate replaced: the finite-difference points x +- h e_i of the gradient
// (_sda.py:328-339) without a private copy of x per lane
struct PerturbedView {
    const double *x;
    int index;
    double value;
    __device__ __forceinline__ double operator[](int k) const { return k == index ? value : x[k]; }
};

// ---- separable objectives: f = F(sum_k ta(x_k), sum_k tb(x_k)) -------------------------------------
// Used by the finite-difference gradient of the local search (sda_lbfgsb.cuh): the 2 D perturbed points
// of a gradient differ from the base point in ONE coordinate each, so all but one of the D terms of
// every perturbed evaluation are the base point's terms.  They are computed once, and every perturbed
// value is then summed from them IN THE SAME ORDER as the one-lane evaluation of eval_objective sums
// (k = 0 .. D-1 from 0.0), so each of the 2 D values is bit-identical to the full evaluation -- the
// expressions below are the same expressions as in eval_objective, and the library is built with
// --fmad=false.  2 D + 1 evaluations of D transcendental terms become 3 D terms plus 2 D^2 additions.
__device__ __forceinline__ bool objective_is_separable(int fid)
{
    return fid == SDA_FN_RASTRIGIN || fid == SDA_FN_SHIFTED_RASTRIGIN || fid == SDA_FN_ACKLEY01 ||
           fid == SDA_FN_SCHWEFEL01 || fid == SDA_FN_SCHWEFEL26 || fid == SDA_FN_EXPONENTIAL ||
           fid == SDA_FN_SPHERE;
}
__device__ __forceinline__ bool separable_has_second_sum(int fid) { return fid == SDA_FN_ACKLEY01; }

__device__ __forceinline__ void separable_term(int fid, double v, const double *params, double &ta, double &tb)
{
    tb = 0.0;
    switch (fid) {
    case SDA_FN_RASTRIGIN: ta = v * v - 10.0 * cos_2pi(v); break;
    case SDA_FN_SHIFTED_RASTRIGIN: { const double d = v - params[0]; ta = d * d - 10.0 * cos_2pi(d); break; }
    case SDA_FN_ACKLEY01: ta = v * v; tb = cos_2pi(v); break;
    case SDA_FN_SCHWEFEL26: ta = v * sin(sqrt(fabs(v))); break;
    default: ta = v * v; break;                 // Schwefel01, Exponential, Sphere
    }
}

__device__ __forceinline__ double separable_finish(int fid, double A, double B, int dim)
{
    switch (fid) {
    case SDA_FN_RASTRIGIN: return 10.0 * dim + A;
    case SDA_FN_SHIFTED_RASTRIGIN: return A + 10.0 * dim;
    case SDA_FN_ACKLEY01: return -20. * exp(-0.2 * sqrt(A / dim)) - exp(B / dim) + 20. + 2.718281828459045;
    case SDA_FN_SCHWEFEL01: return pow(A, 1.7724538509055159);
    case SDA_FN_SCHWEFEL26: return 418.982887 * dim - A;
    case SDA_FN_EXPONENTIAL: return -exp(-0.5 * A);
    default: return A;                          // Sphere
    }
}

// Out-of-line copy for the cold callers (local search, (re)initialisation): the body with all
// registered objectives is ~4.5 K instructions, and inlining it at every call site grew the
// L-BFGS-B code to 750 KB -- ncu showed 85 % of its cycles stalled on instruction fetch.
__device__ __noinline__ double eval_objective_ool(int fid, const double *x, int dim,
                                                  const double *params, int G, int lane,
                                                  unsigned long long salt)
{
    return eval_objective(fid, x, dim, params, make_group(G, lane), salt);
}

// One lane evaluates one perturbed point by itself (G = 1): the 2*D finite-difference evaluations
// of a gradient run 32 at a time.
__device__ __noinline__ double eval_objective_solo_ool(int fid, const double *x, int index,
                                                       double value, int dim, const double *params,
                                                       int lane, unsigned long long salt)
{
    const PerturbedView v = { x, index, value };
    return eval_objective(fid, v, dim, params, make_group(1, lane), salt);
}

}  // namespace sda

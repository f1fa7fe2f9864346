// sda_objectives.cuh -- K3: device-side ports of the benchmark objectives.
//
// One chain group (G lanes, sda_device.cuh) evaluates one point: lane gl handles terms gl, gl+G,
// ...; partial sums meet in an xor-butterfly so all lanes of the group return the same double.  x points to shared memory (or any memory
// the whole warp can read).  Reference bodies: sdaopt/benchmark/go_benchmark_functions/*.py and
// sdaopt/benchmark/suttonchen.py; formulas keep the reference's operation order per term, the
// summation order differs (allowed by the 1e-12 relative criterion).
#pragma once

#include "sda_device.cuh"

namespace sda {

// suttonchen.py:23-40  (A=1, C=39.432, M=6, K=9).  Lane L owns atoms L, L+32, ...: for each it
// accumulates sum_j r_ij^-9 and sum_j r_ij^-6 over all other atoms from shared memory.
template <class XV>
__device__ __forceinline__ double eval_sutton_chen(const XV &x, int dim, const Group &g)
{
    const int n = dim / 3;
    double part = 0.0;
    for (int i = g.gl; i < n; i += g.G) {
        const double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
        double rep = 0.0, rho = 0.0;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const double dx = xi - x[3 * j], dy = yi - x[3 * j + 1], dz = zi - x[3 * j + 2];
            const double r2 = dx * dx + dy * dy + dz * dz;
            const double inv2 = 1.0 / r2;
            const double inv6 = inv2 * inv2 * inv2;
            rho += inv6;
            rep += inv6 * inv2 * sqrt(inv2);
        }
        part += 0.5 * rep - 39.432 * sqrt(rho);
    }
    return group_sum(part, g);
}

// Evaluate objective `fid` at x[0..dim).  Warp-uniform arguments; warp-uniform result.
// x is anything indexable: a pointer into shared memory, or a PerturbedView of it.
template <class XV>
__device__ __forceinline__ double eval_objective(int fid, const XV &x, int dim,
                                                 const double *params, const Group &g)
{
    const int lane = g.gl, SDA_STRIDE = g.G;
    double a = 0.0, b = 0.0;
    switch (fid) {
    case SDA_FN_RASTRIGIN:  // go_funcs_R.py:91   10*N + sum(x^2 - 10 cos(2 pi x))
        for (int k = lane; k < dim; k += SDA_STRIDE) {
            const double v = x[k];
            a += v * v - 10.0 * cos(kTwoPi * v);
        }
        return 10.0 * dim + group_sum(a, g);
    case SDA_FN_SHIFTED_RASTRIGIN: {  // README.md:65-66
        const double shift = params[0];
        for (int k = lane; k < dim; k += SDA_STRIDE) {
            const double d = x[k] - shift;
            a += d * d - 10.0 * cos(kTwoPi * d);
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
            b += cos(kTwoPi * v);
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
    default:
        return __longlong_as_double(0x7ff8000000000000LL);
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

// Out-of-line copy for the cold callers (local search, (re)initialisation): the body with all
// registered objectives is ~4.5 K instructions, and inlining it at every call site grew the
// L-BFGS-B code to 750 KB -- ncu showed 85 % of its cycles stalled on instruction fetch.
__device__ __noinline__ double eval_objective_ool(int fid, const double *x, int dim,
                                                  const double *params, int G, int lane)
{
    return eval_objective(fid, x, dim, params, make_group(G, lane));
}

// One lane evaluates one perturbed point by itself (G = 1): the 2*D finite-difference evaluations
// of a gradient run 32 at a time.
__device__ __noinline__ double eval_objective_solo_ool(int fid, const double *x, int index,
                                                       double value, int dim, const double *params,
                                                       int lane)
{
    const PerturbedView v = { x, index, value };
    return eval_objective(fid, v, dim, params, make_group(1, lane));
}

}  // namespace sda

// sda_objectives_cat.cuh -- K3, catalogue part 3: the remaining classes of
// sdaopt/benchmark/go_benchmark_functions/go_funcs_*.py.  Functor ids 60..195.
//
// Same calling convention as eval_objective (sda_objectives.cuh): one chain group evaluates one
// point, x is indexable, every lane of the group returns the same double.  Fixed small-N classes
// are evaluated by every lane redundantly; n-D classes split their terms over the lanes.  Each case
// names the reference class; the formula keeps the reference's per-term operation order, sums are
// taken in group order (covered by the 1e-12 relative parity criterion against the golden points
// generated from the live reference, tests/golden/objective_points_cat3.npz).
#pragma once

#include "sda_device.cuh"

namespace sda {

__device__ __forceinline__ double group_max(double v, const Group &g)
{
    for (int o = g.G >> 1; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(g.mask, v, o));
    return v;
}
__device__ __forceinline__ double sq(double v) { return v * v; }
__device__ __forceinline__ double cube(double v) { return v * v * v; }
__device__ __forceinline__ double quart(double v) { const double t = v * v; return t * t; }
__device__ __forceinline__ double np_sign(double v) { return (double)((v > 0.0) - (v < 0.0)); }
__device__ __forceinline__ double nan_value() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---- objectives that draw random numbers inside fun() --------------------------------------------
// Stochastic (go_funcs_S.py:1274-1280) and XinSheYang01 (go_funcs_X.py:47-51) multiply every term by
// eps_i ~ U[0,1) drawn from NumPy's GLOBAL generator at every call, so their value at a point is not
// a function of the point and no call-by-call parity exists.  Here eps is counter-based noise:
// Philox4x32-10 with counter (hash64(x) + salt * phi64, i) -> the 53-bit uniform of numpy's
// random_sample().  `salt` is the chain's evaluation counter (ChainState::ncall; the 2 D points of
// a gradient carry their own call numbers), so every call draws fresh factors exactly as in the
// reference -- including repeated requests near one point, which is how the reference's line
// searches eventually hit f <= 1e-8 on these functions: with noise addressed by the point alone the
// reference itself drops from 29 % to 0 % success on Stochastic (measured, oracle/gen_suite_ref_rng.py
// header).  sda_eval passes salt = 0: a pure function of the point, restated on the CPU in
// oracle/rng_objectives.py.  Runs stay reproducible: the noise is a function of (x, call number).
__device__ __forceinline__ uint64_t mix64(uint64_t z)     // splitmix64 finaliser
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <class XV>
__device__ __forceinline__ uint64_t point_hash(const XV &x, int dim, const Group &g, uint64_t salt)
{
    uint64_t h = g.gl == 0 ? salt * 0x9E3779B97F4A7C15ull : 0ull;
    for (int k = g.gl; k < dim; k += g.G)
        h += mix64((uint64_t)__double_as_longlong(x[k]) + (uint64_t)(k + 1) * 0x9E3779B97F4A7C15ull);
    for (int o = g.G >> 1; o > 0; o >>= 1) h += __shfl_xor_sync(g.mask, h, o);
    return h;
}
__device__ __forceinline__ double point_noise(uint64_t h, int k, uint32_t tag)
{
    Philox ph;
    ph.k0 = 0x53444131u;   // "SDA1"
    ph.k1 = tag;
    uint32_t o[4];
    ph.block((uint32_t)h, (uint32_t)(h >> 32), (uint32_t)k, 0u, o);
    return u53(o[0], o[1]);
}

// ---- data tables of the fixed-size classes ------------------------------------------------------
__device__ const double kColaD[45] = {   // lower triangle of Cola.d, row i = 1..9, columns j < i
    1.27,
    1.69, 1.43,
    2.04, 2.35, 2.43,
    3.09, 3.18, 3.26, 2.85,
    3.20, 3.22, 3.27, 2.88, 1.55,
    2.86, 2.56, 2.58, 2.59, 3.12, 3.06,
    3.17, 3.18, 3.18, 3.12, 1.31, 1.64, 3.00,
    3.21, 3.18, 3.18, 3.17, 1.70, 1.36, 2.95, 1.32,
    2.38, 2.31, 2.42, 1.94, 2.85, 2.81, 2.56, 2.91, 2.97};
__device__ const double kEckerleA[35] = {
    1.5750000E-04, 1.6990000E-04, 2.3500000E-04, 3.1020000E-04, 4.9170000E-04, 8.7100000E-04,
    1.7418000E-03, 4.6400000E-03, 6.5895000E-03, 9.7302000E-03, 1.4900200E-02, 2.3731000E-02,
    4.0168300E-02, 7.1255900E-02, 1.2644580E-01, 2.0734130E-01, 2.9023660E-01, 3.4456230E-01,
    3.6980490E-01, 3.6685340E-01, 3.1067270E-01, 2.0781540E-01, 1.1643540E-01, 6.1676400E-02,
    3.3720000E-02, 1.9402300E-02, 1.1783100E-02, 7.4357000E-03, 2.2732000E-03, 8.8000000E-04,
    4.5790000E-04, 2.3450000E-04, 1.5860000E-04, 1.1430000E-04, 7.1000000E-05};
__device__ const double kEckerleB[35] = {
    400.0, 405.0, 410.0, 415.0, 420.0, 425.0, 430.0, 435.0, 436.5, 438.0, 439.5, 441.0,
    442.5, 444.0, 445.5, 447.0, 448.5, 450.0, 451.5, 453.0, 454.5, 456.0, 457.5, 459.0,
    460.5, 462.0, 463.5, 465.0, 470.0, 475.0, 480.0, 485.0, 490.0, 495.0, 500.0};
__device__ const double kHart3A[12] = {3.0, 10., 30., 0.1, 10., 35., 3.0, 10., 30., 0.1, 10., 35.};
__device__ const double kHart3P[12] = {0.3689, 0.1170, 0.2673, 0.4699, 0.4387, 0.7470,
                                       0.1091, 0.8732, 0.5547, 0.03815, 0.5743, 0.8828};
__device__ const double kHart6A[24] = {10., 3., 17., 3.5, 1.7, 8., 0.05, 10., 17., 0.1, 8., 14.,
                                       3., 3.5, 1.7, 10., 17., 8., 17., 8., 0.05, 10., 0.1, 14.};
__device__ const double kHart6P[24] = {0.1312, 0.1696, 0.5569, 0.0124, 0.8283, 0.5886,
                                       0.2329, 0.4135, 0.8307, 0.3736, 0.1004, 0.9991,
                                       0.2348, 0.1451, 0.3522, 0.2883, 0.3047, 0.665,
                                       0.4047, 0.8828, 0.8732, 0.5743, 0.1091, 0.0381};
__device__ const double kHartC[4] = {1.0, 1.2, 3.0, 3.2};
__device__ const double kJudgeC[20] = {4.284, 4.149, 3.877, 0.533, 2.211, 2.389, 2.145, 3.231, 1.998, 1.379,
                                       2.106, 1.428, 1.011, 2.179, 2.858, 1.388, 1.651, 1.593, 1.046, 2.152};
__device__ const double kJudgeA[20] = {0.286, 0.973, 0.384, 0.276, 0.973, 0.543, 0.957, 0.948, 0.543, 0.797,
                                       0.936, 0.889, 0.006, 0.828, 0.399, 0.617, 0.939, 0.784, 0.072, 0.889};
__device__ const double kJudgeB[20] = {0.645, 0.585, 0.310, 0.058, 0.455, 0.779, 0.259, 0.202, 0.028, 0.099,
                                       0.142, 0.296, 0.175, 0.180, 0.842, 0.039, 0.103, 0.620, 0.158, 0.704};
__device__ const double kKowalikA[11] = {4.0, 2.0, 1.0, 1 / 2.0, 1 / 4.0, 1 / 6.0, 1 / 8.0, 1 / 10.0,
                                         1 / 12.0, 1 / 14.0, 1 / 16.0};
__device__ const double kKowalikB[11] = {0.1957, 0.1947, 0.1735, 0.1600, 0.0844, 0.0627,
                                         0.0456, 0.0342, 0.0323, 0.0235, 0.0246};
__device__ const double kLangA[5] = {3, 5, 2, 1, 7};
__device__ const double kLangB[5] = {5, 2, 1, 4, 9};
__device__ const double kLangC[5] = {1, 2, 5, 2, 3};
// The data vector y of the two De Villiers-Glasser fits is a constant of the class (go_funcs_D.py:406-407,
// :452-454: recomputed by the reference at every call).  Tabulated from the reference's own numpy
// expressions (t = 0.1 * arange(n)): bit-identical to what the reference subtracts, and half of the
// transcendental work of an evaluation gone -- DeVilliersGlasser02 fails 87 % of its suite runs, i.e.
// its 1e6-evaluation chains are the longest pole of the whole suite (17.6 s per chain before).
__constant__ double kDVG01Y[24] = {
    59.052474262824902, 54.425183696648972, 44.044493267698726, 28.575486882609923,
    9.2362884771820877, -12.287935792260145, -33.983362745094588, -53.687834827260026,
    -69.297788526979815, -78.982792932874744, -81.386703064317203, -75.794405771521681,
    -62.245213293284969, -41.578069950386947, -15.399601629767911, 14.026817512187025,
    43.965013627056173, 71.449306590735944, 93.566642508795567, 107.75183312309555,
    112.06718072085815, 105.43746034050714, 87.813812136936633, 60.245495279237169};
__constant__ double kDVG02Y[16] = {
    0, 25.652980123570309, 40.985849081995291, 45.968797241301942,
    44.800862962842317, 40.224284605703971, 33.489336334181722, 25.171629131828926,
    15.612834465347113, 5.0941751865351943, -6.1031686956456124, -17.680274167985303,
    -29.318700343916561, -40.685106311532131, -51.438705611090981, -61.240115692605052};
__device__ const double kMeyerA[16] = {3.478E+04, 2.861E+04, 2.365E+04, 1.963E+04, 1.637E+04, 1.372E+04,
                                       1.154E+04, 9.744E+03, 8.261E+03, 7.030E+03, 6.005E+03, 5.147E+03,
                                       4.427E+03, 3.820E+03, 3.307E+03, 2.872E+03};
__device__ const double kOddSquareB[20] = {1, 1.3, 0.8, -0.4, -1.3, 1.6, -0.2, -0.6, 0.5, 1.4,
                                           1, 1.3, 0.8, -0.4, -1.3, 1.6, -0.2, -0.6, 0.5, 1.4};
__device__ const double kPowerSumB[4] = {8.0, 18.0, 44.0, 114.0};
__device__ const double kRat1A[15] = {16.08, 33.83, 65.80, 97.20, 191.55, 326.20, 386.87, 520.53,
                                      590.03, 651.92, 724.93, 699.56, 689.96, 637.56, 717.41};
__device__ const double kRat2A[9] = {8.93, 10.8, 18.59, 22.33, 39.35, 56.11, 61.73, 64.62, 67.08};
__device__ const double kRat2B[9] = {9., 14., 21., 28., 42., 57., 63., 70., 79.};
__device__ const double kShekelA[40] = {4.0, 4.0, 4.0, 4.0, 1.0, 1.0, 1.0, 1.0, 8.0, 8.0, 8.0, 8.0,
                                        6.0, 6.0, 6.0, 6.0, 3.0, 7.0, 3.0, 7.0, 2.0, 9.0, 2.0, 9.0,
                                        5.0, 5.0, 3.0, 3.0, 8.0, 1.0, 8.0, 1.0, 6.0, 2.0, 6.0, 2.0,
                                        7.0, 3.6, 7.0, 3.6};
__device__ const double kShekelC[10] = {0.1, 0.2, 0.2, 0.4, 0.4, 0.6, 0.3, 0.7, 0.5, 0.5};
__device__ const double kThurberA[37] = {
    80.574, 84.248, 87.264, 87.195, 89.076, 89.608, 89.868, 90.101, 92.405, 95.854, 100.696, 101.06,
    401.672, 390.724, 567.534, 635.316, 733.054, 759.087, 894.206, 990.785, 1090.109, 1080.914,
    1122.643, 1178.351, 1260.531, 1273.514, 1288.339, 1327.543, 1353.863, 1414.509, 1425.208,
    1421.384, 1442.962, 1464.350, 1468.705, 1447.894, 1457.628};
__device__ const double kThurberB[37] = {
    -3.067, -2.981, -2.921, -2.912, -2.840, -2.797, -2.702, -2.699, -2.633, -2.481, -2.363, -2.322,
    -1.501, -1.460, -1.274, -1.212, -1.100, -1.046, -0.915, -0.714, -0.566, -0.545, -0.400, -0.309,
    -0.109, -0.103, 0.010, 0.119, 0.377, 0.790, 0.963, 1.006, 1.115, 1.572, 1.841, 2.047, 2.200};

// Shekel family: -sum_i 1 / (sum_j (x_j - A_ij)^2 + C_i), m rows
template <class XV>
__device__ __forceinline__ double eval_shekel(const XV &x, int m)
{
    double s = 0.0;
    for (int i = 0; i < m; ++i) {
        double d = 0.0;
        for (int j = 0; j < 4; ++j) d += sq(x[j] - kShekelA[4 * i + j]);
        s += 1 / (d + kShekelC[i]);
    }
    return -s;
}

// Hartmann family: -sum_i c_i exp(-sum_j a_ij (x_j - p_ij)^2)
template <class XV>
__device__ __forceinline__ double eval_hartmann(const XV &x, int n, const double *A, const double *P)
{
    double s = 0.0;
    for (int i = 0; i < 4; ++i) {
        double d = 0.0;
        for (int j = 0; j < n; ++j) d += A[n * i + j] * sq(x[j] - P[n * i + j]);
        s += kHartC[i] * exp(-d);
    }
    return -s;
}

// Levy03 / Penalty01 share sum_{i<N-1} (y_i - 1)^2 (1 + 10 sin^2(pi y_{i+1})), y = 1 + (x + c) / 4
template <class XV>
__device__ __forceinline__ double levy_chain(const XV &x, int dim, double c, const Group &g)
{
    double a = 0.0;
    for (int k = g.gl; k + 1 < dim; k += g.G) {
        const double y0 = 1.0 + (x[k] + c) / 4.0, y1 = 1.0 + (x[k + 1] + c) / 4.0;
        const double s1 = sin(kPi * y1);
        a += sq(y0 - 1.0) * (1.0 + 10.0 * (s1 * s1));
    }
    return group_sum(a, g);
}

template <class XV>
__device__ __noinline__ double eval_catalogue(int fid, const XV &x, int dim, const Group &g,
                                             unsigned long long salt)
{
    const int lane = g.gl, S = g.G;
    double a = 0.0, b = 0.0, c = 0.0;
    switch (fid) {
    case SDA_FN_AMGM: {  // go_funcs_A.py  ((sum/N) - prod^(1/N))^2
        b = 1.0;
        for (int k = lane; k < dim; k += S) { a += x[k]; b *= x[k]; }
        const double f1 = group_sum(a, g) / dim, f2 = pow(group_prod(b, g), 1.0 / dim);
        return sq(f1 - f2);
    }
    case SDA_FN_ALPINE02:  // prod(sqrt(x) sin(x))
        b = 1.0;
        for (int k = lane; k < dim; k += S) b *= sqrt(x[k]) * sin(x[k]);
        return group_prod(b, g);
    case SDA_FN_BARTELS_CONN: { const double p = x[0], q = x[1]; return fabs(p * p + q * q + p * q) + fabs(sin(p)) + fabs(cos(q)); }
    case SDA_FN_BIGGS_EXP02:
        for (int i = 1; i <= 10; ++i) {
            const double t = i * 0.1, y = exp(-t) - 5 * exp(-10 * t);
            a += sq(exp(-t * x[0]) - 5 * exp(-t * x[1]) - y);
        }
        return a;
    case SDA_FN_BIGGS_EXP03:
        for (int i = 1; i <= 10; ++i) {
            const double t = i * 0.1, y = exp(-t) - 5 * exp(-10 * t);
            a += sq(exp(-t * x[0]) - x[2] * exp(-t * x[1]) - y);
        }
        return a;
    case SDA_FN_BIGGS_EXP04:
        for (int i = 1; i <= 10; ++i) {
            const double t = i * 0.1, y = exp(-t) - 5 * exp(-10 * t);
            a += sq(x[2] * exp(-t * x[0]) - x[3] * exp(-t * x[1]) - y);
        }
        return a;
    case SDA_FN_BIGGS_EXP05:
        for (int i = 1; i <= 11; ++i) {
            const double t = i * 0.1, y = exp(-t) - 5 * exp(-10 * t) + 3 * exp(-4 * t);
            a += sq(x[2] * exp(-t * x[0]) - x[3] * exp(-t * x[1]) + 3 * exp(-t * x[4]) - y);
        }
        return a;
    case SDA_FN_BOHACHEVSKY2: { const double p = x[0], q = x[1]; return p * p + 2 * (q * q) - 0.3 * cos(3 * kPi * p) * cos(4 * kPi * q) + 0.3; }
    case SDA_FN_BOHACHEVSKY3: { const double p = x[0], q = x[1]; return p * p + 2 * (q * q) - 0.3 * cos(3 * kPi * p + 4 * kPi * q) + 0.3; }
    case SDA_FN_BOX_BETTS:
        for (int i = 1; i <= 10; ++i) {
            const double w = -0.1 * i;
            const double t = exp(w * x[0]) - exp(w * x[1]) - (exp(w) - exp(-(double)i)) * x[2];
            a += t * t;
        }
        return a;
    case SDA_FN_BRANIN01: {
        const double p = x[0], q = x[1];
        return sq(q - (5.1 / (4 * (kPi * kPi))) * (p * p) + 5 * p / kPi - 6) + 10 * (1 - 1 / (8 * kPi)) * cos(p) + 10;
    }
    case SDA_FN_BRANIN02: {
        const double p = x[0], q = x[1];
        return sq(q - (5.1 / (4 * (kPi * kPi))) * (p * p) + 5 * p / kPi - 6)
               + 10 * (1 - 1 / (8 * kPi)) * cos(p) * cos(q) + log(p * p + q * q + 1.0) + 10;
    }
    case SDA_FN_BUKIN02: { const double p = x[0], q = x[1]; return 100 * (q * q - 0.01 * (p * p) + 1.0) + 0.01 * sq(p + 10.0); }
    case SDA_FN_BUKIN04: return 100 * (x[1] * x[1]) + 0.01 * fabs(x[0] + 10);
    case SDA_FN_CARROM_TABLE: {
        const double p = x[0], q = x[1];
        const double u = cos(p) * cos(q), v = sqrt(p * p + q * q);
        return -sq(u * exp(fabs(1 - v / kPi))) / 30.;
    }
    case SDA_FN_CHICHINADZE: {
        const double p = x[0], q = x[1];
        return p * p - 12 * p + 11 + 10 * cos(kPi * p / 2) + 8 * sin(5 * kPi * p / 2)
               - 1.0 / sqrt(5.0) * exp(-sq(q - 0.5) / 2);
    }
    case SDA_FN_COLA: {  // 10 planar points: (0,0), (x0,0), (x[2i-3], x[2i-2]) for i >= 2
        int t = 0;
        for (int i = 1; i < 10; ++i) {
            const double xi = (i == 1) ? x[0] : x[2 * i - 3], yi = (i == 1) ? 0.0 : x[2 * i - 2];
            for (int j = 0; j < i; ++j, ++t) {
                const double xj = (j == 0) ? 0.0 : (j == 1) ? x[0] : x[2 * j - 3];
                const double yj = (j < 2) ? 0.0 : x[2 * j - 2];
                a += sq(sqrt(sq(xi - xj) + sq(yi - yj)) - kColaD[t]);
            }
        }
        return a;
    }
    case SDA_FN_COLVILLE:
        return 100 * sq(x[0] - x[1] * x[1]) + sq(1 - x[0]) + sq(1 - x[2]) + 90 * sq(x[3] - x[2] * x[2])
               + 10.1 * (sq(x[1] - 1) + sq(x[3] - 1)) + 19.8 * (x[1] - 1) * (x[3] - 1);
    case SDA_FN_CORANA: {
        const double d[4] = {1., 1000., 10., 100.};
        for (int j = 0; j < 4; ++j) {
            const double v = x[j];
            const double zj = floor(fabs(v / 0.2) + 0.49999) * np_sign(v) * 0.2;
            if (fabs(v - zj) < 0.05) a += 0.15 * sq(zj - 0.05 * np_sign(zj)) * d[j];
            else a += d[j] * v * v;
        }
        return a;
    }
    case SDA_FN_CROSS_IN_TRAY: {
        const double p = x[0], q = x[1];
        return -0.0001 * pow(fabs(sin(p) * sin(q) * exp(fabs(100 - sqrt(p * p + q * q) / kPi))) + 1, 0.1);
    }
    case SDA_FN_CROSS_LEG_TABLE: {
        const double p = x[0], q = x[1];
        const double u = 100 - sqrt(p * p + q * q) / kPi, v = sin(p) * sin(q);
        return -pow(fabs(v * exp(fabs(u))) + 1, -0.1);
    }
    case SDA_FN_CROWNED_CROSS: {
        const double p = x[0], q = x[1];
        const double u = 100 - sqrt(p * p + q * q) / kPi, v = sin(p) * sin(q);
        return 0.0001 * pow(fabs(v * exp(fabs(u))) + 1, 0.1);
    }
    case SDA_FN_CSENDES:  // sum x^6 (2 + sin(1/x)); nan at x = 0 like the reference
        for (int k = lane; k < dim; k += S) { const double v = x[k]; a += pow(v, 6.0) * (2.0 + sin(1.0 / v)); }
        return group_sum(a, g);
    case SDA_FN_DAMAVANDI: {
        const double p = x[0], q = x[1];
        const double num = sin(kPi * (p - 2.0)) * sin(kPi * (q - 2.0));
        const double den = (kPi * kPi) * (p - 2.0) * (q - 2.0);
        const double f1 = 1.0 - pow(fabs(num / den), 5.0);
        const double f2 = 2 + sq(p - 7.0) + 2 * sq(q - 7.0);
        return f1 * f2;
    }
    case SDA_FN_DE_VILLIERS_GLASSER01:
        for (int i = 0; i < 24; ++i) {
            const double t = 0.1 * i;
            a += sq(x[0] * pow(x[1], t) * sin(x[2] * t + x[3]) - kDVG01Y[i]);
        }
        return a;
    case SDA_FN_DE_VILLIERS_GLASSER02:
    {
        const double e4 = exp(x[4]);
        for (int i = 0; i < 16; ++i) {
            const double t = 0.1 * i;
            a += sq(x[0] * pow(x[1], t) * tanh(x[2] * t + sin(x[3] * t)) * cos(t * e4) - kDVG02Y[i]);
        }
        return a;
    }
    case SDA_FN_DEB03:
        for (int k = lane; k < dim; k += S) a += pow(sin(5 * kPi * (pow(x[k], 0.75) - 0.05)), 6.0);
        return -(1.0 / dim) * group_sum(a, g);
    case SDA_FN_DECANOMIAL: {
        const double p = x[0], q = x[1];
        const double v1 = pow(q, 4.0) + 12 * cube(q) + 54 * (q * q) + 108 * q + 81.0;
        double v2 = pow(p, 10.) - 20 * pow(p, 9.0) + 180 * pow(p, 8.0) - 960 * pow(p, 7.0);
        v2 += 3360 * pow(p, 6.0) - 8064 * pow(p, 5.0) + 13340 * pow(p, 4.0);
        v2 += -15360 * cube(p) + 11520 * (p * p) - 5120 * p + 2624;
        return 0.001 * sq(fabs(v1) + fabs(v2));
    }
    case SDA_FN_DECEPTIVE:
        for (int k = lane; k < dim; k += S) {
            const double al = (k + 1.0) / (dim + 1.0), v = x[k];
            double gi;
            if (v <= 0.0) gi = v;
            else if (v < 0.8 * al) gi = -v / al + 0.8;
            else if (v < al) gi = 5.0 * v / al - 4.0;
            else if (v < (1.0 + 4 * al) / 5.0) gi = 5.0 * (v - al) / (al - 1.0) + 1.0;
            else if (v <= 1.0) gi = (v - 1.0) / (1.0 - al) + 4.0 / 5.0;
            else gi = v - 1.0;
            a += gi;
        }
        return -sq((1.0 / dim) * group_sum(a, g));
    case SDA_FN_DECKKERS_AARTS: {
        const double p2 = x[0] * x[0], q2 = x[1] * x[1], r = p2 + q2;
        return 1.e5 * p2 + q2 - r * r + 1.e-5 * quart(r);
    }
    case SDA_FN_DEFLECTED_CORRUGATED_SPRING: {
        for (int k = lane; k < dim; k += S) a += sq(x[k] - 5.0);
        const double s = group_sum(a, g);
        return -cos(5.0 * sqrt(s)) + 0.1 * s;
    }
    case SDA_FN_DOLAN:
        return fabs((x[0] + 1.7 * x[1]) * sin(x[0]) - 1.5 * x[2] - 0.1 * x[3] * cos(x[3] + x[4] - x[0])
                    + 0.2 * (x[4] * x[4]) - x[1] - 1);
    case SDA_FN_ECKERLE4:
        for (int i = 0; i < 35; ++i) {
            const double v = x[0] / x[1] * exp(-sq(kEckerleB[i] - x[2]) / (2 * (x[1] * x[1])));
            a += sq(kEckerleA[i] - v);
        }
        return a;
    case SDA_FN_EGG_HOLDER:
        for (int k = lane; k + 1 < dim; k += S) {
            const double u = x[k], v = x[k + 1];
            a += -(v + 47) * sin(sqrt(fabs(v + u / 2. + 47))) - u * sin(sqrt(fabs(u - (v + 47))));
        }
        return group_sum(a, g);
    case SDA_FN_EL_ATTAR_VIDYASAGAR_DUTTA: {
        const double p = x[0], q = x[1];
        return sq(p * p + q - 10) + sq(p + q * q - 7) + sq(p * p + cube(q) - 1);
    }
    case SDA_FN_EXP2:
        for (int i = 0; i < 10; ++i) {
            const double t = exp(-i * x[0] / 10.) - 5 * exp(-i * x[1] / 10.) - exp(-i / 10.) + 5 * exp(-(double)i);
            a += t * t;
        }
        return a;
    case SDA_FN_FREUDENSTEIN_ROTH: {
        const double p = x[0], q = x[1];
        return sq(-13.0 + p + ((5.0 - q) * q - 2.0) * q) + sq(-29.0 + p + ((q + 1.0) * q - 14.0) * q);
    }
    case SDA_FN_GEAR: return sq(1. / 6.931 - floor(x[0]) * floor(x[1]) / floor(x[2]) / floor(x[3]));
    case SDA_FN_GIUNTA:
        for (int k = lane; k < dim; k += S) {
            const double arg = 16 * x[k] / 15.0 - 1, s1 = sin(arg);
            a += s1 + s1 * s1 + sin(4 * arg) / 50.;
        }
        return 0.6 + group_sum(a, g);
    case SDA_FN_GULF:
        for (int i = lane + 1; i <= 99; i += S) {
            const double u = 25 + pow(-50 * log(i / 100.), 2 / 3.);
            const double v = exp(-(pow(fabs(u - x[1]), x[2]) / x[0])) - i / 100.;
            a += v * v;
        }
        return group_sum(a, g);
    case SDA_FN_HANSEN:
        for (int i = 0; i < 5; ++i) {
            a += (i + 1) * cos(i * x[0] + i + 1);
            b += (i + 1) * cos((i + 2) * x[1] + i + 1);
        }
        return a * b;
    case SDA_FN_HARTMANN3: return eval_hartmann(x, 3, kHart3A, kHart3P);
    case SDA_FN_HARTMANN6: return eval_hartmann(x, 6, kHart6A, kHart6P);
    case SDA_FN_HELICAL_VALLEY: {
        const double r = sqrt(x[0] * x[0] + x[1] * x[1]);
        const double theta = 1 / (2. * kPi) * atan2(x[1], x[0]);
        return x[2] * x[2] + 100 * (sq(x[2] - 10 * theta) + sq(r - 1));
    }
    case SDA_FN_HOLDER_TABLE: {
        const double p = x[0], q = x[1];
        return -fabs(sin(p) * cos(q) * exp(fabs(1 - sqrt(p * p + q * q) / kPi)));
    }
    case SDA_FN_HOSAKI: {
        const double p = x[0], q = x[1];
        const double v = 1 - 8 * p + 7 * (p * p) - 7 / 3. * cube(p) + 0.25 * quart(p);
        return v * (q * q) * exp(-q);
    }
    case SDA_FN_INFINITY:
        for (int k = lane; k < dim; k += S) { const double v = x[k]; a += pow(v, 6.0) * (sin(1.0 / v) + 2.0); }
        return group_sum(a, g);
    case SDA_FN_JENNRICH_SAMPSON:
        for (int i = 1; i <= 10; ++i) a += sq(2 + 2 * i - (exp(i * x[0]) + exp(i * x[1])));
        return a;
    case SDA_FN_JUDGE:
        for (int i = 0; i < 20; ++i) a += sq((x[0] + x[1] * kJudgeA[i] + (x[1] * x[1]) * kJudgeB[i]) - kJudgeC[i]);
        return a;
    case SDA_FN_KATSUURA:  // prod_j (1 + (j+1) sum_{k=1..32} rint(2^k x_j) 2^-k)
        b = 1.0;
        for (int j = lane; j < dim; j += S) {
            double s = 0.0, p2 = 1.0;
            for (int k = 1; k <= 32; ++k) { p2 *= 2.0; s += rint(p2 * x[j]) * (1.0 / p2); }
            b *= s * (j + 1.0) + 1;
        }
        return group_prod(b, g);
    case SDA_FN_KEANE: {
        const double p = x[0], q = x[1];
        return sq(sin(p - q)) * sq(sin(p + q)) / sqrt(p * p + q * q);
    }
    case SDA_FN_KOWALIK:
        for (int i = 0; i < 11; ++i) {
            const double ai = kKowalikA[i];
            a += sq(kKowalikB[i] - (x[0] * (ai * ai + ai * x[1]) / (ai * ai + ai * x[2] + x[3])));
        }
        return a;
    case SDA_FN_LANGERMANN:
        for (int i = 0; i < 5; ++i) {
            const double r = sq(x[0] - kLangA[i]) + sq(x[1] - kLangB[i]);
            a += kLangC[i] * exp(-(1 / kPi) * r) * cos(kPi * r);
        }
        return -a;
    case SDA_FN_LENNARD_JONES: {  // pairs (i, j > i) of the dim/3 atoms; lane owns atom i
        const int n = dim / 3;
        for (int i = lane; i + 1 < n; i += S) {
            for (int j = i + 1; j < n; ++j) {
                const double xd = x[3 * i] - x[3 * j], yd = x[3 * i + 1] - x[3 * j + 1], zd = x[3 * i + 2] - x[3 * j + 2];
                const double ed = xd * xd + yd * yd + zd * zd, ud = ed * ed * ed;
                if (ed > 0.0) a += (1.0 / ud - 2.0) / ud;
            }
        }
        return group_sum(a, g);
    }
    case SDA_FN_LEVY03: {
        const double y0 = 1 + (x[0] - 1) / 4, yl = 1 + (x[dim - 1] - 1) / 4;
        return sq(sin(kPi * y0)) + levy_chain(x, dim, -1.0, g) + sq(yl - 1);
    }
    case SDA_FN_LEVY05:
        for (int i = 1; i <= 5; ++i) {
            a += i * cos((i - 1) * x[0] + i);
            b += i * cos((i + 1) * x[1] + i);
        }
        return a * b + sq(x[0] + 1.42513) + sq(x[1] + 0.80032);
    case SDA_FN_MEYER:
        for (int i = 0; i < 16; ++i) a += sq(kMeyerA[i] - x[0] * exp(x[1] / ((50.0 + 5.0 * i) + x[2])));
        return a;
    case SDA_FN_MICHALEWICZ:
        for (int k = lane; k < dim; k += S) { const double v = x[k]; a += sin(v) * pow(sin((k + 1) * (v * v) / kPi), 20.0); }
        return -group_sum(a, g);
    case SDA_FN_MIELE_CANTRELL:
        return quart(exp(-x[0]) - x[1]) + 100 * pow(x[1] - x[2], 6.0) + quart(tan(x[2] - x[3])) + pow(x[0], 8.0);
    case SDA_FN_MISHRA01: {
        for (int k = lane; k + 1 < dim; k += S) a += x[k];
        const double xn = dim - group_sum(a, g);
        return pow(1 + xn, xn);
    }
    case SDA_FN_MISHRA02: {
        for (int k = lane; k + 1 < dim; k += S) a += (x[k] + x[k + 1]) / 2.0;
        const double xn = dim - group_sum(a, g);
        return pow(1 + xn, xn);
    }
    case SDA_FN_MISHRA03: { const double p = x[0], q = x[1]; return 0.01 * (p + q) + sqrt(fabs(cos(sqrt(fabs(p * p + q * q))))); }
    case SDA_FN_MISHRA04: { const double p = x[0], q = x[1]; return 0.01 * (p + q) + sqrt(fabs(sin(sqrt(fabs(p * p + q * q))))); }
    case SDA_FN_MISHRA05: {
        const double p = x[0], q = x[1];
        return 0.01 * p + 0.1 * q + sq(sq(sin(sq(cos(p) + cos(q)))) + sq(cos(sq(sin(p) + sin(q)))) + p);
    }
    case SDA_FN_MISHRA06: {
        const double p = x[0], q = x[1];
        const double w = 0.1 * (sq(p - 1) + sq(q - 1));
        const double u = sq(cos(p) + cos(q)), v = sq(sin(p) + sin(q));
        return w - log(sq(sq(sin(u)) - sq(cos(v)) + p));
    }
    case SDA_FN_MISHRA07: {
        b = 1.0;
        double fact = 1.0;
        for (int k = 2; k <= dim; ++k) fact *= k;
        for (int k = lane; k < dim; k += S) b *= x[k];
        return sq(group_prod(b, g) - fact);
    }
    case SDA_FN_MISHRA08: {
        const double p = x[0], q = x[1];
        double v = fabs(pow(p, 10.0) - 20 * pow(p, 9.0) + 180 * pow(p, 8.0) - 960 * pow(p, 7.0) + 3360 * pow(p, 6.0)
                        - 8064 * pow(p, 5.0) + 13340 * pow(p, 4.0) - 15360 * cube(p) + 11520 * (p * p) - 5120 * p + 2624);
        v += fabs(pow(q, 4.0) + 12 * cube(q) + 54 * (q * q) + 108 * q + 81);
        return 0.001 * (v * v);
    }
    case SDA_FN_MISHRA09: {
        const double p = x[0], q = x[1], r = x[2];
        const double u = 2 * cube(p) + 5 * p * q + 4 * r - 2 * (p * p) * r - 18;
        const double v = p + cube(q) + p * (q * q) + p * (r * r) - 22.0;
        const double w = 8 * (p * p) + 2 * q * r + 2 * (q * q) + 3 * cube(q) - 52;
        return sq(u * w * (v * v) + u * v * (w * w) + v * v + sq(p + q - r));
    }
    case SDA_FN_MISHRA10: {
        const double p = trunc(x[0]), q = trunc(x[1]);
        return sq((p + q) - p * q);
    }
    case SDA_FN_MISHRA11: {  // ((1/N) sum|x| - prod|x| ** 1.0 / N)^2 : the power binds before the division
        b = 1.0;
        for (int k = lane; k < dim; k += S) { const double v = fabs(x[k]); a += v; b *= v; }
        return sq((1.0 / dim) * group_sum(a, g) - group_prod(b, g) / dim);
    }
    case SDA_FN_MULTI_MODAL:
        b = 1.0;
        for (int k = lane; k < dim; k += S) { const double v = fabs(x[k]); a += v; b *= v; }
        return group_sum(a, g) * group_prod(b, g);
    case SDA_FN_NEEDLE_EYE: {
        for (int k = lane; k < dim; k += S) {
            const double v = fabs(x[k]);
            if (v >= 0.0001) { b = 1.0; a += 100.0 + v; }
            else a += 1.0;
        }
        double f = group_sum(a, g);
        const double fp = group_max(b, g);
        if (fp < 1e-6) f = f / dim;
        return f;
    }
    case SDA_FN_NEW_FUNCTION01: { const double p = x[0], q = x[1]; return sqrt(fabs(cos(sqrt(fabs(p * p + q))))) + 0.01 * (p + q); }
    case SDA_FN_NEW_FUNCTION02: { const double p = x[0], q = x[1]; return sqrt(fabs(sin(sqrt(fabs(p * p + q))))) + 0.01 * (p + q); }
    case SDA_FN_ODD_SQUARE: {
        for (int k = lane; k < dim; k += S) { const double t = sq(x[k] - kOddSquareB[k % 20]); a += t; b = fmax(b, t); }
        const double h = group_sum(a, g), d = dim * group_max(b, g);
        return -exp(-d / (2.0 * kPi)) * cos(kPi * d) * (1.0 + 0.02 * h / (d + 0.01));
    }
    case SDA_FN_PARSOPOULOS: return sq(cos(x[0])) + sq(sin(x[1]));
    case SDA_FN_PAVIANI:
        b = 1.0;
        for (int k = lane; k < dim; k += S) { const double v = x[k]; a += sq(log(v - 2)) + sq(log(10.0 - v)); b *= v; }
        return group_sum(a, g) - pow(group_prod(b, g), 0.2);
    case SDA_FN_PEN_HOLDER: {
        const double p = x[0], q = x[1];
        const double w = fabs(1. - (sqrt(p * p + q * q) / kPi));
        const double v = cos(p) * cos(q) * exp(w);
        return -exp(-1.0 / fabs(v));
    }
    case SDA_FN_PENALTY01: {
        for (int k = lane; k < dim; k += S) { const double v = fabs(x[k]); if (v > 10.0) a += 100.0 * quart(v - 10.0); }
        const double u = group_sum(a, g);
        const double y0 = 1.0 + (x[0] + 1.0) / 4.0, yl = 1.0 + (x[dim - 1] + 1.0) / 4.0;
        return u + (kPi / 30.0) * (10.0 * sq(sin(kPi * y0)) + levy_chain(x, dim, 1.0, g) + sq(yl - 1));
    }
    case SDA_FN_PENALTY02: {
        for (int k = lane; k < dim; k += S) { const double v = fabs(x[k]); if (v > 5.0) a += 100.0 * quart(v - 5.0); }
        const double u = group_sum(a, g);
        for (int k = lane; k + 1 < dim; k += S) b += sq(x[k] - 1.0) * (1.0 + sq(sin(3 * kPi * x[k + 1])));
        const double mid = group_sum(b, g), xl = x[dim - 1];
        return u + 0.1 * (10 * sq(sin(3.0 * kPi * x[0])) + mid + sq(xl - 1) * (1 + sq(sin(2 * kPi * xl))));
    }
    case SDA_FN_PERM_FUNCTION01:  // sum_k (sum_j (j^k + 0.5)((x_j / j)^k - 1))^2, k, j = 1..N
        for (int k = lane + 1; k <= dim; k += S) {
            double in = 0.0;
            for (int j = 1; j <= dim; ++j) in += (pow((double)j, (double)k) + 0.5) * (pow(x[j - 1] / j, (double)k) - 1);
            a += in * in;
        }
        return group_sum(a, g);
    case SDA_FN_PERM_FUNCTION02:  // sum_k (sum_j (j + 10)(x_j^k - (1/j)^k))^2
        for (int k = lane + 1; k <= dim; k += S) {
            double in = 0.0;
            for (int j = 1; j <= dim; ++j) in += (j + 10) * (pow(x[j - 1], (double)k) - pow(1. / j, (double)k));
            a += in * in;
        }
        return group_sum(a, g);
    case SDA_FN_PINTER:
        for (int k = lane; k < dim; k += S) {
            const double i = k + 1.0;
            const double xm = x[k == 0 ? dim - 1 : k - 1], xc = x[k], xp = x[k == dim - 1 ? 0 : k + 1];
            const double A = xm * sin(xc) + sin(xp);
            const double B = xm * xm - 2 * xc + 3 * xp - cos(xc) + 1;
            a += i * (xc * xc);
            b += 20 * i * sq(sin(A));
            c += i * log10(1 + i * (B * B));
        }
        return group_sum(a, g) + group_sum(b, g) + group_sum(c, g);
    case SDA_FN_POWELL:
        return sq(x[0] + 10 * x[1]) + 5 * sq(x[2] - x[3]) + quart(x[1] - 2 * x[2]) + 10 * quart(x[0] - x[3]);
    case SDA_FN_POWER_SUM:
    {   // sum_k (sum_j x_j^k - b_k)^2, k = 1..4: x ** k as products (pow(x, 1) and pow(x, 2) are exactly x
        // and x*x; x^3, x^4 within 1 ulp of it) -- 16 pow calls per evaluation were 80 % of this
        // job's 14.7 s per 1e6-evaluation chain
        double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
        for (int j = 0; j < dim; ++j) {
            const double v = x[j], v2 = v * v;
            s1 += v; s2 += v2; s3 += v2 * v; s4 += v2 * v2;
        }
        a += sq(s1 - kPowerSumB[0]);
        a += sq(s2 - kPowerSumB[1]);
        a += sq(s3 - kPowerSumB[2]);
        a += sq(s4 - kPowerSumB[3]);
        return a;
    }
    case SDA_FN_PRICE02:
        for (int k = lane; k < dim; k += S) a += sq(sin(x[k]));
        return 1.0 + group_sum(a, g) - 0.1 * exp(-(x[0] * x[0]) - x[1] * x[1]);
    case SDA_FN_PRICE03: { const double p = x[0], q = x[1]; return 100 * sq(q - p * p) + sq(6.4 * sq(q - 0.5) - p - 0.6); }
    case SDA_FN_PRICE04: { const double p = x[0], q = x[1]; return sq(2.0 * q * cube(p) - cube(q)) + sq(6.0 * p - q * q + q); }
    case SDA_FN_QUADRATIC: {
        const double p = x[0], q = x[1];
        return -3803.84 - 138.08 * p - 232.92 * q + 128.08 * (p * p) + 203.64 * (q * q) + 182.25 * p * q;
    }
    case SDA_FN_RATKOWSKY01:
        for (int i = 0; i < 15; ++i) a += sq(kRat1A[i] - x[0] / pow(1 + exp(x[1] - x[2] * (i + 1.0)), 1 / x[3]));
        return a;
    case SDA_FN_RATKOWSKY02:
        for (int i = 0; i < 9; ++i) a += sq(kRat2A[i] - x[0] / (1 + exp(x[1] - x[2] * kRat2B[i])));
        return a;
    case SDA_FN_RIPPLE01:
        for (int k = lane; k < dim; k += S) {
            const double v = x[k];
            const double u = -2.0 * log(2.0) * sq((v - 0.1) / 0.8);
            const double w = pow(sin(5.0 * kPi * v), 6.0) + 0.1 * sq(cos(500.0 * kPi * v));
            a += -exp(u) * w;
        }
        return group_sum(a, g);
    case SDA_FN_RIPPLE25:
        for (int k = lane; k < dim; k += S) {
            const double v = x[k];
            const double u = -2.0 * log(2.0) * sq((v - 0.1) / 0.8);
            a += -exp(u) * pow(sin(5.0 * kPi * v), 6.0);
        }
        return group_sum(a, g);
    case SDA_FN_ROSENBROCK_MODIFIED: {
        const double p = x[0], q = x[1];
        double r = 74 + 100. * sq(q - p * p) + sq(1 - p);
        r -= 400 * exp(-(sq(p + 1.) + sq(q + 1.)) / 0.1);
        return r;
    }
    case SDA_FN_ROTATED_ELLIPSE01: { const double p = x[0], q = x[1]; return 7.0 * (p * p) - 6.0 * sqrt(3.0) * p * q + 13 * (q * q); }
    case SDA_FN_ROTATED_ELLIPSE02: { const double p = x[0], q = x[1]; return p * p - p * q + q * q; }
    case SDA_FN_SARGAN: {  // sum_i N (x_i^2 + 0.4 sum_{j<N-1} x_j x_{j+1})
        for (int k = lane; k + 1 < dim; k += S) b += x[k] * x[k + 1];
        const double cross = 0.4 * group_sum(b, g);
        for (int k = lane; k < dim; k += S) a += dim * (x[k] * x[k] + cross);
        return group_sum(a, g);
    }
    case SDA_FN_SCHAFFER02: {
        const double p2 = x[0] * x[0], q2 = x[1] * x[1];
        return 0.5 + (sq(sin(p2 - q2)) - 0.5) / sq(1 + 0.001 * (p2 + q2));
    }
    case SDA_FN_SCHAFFER03: {
        const double p2 = x[0] * x[0], q2 = x[1] * x[1];
        return 0.5 + (sq(sin(cos(fabs(p2 - q2)))) - 0.5) / sq(1 + 0.001 * (p2 + q2));
    }
    case SDA_FN_SCHAFFER04: {
        const double p2 = x[0] * x[0], q2 = x[1] * x[1];
        return 0.5 + (sq(cos(sin(fabs(p2 - q2)))) - 0.5) / sq(1 + 0.001 * (p2 + q2));
    }
    case SDA_FN_SCHWEFEL02:  // sum_i (sum_{j<=i} x_j)^2
        for (int i = lane; i < dim; i += S) {
            double in = 0.0;
            for (int j = 0; j <= i; ++j) in += x[j];
            a += in * in;
        }
        return group_sum(a, g);
    case SDA_FN_SCHWEFEL04: {
        const double x0 = x[0];
        for (int k = lane; k < dim; k += S) { const double v = x[k]; a += sq(v - 1.0) + sq(x0 - v * v); }
        return group_sum(a, g);
    }
    case SDA_FN_SCHWEFEL06: return fmax(fabs(x[0] + 2 * x[1] - 7), fabs(2 * x[0] + x[1] - 5));
    case SDA_FN_SCHWEFEL36: return -x[0] * x[1] * (72 - 2 * x[0] - 2 * x[1]);
    case SDA_FN_SHEKEL05: return eval_shekel(x, 5);
    case SDA_FN_SHEKEL07: return eval_shekel(x, 7);
    case SDA_FN_SHEKEL10: return eval_shekel(x, 10);
    case SDA_FN_SHUBERT01:
        b = 1.0;
        for (int k = lane; k < dim; k += S) {
            double s = 0.0;
            for (int j = 1; j <= 5; ++j) s += j * cos((j + 1) * x[k] + j);
            b *= s;
        }
        return group_prod(b, g);
    case SDA_FN_SHUBERT03:
        for (int k = lane; k < dim; k += S)
            for (int j = 1; j <= 5; ++j) a += -j * sin((j + 1) * x[k] + j);
        return group_sum(a, g);
    case SDA_FN_SHUBERT04:
        for (int k = lane; k < dim; k += S)
            for (int j = 1; j <= 5; ++j) a += -j * cos((j + 1) * x[k] + j);
        return group_sum(a, g);
    case SDA_FN_STRETCHED_V:
        for (int k = lane; k + 1 < dim; k += S) {
            const double t = x[k + 1] * x[k + 1] + x[k] * x[k];
            a += pow(t, 0.25) * sq(sin(50.0 * pow(t, 0.1) + 1));
        }
        return group_sum(a, g);
    case SDA_FN_TEST_TUBE_HOLDER: {
        const double p = x[0], q = x[1];
        const double u = sin(p) * cos(q), v = (p * p + q * q) / 200;
        return -4 * fabs(u * exp(fabs(cos(v))));
    }
    case SDA_FN_THURBER:
        #pragma unroll 4
        for (int i = 0; i < 37; ++i) {
            const double t = kThurberB[i], t2 = t * t, t3 = t * t * t;
            double v = x[0] + x[1] * t + x[2] * t2 + x[3] * t3;
            v /= 1 + x[4] * t + x[5] * t2 + x[6] * t3;
            a += sq(kThurberA[i] - v);
        }
        return a;
    case SDA_FN_TRECCANI: { const double p = x[0], q = x[1]; return quart(p) + 4.0 * cube(p) + 4.0 * (p * p) + q * q; }
    case SDA_FN_TREFETHEN: {
        const double p = x[0], q = x[1];
        double v = 0.25 * (p * p) + 0.25 * (q * q);
        v += exp(sin(50. * p)) - sin(10 * p + 10 * q);
        v += sin(60 * exp(q));
        v += sin(70 * sin(p));
        v += sin(sin(80 * q));
        return v;
    }
    case SDA_FN_TRIGONOMETRIC01: {  // sum_i (N - sum_j [cos x_j + i (1 - cos x_j - sin x_j)])^2
        for (int i = lane + 1; i <= dim; i += S) {
            double in = 0.0;
            for (int j = 0; j < dim; ++j) { const double cj = cos(x[j]); in += cj + i * (1 - cj - sin(x[j])); }
            a += sq(dim - in);
        }
        return group_sum(a, g);
    }
    case SDA_FN_TRIPOD: {
        const double p = x[0], q = x[1];
        const double p1 = (p >= 0) ? 1.0 : 0.0, p2 = (q >= 0) ? 1.0 : 0.0;
        return p2 * (1.0 + p1) + fabs(p + 50.0 * p2 * (1.0 - 2.0 * p1)) + fabs(q + 50.0 * (1.0 - 2.0 * p2));
    }
    case SDA_FN_URSEM01: return -sin(2 * x[0] - 0.5 * kPi) - 3.0 * cos(x[1]) - 0.5 * x[0];
    case SDA_FN_URSEM03: {
        const double p = x[0], q = x[1];
        const double u = -(sin(2.2 * kPi * p + 0.5 * kPi) * ((2.0 - fabs(p)) / 2.0) * ((3.0 - fabs(p)) / 2));
        const double v = -(sin(2.2 * kPi * q + 0.5 * kPi) * ((2.0 - fabs(q)) / 2) * ((3.0 - fabs(q)) / 2));
        return u + v;
    }
    case SDA_FN_URSEM04: {
        const double p = x[0], q = x[1];
        return -3 * sin(0.5 * kPi * p + 0.5 * kPi) * (2 - sqrt(p * p + q * q)) / 4;
    }
    case SDA_FN_URSEM_WAVES: {
        const double p = x[0], q = x[1];
        const double u = -0.9 * (p * p);
        const double v = (q * q - 4.5 * (q * q)) * p * q;
        const double w = 4.7 * cos(3 * p - (q * q) * (2 + p)) * sin(2.5 * kPi * p);
        return u + v + w;
    }
    case SDA_FN_VENTER_SOBIEZCCZANSKI_SOBIESKI: {
        const double p = x[0], q = x[1];
        const double u = p * p - 100.0 * sq(cos(p));
        const double v = -100.0 * cos(p * p / 30.0) + q * q;
        const double w = -100.0 * sq(cos(q)) - 100.0 * cos(q * q / 30.0);
        return u + v + w;
    }
    case SDA_FN_VINCENT:
        for (int k = lane; k < dim; k += S) a += sin(10.0 * log(x[k]));
        return -group_sum(a, g);
    case SDA_FN_WATSON:
        for (int i = lane; i < 30; i += S) {
            const double ai = i / 29.;
            double t1 = 0.0, t2 = 0.0;
            for (int j = 0; j < 5; ++j) t1 += (j + 1) * pow(ai, (double)j) * x[j + 1];
            for (int k = 0; k < 6; ++k) t2 += pow(ai, (double)k) * x[k];
            a += sq(t1 - t2 * t2 - 1);
        }
        return group_sum(a, g) + x[0] * x[0];
    case SDA_FN_WAYBURN_SEADER01: { const double p = x[0], q = x[1]; return sq(pow(p, 6.0) + quart(q) - 17) + sq(2 * p + q - 4); }
    case SDA_FN_WAYBURN_SEADER02: {
        const double p = x[0], q = x[1];
        return sq(1.613 - 4 * sq(p - 0.3125) - 4 * sq(q - 1.625)) + sq(q - 1);
    }
    case SDA_FN_WEIERSTRASS: {  // sum_i sum_{k<=20} 0.5^k cos(2 pi 3^k (x_i + 0.5)) - N sum_k 0.5^k cos(pi 3^k)
        for (int i = lane; i < dim; i += S) {
            double ak = 1.0, bk = 1.0;
            for (int k = 0; k <= 20; ++k) { a += ak * cos(2 * kPi * bk * (x[i] + 0.5)); ak *= 0.5; bk *= 3.0; }
        }
        double ak = 1.0, bk = 1.0;
        for (int k = 0; k <= 20; ++k) { b += ak * cos(kPi * bk); ak *= 0.5; bk *= 3.0; }
        return group_sum(a, g) - dim * b;
    }
    case SDA_FN_WHITLEY:  // sum_ij t^2/4000 - cos t + 1, t = 100 (x_i^2 - x_j) + (1 - x_j)^2
        for (int j = lane; j < dim; j += S) {
            const double xj = x[j];
            for (int i = 0; i < dim; ++i) {
                const double t = 100.0 * (x[i] * x[i] - xj) + sq(1.0 - xj);
                a += (t * t / 4000.0) - cos(t) + 1.0;
            }
        }
        return group_sum(a, g);
    case SDA_FN_WOLFE: return 4.0 / 3.0 * pow(x[0] * x[0] + x[1] * x[1] - x[0] * x[1], 0.75) + x[2];
    case SDA_FN_XIN_SHE_YANG03: {
        c = 1.0;
        for (int k = lane; k < dim; k += S) { const double v = x[k]; a += pow(v / 15.0, 10.0); b += v * v; c *= sq(cos(v)); }
        return exp(-group_sum(a, g)) - 2 * exp(-group_sum(b, g)) * group_prod(c, g);
    }
    case SDA_FN_XIN_SHE_YANG04: {
        for (int k = lane; k < dim; k += S) { const double v = x[k]; a += sq(sin(v)); b += v * v; c += sq(sin(sqrt(fabs(v)))); }
        return (group_sum(a, g) - exp(-group_sum(b, g))) * exp(-group_sum(c, g));
    }
    case SDA_FN_XOR: {
        const double F11 = x[6] / (1.0 + exp(-x[0] - x[1] - x[4]));
        const double F12 = x[7] / (1.0 + exp(-x[2] - x[3] - x[5]));
        const double F1 = 1.0 / sq(1.0 + exp(-F11 - F12 - x[8]));
        const double F21 = x[6] / (1.0 + exp(-x[4]));
        const double F22 = x[7] / (1.0 + exp(-x[5]));
        const double F2 = 1.0 / sq(1.0 + exp(-F21 - F22 - x[8]));
        const double F31 = x[6] / (1.0 + exp(-x[0] - x[4]));
        const double F32 = x[7] / (1.0 + exp(-x[2] - x[5]));
        const double F3 = sq(1.0 - 1.0 / (1.0 + exp(-F31 - F32 - x[8])));
        const double F41 = x[6] / (1.0 + exp(-x[1] - x[4]));
        const double F42 = x[7] / (1.0 + exp(-x[3] - x[5]));
        const double F4 = sq(1.0 - 1.0 / (1.0 + exp(-F41 - F42 - x[8])));
        return F1 + F2 + F3 + F4;
    }
    case SDA_FN_YAO_LIU04:
        for (int k = lane; k < dim; k += S) a = fmax(a, fabs(x[k]));
        return group_max(a, g);
    case SDA_FN_YAO_LIU09:
        for (int k = lane; k < dim; k += S) { const double v = x[k]; a += v * v - 10.0 * cos_2pi(v) + 10; }
        return group_sum(a, g);
    case SDA_FN_ZERO_SUM: {
        for (int k = lane; k < dim; k += S) a += x[k];
        const double s = fabs(group_sum(a, g));
        if (s < 3e-16) return 0.0;
        return 1.0 + sqrt(10000.0 * s);
    }
    case SDA_FN_ZIMMERMAN: {
        const double p = x[0], q = x[1];
        const double h1 = 9.0 - p - q;
        const double h2 = sq(p - 3.0) + sq(q - 2.0) - 16.0;
        const double h3 = p * q - 14.0;
        double v = h1;
        v = fmax(v, 100.0 * (1.0 + h2) * np_sign(h2));
        v = fmax(v, 100.0 * (1.0 + h3) * np_sign(h3));
        v = fmax(v, 100.0 * (1.0 - p) * np_sign(p));
        v = fmax(v, 100.0 * (1.0 - q) * np_sign(q));
        return v;
    }
    case SDA_FN_STOCHASTIC: {  // sum(rnd * abs(x - 1/i)), i = 1..N
        const uint64_t h = point_hash(x, dim, g, salt);
        for (int k = lane; k < dim; k += S) a += point_noise(h, k, 0x53544F43u) * fabs(x[k] - 1.0 / (double)(k + 1));
        return group_sum(a, g);
    }
    case SDA_FN_XIN_SHE_YANG01: {  // sum(rnd * abs(x) ** i), i = 1..N
        const uint64_t h = point_hash(x, dim, g, salt);
        for (int k = lane; k < dim; k += S) a += point_noise(h, k, 0x58535931u) * pow(fabs(x[k]), (double)(k + 1));
        return group_sum(a, g);
    }
    default:
        return nan_value();
    }
}

}  // namespace sda

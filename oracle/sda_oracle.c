/*
 * sda_oracle.c -- CPU ORACLE (test infrastructure, see sda_oracle.h).
 *
 * Single-chain, scalar, line-by-line restatement of /root/reference/sdaopt/_sda.py.
 * Every function cites the reference lines it follows.  Uniform draws come either from a
 * recorded tape of numpy RandomState.random_sample() values (exact replay of the reference
 * draw protocol) or from the Philox4x32-10 addressing used by the CUDA
 * production path (same algorithm, different uniform source).
 */
#include "sda_oracle.h"
#include "lbfgsb_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BIG_VALUE 1e16          /* _sda.py:17 */
#define TAIL_LIMIT 1.e8         /* _sda.py:27 */
#define MIN_VISIT_BOUND 1.e-10  /* _sda.py:28 */
#define MAX_REINIT_COUNT 1000   /* _sda.py:116 */

/* ------------------------------------------------------------------------------------ */
/* Philox4x32-10                                                                          */
/* ------------------------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 53-bit uniform in [0,1) from two words, the construction of numpy's legacy
 * random_sample (genrand_res53): (a>>5, b>>6) -> (a*2^26+b)/2^53 */
static double u53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

/* Philox draw addressing shared with the CUDA path (include/sdaopt_b200.h, "RNG protocol"):
 *   ctr = { k, attempt, (purpose<<24)|j, iter },  key = seed (lo, hi)
 * one block = two uniforms. */
enum { PUR_POLAR = 0, PUR_CLIP = 1, PUR_ACCEPT = 2, PUR_INIT = 3 };

typedef struct {
    int mode;
    /* tape */
    const double *tape;
    int64_t pos, len;
    int exhausted;
    /* philox */
    uint32_t key[2];
    uint32_t iter;  /* SDARunner._iter of the current step */
    uint32_t j;     /* Markov chain index inside the step  */
} Rng;

static void rng_pair(Rng *r, int purpose, uint32_t k, uint32_t attempt, double *u1, double *u2)
{
    if (r->mode == ORC_RNG_TAPE) {
        if (r->pos + 2 > r->len) { r->exhausted = 1; *u1 = 0.25; *u2 = 0.25; return; }
        *u1 = r->tape[r->pos++];
        *u2 = r->tape[r->pos++];
    } else {
        uint32_t ctr[4] = { k, attempt, ((uint32_t)purpose << 24) | r->j, r->iter };
        uint32_t o[4];
        orc_philox4x32_10(ctr, r->key, o);
        *u1 = u53(o[0], o[1]);
        *u2 = u53(o[2], o[3]);
    }
}

static double rng_one(Rng *r, int purpose, uint32_t k, uint32_t attempt)
{
    if (r->mode == ORC_RNG_TAPE) {
        if (r->pos + 1 > r->len) { r->exhausted = 1; return 0.25; }
        return r->tape[r->pos++];
    } else {
        double a, b;
        rng_pair(r, purpose, k, attempt, &a, &b);
        return a;
    }
}

/* ------------------------------------------------------------------------------------ */
/* K1: temperature and sigmax                                                             */
/* ------------------------------------------------------------------------------------ */
/* _sda.py:396-402 */
double orc_temperature(double t0, double qv, int64_t i)
{
    double t1 = exp((qv - 1) * log(2.0)) - 1.0;
    double s = (double)i + 2.0;
    double t2 = exp((qv - 1) * log(s)) - 1.0;
    return t0 * t1 / t2;
}

/* _sda.py:76-86 */
double orc_sigmax(double temperature, double qv)
{
    double factor1 = exp(log(temperature) / (qv - 1.0));
    double factor2 = exp((4.0 - qv) * log(qv - 1.0));
    double factor3 = exp((2.0 - qv) * log(2.0) / (qv - 1.0));
    double factor4 = sqrt(M_PI) * factor1 * factor2 / (factor3 * (3.0 - qv));
    double factor5 = 1.0 / (qv - 1.0) - 0.5;
    double d1 = 2.0 - factor5;
    double factor6 = M_PI * (1.0 - factor5) / sin(M_PI * (1.0 - factor5)) / exp(lgamma(d1));
    return exp(-(qv - 1.0) * log(factor6 / factor4) / (3.0 - qv));
}

/* ------------------------------------------------------------------------------------ */
/* objectives                                                                             */
/* ------------------------------------------------------------------------------------ */
static double sutton_chen(const double *x, int dim)
{
    /* suttonchen.py:23-40: A=1, C=39.432, M=6, K=9 */
    int n = dim / 3;
    double total = 0.0;
    for (int i = 0; i < n; ++i) {
        double rep = 0.0, rho = 0.0;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            double dx = x[3 * i] - x[3 * j];
            double dy = x[3 * i + 1] - x[3 * j + 1];
            double dz = x[3 * i + 2] - x[3 * j + 2];
            double r = sqrt(dx * dx + dy * dy + dz * dz);
            rep += 1.0 / pow(r, 9);
            rho += 1.0 / pow(r, 6);
        }
        total += 0.5 * rep - 39.432 * sqrt(rho);
    }
    return total;
}

/* numpy.sum over a contiguous float64 vector: pairwise summation with an 8-way unrolled base
 * case (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum).  Restated so that the
 * oracle's energies carry the same rounding as the reference's np.sum. */
static double np_sum(const double *a, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int i;
        for (i = 0; i < 8; ++i) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_sum(a, n2) + np_sum(a + n2, n - n2);
    }
}

#define ORC_MAX_STACK_DIM 4096

double orc_eval(const OrcProblem *p, const double *x)
{
    const int n = p->dim;
    double s = 0.0, pr = 1.0;
    double tbuf[ORC_MAX_STACK_DIM], tbuf2[ORC_MAX_STACK_DIM];
    double *t = tbuf, *t2 = tbuf2;
    if (n > ORC_MAX_STACK_DIM) return NAN;
    switch (p->functor) {
    case ORC_FN_RASTRIGIN: /* go_funcs_R.py:91 */
        for (int i = 0; i < n; ++i) t[i] = x[i] * x[i] - 10.0 * cos(2.0 * M_PI * x[i]);
        return 10.0 * n + np_sum(t, n);
    case ORC_FN_SHIFTED_RASTRIGIN: { /* README.md:65-66 */
        double shift = p->params[0];
        for (int i = 0; i < n; ++i) {
            double d = x[i] - shift;
            t[i] = d * d - 10 * cos(2 * M_PI * d);
        }
        return np_sum(t, n) + 10 * n;
    }
    case ORC_FN_ROSENBROCK: /* scipy.optimize.rosen */
        for (int i = 0; i + 1 < n; ++i) {
            double a = x[i + 1] - x[i] * x[i];
            double b = 1 - x[i];
            t[i] = 100.0 * (a * a) + b * b;
        }
        return np_sum(t, n - 1);
    case ORC_FN_ACKLEY01: /* go_funcs_A.py:45-48 */
        for (int i = 0; i < n; ++i) { t[i] = x[i] * x[i]; t2[i] = cos(2 * M_PI * x[i]); }
        return -20. * exp(-0.2 * sqrt(np_sum(t, n) / n)) - exp(np_sum(t2, n) / n) + 20. + exp(1.);
    case ORC_FN_SCHWEFEL01: /* go_funcs_S.py:335-336 */
        for (int i = 0; i < n; ++i) t[i] = x[i] * x[i];
        return pow(np_sum(t, n), sqrt(M_PI));
    case ORC_FN_SCHWEFEL26: /* go_funcs_S.py:616 */
        for (int i = 0; i < n; ++i) t[i] = x[i] * sin(sqrt(fabs(x[i])));
        return 418.982887 * n - np_sum(t, n);
    case ORC_FN_EXPONENTIAL: /* go_funcs_E.py:306 */
        for (int i = 0; i < n; ++i) t[i] = x[i] * x[i];
        return -exp(-0.5 * np_sum(t, n));
    case ORC_FN_SUTTON_CHEN:
        return sutton_chen(x, n);
    case ORC_FN_CONSTANT:
        return p->params[0];
    case ORC_FN_SPHERE: /* go_funcs_S.py:1159 */
        for (int i = 0; i < n; ++i) t[i] = x[i] * x[i];
        return np_sum(t, n);
    case ORC_FN_GRIEWANK: /* go_funcs_G.py:173-175 */
        for (int i = 0; i < n; ++i) {
            t[i] = x[i] * x[i];
            pr *= cos(x[i] / sqrt((double)(i + 1)));
        }
        s = np_sum(t, n);
        return s / 4000.0 - pr + 1.0;

    /* ---- catalogue widening: go_funcs_<initial>.py, one case per class `fun` ---- */
    case ORC_FN_ALPINE01:
        for (int i = 0; i < n; ++i) t[i] = fabs(x[i] * sin(x[i]) + 0.1 * x[i]);
        return np_sum(t, n);
    case ORC_FN_BROWN:
        for (int i = 0; i + 1 < n; ++i) {
            double x0 = x[i] * x[i], x1 = x[i + 1] * x[i + 1];
            t[i] = pow(x0, x1 + 1.0) + pow(x1, x0 + 1.0);
        }
        return np_sum(t, n - 1);
    case ORC_FN_CIGAR:
        for (int i = 1; i < n; ++i) t[i - 1] = x[i] * x[i];
        return x[0] * x[0] + 1e6 * np_sum(t, n - 1);
    case ORC_FN_COSINE_MIXTURE:
        for (int i = 0; i < n; ++i) { t[i] = cos(5.0 * M_PI * x[i]); t2[i] = x[i] * x[i]; }
        return -0.1 * np_sum(t, n) - np_sum(t2, n);
    case ORC_FN_DEB01:
        for (int i = 0; i < n; ++i) t[i] = pow(sin(5 * M_PI * x[i]), 6.0);
        return -(1.0 / n) * np_sum(t, n);
    case ORC_FN_DIXON_PRICE:
        for (int i = 1; i < n; ++i) { double d = 2.0 * x[i] * x[i] - x[i - 1]; t[i - 1] = (i + 1) * (d * d); }
        return np_sum(t, n - 1) + (x[0] - 1.0) * (x[0] - 1.0);
    case ORC_FN_QING:
        for (int i = 0; i < n; ++i) { double d = x[i] * x[i] - (i + 1); t[i] = d * d; }
        return np_sum(t, n);
    case ORC_FN_QUINTIC:
        for (int i = 0; i < n; ++i) {
            double v = x[i];
            t[i] = fabs(pow(v, 5) - 3 * pow(v, 4) + 4 * pow(v, 3) + 2 * v * v - 10 * v - 4);
        }
        return np_sum(t, n);
    case ORC_FN_SALOMON: {
        for (int i = 0; i < n; ++i) t[i] = x[i] * x[i];
        double u = sqrt(np_sum(t, n));
        return 1 - cos(2 * M_PI * u) + 0.1 * u;
    }
    case ORC_FN_SCHWEFEL20:
        for (int i = 0; i < n; ++i) t[i] = fabs(x[i]);
        return np_sum(t, n);
    case ORC_FN_SCHWEFEL21:
        for (int i = 0; i < n; ++i) s = fmax(s, fabs(x[i]));
        return s;
    case ORC_FN_SCHWEFEL22:
        for (int i = 0; i < n; ++i) { t[i] = fabs(x[i]); pr *= fabs(x[i]); }
        return np_sum(t, n) + pr;
    case ORC_FN_STEP:
        for (int i = 0; i < n; ++i) t[i] = floor(fabs(x[i]));
        return np_sum(t, n);
    case ORC_FN_STYBLINSKI_TANG:
        for (int i = 0; i < n; ++i) t[i] = pow(x[i], 4) - 16 * x[i] * x[i] + 5 * x[i];
        return np_sum(t, n) / 2;
    case ORC_FN_TRID:
        for (int i = 0; i < n; ++i) t[i] = (x[i] - 1.0) * (x[i] - 1.0);
        for (int i = 1; i < n; ++i) t2[i - 1] = x[i] * x[i - 1];
        return np_sum(t, n) - np_sum(t2, n - 1);
    case ORC_FN_ZACHAROV: {
        for (int i = 0; i < n; ++i) { t[i] = x[i] * x[i]; t2[i] = (i + 1) * x[i]; }
        double u = np_sum(t, n), v = np_sum(t2, n);
        return u + pow(0.5 * v, 2) + pow(0.5 * v, 4);
    }
    case ORC_FN_WAVY:
        for (int i = 0; i < n; ++i) t[i] = cos(10 * x[i]) * exp(-(x[i] * x[i]) / 2.0);
        return 1.0 - (1.0 / n) * np_sum(t, n);
    case ORC_FN_STEP2:
        for (int i = 0; i < n; ++i) { double d = floor(x[i]) + 0.5; t[i] = d * d; }
        return np_sum(t, n);
    case ORC_FN_PLATEAU:
        for (int i = 0; i < n; ++i) t[i] = floor(fabs(x[i]));
        return 30.0 + np_sum(t, n);
    case ORC_FN_SODP:
        for (int i = 0; i < n; ++i) t[i] = pow(fabs(x[i]), i + 2);
        return np_sum(t, n);
    case ORC_FN_TRIGONOMETRIC02:
        for (int i = 0; i < n; ++i) {
            double d = (x[i] - 0.9) * (x[i] - 0.9);
            t[i] = 8 * pow(sin(7 * d), 2) + 6 * pow(sin(14 * d), 2) + d;
        }
        return 1.0 + np_sum(t, n);
    case ORC_FN_XIN_SHE_YANG02:
        for (int i = 0; i < n; ++i) { t[i] = fabs(x[i]); t2[i] = sin(x[i] * x[i]); }
        return np_sum(t, n) * exp(-np_sum(t2, n));
    case ORC_FN_SINE_ENVELOPE:
        for (int i = 0; i + 1 < n; ++i) {
            double q = x[i] * x[i] + x[i + 1] * x[i + 1];
            t[i] = (pow(sin(sqrt(q)), 2) - 0.5) / pow(1 + 0.001 * q, 2) + 0.5;
        }
        return np_sum(t, n - 1);
    case ORC_FN_PATHOLOGICAL:
        for (int i = 0; i + 1 < n; ++i) {
            double a0 = x[i], a1 = x[i + 1];
            t[i] = 0.5 + (pow(sin(sqrt(100 * a0 * a0 + a1 * a1)), 2) - 0.5) /
                             (1. + 0.001 * pow(a0 * a0 - 2 * a0 * a1 + a1 * a1, 2));
        }
        return np_sum(t, n - 1);
    case ORC_FN_RANA:
        for (int i = 0; i + 1 < n; ++i) {
            double a0 = x[i], a1 = x[i + 1];
            double t1 = sqrt(fabs(a1 + a0 + 1)), tt2 = sqrt(fabs(a1 - a0 + 1));
            t[i] = (a1 + 1) * cos(tt2) * sin(t1) + a0 * cos(t1) * sin(tt2);
        }
        return np_sum(t, n - 1);
    case ORC_FN_EASOM: {
        double a = pow(x[0] - M_PI, 2) + pow(x[1] - M_PI, 2);
        return -cos(x[0]) * cos(x[1]) * exp(-a);
    }
    case ORC_FN_SIX_HUMP_CAMEL:
        return (4 - 2.1 * x[0] * x[0] + pow(x[0], 4) / 3) * x[0] * x[0] + x[0] * x[1] +
               (4 * x[1] * x[1] - 4) * x[1] * x[1];
    case ORC_FN_BEALE:
        return pow(1.5 - x[0] + x[0] * x[1], 2) + pow(2.25 - x[0] + x[0] * x[1] * x[1], 2) +
               pow(2.625 - x[0] + x[0] * pow(x[1], 3), 2);
    case ORC_FN_HIMMEL_BLAU:
        return pow(x[0] * x[0] + x[1] - 11, 2) + pow(x[0] + x[1] * x[1] - 7, 2);
    case ORC_FN_MATYAS:
        return 0.26 * (x[0] * x[0] + x[1] * x[1]) - 0.48 * x[0] * x[1];
    case ORC_FN_MC_CORMICK:
        return sin(x[0] + x[1]) + pow(x[0] - x[1], 2) - 1.5 * x[0] + 2.5 * x[1] + 1;
    case ORC_FN_GOLDSTEIN_PRICE: {
        double a = 1 + pow(x[0] + x[1] + 1, 2) * (19 - 14 * x[0] + 3 * x[0] * x[0] - 14 * x[1] +
                                                  6 * x[0] * x[1] + 3 * x[1] * x[1]);
        double b = 30 + pow(2 * x[0] - 3 * x[1], 2) * (18 - 32 * x[0] + 12 * x[0] * x[0] + 48 * x[1] -
                                                      36 * x[0] * x[1] + 27 * x[1] * x[1]);
        return a * b;
    }
    case ORC_FN_LEVY13: {
        double u = pow(sin(3 * M_PI * x[0]), 2);
        double v = pow(x[0] - 1, 2) * (1 + pow(sin(3 * M_PI * x[1]), 2));
        double w = pow(x[1] - 1, 2) * (1 + pow(sin(2 * M_PI * x[1]), 2));
        return u + v + w;
    }
    case ORC_FN_BOHACHEVSKY1:
        return x[0] * x[0] + 2 * x[1] * x[1] - 0.3 * cos(3 * M_PI * x[0]) - 0.4 * cos(4 * M_PI * x[1]) + 0.7;
    case ORC_FN_BUKIN06:
        return 100 * sqrt(fabs(x[1] - 0.01 * x[0] * x[0])) + 0.01 * fabs(x[0] + 10);
    case ORC_FN_EGG_CRATE:
        return x[0] * x[0] + x[1] * x[1] + 25 * (pow(sin(x[0]), 2) + pow(sin(x[1]), 2));
    case ORC_FN_SCHAFFER01: {
        double u = x[0] * x[0] + x[1] * x[1];
        return 0.5 + (pow(sin(u), 2) - 0.5) / pow(1 + 0.001 * u, 2);
    }
    case ORC_FN_DROP_WAVE: {
        double nx = x[0] * x[0] + x[1] * x[1];
        return -(1 + cos(12 * sqrt(nx))) / (0.5 * nx + 2);
    }
    case ORC_FN_THREE_HUMP_CAMEL:
        return 2.0 * x[0] * x[0] - 1.05 * pow(x[0], 4.0) + pow(x[0], 6) / 6.0 + x[0] * x[1] + x[1] * x[1];
    case ORC_FN_ZETTL:
        return pow(x[0] * x[0] + x[1] * x[1] - 2 * x[0], 2) + 0.25 * x[0];
    case ORC_FN_LEON:
        return 100. * pow(x[1] - x[0] * x[0], 2.0) + pow(1 - x[0], 2.0);
    case ORC_FN_CUBE:
        return 100.0 * pow(x[1] - pow(x[0], 3.0), 2.0) + pow(1.0 - x[0], 2.0);
    case ORC_FN_BRENT:
        return pow(x[0] + 10.0, 2.0) + pow(x[1] + 10.0, 2.0) + exp(-x[0] * x[0] - x[1] * x[1]);
    case ORC_FN_PRICE01:
        return pow(fabs(x[0]) - 5.0, 2.0) + pow(fabs(x[1]) - 5.0, 2.0);
    case ORC_FN_BIRD:
        return sin(x[0]) * exp(pow(1 - cos(x[1]), 2)) + cos(x[1]) * exp(pow(1 - sin(x[0]), 2)) +
               pow(x[0] - x[1], 2);
    case ORC_FN_ACKLEY02:
        return -200 * exp(-0.02 * sqrt(x[0] * x[0] + x[1] * x[1]));
    case ORC_FN_ACKLEY03:
        return -200 * exp(-0.02 * sqrt(x[0] * x[0] + x[1] * x[1])) + 5 * exp(cos(3 * x[0]) + sin(3 * x[1]));
    case ORC_FN_ZIRILLI:
        return 0.25 * pow(x[0], 4) - 0.5 * x[0] * x[0] + 0.1 * x[0] + 0.5 * x[1] * x[1];
    case ORC_FN_ADJIMAN:
        return cos(x[0]) * sin(x[1]) - x[0] / (x[1] * x[1] + 1);
    default:
        return NAN;
    }
}

/* ------------------------------------------------------------------------------------ */
/* chain state                                                                            */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    const OrcProblem *p;
    const OrcConfig *c;
    int dim;
    Rng rng;
    /* EnergyState _sda.py:118-124 */
    double *x_cur, *xbest;
    double e_cur, ebest;
    int have_best;
    /* MarkovChain _sda.py:174-190 */
    double *xmin;
    int xmin_valid; /* xmin = np.array(None) after a failed LS, _sda.py:264-265 */
    double emin;
    int64_t not_improved_idx, not_improved_max_idx;
    double temperature_step;
    int state_improved;
    /* ObjectiveFunWrapper _sda.py:284 */
    int64_t nb_fun_call;
    int64_t nb_ls_call, n_ls;
    int in_ls;
    /* suite monitor bench.py:292-321 */
    int64_t mon_calls;
    int stop;       /* an exception left sda(): hit or budget */
    int hit;
    int64_t ncall_hit;
    double f_hit;
    /* scratch */
    double *x_visit, *visits, *gx1, *gx2;
    /* scipy ScalarFunction cache (optimize/_differentiable_functions.py fun_and_grad): a request
     * at an x bit-identical to the last evaluated one returns the stored f, g without calling
     * the objective */
    double *ls_x, *ls_g;
    double ls_f;
    int ls_cache_valid;
    /* gaussian_fn state _sda.py:36-38 */
    double x_gauss, s_gauss, root_gauss;
    /* trace */
    double *trace_e;
    uint8_t *trace_d;
    int64_t trace_pos;
    int status;
} Chain;

/* the callable handed to sda(): Benchmark.fun, wrapped by Algo._funcwrapped in suite mode
 * (bench.py:292-321).  Returns the value; sets ch->stop when the wrapper would raise. */
static double call_func(Chain *ch, const double *x)
{
    double res = orc_eval(ch->p, x);
    if (ch->c->max_calls > 0 && !ch->stop) {
        ch->mon_calls += 1;                                   /* bench.py:300 */
        if (ch->mon_calls >= ch->c->max_calls) {              /* bench.py:301-304 */
            ch->stop = 1;
        } else if (res <= ch->c->fglob + ch->c->tol) {        /* bench.py:314-320 */
            ch->hit = 1;
            ch->ncall_hit = ch->mon_calls;
            ch->f_hit = res;
            ch->stop = 1;
        }
    }
    return res;
}

/* _sda.py:321-323 */
static double func_wrapper(Chain *ch, const double *x)
{
    ch->nb_fun_call += 1;
    if (ch->in_ls) ch->nb_ls_call += 1;
    return call_func(ch, x);
}

/* call libm's pow through a volatile pointer: gcc folds pow(x, 2.0) into x*x otherwise */
static double (*volatile libm_pow)(double, double) = pow;

/* _sda.py:93-107, axis == 1 */
static double gaussian_fn1(Chain *ch, uint32_t k)
{
    int enter = 1;
    double y_gauss = 0.0;
    uint32_t attempt = 0;
    while (enter || (ch->s_gauss <= 0 || ch->s_gauss >= 1)) {
        enter = 0;
        double s1, s2;
        rng_pair(&ch->rng, PUR_POLAR, k, attempt++, &s1, &s2);
        if (ch->rng.exhausted) { ch->s_gauss = 0.5; break; }
        ch->x_gauss = s1 * 2.0 - 1.0;
        y_gauss = s2 * 2.0 - 1.0;
        /* `self.x_gauss ** 2` on Python floats is libm pow(x, 2.0) (float_pow), which is not
         * always the correctly rounded x*x; follow the reference's call. */
        ch->s_gauss = libm_pow(ch->x_gauss, 2.0) + libm_pow(y_gauss, 2.0);
    }
    ch->root_gauss = sqrt(-2.0 / ch->s_gauss * log(ch->s_gauss));
    return ch->root_gauss * y_gauss;
}

/* _sda.py:75-91 with the temperature-only constant hoisted (K1) */
static double visit_fn(Chain *ch, double sigmax, double qv, uint32_t k)
{
    double x = sigmax * gaussian_fn1(ch, k);
    double y = ch->root_gauss * ch->x_gauss;   /* gaussian_fn(0), _sda.py:106-107 */
    double den = exp((qv - 1.0) * log(fabs(y)) / (3.0 - qv));
    return x / den;
}

/* _sda.py:50-55 / :65-72 */
static double wrap_coord(double xv, double lower, double range)
{
    double a = xv - lower;
    double b = fmod(a, range) + range;
    double r = fmod(b, range) + lower;
    if (fabs(r - lower) < MIN_VISIT_BOUND) r += 1.e-10;
    return r;
}

/* _sda.py:40-73 */
static void visiting(Chain *ch, const double *x, int64_t step, double sigmax, double qv,
                     double *x_visit)
{
    const int dim = ch->dim;
    const double *lo = ch->p->lower, *up = ch->p->upper;
    if (step < dim) {
        for (int k = 0; k < dim; ++k) ch->visits[k] = visit_fn(ch, sigmax, qv, (uint32_t)k);
        double upper_sample, lower_sample;
        if (ch->rng.mode == ORC_RNG_TAPE) {
            upper_sample = rng_one(&ch->rng, PUR_CLIP, 0, 0);
            lower_sample = rng_one(&ch->rng, PUR_CLIP, 0, 0);
        } else {
            rng_pair(&ch->rng, PUR_CLIP, 0, 0, &upper_sample, &lower_sample);
        }
        for (int k = 0; k < dim; ++k) {
            double v = ch->visits[k];
            if (v > TAIL_LIMIT) v = TAIL_LIMIT * upper_sample;
            else if (v < -TAIL_LIMIT) v = -TAIL_LIMIT * lower_sample;
            x_visit[k] = wrap_coord(v + x[k], lo[k], up[k] - lo[k]);
        }
    } else {
        int index = (int)(step - dim);
        memcpy(x_visit, x, sizeof(double) * dim);
        double visit = visit_fn(ch, sigmax, qv, (uint32_t)index);
        if (visit > TAIL_LIMIT) visit = TAIL_LIMIT * rng_one(&ch->rng, PUR_CLIP, 0, 0);
        else if (visit < -TAIL_LIMIT) visit = -TAIL_LIMIT * rng_one(&ch->rng, PUR_CLIP, 0, 0);
        x_visit[index] = wrap_coord(visit + x[index], lo[index], up[index] - lo[index]);
    }
}

static void draw_location(Chain *ch, uint32_t attempt)
{
    /* lower + rs.random_sample(D) * (upper - lower), _sda.py:137-138 / :157-158 */
    for (int k = 0; k < ch->dim; ++k) {
        double u = rng_one(&ch->rng, PUR_INIT, (uint32_t)k, attempt);
        ch->x_cur[k] = ch->p->lower[k] + u * (ch->p->upper[k] - ch->p->lower[k]);
    }
}

/* EnergyState.reset, _sda.py:135-166.  Returns 0 or an ORC_ERR_* code. */
static int energy_reset(Chain *ch, const double *x0)
{
    uint32_t attempt = 0;
    if (x0 == NULL) draw_location(ch, attempt++);
    else memcpy(ch->x_cur, x0, sizeof(double) * ch->dim);
    int reinit_counter = 0;
    for (;;) {
        int done = 0;
        ch->e_cur = call_func(ch, ch->x_cur);   /* owf.func: NOT counted, _sda.py:144 */
        if (ch->stop) {
            /* the suite monitor ended the run at its first evaluation (the reference leaves sda()
             * through the monitor's exception): report that point */
            if (!ch->have_best) {
                ch->ebest = ch->e_cur;
                memcpy(ch->xbest, ch->x_cur, sizeof(double) * ch->dim);
                ch->have_best = 1;
            }
            return 0;
        }
        if (ch->e_cur >= BIG_VALUE || isnan(ch->e_cur)) {
            if (reinit_counter >= MAX_REINIT_COUNT) return ORC_ERR_REINIT;
            draw_location(ch, attempt++);
            reinit_counter += 1;
        } else {
            done = 1;
        }
        if (!ch->have_best) {                   /* _sda.py:163-165 */
            ch->ebest = ch->e_cur;
            memcpy(ch->xbest, ch->x_cur, sizeof(double) * ch->dim);
            ch->have_best = 1;
        }
        if (done) break;
    }
    return 0;
}

static void trace(Chain *ch, double e, int dec)
{
    if (ch->trace_e && ch->trace_pos < ch->c->trace_len) {
        ch->trace_e[ch->trace_pos] = e;
        ch->trace_d[ch->trace_pos] = (uint8_t)dec;
    }
    ch->trace_pos += 1;
}

/* MarkovChain.run, _sda.py:192-234 */
static void markov_run(Chain *ch, int64_t step, double temperature, double sigmax)
{
    const int dim = ch->dim;
    const double qa = ch->c->accept, qv = ch->c->visit;
    ch->temperature_step = temperature / (double)(step + 1);
    ch->not_improved_idx += 1;
    for (int64_t j = 0; j < 2 * (int64_t)dim; ++j) {
        if (j == 0) ch->state_improved = 0;
        if (step == 0 && j == 0) ch->state_improved = 1;
        ch->rng.j = (uint32_t)j;
        visiting(ch, ch->x_cur, j, sigmax, qv, ch->x_visit);
        double e = func_wrapper(ch, ch->x_visit);
        if (ch->stop || ch->rng.exhausted) return;
        if (e < ch->e_cur) {
            int dec = ORC_DEC_IMPROVE;
            ch->e_cur = e;
            memcpy(ch->x_cur, ch->x_visit, sizeof(double) * dim);
            if (e < ch->ebest) {
                ch->ebest = e;
                memcpy(ch->xbest, ch->x_visit, sizeof(double) * dim);
                ch->state_improved = 1;
                ch->not_improved_idx = 0;
                dec = ORC_DEC_NEW_BEST;
            }
            trace(ch, e, dec);
        } else {
            int dec = ORC_DEC_REJECT;
            double pqa_temp = (qa - 1.0) * (e - ch->e_cur) / ch->temperature_step + 1.;
            double pqa, r;
            /* Reference: r is drawn first, always (_sda.py:216).  The tape replays that.  The
             * Philox protocol shared with the CUDA path addresses draws, so the draw can be
             * skipped when pqa == 0: r <= 0 needs r == 0.0 exactly (2^-53), defined as reject. */
            if (ch->rng.mode == ORC_RNG_TAPE) r = rng_one(&ch->rng, PUR_ACCEPT, 0, 0);
            else r = 2.0;
            if (pqa_temp < 0.) pqa = 0.;
            else {
                pqa = exp(log(pqa_temp) / (1. - qa));
                if (ch->rng.mode != ORC_RNG_TAPE) r = rng_one(&ch->rng, PUR_ACCEPT, 0, 0);
            }
            if (r <= pqa) {
                ch->e_cur = e;
                memcpy(ch->x_cur, ch->x_visit, sizeof(double) * dim);
                memcpy(ch->xmin, ch->x_cur, sizeof(double) * dim);
                ch->xmin_valid = 1;
                dec = ORC_DEC_METRO_ACCEPT;
            }
            if (ch->not_improved_idx >= ch->not_improved_max_idx) {
                if (j == 0 || ch->e_cur < ch->emin) {
                    ch->emin = ch->e_cur;
                    memcpy(ch->xmin, ch->x_cur, sizeof(double) * dim);
                    ch->xmin_valid = 1;
                }
            }
            trace(ch, e, dec);
        }
    }
}

/* ObjectiveFunWrapper.gradient, _sda.py:325-345 */
static int fun_and_grad(void *ctx, const double *x, double *f, double *g)
{
    Chain *ch = (Chain *)ctx;
    const int n = ch->dim;
    const double reps = 1.e-6;  /* _sda.py:315 */
    if (ch->ls_cache_valid && memcmp(ch->ls_x, x, sizeof(double) * n) == 0) {
        *f = ch->ls_f;
        memcpy(g, ch->ls_g, sizeof(double) * n);
        return 0;
    }
    *f = func_wrapper(ch, x);   /* ScalarFunction.fun_and_grad: f first, then jac(x) */
    if (getenv("ORC_TRACE")) {  /* debugging aid: dump the fun_and_grad request points */
        FILE *fp = fopen("/tmp/orc_trace.txt", ch->nb_fun_call == 1 ? "w" : "a");
        if (fp) { for (int i = 0; i < n; ++i) fprintf(fp, "%.17g ", x[i]); fprintf(fp, "\n"); fclose(fp); }
    }
    if (ch->stop) return 1;
    for (int i = 0; i < n; ++i) {
        memcpy(ch->gx1, x, sizeof(double) * n);
        memcpy(ch->gx2, x, sizeof(double) * n);
        double respl = reps, respr = reps;
        ch->gx1[i] = x[i] + respr;
        if (ch->gx1[i] > ch->p->upper[i]) {
            ch->gx1[i] = ch->p->upper[i];
            respr = ch->gx1[i] - x[i];
        }
        ch->gx2[i] = x[i] - respl;
        if (ch->gx2[i] < ch->p->lower[i]) {
            ch->gx2[i] = ch->p->lower[i];
            respl = x[i] - ch->gx2[i];
        }
        double f1 = func_wrapper(ch, ch->gx1);
        if (ch->stop) return 1;
        double f2 = func_wrapper(ch, ch->gx2);
        if (ch->stop) return 1;
        g[i] = (f1 - f2) / (respl + respr);
        if (isnan(g[i]) || isinf(g[i])) g[i] = 101.0;
    }
    memcpy(ch->ls_x, x, sizeof(double) * n);
    memcpy(ch->ls_g, g, sizeof(double) * n);
    ch->ls_f = *f;
    ch->ls_cache_valid = 1;
    return 0;
}

/* ObjectiveFunWrapper.local_search, _sda.py:347-351 with the option dict of :300-311 */
static int ofw_local_search(Chain *ch, const double *x_in, double *x_out, double *e_out)
{
    const int n = ch->dim;
    OrcLbfgsbOpts o;
    o.m = 10;                                  /* maxcor */
    o.factr = 2.2204460492503131e-09 / 2.220446049250313e-16; /* scipy ftol default / eps */
    o.pgtol = 1e-6;                            /* gtol */
    o.maxls = 100;
    int ls_max_iter = n * 6;                   /* _sda.py:291-296 */
    if (ls_max_iter < 100) ls_max_iter = 100;
    if (ls_max_iter > 1000) ls_max_iter = 1000;
    o.maxiter = ls_max_iter;
    o.maxfun = 15000;
    memcpy(x_out, x_in, sizeof(double) * n);
    int nit = 0, nfev = 0;
    ch->in_ls = 1;
    ch->n_ls += 1;
    ch->ls_cache_valid = 0;
    int warnflag = orc_lbfgsb_minimize(n, ch->p->lower, ch->p->upper, x_out, e_out,
                                       fun_and_grad, ch, &o, &nit, &nfev);
    ch->in_ls = 0;
    if (getenv("ORC_TRACE_LS")) {
        fprintf(stderr, "LSX");
        for (int i = 0; i < n; ++i) fprintf(stderr, " %.17g", x_in[i]);
        fprintf(stderr, "\n");
    }
    if (getenv("ORC_TRACE_LS"))
        fprintf(stderr, "LS ncall=%lld warn=%d nit=%d nfev=%d e=%.17g\n", (long long)ch->nb_fun_call,
                warnflag, nit, nfev, *e_out);
    if (warnflag != 0) {                       /* not mres.success -> (BIG_VALUE, None) */
        *e_out = BIG_VALUE;
        return 0;
    }
    return 1;
}

/* MarkovChain.local_search, _sda.py:236-273 */
static void markov_local_search(Chain *ch, double *xls)
{
    const int dim = ch->dim;
    double e;
    if (ch->state_improved) {
        int ok = ofw_local_search(ch, ch->xbest, xls, &e);
        if (ch->stop) return;
        if (e < ch->ebest) {   /* only possible when ok */
            (void)ok;
            ch->not_improved_idx = 0;
            ch->ebest = e;
            memcpy(ch->xbest, xls, sizeof(double) * dim);
            ch->e_cur = e;
            memcpy(ch->x_cur, xls, sizeof(double) * dim);
            return;
        }
    }
    int do_ls = 0;
    /* _sda.py:254-258 is dead code: K = 100*D is never < 90*D, no uniform is drawn */
    if (ch->not_improved_idx >= ch->not_improved_max_idx) do_ls = 1;
    if (do_ls) {
        /* NOTE: after a failed LS xmin is array(None); the reference would then pass None to
         * scipy on the next trigger unless the Markov chain refreshed xmin in between.  The
         * chain always refreshes it (not_improved_idx>=max at j==0 of the next run()) before
         * this point can be reached again, so xmin_valid is informational only. */
        int ok = ofw_local_search(ch, ch->xmin, xls, &e);
        if (ch->stop) return;
        if (ok) { memcpy(ch->xmin, xls, sizeof(double) * dim); ch->xmin_valid = 1; }
        else ch->xmin_valid = 0;
        ch->emin = e;
        ch->not_improved_idx = 0;
        ch->not_improved_max_idx = dim;
        if (e < ch->ebest) {
            ch->ebest = ch->emin;
            memcpy(ch->xbest, ch->xmin, sizeof(double) * dim);
            ch->e_cur = e;
            memcpy(ch->x_cur, xls, sizeof(double) * dim);
        }
    }
}

static int run_chain(const OrcProblem *p, const OrcConfig *c, int64_t chain, const double *x0,
                     const uint64_t *seeds, const double *tape, int64_t tape_len, OrcResult *out)
{
    const int dim = p->dim;
    Chain ch;
    memset(&ch, 0, sizeof(ch));
    ch.p = p; ch.c = c; ch.dim = dim;
    double *buf = (double *)calloc((size_t)dim * 10, sizeof(double));
    if (!buf) return -3;
    ch.x_cur = buf; ch.xbest = buf + dim; ch.xmin = buf + 2 * dim; ch.x_visit = buf + 3 * dim;
    ch.visits = buf + 4 * dim; ch.gx1 = buf + 5 * dim; ch.gx2 = buf + 6 * dim;
    ch.ls_x = buf + 8 * dim; ch.ls_g = buf + 9 * dim;
    double *xls = buf + 7 * dim;
    ch.rng.mode = c->rng_mode;
    if (c->rng_mode == ORC_RNG_TAPE) {
        ch.rng.tape = tape + chain * tape_len;
        ch.rng.len = tape_len;
    } else {
        uint64_t s = seeds ? seeds[chain] : (uint64_t)chain;
        ch.rng.key[0] = (uint32_t)s;
        ch.rng.key[1] = (uint32_t)(s >> 32);
    }
    ch.rng.iter = 0;
    ch.rng.j = 0xFFFFFFu;
    if (out->trace_e && c->trace_len > 0) {
        ch.trace_e = out->trace_e + chain * c->trace_len;
        ch.trace_d = out->trace_d + chain * c->trace_len;
    }
    ch.ncall_hit = c->max_calls;      /* _fcall_success = MAX_FN_CALL, bench.py:269 */
    ch.f_hit = INFINITY;              /* bench.py:272 */

    /* SDARunner.__init__ :375-376 */
    int err = energy_reset(&ch, x0 ? x0 + chain * dim : NULL);
    int64_t iter = 0;
    if (!err && !ch.stop) {
        /* MarkovChain.__init__ :176-190 */
        ch.emin = ch.e_cur;
        memcpy(ch.xmin, ch.x_cur, sizeof(double) * dim);
        ch.xmin_valid = 1;
        ch.not_improved_idx = 0;
        ch.not_improved_max_idx = 1000;
        /* SDARunner.search :393-418 */
        int max_steps_reached = 0;
        const double qv = c->visit;
        double t1 = exp((qv - 1) * log(2.0)) - 1.0;
        while (!max_steps_reached && !ch.stop && !err && !ch.rng.exhausted) {
            for (int64_t i = 0; i < c->maxiter; ++i) {
                double s = (double)i + 2.0;
                double t2 = exp((qv - 1) * log(s)) - 1.0;
                double temperature = c->initial_temp * t1 / t2;
                int tabled = c->temp_table && c->sigmax_table && i < c->table_len;
                if (tabled) temperature = c->temp_table[i];
                iter += 1;
                ch.rng.iter = (uint32_t)iter;
                ch.rng.j = 0xFFFFFFu;
                if (iter == c->maxiter) { max_steps_reached = 1; break; }
                if (temperature < c->temperature_restart) {
                    err = energy_reset(&ch, NULL);
                    break;
                }
                markov_run(&ch, i, temperature, tabled ? c->sigmax_table[i] : orc_sigmax(temperature, qv));
                if (ch.stop || ch.rng.exhausted) break;
                if ((double)ch.nb_fun_call >= c->maxfun) break;
                if (!c->pure_sa) {
                    markov_local_search(&ch, xls);
                    if (ch.stop) break;
                    if ((double)ch.nb_fun_call >= c->maxfun) break;
                }
            }
            if (c->maxiter <= 0) break;
        }
    }
    if (ch.rng.exhausted) err |= ORC_ERR_TAPE_EXHAUSTED;
    memcpy(out->x + chain * dim, ch.xbest, sizeof(double) * dim);
    out->fun[chain] = ch.ebest;
    out->nit[chain] = iter;
    out->ncall[chain] = ch.nb_fun_call;
    if (out->ncall_ls) out->ncall_ls[chain] = ch.nb_ls_call;
    if (out->n_ls) out->n_ls[chain] = ch.n_ls;
    out->status[chain] = err;
    if (out->hit) out->hit[chain] = ch.hit;
    if (out->ncall_hit) out->ncall_hit[chain] = ch.ncall_hit;
    if (out->f_hit) out->f_hit[chain] = ch.f_hit;
    if (out->tape_used) out->tape_used[chain] = ch.rng.pos;
    free(buf);
    return 0;
}

static int check_args(const OrcProblem *p, const OrcConfig *c, const double *tape)
{
    if (!p || !c || p->dim <= 0 || p->functor < 0 || p->functor >= ORC_FN_COUNT) return -1;
    for (int k = 0; k < p->dim; ++k)
        if (!(p->lower[k] < p->upper[k])) return -2;   /* _sda.py:366-367 */
    if (c->rng_mode == ORC_RNG_TAPE && !tape) return -1;
    return 0;
}

int orc_run(const OrcProblem *p, const OrcConfig *c, int64_t n_chains, const double *x0,
            const uint64_t *seeds, const double *tape, int64_t tape_len, OrcResult *out)
{
    int rc = check_args(p, c, tape);
    if (rc) return rc;
    for (int64_t ch = 0; ch < n_chains; ++ch) {
        rc = run_chain(p, c, ch, x0, seeds, tape, tape_len, out);
        if (rc) return rc;
    }
    return 0;
}

typedef struct {
    const OrcProblem *p; const OrcConfig *c; int64_t begin, end; const double *x0;
    const uint64_t *seeds; const double *tape; int64_t tape_len; OrcResult *out; int rc;
} MtArg;

static void *mt_worker(void *v)
{
    MtArg *a = (MtArg *)v;
    for (int64_t ch = a->begin; ch < a->end; ++ch) {
        int rc = run_chain(a->p, a->c, ch, a->x0, a->seeds, a->tape, a->tape_len, a->out);
        if (rc) { a->rc = rc; break; }
    }
    return NULL;
}

int orc_run_mt(const OrcProblem *p, const OrcConfig *c, int64_t n_chains, const double *x0,
               const uint64_t *seeds, const double *tape, int64_t tape_len, OrcResult *out,
               int n_threads)
{
    int rc = check_args(p, c, tape);
    if (rc) return rc;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    MtArg *args = (MtArg *)calloc((size_t)n_threads, sizeof(MtArg));
    for (int t = 0; t < n_threads; ++t) {
        MtArg a = { p, c, n_chains * t / n_threads, n_chains * (t + 1) / n_threads, x0, seeds,
                    tape, tape_len, out, 0 };
        args[t] = a;
        pthread_create(&th[t], NULL, mt_worker, &args[t]);
    }
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        if (args[t].rc) rc = args[t].rc;
    }
    free(th); free(args);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* unit-test hooks                                                                        */
/* ------------------------------------------------------------------------------------ */
static void hook_chain(Chain *ch, const OrcProblem *p, OrcConfig *c, const double *tape,
                       int64_t tape_len, double *scratch)
{
    memset(ch, 0, sizeof(*ch));
    memset(c, 0, sizeof(*c));
    ch->p = p; ch->c = c; ch->dim = p ? p->dim : 1;
    ch->rng.mode = ORC_RNG_TAPE; ch->rng.tape = tape; ch->rng.len = tape_len;
    ch->visits = scratch;
}

int64_t orc_visiting(const OrcProblem *p, double qv, const double *x, int64_t step,
                     double temperature, const double *tape, int64_t tape_len, double *x_visit)
{
    Chain ch; OrcConfig c;
    double *scratch = (double *)calloc((size_t)p->dim, sizeof(double));
    hook_chain(&ch, p, &c, tape, tape_len, scratch);
    visiting(&ch, x, step, orc_sigmax(temperature, qv), qv, x_visit);
    free(scratch);
    return ch.rng.exhausted ? -1 : ch.rng.pos;
}

int64_t orc_visit_fn(double qv, double temperature, const double *tape, int64_t tape_len,
                     double *out, int64_t n)
{
    Chain ch; OrcConfig c;
    hook_chain(&ch, NULL, &c, tape, tape_len, NULL);
    double sigmax = orc_sigmax(temperature, qv);
    for (int64_t i = 0; i < n; ++i) out[i] = visit_fn(&ch, sigmax, qv, 0);
    return ch.rng.exhausted ? -1 : ch.rng.pos;
}

typedef struct { const OrcProblem *p; double *g1, *g2; int64_t nfev; } LsHookCtx;

static int hook_fg(void *ctx, const double *x, double *f, double *g)
{
    Chain *ch = (Chain *)ctx;
    return fun_and_grad(ch, x, f, g);
}

int orc_local_search(const OrcProblem *p, const double *x_in, double *x_out, double *f_out,
                     int64_t *nfev, int64_t *nit)
{
    Chain ch; OrcConfig c;
    double *scratch = (double *)calloc((size_t)p->dim * 4, sizeof(double));
    hook_chain(&ch, p, &c, NULL, 0, NULL);
    ch.gx1 = scratch; ch.gx2 = scratch + p->dim;
    ch.ls_x = scratch + 2 * p->dim; ch.ls_g = scratch + 3 * p->dim;
    (void)hook_fg;
    int ok = ofw_local_search(&ch, x_in, x_out, f_out);
    if (nfev) *nfev = ch.nb_fun_call;
    if (nit) *nit = 0;
    free(scratch);
    return ok;
}

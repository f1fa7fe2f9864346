/*
 * sdaopt_b200.h -- C ABI of the B200-native dual-annealing hot path (libsdaopt_b200.so).
 *
 * Drop-in boundary for sgubianpm/sdaopt's `sda()` / `SDARunner.search()` path.  The reference is
 * pure Python, so the "FFI" a maintainer binds is ctypes; every entry point
 * below names the reference interface it replaces (file:line under the reference checkout).
 *
 * Conventions
 *  - plain C types only; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *  - "_host" entry points take HOST pointers and perform allocation, H2D, launch, D2H and a stream
 *    synchronize internally; the others take DEVICE pointers, never allocate and never synchronize
 *  - every function returns SDA_OK (0) or a negative SDA_E_* code; sda_error_string() decodes it;
 *    there is no CPU fallback: without a usable CUDA device the calls fail with SDA_E_CUDA
 *  - all arithmetic is IEEE double; chains are independent; one call = one problem.  Large Philox
 *    batches with local search run as bounded rounds of { annealing launch, local-search launch }
 *    on the caller's stream (SDA_PHASED=0 forces the single persistent launch); both paths give
 *    bit-identical results
 *
 * RNG protocol (production mode, SDA_RNG_PHILOX): every uniform is a pure function of
 *     Philox4x32-10( ctr = { k, attempt, (purpose << 24) | j, iter }, key = seed[chain] )
 * with  k       coordinate index (0 where not applicable)
 *       attempt polar-rejection retry / re-initialisation counter
 *       purpose 0 polar pair, 1 tail-clip uniforms, 2 acceptance uniform, 3 (re)initial location
 *       j       Markov-chain index inside the temperature step (0xFFFFFF outside a step)
 *       iter    SDARunner._iter (_sda.py:403), 0 before the first step
 * One block yields two 53-bit uniforms ((w0>>5)*2^26 + (w1>>6))/2^53, ((w2>>5)*2^26 + (w3>>6))/2^53
 * -- the construction of numpy's RandomState.random_sample().  In SDA_RNG_TAPE mode the uniforms
 * are consumed sequentially from a caller-supplied tape in exactly the reference's draw order
 * (SURVEY.md App. A.1), which reproduces a reference trajectory.
 */
#ifndef SDAOPT_B200_H
#define SDAOPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDA_ABI_VERSION 1

/* ---- status codes ---------------------------------------------------------------------- */
#define SDA_OK 0
#define SDA_E_INVALID (-1)    /* bad argument */
#define SDA_E_BOUNDS (-2)     /* not all(lower < upper): ValueError of _sda.py:366-367 */
#define SDA_E_CUDA (-3)       /* CUDA runtime error / no device (see sda_last_cuda_error) */
#define SDA_E_WORKSPACE (-4)  /* workspace too small */
#define SDA_E_UNSUPPORTED (-5)

/* ---- registered device objectives (replace the Python callable `func`, _sda.py:440-444) ----- */
enum {
    SDA_FN_RASTRIGIN = 0,         /* go_benchmark_functions/go_funcs_R.py:88-91 */
    SDA_FN_SHIFTED_RASTRIGIN = 1, /* README.md:61-71; params[0] = shift */
    SDA_FN_ROSENBROCK = 2,        /* go_funcs_R.py:288-291 */
    SDA_FN_ACKLEY01 = 3,          /* go_funcs_A.py:43-48 */
    SDA_FN_SCHWEFEL01 = 4,        /* go_funcs_S.py:332-336 */
    SDA_FN_SCHWEFEL26 = 5,        /* go_funcs_S.py:613-616 */
    SDA_FN_EXPONENTIAL = 6,       /* go_funcs_E.py:303-306 */
    SDA_FN_SUTTON_CHEN = 7,       /* benchmark/suttonchen.py:23-40 */
    SDA_FN_CONSTANT = 8,          /* tests/test_sda.py:34 `weirdfunc`; params[0] = value */
    SDA_FN_SPHERE = 9,            /* go_funcs_S.py:1156-1159 */
    SDA_FN_GRIEWANK = 10,         /* go_funcs_G.py:171-175 */
    /* catalogue widening: go_benchmark_functions/go_funcs_*.py, same class names */
    SDA_FN_ALPINE01 = 11,
    SDA_FN_BROWN = 12,
    SDA_FN_CIGAR = 13,
    SDA_FN_COSINE_MIXTURE = 14,
    SDA_FN_DEB01 = 15,
    SDA_FN_DIXON_PRICE = 16,
    SDA_FN_EASOM = 17,
    SDA_FN_QING = 18,
    SDA_FN_QUINTIC = 19,
    SDA_FN_SALOMON = 20,
    SDA_FN_SCHWEFEL20 = 21,
    SDA_FN_SCHWEFEL21 = 22,
    SDA_FN_SCHWEFEL22 = 23,
    SDA_FN_STEP = 24,
    SDA_FN_STYBLINSKI_TANG = 25,
    SDA_FN_TRID = 26,
    SDA_FN_ZACHAROV = 27,
    SDA_FN_SIX_HUMP_CAMEL = 28,
    SDA_FN_BEALE = 29,
    SDA_FN_HIMMEL_BLAU = 30,
    SDA_FN_MATYAS = 31,
    SDA_FN_MC_CORMICK = 32,
    SDA_FN_GOLDSTEIN_PRICE = 33,
    SDA_FN_LEVY13 = 34,
    SDA_FN_BOHACHEVSKY1 = 35,
    SDA_FN_BUKIN06 = 36,
    SDA_FN_EGG_CRATE = 37,
    SDA_FN_SCHAFFER01 = 38,
    SDA_FN_DROP_WAVE = 39,
    SDA_FN_THREE_HUMP_CAMEL = 40,
    SDA_FN_ZETTL = 41,
    SDA_FN_LEON = 42,
    SDA_FN_CUBE = 43,
    SDA_FN_BRENT = 44,
    SDA_FN_PRICE01 = 45,
    SDA_FN_BIRD = 46,
    SDA_FN_ACKLEY02 = 47,
    SDA_FN_ACKLEY03 = 48,
    SDA_FN_ZIRILLI = 49,
    SDA_FN_WAVY = 50,
    SDA_FN_STEP2 = 51,
    SDA_FN_PLATEAU = 52,
    SDA_FN_SODP = 53,
    SDA_FN_TRIGONOMETRIC02 = 54,
    SDA_FN_XIN_SHE_YANG02 = 55,
    SDA_FN_SINE_ENVELOPE = 56,
    SDA_FN_PATHOLOGICAL = 57,
    SDA_FN_RANA = 58,
    SDA_FN_ADJIMAN = 59,
    /* catalogue part 3: the remaining go_funcs_*.py classes (ids follow class-name order) */
    SDA_FN_AMGM = 60,
    SDA_FN_ALPINE02 = 61,
    SDA_FN_BARTELS_CONN = 62,
    SDA_FN_BIGGS_EXP02 = 63,
    SDA_FN_BIGGS_EXP03 = 64,
    SDA_FN_BIGGS_EXP04 = 65,
    SDA_FN_BIGGS_EXP05 = 66,
    SDA_FN_BOHACHEVSKY2 = 67,
    SDA_FN_BOHACHEVSKY3 = 68,
    SDA_FN_BOX_BETTS = 69,
    SDA_FN_BRANIN01 = 70,
    SDA_FN_BRANIN02 = 71,
    SDA_FN_BUKIN02 = 72,
    SDA_FN_BUKIN04 = 73,
    SDA_FN_CARROM_TABLE = 74,
    SDA_FN_CHICHINADZE = 75,
    SDA_FN_COLA = 76,
    SDA_FN_COLVILLE = 77,
    SDA_FN_CORANA = 78,
    SDA_FN_CROSS_IN_TRAY = 79,
    SDA_FN_CROSS_LEG_TABLE = 80,
    SDA_FN_CROWNED_CROSS = 81,
    SDA_FN_CSENDES = 82,
    SDA_FN_DAMAVANDI = 83,
    SDA_FN_DE_VILLIERS_GLASSER01 = 84,
    SDA_FN_DE_VILLIERS_GLASSER02 = 85,
    SDA_FN_DEB03 = 86,
    SDA_FN_DECANOMIAL = 87,
    SDA_FN_DECEPTIVE = 88,
    SDA_FN_DECKKERS_AARTS = 89,
    SDA_FN_DEFLECTED_CORRUGATED_SPRING = 90,
    SDA_FN_DOLAN = 91,
    SDA_FN_ECKERLE4 = 92,
    SDA_FN_EGG_HOLDER = 93,
    SDA_FN_EL_ATTAR_VIDYASAGAR_DUTTA = 94,
    SDA_FN_EXP2 = 95,
    SDA_FN_FREUDENSTEIN_ROTH = 96,
    SDA_FN_GEAR = 97,
    SDA_FN_GIUNTA = 98,
    SDA_FN_GULF = 99,
    SDA_FN_HANSEN = 100,
    SDA_FN_HARTMANN3 = 101,
    SDA_FN_HARTMANN6 = 102,
    SDA_FN_HELICAL_VALLEY = 103,
    SDA_FN_HOLDER_TABLE = 104,
    SDA_FN_HOSAKI = 105,
    SDA_FN_INFINITY = 106,
    SDA_FN_JENNRICH_SAMPSON = 107,
    SDA_FN_JUDGE = 108,
    SDA_FN_KATSUURA = 109,
    SDA_FN_KEANE = 110,
    SDA_FN_KOWALIK = 111,
    SDA_FN_LANGERMANN = 112,
    SDA_FN_LENNARD_JONES = 113,
    SDA_FN_LEVY03 = 114,
    SDA_FN_LEVY05 = 115,
    SDA_FN_MEYER = 116,
    SDA_FN_MICHALEWICZ = 117,
    SDA_FN_MIELE_CANTRELL = 118,
    SDA_FN_MISHRA01 = 119,
    SDA_FN_MISHRA02 = 120,
    SDA_FN_MISHRA03 = 121,
    SDA_FN_MISHRA04 = 122,
    SDA_FN_MISHRA05 = 123,
    SDA_FN_MISHRA06 = 124,
    SDA_FN_MISHRA07 = 125,
    SDA_FN_MISHRA08 = 126,
    SDA_FN_MISHRA09 = 127,
    SDA_FN_MISHRA10 = 128,
    SDA_FN_MISHRA11 = 129,
    SDA_FN_MULTI_MODAL = 130,
    SDA_FN_NEEDLE_EYE = 131,
    SDA_FN_NEW_FUNCTION01 = 132,
    SDA_FN_NEW_FUNCTION02 = 133,
    SDA_FN_ODD_SQUARE = 134,
    SDA_FN_PARSOPOULOS = 135,
    SDA_FN_PAVIANI = 136,
    SDA_FN_PEN_HOLDER = 137,
    SDA_FN_PENALTY01 = 138,
    SDA_FN_PENALTY02 = 139,
    SDA_FN_PERM_FUNCTION01 = 140,
    SDA_FN_PERM_FUNCTION02 = 141,
    SDA_FN_PINTER = 142,
    SDA_FN_POWELL = 143,
    SDA_FN_POWER_SUM = 144,
    SDA_FN_PRICE02 = 145,
    SDA_FN_PRICE03 = 146,
    SDA_FN_PRICE04 = 147,
    SDA_FN_QUADRATIC = 148,
    SDA_FN_RATKOWSKY01 = 149,
    SDA_FN_RATKOWSKY02 = 150,
    SDA_FN_RIPPLE01 = 151,
    SDA_FN_RIPPLE25 = 152,
    SDA_FN_ROSENBROCK_MODIFIED = 153,
    SDA_FN_ROTATED_ELLIPSE01 = 154,
    SDA_FN_ROTATED_ELLIPSE02 = 155,
    SDA_FN_SARGAN = 156,
    SDA_FN_SCHAFFER02 = 157,
    SDA_FN_SCHAFFER03 = 158,
    SDA_FN_SCHAFFER04 = 159,
    SDA_FN_SCHWEFEL02 = 160,
    SDA_FN_SCHWEFEL04 = 161,
    SDA_FN_SCHWEFEL06 = 162,
    SDA_FN_SCHWEFEL36 = 163,
    SDA_FN_SHEKEL05 = 164,
    SDA_FN_SHEKEL07 = 165,
    SDA_FN_SHEKEL10 = 166,
    SDA_FN_SHUBERT01 = 167,
    SDA_FN_SHUBERT03 = 168,
    SDA_FN_SHUBERT04 = 169,
    SDA_FN_STRETCHED_V = 170,
    SDA_FN_TEST_TUBE_HOLDER = 171,
    SDA_FN_THURBER = 172,
    SDA_FN_TRECCANI = 173,
    SDA_FN_TREFETHEN = 174,
    SDA_FN_TRIGONOMETRIC01 = 175,
    SDA_FN_TRIPOD = 176,
    SDA_FN_URSEM01 = 177,
    SDA_FN_URSEM03 = 178,
    SDA_FN_URSEM04 = 179,
    SDA_FN_URSEM_WAVES = 180,
    SDA_FN_VENTER_SOBIEZCCZANSKI_SOBIESKI = 181,
    SDA_FN_VINCENT = 182,
    SDA_FN_WATSON = 183,
    SDA_FN_WAYBURN_SEADER01 = 184,
    SDA_FN_WAYBURN_SEADER02 = 185,
    SDA_FN_WEIERSTRASS = 186,
    SDA_FN_WHITLEY = 187,
    SDA_FN_WOLFE = 188,
    SDA_FN_XIN_SHE_YANG03 = 189,
    SDA_FN_XIN_SHE_YANG04 = 190,
    SDA_FN_XOR = 191,
    SDA_FN_YAO_LIU04 = 192,
    SDA_FN_YAO_LIU09 = 193,
    SDA_FN_ZERO_SUM = 194,
    SDA_FN_ZIMMERMAN = 195,
    /* the two classes that draw random numbers inside fun() (go_funcs_S.py:1274-1280,
     * go_funcs_X.py:47-51): counter-based noise addressed by the point, see sda_objectives_cat.cuh */
    SDA_FN_STOCHASTIC = 196,
    SDA_FN_XIN_SHE_YANG01 = 197,
    SDA_FN_COUNT,
    /* user-supplied CUDA device functor (north star; replaces an arbitrary Python `func`,
     * _sda.py:440-444).  Available only in a library built with -DSDA_USER_FUNCTOR_HEADER="<file>"
     * where <file> defines
     *   template <class X> __device__ double sda_user_objective(const X &x, int dim, const double *params);
     * (x[k] is the k-th coordinate).  sdaopt_b200.functors.compile_user_functor() does the build. */
    SDA_FN_USER = 255
};

enum { SDA_RNG_PHILOX = 0, SDA_RNG_TAPE = 1 };

/* decision codes of the optional per-evaluation trace */
enum { SDA_DEC_IMPROVE = 1, SDA_DEC_NEW_BEST = 2, SDA_DEC_METRO_ACCEPT = 3, SDA_DEC_REJECT = 4 };

/* per-chain status bits (SdaResult.status) */
enum {
    SDA_CHAIN_OK = 0,
    SDA_CHAIN_REINIT_FAILED = 1, /* ValueError of _sda.py:149-156 (1000 re-inits) */
    SDA_CHAIN_TAPE_EXHAUSTED = 2
};

/* The optimisation problem: replaces (func, bounds) of sda(), _sda.py:431-452. */
typedef struct SdaProblem {
    int32_t functor;      /* SDA_FN_* */
    int32_t dim;          /* len(bounds) */
    const double *lower;  /* [dim] */
    const double *upper;  /* [dim] */
    const double *params; /* functor parameters or NULL */
    int32_t n_params;
} SdaProblem;

/* Replaces the keyword arguments of sda() (_sda.py:431-433) plus the constants the reference
 * hard-codes (temperature_restart :385; L-BFGS-B option dict :300-311) and the suite's first-hit
 * monitor (benchmark/bench.py:292-321). */
typedef struct SdaConfig {
    int64_t maxiter;            /* 1000 */
    double initial_temp;        /* 5230. */
    double visit;               /* qv 2.62 */
    double accept;              /* qa -5.0 */
    double maxfun;              /* 1e7 (soft limit with the reference's semantics) */
    int32_t pure_sa;            /* no local search */
    double temperature_restart; /* 0.1 */
    int32_t rng_mode;           /* SDA_RNG_* */
    int64_t max_calls;          /* suite budget MAX_FN_CALL (0 = monitor off) */
    double fglob;               /* suite: known optimum */
    double tol;                 /* suite: DEFAULT_TOL 1e-8 */
    int64_t trace_len;          /* evaluations traced per chain (0 = off) */
    /* K1 tables indexed by the step index i of _sda.py:398 (host pointers, always): T_i and
     * sigmax(T_i) computed by the caller with the reference's own formulas.  NULL -> the library
     * computes them with libm. */
    const double *temp_table;
    const double *sigmax_table;
    int64_t table_len;
} SdaConfig;

/* Replaces SDARunner.result (_sda.py:420-428) for a batch of chains, plus the suite's per-run
 * record (bench.py:184-188).  Any pointer except x/fun/nit/ncall/status may be NULL. */
typedef struct SdaResult {
    double *x;           /* [C*dim] res.x   */
    double *fun;         /* [C]     res.fun */
    int64_t *nit;        /* [C]     res.nit */
    int64_t *ncall;      /* [C]     res.ncall */
    int64_t *ncall_ls;   /* [C] evaluations spent inside local searches */
    int64_t *n_ls;       /* [C] local searches started */
    int32_t *status;     /* [C] SDA_CHAIN_* bits */
    int32_t *hit;        /* [C] suite: success */
    int64_t *ncall_hit;  /* [C] suite: call index of the first hit (max_calls on failure) */
    double *f_hit;       /* [C] suite: value at the first hit (inf on failure) */
    double *trace_e;     /* [C*trace_len] */
    uint8_t *trace_d;    /* [C*trace_len] */
    int64_t *tape_used;  /* [C] uniforms consumed (tape mode) */
} SdaResult;

/* ---- library ---------------------------------------------------------------------------- */
int sda_abi_version(void);
const char *sda_error_string(int code);
const char *sda_last_cuda_error(void);
/* functor id for a registered name ("Rastrigin", ...), or SDA_E_INVALID */
int sda_functor_id(const char *name);
const char *sda_functor_name(int functor);
/* number of coordinates a fixed-arity catalogue class indexes (its `N` in go_funcs_*.py with
 * change_dimensionality = False); 0 for the n-D classes; SDA_E_INVALID for an unknown id.  Every entry
 * point rejects (SDA_E_INVALID) a problem whose dim differs from a non-zero arity -- the reference's
 * fun() would raise IndexError there. */
int sda_functor_arity(int functor);
/* 1 if this build contains a user functor (SDA_FN_USER), else 0 */
int sda_has_user_functor(void);
/* number of CUDA devices visible, or SDA_E_CUDA */
int sda_device_count(void);

/* ---- the hot path: replaces sda(...) -> SDARunner.__init__ + search() + result
 *      (_sda.py:585-589, :357-428) for n_chains independent runs ------------------------------- */
/* bytes of device scratch sda_batch_run needs for this (problem, config, n_chains) */
size_t sda_workspace_bytes(const SdaProblem *p, const SdaConfig *c, int64_t n_chains);

/* DEVICE-pointer form.  p->lower/upper/params, x0, seeds, tape, out->* and workspace are device
 * pointers (c->temp_table / sigmax_table stay host pointers).  Asynchronous on `stream`.
 *   x0    [C*dim] or NULL -> uniform initial location, _sda.py:137-138
 *   seeds [C]     Philox keys (NULL -> key = chain index); ignored in tape mode
 *   tape  [C*tape_len] uniforms (tape mode) */
int sda_batch_run(const SdaProblem *p, const SdaConfig *c, int64_t n_chains, const double *x0,
                  const uint64_t *seeds, const double *tape, int64_t tape_len, const SdaResult *out,
                  void *workspace, size_t workspace_bytes, void *stream);

/* kernels launched by the last sda_batch_run on the calling thread (1 for the persistent path,
 * 4 per round + 1 for the phased path) */
long long sda_last_launch_count(void);

/* Per-kernel device time of sda_batch_run calls, measured with CUDA events on the caller's stream.
 * sda_kernel_timing(1) arms it for the calling thread (every following sda_batch_run records an
 * event pair around each of its launches), sda_kernel_timing(0) disarms.  sda_kernel_times()
 * waits for the recorded events, returns the accumulated milliseconds and launch counts since the
 * last call -- index 0: annealing launches (persistent, pure-SA or phased-annealing instantiation),
 * 1: local-search launches, 2: the final resume launch of the phased path -- and resets them.
 * Used by bench.py for the live per-kernel roofline figures; off by default. */
int sda_kernel_timing(int enable);
int sda_kernel_times(double ms[3], long long launches[3]);

/* HOST-pointer form of the same call (what the ctypes binding of sda() uses): allocates, copies
 * in, runs on `device`, copies out, synchronizes. */
int sda_batch_run_host(const SdaProblem *p, const SdaConfig *c, int64_t n_chains, const double *x0,
                       const uint64_t *seeds, const double *tape, int64_t tape_len,
                       const SdaResult *out, int device);

/* Accuracy probe of the device elementary functions of csrc/sda_math.cuh (host pointers):
 * kind 0: cos_2pi(x) = cos(2 pi x), 1: log_pos(x), 2: exp_any(x), 3: sqrt_normal(x),
 * 4: div_normal(1.2345678901234567, x); out[i] = f(x[i]). */
int sda_math_probe_host(int kind, const double *x, int64_t n, double *out, int device);

/* ---- objective evaluation: replaces Benchmark.fun(x) / sutton_chen(x)
 *      (go_benchmark_functions/go_benchmark.py:15, suttonchen.py:23) for n points ----------------- */
int sda_eval(const SdaProblem *p, const double *x, int64_t n, double *f, void *stream);
int sda_eval_host(const SdaProblem *p, const double *x, int64_t n, double *f, int device);

/* ---- unit-test surface mirroring sdaopt/__init__.py:9-13 (host pointers) ------------------------ */
/* VisitingDistribution.visiting(x, step, temperature) (_sda.py:40-73) driven by a tape of
 * uniforms; writes x_visit[dim] and the number of uniforms consumed */
int sda_visiting_host(const SdaProblem *p, double qv, const double *x, int64_t step,
                      double temperature, double sigmax, const double *tape, int64_t tape_len,
                      double *x_visit, int64_t *tape_used, int device);
/* n successive VisitingDistribution.visit_fn(temperature) values (_sda.py:75-91) */
int sda_visit_fn_host(double qv, double sigmax, const double *tape, int64_t tape_len, double *out,
                      int64_t n, int64_t *tape_used, int device);
/* ObjectiveFunWrapper.local_search(x) (_sda.py:347-351) for n start points: batched projected
 * L-BFGS-B with the reference's option dict (:300-311) and clipped central-difference gradient
 * (:325-345).  success[i] mirrors mres.success; nfev[i] counts objective calls. */
int sda_local_search_host(const SdaProblem *p, const double *x_in, int64_t n, double *x_out,
                          double *f_out, int32_t *success, int64_t *nfev, int device);

/* EnergyState.reset(owf, rs, x0) (_sda.py:135-166) driven by a tape of uniforms: (re)initial location
 * (x0 or lower + u*(upper-lower)), its energy through owf.func, up to 1000 re-draws while the energy
 * is >= 1e16 or NaN (*status = SDA_CHAIN_REINIT_FAILED after that: the ValueError of :149-156), best
 * retained when have_best != 0 (in: xbest[dim], *ebest) and set on the first evaluation otherwise
 * (:163-165).  *tape_used = uniforms consumed, -1 if the tape was too short. */
int sda_energy_reset_host(const SdaProblem *p, const double *x0, const double *tape, int64_t tape_len,
                          int32_t have_best, double *current_location, double *current_energy,
                          double *xbest, double *ebest, int32_t *status, int64_t *tape_used,
                          int device);

/* ---- measurement helper: sustained FP64 FMA rate of the device (roofline denominator) ----------- */
int sda_fp64_peak(double *tflops, double *ms, int device);

#ifdef __cplusplus
}
#endif
#endif

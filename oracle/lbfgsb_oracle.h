/*
 * lbfgsb_oracle.h -- CPU ORACLE (test infrastructure): restatement of L-BFGS-B.
 *
 * The reference calls scipy.optimize.minimize(method='L-BFGS-B') (_sda.py:348).  SciPy is a
 * third-party dependency that is NOT vendored under /root/reference (setup.py:11, unpinned;
 * scipy 1.18.1 in this image, whose _lbfgsb is a compiled port of L-BFGS-B 3.0).  This file
 * restates the published algorithm: Byrd, Lu, Nocedal, Zhu, "A limited memory algorithm for
 * bound constrained optimization" (SIAM J. Sci. Comput. 16, 1995); Zhu, Byrd, Lu, Nocedal,
 * "Algorithm 778: L-BFGS-B" (ACM TOMS 23, 1997); Morales, Nocedal, "Remark on Algorithm 778"
 * (ACM TOMS 38, 2011: subspace-minimisation projection step); More', Thuente line search
 * (dcsrch/dcstep, MINPACK-2) -- plus the termination logic of the SciPy Python driver
 * (scipy/optimize/_lbfgsb_py.py:272-461).
 *
 * Parity status: PINNED against SciPy itself -- tests/golden/lbfgsb_*.npz hold the sequence of
 * evaluation points, the final (x, f), nit, nfev and success flag that scipy 1.18.1 produced
 * (oracle/gen_golden.py), and tests/test_oracle_lbfgsb.py replays them.
 */
#ifndef LBFGSB_ORACLE_H
#define LBFGSB_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int m;          /* maxcor */
    double factr;   /* ftol / eps */
    double pgtol;   /* gtol */
    int maxls;
    int maxiter;
    int maxfun;
} OrcLbfgsbOpts;

/* f and g at x; return non-zero to abort the minimisation (suite monitor stop) */
typedef int (*orc_fg_fn)(void *ctx, const double *x, double *f, double *g);

/*
 * Minimise inside [l,u] starting from x (clipped into the box first, _lbfgsb_py.py:362).
 * Returns the SciPy warnflag: 0 converged (success), 1 maxiter/maxfun, 2 abnormal, -1 aborted
 * by the callback.  x and *f hold the final iterate.  nfev counts fun_and_grad requests.
 */
int orc_lbfgsb_minimize(int n, const double *l, const double *u, double *x, double *f,
                        orc_fg_fn fg, void *ctx, const OrcLbfgsbOpts *opts, int *nit, int *nfev);

#ifdef __cplusplus
}
#endif
#endif

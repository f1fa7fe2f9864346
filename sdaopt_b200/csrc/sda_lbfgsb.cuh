// sda_lbfgsb.cuh -- K6/K7: projected L-BFGS-B local search, one chain per warp.
//
// Replaces ObjectiveFunWrapper.local_search -> scipy.optimize.minimize(method='L-BFGS-B',
// jac=self.gradient, bounds=bounds, options={maxls:100, gtol:1e-6, eps:1e-6,
// maxiter:clamp(6D,100,1000), maxcor:10, maxfun:15000})   (_sda.py:300-311, :347-351)
// and ObjectiveFunWrapper.gradient (:325-345).
//
// Algorithm: Byrd-Lu-Nocedal-Zhu limited-memory BFGS for box constraints (generalized Cauchy
// point, subspace minimisation with the Morales-Nocedal projection, More'-Thuente line search),
// with the termination logic of SciPy's driver (scipy/optimize/_lbfgsb_py.py:272-461).
//
// Warp mapping: the history S, Y (n x m), the 2m x 2m middle matrices and the work vectors live in
// a per-warp slot of global memory (L1/L2 resident).  Objective evaluations (1 + 2n per request)
// are warp-cooperative (K3).  The O(m^2)..O(m^3) dense algebra is executed redundantly by all 32
// lanes on warp-uniform addresses -- in SIMT that costs exactly one lane's instruction stream and
// one memory transaction per access, and needs no broadcasts.  n-sized loops that dominate
// (history dot products, axpys) are striped across lanes with butterfly reductions.
#pragma once

#include "sda_chain.cuh"

namespace sda {

constexpr int kLsM = 10;             // maxcor
constexpr double kLsEps = 2.220446049250313e-16;
constexpr double kLsFactr = 2.2204460492503131e-09 / 2.220446049250313e-16;  // scipy ftol / eps
constexpr double kLsPgtol = 1e-6;    // gtol
constexpr int kLsMaxls = 100;
constexpr int kLsMaxfun = 15000;
constexpr double kGradEps = 1.e-6;   // self.reps                                   :315

// global (L1/L2) part of a warp's workspace: ws, wy | z r d t xp g x xl gl | 3 int arrays | snd
__host__ __device__ inline long long ls_workspace_doubles(int n)
{
    const int m = kLsM;
    return 2LL * m * n + 9LL * n + 2LL * n + 2LL * n + 8 + 4LL * m * m;   // (+ 2 n: memoised terms ta, tb)
}
// shared-memory part: sy ss wt (m x m) | wn (2m x 2m) | wa (8m).  These are updated in place
// element by element (Cholesky, triangular solves, running sums); through global memory every
// read-after-write would pay an L2 round trip because stores write through and drop the L1 line.
// (snd, the 2m x 2m accumulator of formk, is only read-modify-written once per element: global.)
__host__ __device__ constexpr int ls_smem_doubles() { return 3 * kLsM * kLsM + 4 * kLsM * kLsM + 8 * kLsM; }

struct LsWork {
    int n;
    double *ws, *wy, *sy, *ss, *wt, *wn, *snd, *z, *r, *d, *t, *xp, *g, *x, *wa;
    // SciPy's ScalarFunction keeps the last evaluated point: a request at a bit-identical x is
    // served from (xl, fl, gl) without calling the objective (no ncall, no monitor tick)
    double *xl, *gl;
    double *ta, *tb;     // terms of the base point of a gradient (separable objectives, ls_fun_and_grad)
    double fl;
    int cache_valid;
    int *index, *iwhere, *indx2;
    __device__ void bind(double *base, double *sbase, int n_)
    {
        const int m = kLsM;
        n = n_;
        double *q = sbase;
        sy = q; q += m * m;  ss = q; q += m * m;  wt = q; q += m * m;
        wn = q; q += 4 * m * m;
        wa = q;
        q = base;
        ws = q; q += (long long)m * n;  wy = q; q += (long long)m * n;
        z = q; q += n;  r = q; q += n;  d = q; q += n;  t = q; q += n;  xp = q; q += n;
        g = q; q += n;  x = q; q += n;  xl = q; q += n;  gl = q; q += n;
        ta = q; q += n;  tb = q; q += n;
        cache_valid = 0;
        index = reinterpret_cast<int *>(q);
        iwhere = index + n;
        indx2 = iwhere + n;
        snd = q + 2 * n + 4;
    }
};

// 1-based column-major views, as in the published description of the algorithm
#define LWS(i, j) w.ws[((j) - 1) * n + ((i) - 1)]
#define LWY(i, j) w.wy[((j) - 1) * n + ((i) - 1)]
#define LSY(i, j) w.sy[((j) - 1) * m + ((i) - 1)]
#define LSS(i, j) w.ss[((j) - 1) * m + ((i) - 1)]
#define LWT(i, j) w.wt[((j) - 1) * m + ((i) - 1)]
#define LWN(i, j) w.wn[((j) - 1) * 2 * m + ((i) - 1)]
#define LWN1(i, j) w.snd[((j) - 1) * 2 * m + ((i) - 1)]

// ---- dense helpers on warp-uniform data (executed redundantly by every lane) -----------------
__device__ __noinline__ int chol_upper(double *a, int lda, int nn)
{
    #pragma unroll 1
    for (int j = 1; j <= nn; ++j) {
        double s = 0.0;
        #pragma unroll 1
        for (int k = 1; k <= j - 1; ++k) {
            double dot = 0.0;
            #pragma unroll 1
            for (int i = 1; i <= k - 1; ++i) dot += a[(k - 1) * lda + i - 1] * a[(j - 1) * lda + i - 1];
            double tt = a[(j - 1) * lda + k - 1] - dot;
            tt = tt / a[(k - 1) * lda + k - 1];
            a[(j - 1) * lda + k - 1] = tt;
            s += tt * tt;
        }
        s = a[(j - 1) * lda + j - 1] - s;
        if (s <= 0.0) return j;
        a[(j - 1) * lda + j - 1] = sqrt(s);
    }
    return 0;
}

__device__ __noinline__ int tri_solve_upper(const double *t, int ldt, int nn, double *b, int trans)
{
    #pragma unroll 1
    for (int j = 1; j <= nn; ++j)
        if (t[(j - 1) * ldt + j - 1] == 0.0) return j;
    if (trans) {
        b[0] = b[0] / t[0];
        #pragma unroll 1
        for (int j = 2; j <= nn; ++j) {
            double dot = 0.0;
            #pragma unroll 1
            for (int i = 1; i <= j - 1; ++i) dot += t[(j - 1) * ldt + i - 1] * b[i - 1];
            b[j - 1] = (b[j - 1] - dot) / t[(j - 1) * ldt + j - 1];
        }
    } else {
        b[nn - 1] = b[nn - 1] / t[(nn - 1) * ldt + nn - 1];
        #pragma unroll 1
        for (int jj = 2; jj <= nn; ++jj) {
            const int j = nn - jj + 1;
            const double temp = -b[j];
            #pragma unroll 1
            for (int i = 1; i <= j; ++i) b[i - 1] += temp * t[j * ldt + i - 1];
            b[j - 1] = b[j - 1] / t[(j - 1) * ldt + j - 1];
        }
    }
    return 0;
}

// product of the 2m x 2m middle matrix of the compact representation with v
__device__ __noinline__ int ls_bmv(LsWork &w, int col, const double *v, double *p)
{
    const int m = kLsM;
    if (col == 0) return 0;
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane == 0) p[col] = v[col];
    if (lane >= 1 && lane < col) {           // row i = lane + 1 of L D^-1 v1 (serial formula per row)
        const int i = lane + 1, i2 = col + i;
        double sum = 0.0;
        #pragma unroll 1
        for (int k = 1; k <= i - 1; ++k) sum += LSY(i, k) * v[k - 1] / LSY(k, k);
        p[i2 - 1] = v[i2 - 1] + sum;
    }
    __syncwarp();
    if (tri_solve_upper(w.wt, m, col, p + col, 1)) return 1;
    #pragma unroll 1
    for (int i = 1; i <= col; ++i) p[i - 1] = v[i - 1] / sqrt(LSY(i, i));
    if (tri_solve_upper(w.wt, m, col, p + col, 0)) return 1;
    __syncwarp();
    if (lane < col) {
        const int i = lane + 1;
        double pi = -p[i - 1] / sqrt(LSY(i, i));
        double sum = 0.0;
        #pragma unroll 1
        for (int k = i + 1; k <= col; ++k) sum += LSY(k, i) * p[col + k - 1] / LSY(i, i);
        p[i - 1] = pi + sum;
    }
    __syncwarp();
    return 0;
}

__device__ __noinline__ void heap_least(int nn, double *t, int *iorder, int iheap)
{
    double *T = t - 1;
    int *IO = iorder - 1;
    if (iheap == 0) {
        #pragma unroll 1
        for (int k = 2; k <= nn; ++k) {
            const double ddum = T[k];
            const int indxin = IO[k];
            int i = k;
            #pragma unroll 1
            while (i > 1) {
                const int j = i / 2;
                if (ddum < T[j]) { T[i] = T[j]; IO[i] = IO[j]; i = j; }
                else break;
            }
            T[i] = ddum; IO[i] = indxin;
        }
    }
    if (nn > 1) {
        int i = 1;
        const double out = T[1];
        const int indxou = IO[1];
        const double ddum = T[nn];
        const int indxin = IO[nn];
        #pragma unroll 1
        for (;;) {
            int j = i + i;
            if (j <= nn - 1) {
                if (T[j + 1] < T[j]) j = j + 1;
                if (T[j] < ddum) { T[i] = T[j]; IO[i] = IO[j]; i = j; }
                else break;
            } else break;
        }
        T[i] = ddum; IO[i] = indxin;
        T[nn] = out; IO[nn] = indxou;
    }
}

// generalized Cauchy point (warp-uniform, serial: the breakpoint bookkeeping is order dependent)
__device__ __noinline__ int ls_cauchy(LsWork &w, const double *x, const double *l, const double *u,
                                const double *g, double theta, int col, int head, double sbgnrm,
                                int lane)
{
    const int n = w.n, m = kLsM;
    int *iorder = w.indx2, *iwhere = w.iwhere;
    double *t = w.t, *d = w.d, *xcp = w.z;
    double *p = w.wa, *c = w.wa + 2 * m, *wbp = w.wa + 4 * m, *v = w.wa + 6 * m;
    if (sbgnrm <= 0.0) { for (int i = 0; i < n; ++i) xcp[i] = x[i]; return 0; }
    int bnded = 1, nfree = n + 1, nbreak = 0, ibkmin = 0;
    double bkmin = 0.0, f1 = 0.0;
    const int col2 = 2 * col;
    #pragma unroll 1
    for (int i = 0; i < col2; ++i) p[i] = 0.0;
    #pragma unroll 1
    for (int i = 1; i <= n; ++i) {
        const double neggi = -g[i - 1];
        const double tl = x[i - 1] - l[i - 1];
        const double tu = u[i - 1] - x[i - 1];
        if (iwhere[i - 1] != 3 && iwhere[i - 1] != -1) {
            const int xlower = tl <= 0.0, xupper = tu <= 0.0;
            int iw = 0;
            if (xlower) { if (neggi <= 0.0) iw = 1; }
            else if (xupper) { if (neggi >= 0.0) iw = 2; }
            else { if (fabs(neggi) <= 0.0) iw = -3; }
            iwhere[i - 1] = iw;
        }
        if (iwhere[i - 1] != 0 && iwhere[i - 1] != -1) {
            d[i - 1] = 0.0;
        } else {
            d[i - 1] = neggi;
            f1 -= neggi * neggi;
            if (neggi < 0.0) {
                nbreak++; iorder[nbreak - 1] = i; t[nbreak - 1] = tl / (-neggi);
                if (nbreak == 1 || t[nbreak - 1] < bkmin) { bkmin = t[nbreak - 1]; ibkmin = nbreak; }
            } else if (neggi > 0.0) {
                nbreak++; iorder[nbreak - 1] = i; t[nbreak - 1] = tu / neggi;
                if (nbreak == 1 || t[nbreak - 1] < bkmin) { bkmin = t[nbreak - 1]; ibkmin = nbreak; }
            } else {
                nfree--; iorder[nfree - 1] = i;
                if (fabs(neggi) > 0.0) bnded = 0;
            }
        }
    }
    // p = W' d: lane j owns p[j] (columns of Y for j < col, of S for j >= col); d is zero on the
    // variables that do not move, so the sum runs over all i in the published order
    __syncwarp();
    if (lane < col2) {
        const int jj = lane < col ? lane : lane - col;
        const int pointr = (head - 1 + jj) % m + 1;
        const double *wc = lane < col ? &LWY(1, pointr) : &LWS(1, pointr);
        double acc = 0.0;
        #pragma unroll 1
        for (int i = 0; i < n; ++i) acc += wc[i] * d[i];
        p[lane] = (lane >= col && theta != 1.0) ? acc * theta : acc;
    }
    __syncwarp();
    for (int i = lane; i < n; i += SDA_WARP) xcp[i] = x[i];
    __syncwarp();
    if (nbreak == 0 && nfree == n + 1) return 0;
    #pragma unroll 1
    for (int j = 0; j < col2; ++j) c[j] = 0.0;
    double f2 = -theta * f1;
    const double f2_org = f2;
    if (col > 0) {
        if (ls_bmv(w, col, p, v)) return 1;
        double dot = 0.0;
        #pragma unroll 1
        for (int j = 0; j < col2; ++j) dot += v[j] * p[j];
        f2 -= dot;
    }
    double dtm = -f1 / f2;
    double tsum = 0.0;
    bool skip_tail = false;
    if (nbreak != 0) {
        int nleft = nbreak, iter = 1;
        double tj = 0.0;
        #pragma unroll 1
        for (;;) {
            const double tj0 = tj;
            int ibp;
            if (iter == 1) {
                tj = bkmin; ibp = iorder[ibkmin - 1];
            } else {
                if (iter == 2 && ibkmin != nbreak) {
                    t[ibkmin - 1] = t[nbreak - 1];
                    iorder[ibkmin - 1] = iorder[nbreak - 1];
                }
                heap_least(nleft, t, iorder, iter - 2);
                tj = t[nleft - 1]; ibp = iorder[nleft - 1];
            }
            const double dt = tj - tj0;
            if (dtm < dt) break;
            tsum += dt; nleft--; iter++;
            const double dibp = d[ibp - 1];
            d[ibp - 1] = 0.0;
            double zibp;
            if (dibp > 0.0) { zibp = u[ibp - 1] - x[ibp - 1]; xcp[ibp - 1] = u[ibp - 1]; iwhere[ibp - 1] = 2; }
            else { zibp = l[ibp - 1] - x[ibp - 1]; xcp[ibp - 1] = l[ibp - 1]; iwhere[ibp - 1] = 1; }
            if (nleft == 0 && nbreak == n) { dtm = dt; skip_tail = true; break; }
            const double dibp2 = dibp * dibp;
            f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
            f2 = f2 - theta * dibp2;
            if (col > 0) {
                #pragma unroll 1
                for (int j = 0; j < col2; ++j) c[j] += dt * p[j];
                int pointr = head;
                #pragma unroll 1
                for (int j = 1; j <= col; ++j) {
                    wbp[j - 1] = LWY(ibp, pointr);
                    wbp[col + j - 1] = theta * LWS(ibp, pointr);
                    pointr = pointr % m + 1;
                }
                if (ls_bmv(w, col, wbp, v)) return 1;
                double wmc = 0.0, wmp = 0.0, wmw = 0.0;
                #pragma unroll 1
                for (int j = 0; j < col2; ++j) { wmc += c[j] * v[j]; wmp += p[j] * v[j]; wmw += wbp[j] * v[j]; }
                #pragma unroll 1
                for (int j = 0; j < col2; ++j) p[j] -= dibp * wbp[j];
                f1 += dibp * wmc;
                f2 += 2.0 * dibp * wmp - dibp2 * wmw;
            }
            f2 = fmax(kLsEps * f2_org, f2);
            if (nleft > 0) { dtm = -f1 / f2; continue; }
            else if (bnded) { f1 = 0.0; f2 = 0.0; dtm = 0.0; }
            else dtm = -f1 / f2;
            break;
        }
    }
    if (!skip_tail) {
        if (dtm <= 0.0) dtm = 0.0;
        tsum += dtm;
        __syncwarp();
        for (int i = lane; i < n; i += SDA_WARP) xcp[i] += tsum * d[i];
        __syncwarp();
    }
    if (col > 0) for (int j = 0; j < col2; ++j) c[j] += dtm * p[j];
    return 0;
}

__device__ __noinline__ void ls_freev(LsWork &w, int &nfree, int &nenter, int &ileave, int &wrk,
                                int updatd, int iter)
{
    const int n = w.n;
    int *index = w.index, *indx2 = w.indx2, *iwhere = w.iwhere;
    nenter = 0; ileave = n + 1;
    if (iter > 0) {
        #pragma unroll 1
        for (int i = 1; i <= nfree; ++i) {
            const int k = index[i - 1];
            if (iwhere[k - 1] > 0) { ileave--; indx2[ileave - 1] = k; }
        }
        #pragma unroll 1
        for (int i = 1 + nfree; i <= n; ++i) {
            const int k = index[i - 1];
            if (iwhere[k - 1] <= 0) { nenter++; indx2[nenter - 1] = k; }
        }
    }
    wrk = (ileave < n + 1) || (nenter > 0) || updatd;
    nfree = 0;
    int iact = n + 1;
    #pragma unroll 1
    for (int i = 1; i <= n; ++i) {
        if (iwhere[i - 1] <= 0) { nfree++; index[nfree - 1] = i; }
        else { iact--; index[iact - 1] = i; }
    }
}

// lane-striped sum over an index list: sum_{k in list[b..e]} A(k, pa) * B(k, pb)
__device__ __forceinline__ double list_dot(const double *A, const double *B, const int *list, int b,
                                           int e, int lane)
{
    double s = 0.0;
    #pragma unroll 1
    for (int k = b + lane; k <= e; k += SDA_WARP) {
        const int k1 = list[k - 1] - 1;
        s += A[k1] * B[k1];
    }
    return warp_sum(s);
}

// LEL^T factorisation of the indefinite K matrix of the subspace problem
__device__ __noinline__ int ls_formk(LsWork &w, int nsub, int nenter, int ileave, int iupdat, int updatd,
                               double theta, int col, int head, int lane)
{
    const int n = w.n, m = kLsM;
    const int *ind = w.index, *indx2 = w.indx2;
    int upcl;
    if (updatd) {
        if (iupdat > m) {
            #pragma unroll 1
            for (int jy = 1; jy <= m - 1; ++jy) {
                const int js = m + jy;
                #pragma unroll 1
                for (int i = 0; i < m - jy; ++i) LWN1(jy + i, jy) = LWN1(jy + 1 + i, jy + 1);
                #pragma unroll 1
                for (int i = 0; i < m - jy; ++i) LWN1(js + i, js) = LWN1(js + 1 + i, js + 1);
                #pragma unroll 1
                for (int i = 0; i < m - 1; ++i) LWN1(m + 1 + i, jy) = LWN1(m + 2 + i, jy + 1);
            }
        }
        const int pbegin = 1, pend = nsub, dbegin = nsub + 1, dend = n;
        const int iy = col, is = m + col;
        int ipntr = head + col - 1;
        if (ipntr > m) ipntr -= m;
        int jpntr = head;
        #pragma unroll 1
        for (int jy = 1; jy <= col; ++jy) {
            const int js = m + jy;
            const double temp1 = list_dot(&LWY(1, ipntr), &LWY(1, jpntr), ind, pbegin, pend, lane);
            const double temp2 = list_dot(&LWS(1, ipntr), &LWS(1, jpntr), ind, dbegin, dend, lane);
            const double temp3 = list_dot(&LWS(1, ipntr), &LWY(1, jpntr), ind, dbegin, dend, lane);
            LWN1(iy, jy) = temp1; LWN1(is, js) = temp2; LWN1(is, jy) = temp3;
            jpntr = jpntr % m + 1;
        }
        const int jy = col;
        jpntr = head + col - 1;
        if (jpntr > m) jpntr -= m;
        ipntr = head;
        #pragma unroll 1
        for (int i = 1; i <= col; ++i) {
            const int is2 = m + i;
            const double temp3 = list_dot(&LWS(1, ipntr), &LWY(1, jpntr), ind, pbegin, pend, lane);
            ipntr = ipntr % m + 1;
            LWN1(is2, jy) = temp3;
        }
        upcl = col - 1;
    } else {
        upcl = col;
    }
    if (nenter > 0 || ileave <= n) {
        int ipntr = head;
        #pragma unroll 1
        for (int iy = 1; iy <= upcl; ++iy) {
            const int is = m + iy;
            int jpntr = head;
            #pragma unroll 1
            for (int jy = 1; jy <= iy; ++jy) {
                const int js = m + jy;
                const double temp1 = list_dot(&LWY(1, ipntr), &LWY(1, jpntr), indx2, 1, nenter, lane);
                const double temp2 = list_dot(&LWS(1, ipntr), &LWS(1, jpntr), indx2, 1, nenter, lane);
                const double temp3 = list_dot(&LWY(1, ipntr), &LWY(1, jpntr), indx2, ileave, n, lane);
                const double temp4 = list_dot(&LWS(1, ipntr), &LWS(1, jpntr), indx2, ileave, n, lane);
                LWN1(iy, jy) += temp1 - temp3;
                LWN1(is, js) += -temp2 + temp4;
                jpntr = jpntr % m + 1;
            }
            ipntr = ipntr % m + 1;
        }
        ipntr = head;
        #pragma unroll 1
        for (int is = m + 1; is <= m + upcl; ++is) {
            int jpntr = head;
            #pragma unroll 1
            for (int jy = 1; jy <= upcl; ++jy) {
                const double temp1 = list_dot(&LWS(1, ipntr), &LWY(1, jpntr), indx2, 1, nenter, lane);
                const double temp3 = list_dot(&LWS(1, ipntr), &LWY(1, jpntr), indx2, ileave, n, lane);
                if (is <= jy + m) LWN1(is, jy) += temp1 - temp3;
                else LWN1(is, jy) += -temp1 + temp3;
                jpntr = jpntr % m + 1;
            }
            ipntr = ipntr % m + 1;
        }
    }
    const int m2 = 2 * m;
    #pragma unroll 1
    for (int iy = 1; iy <= col; ++iy) {
        const int is = col + iy, is1 = m + iy;
        #pragma unroll 1
        for (int jy = 1; jy <= iy; ++jy) {
            const int js = col + jy, js1 = m + jy;
            LWN(jy, iy) = LWN1(iy, jy) / theta;
            LWN(js, is) = LWN1(is1, js1) * theta;
        }
        #pragma unroll 1
        for (int jy = 1; jy <= iy - 1; ++jy) LWN(jy, is) = -LWN1(is1, jy);
        #pragma unroll 1
        for (int jy = iy; jy <= col; ++jy) LWN(jy, is) = LWN1(is1, jy);
        LWN(iy, iy) += LSY(iy, iy);
    }
    if (chol_upper(w.wn, m2, col)) return -1;
    const int col2 = 2 * col;
    // the col right-hand sides of L^-1 (-L_a' + R_z') and the columns of the (2,2) block are
    // independent: lane c owns column js = col + 1 + c (same operation order per column as the
    // serial sweep, so the factors are bit-identical)
    __syncwarp();
    #pragma unroll 1
    for (int j = 1; j <= col; ++j)
        if (LWN(j, j) == 0.0) return -1;
    if (lane < col) {
        double *b = &LWN(1, col + 1 + lane);
        b[0] = b[0] / LWN(1, 1);
        #pragma unroll 1
        for (int j = 2; j <= col; ++j) {
            double dot = 0.0;
            #pragma unroll 1
            for (int i = 1; i <= j - 1; ++i) dot += LWN(i, j) * b[i - 1];
            b[j - 1] = (b[j - 1] - dot) / LWN(j, j);
        }
    }
    __syncwarp();
    if (lane < col) {
        const int js = col + 1 + lane;
        #pragma unroll 1
        for (int is = col + 1; is <= js; ++is) {
            double dot = 0.0;
            #pragma unroll 1
            for (int k = 1; k <= col; ++k) dot += LWN(k, is) * LWN(k, js);
            LWN(is, js) += dot;
        }
    }
    __syncwarp();
    if (chol_upper(&LWN(col + 1, col + 1), m2, col)) return -2;
    return 0;
}

__device__ __noinline__ int ls_formt(LsWork &w, int col, double theta, int lane)
{
    const int m = kLsM;
    // lane j-1 builds column j of the upper triangle (element formulas unchanged)
    __syncwarp();
    if (lane < col) {
        const int j = lane + 1;
        LWT(1, j) = theta * LSS(1, j);
        #pragma unroll 1
        for (int i = 2; i <= j; ++i) {
            const int k1 = i - 1;
            double ddum = 0.0;
            #pragma unroll 1
            for (int k = 1; k <= k1; ++k) ddum += LSY(i, k) * LSY(j, k) / LSY(k, k);
            LWT(i, j) = ddum + theta * LSS(i, j);
        }
    }
    __syncwarp();
    return chol_upper(w.wt, m, col) ? -3 : 0;
}

// r = -Z'B(xcp - xk) - Z'g   (free variables, lane-striped)
__device__ __noinline__ int ls_cmprlb(LsWork &w, const double *x, const double *g, double theta, int col,
                                int head, int nfree, int lane)
{
    const int n = w.n, m = kLsM;
    double *r = w.r, *z = w.z, *wa = w.wa;
    if (ls_bmv(w, col, wa + 2 * m, wa)) return -8;
    #pragma unroll 1
    for (int i = 1 + lane; i <= nfree; i += SDA_WARP) {
        const int k = w.index[i - 1];
        double ri = -theta * (z[k - 1] - x[k - 1]) - g[k - 1];
        int pointr = head;
        #pragma unroll 1
        for (int j = 1; j <= col; ++j) {
            ri += LWY(k, pointr) * wa[j - 1] + LWS(k, pointr) * (theta * wa[col + j - 1]);
            pointr = pointr % m + 1;
        }
        r[i - 1] = ri;
    }
    __syncwarp();
    return 0;
}

// subspace minimisation + Morales-Nocedal projection
__device__ __noinline__ int ls_subsm(LsWork &w, int nsub, const double *l, const double *u,
                               const double *xx, const double *gg, double theta, int col, int head,
                               int lane)
{
    const int n = w.n, m = kLsM;
    const int *ind = w.index;
    double *x = w.z, *d = w.r, *xp = w.xp, *wv = w.wa;
    if (nsub <= 0) return 0;
    int pointr = head;
    #pragma unroll 1
    for (int i = 1; i <= col; ++i) {
        double temp1 = 0.0, temp2 = 0.0;
        #pragma unroll 1
        for (int j = 1 + lane; j <= nsub; j += SDA_WARP) {
            const int k = ind[j - 1];
            temp1 += LWY(k, pointr) * d[j - 1];
            temp2 += LWS(k, pointr) * d[j - 1];
        }
        temp1 = warp_sum(temp1);
        temp2 = warp_sum(temp2);
        wv[i - 1] = temp1;
        wv[col + i - 1] = theta * temp2;
        pointr = pointr % m + 1;
    }
    const int m2 = 2 * m, col2 = 2 * col;
    if (tri_solve_upper(w.wn, m2, col2, wv, 1)) return 1;
    #pragma unroll 1
    for (int i = 0; i < col; ++i) wv[i] = -wv[i];
    if (tri_solve_upper(w.wn, m2, col2, wv, 0)) return 1;
    __syncwarp();
    #pragma unroll 1
    for (int i = 1 + lane; i <= nsub; i += SDA_WARP) {
        const int k = ind[i - 1];
        double di = d[i - 1];
        pointr = head;
        #pragma unroll 1
        for (int jy = 1; jy <= col; ++jy) {
            di += LWY(k, pointr) * wv[jy - 1] / theta + LWS(k, pointr) * wv[col + jy - 1];
            pointr = pointr % m + 1;
        }
        d[i - 1] = di * (1.0 / theta);
    }
    #pragma unroll 1
    for (int i = lane; i < n; i += SDA_WARP) xp[i] = x[i];
    __syncwarp();
    int iword = 0;
    #pragma unroll 1
    for (int i = 1 + lane; i <= nsub; i += SDA_WARP) {
        const int k = ind[i - 1];
        const double dk = d[i - 1];
        double xk = x[k - 1];
        xk = fmax(l[k - 1], xk + dk);
        xk = fmin(u[k - 1], xk);
        x[k - 1] = xk;
        if (xk == l[k - 1] || xk == u[k - 1]) iword = 1;
    }
    iword = __any_sync(SDA_FULL_MASK, iword);
    __syncwarp();
    if (iword == 0) return 0;
    double dd_p = 0.0;
    #pragma unroll 1
    for (int i = lane; i < n; i += SDA_WARP) dd_p += (x[i] - xx[i]) * gg[i];
    dd_p = warp_sum(dd_p);
    if (dd_p > 0.0) {
        #pragma unroll 1
        for (int i = lane; i < n; i += SDA_WARP) x[i] = xp[i];
        __syncwarp();
        // serial (order dependent) search of the first blocking bound
        double alpha = 1.0, temp1 = alpha;
        int ibd = 0;
        #pragma unroll 1
        for (int i = 1; i <= nsub; ++i) {
            const int k = ind[i - 1];
            const double dk = d[i - 1];
            if (dk < 0.0) {
                const double temp2 = l[k - 1] - x[k - 1];
                if (temp2 >= 0.0) temp1 = 0.0;
                else if (dk * alpha < temp2) temp1 = temp2 / dk;
            } else if (dk > 0.0) {
                const double temp2 = u[k - 1] - x[k - 1];
                if (temp2 <= 0.0) temp1 = 0.0;
                else if (dk * alpha > temp2) temp1 = temp2 / dk;
            }
            if (temp1 < alpha) { alpha = temp1; ibd = i; }
        }
        if (alpha < 1.0) {
            const double dk = d[ibd - 1];
            const int k = ind[ibd - 1];
            if (dk > 0.0) { x[k - 1] = u[k - 1]; d[ibd - 1] = 0.0; }
            else if (dk < 0.0) { x[k - 1] = l[k - 1]; d[ibd - 1] = 0.0; }
        }
        __syncwarp();
        #pragma unroll 1
        for (int i = 1 + lane; i <= nsub; i += SDA_WARP) {
            const int k = ind[i - 1];
            x[k - 1] += alpha * d[i - 1];
        }
        __syncwarp();
    }
    return 0;
}

__device__ __noinline__ void ls_matupd(LsWork &w, int &itail, int iupdat, int &col, int &head,
                                 double &theta, double rr, double dr, double stp, double dtd,
                                 int lane)
{
    const int n = w.n, m = kLsM;
    const double *d = w.d, *r = w.r;
    if (iupdat <= m) {
        col = iupdat;
        itail = (head + iupdat - 2) % m + 1;
    } else {
        itail = itail % m + 1;
        head = head % m + 1;
    }
    #pragma unroll 1
    for (int i = 1 + lane; i <= n; i += SDA_WARP) { LWS(i, itail) = d[i - 1]; LWY(i, itail) = r[i - 1]; }
    __syncwarp();
    theta = rr / dr;
    const int c = col;
    if (iupdat > m) {
        #pragma unroll 1
        for (int j = 1; j <= c - 1; ++j) {
            #pragma unroll 1
            for (int i = 0; i < j; ++i) LSS(1 + i, j) = LSS(2 + i, j + 1);
            #pragma unroll 1
            for (int i = 0; i < c - j; ++i) LSY(j + i, j) = LSY(j + 1 + i, j + 1);
        }
    }
    int pointr = head;
    #pragma unroll 1
    for (int j = 1; j <= c - 1; ++j) {
        double a = 0.0, b = 0.0;
        #pragma unroll 1
        for (int i = 1 + lane; i <= n; i += SDA_WARP) {
            a += d[i - 1] * LWY(i, pointr);
            b += LWS(i, pointr) * d[i - 1];
        }
        LSY(c, j) = warp_sum(a);
        LSS(j, c) = warp_sum(b);
        pointr = pointr % m + 1;
    }
    if (stp == 1.0) LSS(c, c) = dtd;
    else LSS(c, c) = stp * stp * dtd;
    LSY(c, c) = dr;
}

__device__ inline double ls_projgr(int n, const double *l, const double *u, const double *x,
                                   const double *g, int lane)
{
    double s = 0.0;
    #pragma unroll 1
    for (int i = lane; i < n; i += SDA_WARP) {
        double gi = g[i];
        if (gi < 0.0) gi = fmax(x[i] - u[i], gi);
        else gi = fmin(x[i] - l[i], gi);
        s = fmax(s, fabs(gi));
    }
    return warp_max(s);
}

// ---- More'-Thuente line search (dcsrch / dcstep), scalar and warp-uniform ----------------------
struct LsSearch {
    int brackt, stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
};
enum { kLsFG = 0, kLsConv = 1, kLsWarn = 2, kLsError = 3 };

__device__ __noinline__ void ls_dcstep(double &stx, double &fx, double &dx, double &sty, double &fy,
                                 double &dy, double &stp, double fp, double dp, int &brackt,
                                 double stpmin, double stpmax)
{
    const double p66 = 0.66;
    double gamma, p, q, r, s, stpc, stpf, stpq, theta;
    const double sgnd = dp * (dx / fabs(dx));
    if (fp > fx) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc;
        else stpf = stpc + (stpq - stpc) / 2.0;
        brackt = 1;
    } else if (sgnd < 0.0) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
        else stpf = stpq;
        brackt = 1;
    } else if (fabs(dp) < fabs(dx)) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            if (stp > stx) stpf = fmin(stp + p66 * (sty - stp), stpf);
            else stpf = fmax(stp + p66 * (sty - stp), stpf);
        } else {
            if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {
        if (brackt) {
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) stpf = stpmax;
        else stpf = stpmin;
    }
    if (fp > fx) { sty = stp; fy = fp; dy = dp; }
    else {
        if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

__device__ __noinline__ int ls_dcsrch(LsSearch &s, int start, double f, double g, double &stp,
                                double ftol, double gtol, double xtol, double stpmin,
                                double stpmax)
{
    const double p5 = 0.5, p66 = 0.66, xtrapl = 1.1, xtrapu = 4.0;
    if (start) {
        if (stp < stpmin || stp > stpmax || g >= 0.0) return kLsError;
        s.brackt = 0; s.stage = 1;
        s.finit = f; s.ginit = g; s.gtest = ftol * s.ginit;
        s.width = stpmax - stpmin; s.width1 = s.width / p5;
        s.stx = 0.0; s.fx = s.finit; s.gx = s.ginit;
        s.sty = 0.0; s.fy = s.finit; s.gy = s.ginit;
        s.stmin = 0.0; s.stmax = stp + xtrapu * stp;
        return kLsFG;
    }
    const double ftest = s.finit + stp * s.gtest;
    if (s.stage == 1 && f <= ftest && g >= 0.0) s.stage = 2;
    int task = kLsFG;
    if (s.brackt && (stp <= s.stmin || stp >= s.stmax)) task = kLsWarn;
    if (s.brackt && s.stmax - s.stmin <= xtol * s.stmax) task = kLsWarn;
    if (stp == stpmax && f <= ftest && g <= s.gtest) task = kLsWarn;
    if (stp == stpmin && (f > ftest || g >= s.gtest)) task = kLsWarn;
    if (f <= ftest && fabs(g) <= gtol * (-s.ginit)) task = kLsConv;
    if (task != kLsFG) return task;
    if (s.stage == 1 && f <= s.fx && f > ftest) {
        const double fm = f - stp * s.gtest;
        double fxm = s.fx - s.stx * s.gtest;
        double fym = s.fy - s.sty * s.gtest;
        const double gm = g - s.gtest;
        double gxm = s.gx - s.gtest;
        double gym = s.gy - s.gtest;
        ls_dcstep(s.stx, fxm, gxm, s.sty, fym, gym, stp, fm, gm, s.brackt, s.stmin, s.stmax);
        s.fx = fxm + s.stx * s.gtest;
        s.fy = fym + s.sty * s.gtest;
        s.gx = gxm + s.gtest;
        s.gy = gym + s.gtest;
    } else {
        ls_dcstep(s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, stp, f, g, s.brackt, s.stmin, s.stmax);
    }
    if (s.brackt) {
        if (fabs(s.sty - s.stx) >= p66 * s.width1) stp = s.stx + p5 * (s.sty - s.stx);
        s.width1 = s.width;
        s.width = fabs(s.sty - s.stx);
    }
    if (s.brackt) { s.stmin = fmin(s.stx, s.sty); s.stmax = fmax(s.stx, s.sty); }
    else { s.stmin = stp + xtrapl * (stp - s.stx); s.stmax = stp + xtrapu * (stp - s.stx); }
    stp = fmax(stp, stpmin);
    stp = fmin(stp, stpmax);
    if ((s.brackt && (stp <= s.stmin || stp >= s.stmax)) ||
        (s.brackt && s.stmax - s.stmin <= xtol * s.stmax))
        stp = s.stx;
    return kLsFG;
}

// ---- K7: f and the clipped central-difference gradient at x                        :325-345 ----
// xe is the warp's shared-memory evaluation buffer.  Returns true when the suite monitor stopped
// the chain (the reference would have raised out of scipy at that call).
__device__ __noinline__ bool ls_fun_and_grad(const DeviceArgs &a, ChainState &st, LsWork &w,
                                       const double *x, double *xe, const double *lo,
                                       const double *up, double &f, double *g, int lane)
{
    const int n = a.dim;
    if (w.cache_valid) {
        int same = 1;
        for (int k = lane; k < n; k += SDA_WARP)
            same &= (__double_as_longlong(x[k]) == __double_as_longlong(w.xl[k]));
        if (__all_sync(SDA_FULL_MASK, same)) {
            for (int k = lane; k < n; k += SDA_WARP) g[k] = w.gl[k];
            f = w.fl;
            __syncwarp();
            return false;
        }
    }
    for (int k = lane; k < n; k += SDA_WARP) xe[k] = x[k];
    __syncwarp();
    f = func_wrapper<true>(a, st, xe, make_group(SDA_WARP, lane));   // ScalarFunction: f, then jac
    if (st.stop) return true;
    // The 2n finite-difference evaluations f(x + h_r e_i), f(x - h_l e_i), 32 at a time: lane L of
    // round r evaluates point e = 32 r + L by itself (even e: plus side of i = e/2, odd: minus).
    // Call counting and the suite monitor see them in the reference's order (:340-341).
    const long long mc0 = st.mon_calls;
    // separable objective: the base point's terms once, lane-striped (sda_objectives.cuh)
    const bool memo = a.ls_memo && objective_is_separable(a.functor);
    const bool memo2 = memo && separable_has_second_sum(a.functor);
    if (memo) {
        for (int k = lane; k < n; k += SDA_WARP) {
            double ta, tb;
            separable_term(a.functor, xe[k], a.params, ta, tb);
            w.ta[k] = ta;
            if (memo2) w.tb[k] = tb;
        }
        __syncwarp();
    }
    for (int e0 = 0; e0 < 2 * n; e0 += SDA_WARP) {
        const int e = e0 + lane;
        const bool valid = e < 2 * n;
        const int i = valid ? (e >> 1) : 0;
        const bool plus = (e & 1) == 0;
        const double xi = x[i];
        double respr = kGradEps, respl = kGradEps;
        double x1 = xi + respr;
        if (x1 > up[i]) { x1 = up[i]; respr = x1 - xi; }
        double x2 = xi - respl;
        if (x2 < lo[i]) { x2 = lo[i]; respl = xi - x2; }
        double fe = 0.0;
        if (valid && memo) {
            // f at the perturbed point: its own term i, every other term from the base point, summed in
            // the order of the one-lane evaluation -> the same bits as eval_objective_solo_ool
            double pa, pb, A = 0.0, B = 0.0;
            separable_term(a.functor, plus ? x1 : x2, a.params, pa, pb);
            #pragma unroll 4
            for (int k = 0; k < n; ++k) A += (k == i) ? pa : w.ta[k];
            if (memo2) {
                #pragma unroll 4
                for (int k = 0; k < n; ++k) B += (k == i) ? pb : w.tb[k];
            }
            fe = separable_finish(a.functor, A, B, n);
        } else if (valid) fe = eval_objective_solo_ool(a.functor, xe, i, plus ? x1 : x2, n, a.params, lane,
                                                (unsigned long long)(st.ncall + 1 + e));
        if (a.max_calls > 0) {
            const unsigned budget = __ballot_sync(SDA_FULL_MASK, valid && (mc0 + e + 1 >= a.max_calls));
            const unsigned hits = __ballot_sync(SDA_FULL_MASK, valid && (fe <= a.fglob_tol));
            const unsigned any = budget | hits;
            if (any) {
                const int first = __ffs(any) - 1;
                const long long calls = e0 + first + 1;
                st.ncall += calls; st.ncall_ls += calls; st.mon_calls = mc0 + calls;
                if (!((budget >> first) & 1u)) {                 // bench.py:314-320
                    st.hit = 1;
                    st.ncall_hit = st.mon_calls;
                    st.f_hit = __shfl_sync(SDA_FULL_MASK, fe, first);
                }
                st.stop = 1;
                return true;
            }
        }
        const double f2 = __shfl_down_sync(SDA_FULL_MASK, fe, 1);
        if (valid && plus) {
            double gi = (fe - f2) / (respl + respr);
            if (isnan(gi) || isinf(gi)) gi = 101.0;
            g[i] = gi;
            w.gl[i] = gi;
        }
    }
    st.ncall += 2 * n; st.ncall_ls += 2 * n;
    if (a.max_calls > 0) st.mon_calls = mc0 + 2 * n;
    __syncwarp();
    for (int k = lane; k < n; k += SDA_WARP) w.xl[k] = x[k];
    w.fl = f;
    w.cache_valid = 1;
    __syncwarp();
    return false;
}

// ---- the minimiser.  w.x holds the start point on entry and the final iterate on exit.
// Returns SciPy's warnflag: 0 converged, 1 maxiter/maxfun, 2 abnormal, -1 stopped by the monitor.
// Phase clock of the block-synchronous local-search launch (sda_ls_sync_kernel).  The warps of a
// CTA minimise different chains but walk the L-BFGS-B iteration in lock step: a cyclic clock of
// three phases -- kPhDir (Cauchy point, subspace minimisation, line-search set-up), kPhEval
// (function + finite-difference gradient), kPhStep (dcsrch/dcstep, convergence tests, BFGS update)
// -- advanced by CTA barriers.  A warp that has nothing to do in a phase just passes its barrier.
// All warps of an SM then fetch the same ~10-30 KB of code at a time instead of 16 different places
// of the ~120 KB the iteration touches (ncu: 70 % of the unsynchronised launch's stall samples in
// this code were instruction-fetch stalls).  NoClock compiles the barriers away.
enum LsPhase { kPhDir = 0, kPhEval = 1, kPhStep = 2 };
#ifndef SDA_LS_CLOCK_PHASES
#define SDA_LS_CLOCK_PHASES 2     // 2: {algebra, eval} -- kPhStep and kPhDir share a phase; 3: all apart
#endif
struct NoClock {
    __device__ __forceinline__ void enter(int) {}
};
struct PhaseClock {
    int cur;        // clock phase this warp is in (0 .. SDA_LS_CLOCK_PHASES-1)
    int nthreads;   // CTA size (all warps take part in every barrier)
    static __device__ __forceinline__ int slot(int phase)
    {
        return SDA_LS_CLOCK_PHASES == 3 ? phase : (phase == kPhEval ? 1 : 0);
    }
    // warps sign off in sign_off_slot(); the count is read only in the slot after it, i.e. behind a
    // barrier that every sign-off of the cycle precedes, so all warps read the same count
    static __device__ __forceinline__ int sign_off_slot() { return slot(kPhStep); }
    static __device__ __forceinline__ int read_slot() { return (slot(kPhStep) + 1) % SDA_LS_CLOCK_PHASES; }
    __device__ __forceinline__ void tick()
    {
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
        cur = (cur + 1 == SDA_LS_CLOCK_PHASES) ? 0 : cur + 1;
    }
    __device__ __forceinline__ void enter(int phase)
    {
        const int want = slot(phase);
        while (cur != want) tick();
    }
};

template <class Clock>
__device__ inline int lbfgsb_minimize(const DeviceArgs &a, ChainState &st, LsWork &w, double *xe,
                                      const double *l, const double *u, double &f_out, int maxiter,
                                      int lane, Clock &clk)
{
    const int n = w.n;
    double *x = w.x, *g = w.g;
    const double tol = kLsFactr * kLsEps;
    int warnflag = 2, n_iterations = 0, nfev = 0;
    double f = 0.0;

    for (int i = lane; i < n; i += SDA_WARP) {
        double xi = x[i];
        if (xi < l[i]) xi = l[i];
        if (xi > u[i]) xi = u[i];
        x[i] = xi;
        w.iwhere[i] = (u[i] - l[i] <= 0.0) ? 3 : 0;
    }
    __syncwarp();
    int col = 0, head = 1, itail = 0, iupdat = 0, updatd = 0, iback = 0, nfree = n, ifun = 0;
    int iter = 0, nenter = 0, ileave = 0, info = 0, wrk = 0;
    double theta = 1.0, fold = 0.0, sbgnrm, stp = 0.0, gd = 0.0, gdold = 0.0, dtd = 0.0, stpmx = 0.0;

    clk.enter(kPhEval);
    if (ls_fun_and_grad(a, st, w, x, xe, l, u, f, g, lane)) { f_out = f; return -1; }
    nfev = 1;
    sbgnrm = ls_projgr(n, l, u, x, g, lane);
    if (sbgnrm <= kLsPgtol) { f_out = f; return 0; }

    for (;;) {
        // ---------------- generalized Cauchy point ----------------
        clk.enter(kPhDir);
        __syncwarp();
        if (ls_cauchy(w, x, l, u, g, theta, col, head, sbgnrm, lane)) {
            info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
            continue;
        }
        ls_freev(w, nfree, nenter, ileave, wrk, updatd, iter);
        __syncwarp();
        // ---------------- subspace minimisation ----------------
        if (nfree != 0 && col != 0) {
            if (wrk) info = ls_formk(w, nfree, nenter, ileave, iupdat, updatd, theta, col, head, lane);
            else info = 0;
            if (info != 0) {
                info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
                continue;
            }
            info = ls_cmprlb(w, x, g, theta, col, head, nfree, lane);
            if (info == 0) info = ls_subsm(w, nfree, l, u, x, g, theta, col, head, lane);
            if (info != 0) {
                info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
                continue;
            }
        }
        // ---------------- line search along d = z - x ----------------
        __syncwarp();
        {
            double s1 = 0.0;
            for (int i = lane; i < n; i += SDA_WARP) {
                const double di = w.z[i] - x[i];
                w.d[i] = di;
                s1 += di * di;
                w.t[i] = x[i];
                w.r[i] = g[i];
            }
            dtd = warp_sum(s1);
        }
        __syncwarp();
        {
            LsSearch ls;
            bool restart = false, abnormal = false;
            stpmx = 1e10;
            if (iter == 0) stpmx = 1.0;
            else {
                // order dependent (running minimum with products): serial, warp-uniform
                for (int i = 0; i < n; ++i) {
                    const double a1 = w.d[i];
                    if (a1 < 0.0) {
                        const double a2 = l[i] - x[i];
                        if (a2 >= 0.0) stpmx = 0.0;
                        else if (a1 * stpmx < a2) stpmx = a2 / a1;
                    } else if (a1 > 0.0) {
                        const double a2 = u[i] - x[i];
                        if (a2 <= 0.0) stpmx = 0.0;
                        else if (a1 * stpmx > a2) stpmx = a2 / a1;
                    }
                }
            }
            stp = 1.0;   // all variables are boxed
            fold = f; ifun = 0; iback = 0;
            int first = 1;
            for (;;) {
                if (!first) clk.enter(kPhStep);     // the set-up call of dcsrch stays in kPhDir
                double s1 = 0.0;
                for (int i = lane; i < n; i += SDA_WARP) s1 += g[i] * w.d[i];
                gd = warp_sum(s1);
                info = 0;
                int lstask = kLsFG;
                if (ifun == 0) {
                    gdold = gd;
                    if (gd >= 0.0) info = -4;
                }
                if (info == 0) {
                    lstask = ls_dcsrch(ls, first, f, gd, stp, 1e-3, 0.9, 0.1, 0.0, stpmx);
                    first = 0;
                    if (lstask == kLsError) lstask = kLsFG;
                    if (lstask == kLsFG) {
                        ifun++; nfev++; iback = ifun - 1;
                        if (stp == 1.0) { for (int i = lane; i < n; i += SDA_WARP) x[i] = w.z[i]; }
                        else { for (int i = lane; i < n; i += SDA_WARP) x[i] = stp * w.d[i] + w.t[i]; }
                        __syncwarp();
                    }
                }
                if (info != 0 || iback >= kLsMaxls) {
                    for (int i = lane; i < n; i += SDA_WARP) { x[i] = w.t[i]; g[i] = w.r[i]; }
                    __syncwarp();
                    f = fold;
                    if (col == 0) {
                        if (info == 0) { info = -9; nfev--; ifun--; iback--; }
                        abnormal = true;
                        iter++;
                    } else {
                        if (info == 0) nfev--;
                        info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
                        restart = true;
                    }
                    break;
                }
                if (lstask == kLsFG) {
                    clk.enter(kPhEval);
                    if (ls_fun_and_grad(a, st, w, x, xe, l, u, f, g, lane)) { f_out = f; return -1; }
                    continue;
                }
                break;
            }
            if (abnormal) { warnflag = 2; break; }
            if (restart) continue;
        }
        iter++;
        sbgnrm = ls_projgr(n, l, u, x, g, lane);
        // SciPy driver, task NEW_X (_lbfgsb_py.py:415-428)
        n_iterations++;
        if (n_iterations >= maxiter) { warnflag = 1; break; }
        if (nfev > kLsMaxfun) { warnflag = 1; break; }
        if (sbgnrm <= kLsPgtol) { warnflag = 0; break; }
        {
            const double ddum = fmax(fabs(fold), fmax(fabs(f), 1.0));
            if ((fold - f) <= tol * ddum) { warnflag = 0; break; }
        }
        // ---------------- BFGS update ----------------
        {
            double s1 = 0.0;
            for (int i = lane; i < n; i += SDA_WARP) {
                const double ri = g[i] - w.r[i];
                w.r[i] = ri;
                s1 += ri * ri;
            }
            const double rr = warp_sum(s1);
            double dr, ddum;
            if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
            else {
                dr = (gd - gdold) * stp;
                for (int i = lane; i < n; i += SDA_WARP) w.d[i] *= stp;
                ddum = -gdold * stp;
            }
            __syncwarp();
            if (dr <= kLsEps * ddum) {
                updatd = 0;
            } else {
                updatd = 1; iupdat++;
                ls_matupd(w, itail, iupdat, col, head, theta, rr, dr, stp, dtd, lane);
                if (ls_formt(w, col, theta, lane) != 0) {
                    info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
                }
            }
        }
    }
    f_out = f;
    return warnflag;
}

// ObjectiveFunWrapper.local_search                                                 :347-351
// Start point x_start (any memory), result in w.x; returns 1 on success, e = 1e16 otherwise.
template <class Clock>
__device__ __noinline__ int ofw_local_search(const DeviceArgs &a, ChainState &st, LsWork &w,
                                       const double *x_start, double *xe, const double *lo,
                                       const double *up, double &e, int lane, Clock &clk)
{
    for (int k = lane; k < a.dim; k += SDA_WARP) w.x[k] = x_start[k];
    __syncwarp();
    st.n_ls += 1;
    w.cache_valid = 0;
    const int warnflag = lbfgsb_minimize(a, st, w, xe, lo, up, e, a.ls_maxiter, lane, clk);
    __syncwarp();
    if (warnflag != 0) { e = kBigValue; return 0; }
    return 1;
}

#undef LWS
#undef LWY
#undef LSY
#undef LSS
#undef LWT
#undef LWN
#undef LWN1

}  // namespace sda

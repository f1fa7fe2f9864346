/*
 * lbfgsb_oracle.c -- CPU ORACLE (test infrastructure): L-BFGS-B restated from the published
 * algorithm (see lbfgsb_oracle.h for the citations).  Serial, scalar, one problem at a time.
 * Array macros are 1-based / column-major so that the formulas read like the papers'.
 */
#include "lbfgsb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define EPSMCH DBL_EPSILON

typedef struct {
    int n, m;
    double *ws, *wy;          /* n x m : S and Y correction pairs (circular, head/col) */
    double *sy, *ss, *wt;     /* m x m */
    double *wn, *snd;         /* 2m x 2m */
    double *z, *r, *d, *t, *xp, *wa; /* n,n,n,n,n,8m */
    int *index, *iwhere, *indx2;     /* n each */
} Work;

#define WS(i, j) w->ws[((j) - 1) * n + ((i) - 1)]
#define WY(i, j) w->wy[((j) - 1) * n + ((i) - 1)]
#define SY(i, j) w->sy[((j) - 1) * m + ((i) - 1)]
#define SS(i, j) w->ss[((j) - 1) * m + ((i) - 1)]
#define WT(i, j) w->wt[((j) - 1) * m + ((i) - 1)]
#define WN(i, j) w->wn[((j) - 1) * 2 * m + ((i) - 1)]
#define WN1(i, j) w->snd[((j) - 1) * 2 * m + ((i) - 1)]

/* ---- small dense kernels (LINPACK dpofa / dtrsl semantics) --------------------------- */
/* Cholesky A = R'R of the leading nn x nn block, R in the upper triangle; lda = leading dim.
 * returns 0, or j (1-based) when the j-th leading minor is not positive definite. */
static int chol_upper(double *a, int lda, int nn)
{
    for (int j = 1; j <= nn; ++j) {
        double s = 0.0;
        for (int k = 1; k <= j - 1; ++k) {
            double dot = 0.0;
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

/* triangular solves with an upper-triangular T: trans=1 solves T'x=b, trans=0 solves Tx=b */
static int tri_solve_upper(const double *t, int ldt, int nn, double *b, int trans)
{
    for (int j = 1; j <= nn; ++j)
        if (t[(j - 1) * ldt + j - 1] == 0.0) return j;
    if (trans) {
        b[0] = b[0] / t[0];
        for (int j = 2; j <= nn; ++j) {
            double dot = 0.0;
            for (int i = 1; i <= j - 1; ++i) dot += t[(j - 1) * ldt + i - 1] * b[i - 1];
            b[j - 1] = (b[j - 1] - dot) / t[(j - 1) * ldt + j - 1];
        }
    } else {
        b[nn - 1] = b[nn - 1] / t[(nn - 1) * ldt + nn - 1];
        for (int jj = 2; jj <= nn; ++jj) {
            int j = nn - jj + 1;
            double temp = -b[j];
            for (int i = 1; i <= j; ++i) b[i - 1] += temp * t[j * ldt + i - 1];
            b[j - 1] = b[j - 1] / t[(j - 1) * ldt + j - 1];
        }
    }
    return 0;
}

/* ---- product of the 2m x 2m middle matrix of the compact L-BFGS formula with v -------- */
static int bmv(Work *w, int col, const double *v, double *p)
{
    const int m = w->m;
    if (col == 0) return 0;
    p[col] = v[col];
    for (int i = 2; i <= col; ++i) {
        int i2 = col + i;
        double sum = 0.0;
        for (int k = 1; k <= i - 1; ++k) sum += SY(i, k) * v[k - 1] / SY(k, k);
        p[i2 - 1] = v[i2 - 1] + sum;
    }
    if (tri_solve_upper(w->wt, m, col, p + col, 1)) return 1;
    for (int i = 1; i <= col; ++i) p[i - 1] = v[i - 1] / sqrt(SY(i, i));
    if (tri_solve_upper(w->wt, m, col, p + col, 0)) return 1;
    for (int i = 1; i <= col; ++i) p[i - 1] = -p[i - 1] / sqrt(SY(i, i));
    for (int i = 1; i <= col; ++i) {
        double sum = 0.0;
        for (int k = i + 1; k <= col; ++k) sum += SY(k, i) * p[col + k - 1] / SY(i, i);
        p[i - 1] += sum;
    }
    return 0;
}

/* ---- heap of breakpoints: extract the least element of t[1..nn] into t[nn] ------------- */
static void heap_least(int nn, double *t, int *iorder, int iheap)
{
    /* 1-based views */
    double *T = t - 1; int *IO = iorder - 1;
    if (iheap == 0) {
        for (int k = 2; k <= nn; ++k) {
            double ddum = T[k]; int indxin = IO[k]; int i = k;
            while (i > 1) {
                int j = i / 2;
                if (ddum < T[j]) { T[i] = T[j]; IO[i] = IO[j]; i = j; }
                else break;
            }
            T[i] = ddum; IO[i] = indxin;
        }
    }
    if (nn > 1) {
        int i = 1; double out = T[1]; int indxou = IO[1];
        double ddum = T[nn]; int indxin = IO[nn];
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

/* ---- generalized Cauchy point ----------------------------------------------------------- */
static int cauchy(Work *w, const double *x, const double *l, const double *u, const int *nbd,
                  const double *g, double theta, int col, int head, double sbgnrm, int *nseg)
{
    const int n = w->n, m = w->m;
    int *iorder = w->indx2, *iwhere = w->iwhere;
    double *t = w->t, *d = w->d, *xcp = w->z;
    double *p = w->wa, *c = w->wa + 2 * m, *wbp = w->wa + 4 * m, *v = w->wa + 6 * m;

    if (sbgnrm <= 0.0) { memcpy(xcp, x, sizeof(double) * n); return 0; }
    int bnded = 1, nfree = n + 1, nbreak = 0, ibkmin = 0;
    double bkmin = 0.0, f1 = 0.0;
    const int col2 = 2 * col;
    double tl = 0.0, tu = 0.0;
    for (int i = 0; i < col2; ++i) p[i] = 0.0;

    for (int i = 1; i <= n; ++i) {
        double neggi = -g[i - 1];
        if (iwhere[i - 1] != 3 && iwhere[i - 1] != -1) {
            if (nbd[i - 1] <= 2) tl = x[i - 1] - l[i - 1];
            if (nbd[i - 1] >= 2) tu = u[i - 1] - x[i - 1];
            int xlower = nbd[i - 1] <= 2 && tl <= 0.0;
            int xupper = nbd[i - 1] >= 2 && tu <= 0.0;
            iwhere[i - 1] = 0;
            if (xlower) { if (neggi <= 0.0) iwhere[i - 1] = 1; }
            else if (xupper) { if (neggi >= 0.0) iwhere[i - 1] = 2; }
            else { if (fabs(neggi) <= 0.0) iwhere[i - 1] = -3; }
        }
        int pointr = head;
        if (iwhere[i - 1] != 0 && iwhere[i - 1] != -1) {
            d[i - 1] = 0.0;
        } else {
            d[i - 1] = neggi;
            f1 -= neggi * neggi;
            for (int j = 1; j <= col; ++j) {
                p[j - 1] += WY(i, pointr) * neggi;
                p[col + j - 1] += WS(i, pointr) * neggi;
                pointr = pointr % m + 1;
            }
            if (nbd[i - 1] <= 2 && nbd[i - 1] != 0 && neggi < 0.0) {
                nbreak++; iorder[nbreak - 1] = i; t[nbreak - 1] = tl / (-neggi);
                if (nbreak == 1 || t[nbreak - 1] < bkmin) { bkmin = t[nbreak - 1]; ibkmin = nbreak; }
            } else if (nbd[i - 1] >= 2 && neggi > 0.0) {
                nbreak++; iorder[nbreak - 1] = i; t[nbreak - 1] = tu / neggi;
                if (nbreak == 1 || t[nbreak - 1] < bkmin) { bkmin = t[nbreak - 1]; ibkmin = nbreak; }
            } else {
                nfree--; iorder[nfree - 1] = i;
                if (fabs(neggi) > 0.0) bnded = 0;
            }
        }
    }
    if (theta != 1.0) for (int j = 0; j < col; ++j) p[col + j] *= theta;
    memcpy(xcp, x, sizeof(double) * n);
    if (nbreak == 0 && nfree == n + 1) return 0;
    for (int j = 0; j < col2; ++j) c[j] = 0.0;
    double f2 = -theta * f1;
    const double f2_org = f2;
    if (col > 0) {
        if (bmv(w, col, p, v)) return 1;
        double dot = 0.0;
        for (int j = 0; j < col2; ++j) dot += v[j] * p[j];
        f2 -= dot;
    }
    double dtm = -f1 / f2;
    double tsum = 0.0;
    *nseg = 1;
    int skip_tail = 0;
    if (nbreak != 0) {
        int nleft = nbreak, iter = 1;
        double tj = 0.0;
        for (;;) {
            double tj0 = tj;
            int ibp;
            if (iter == 1) {
                tj = bkmin; ibp = iorder[ibkmin - 1];
            } else {
                if (iter == 2) {
                    if (ibkmin != nbreak) { t[ibkmin - 1] = t[nbreak - 1]; iorder[ibkmin - 1] = iorder[nbreak - 1]; }
                }
                heap_least(nleft, t, iorder, iter - 2);
                tj = t[nleft - 1]; ibp = iorder[nleft - 1];
            }
            double dt = tj - tj0;
            if (dtm < dt) break;
            tsum += dt; nleft--; iter++;
            double dibp = d[ibp - 1];
            d[ibp - 1] = 0.0;
            double zibp;
            if (dibp > 0.0) { zibp = u[ibp - 1] - x[ibp - 1]; xcp[ibp - 1] = u[ibp - 1]; iwhere[ibp - 1] = 2; }
            else { zibp = l[ibp - 1] - x[ibp - 1]; xcp[ibp - 1] = l[ibp - 1]; iwhere[ibp - 1] = 1; }
            if (nleft == 0 && nbreak == n) { dtm = dt; skip_tail = 1; break; }
            (*nseg)++;
            double dibp2 = dibp * dibp;
            f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
            f2 = f2 - theta * dibp2;
            if (col > 0) {
                for (int j = 0; j < col2; ++j) c[j] += dt * p[j];
                int pointr = head;
                for (int j = 1; j <= col; ++j) {
                    wbp[j - 1] = WY(ibp, pointr);
                    wbp[col + j - 1] = theta * WS(ibp, pointr);
                    pointr = pointr % m + 1;
                }
                if (bmv(w, col, wbp, v)) return 1;
                double wmc = 0.0, wmp = 0.0, wmw = 0.0;
                for (int j = 0; j < col2; ++j) { wmc += c[j] * v[j]; wmp += p[j] * v[j]; wmw += wbp[j] * v[j]; }
                for (int j = 0; j < col2; ++j) p[j] -= dibp * wbp[j];
                f1 += dibp * wmc;
                f2 += 2.0 * dibp * wmp - dibp2 * wmw;
            }
            f2 = fmax(EPSMCH * f2_org, f2);
            if (nleft > 0) { dtm = -f1 / f2; continue; }
            else if (bnded) { f1 = 0.0; f2 = 0.0; dtm = 0.0; }
            else dtm = -f1 / f2;
            break;
        }
    }
    if (!skip_tail) {
        if (dtm <= 0.0) dtm = 0.0;
        tsum += dtm;
        for (int i = 0; i < n; ++i) xcp[i] += tsum * d[i];
    }
    if (col > 0) for (int j = 0; j < col2; ++j) c[j] += dtm * p[j];
    return 0;
}

/* ---- free / active index sets at the Cauchy point ---------------------------------------- */
static void freev(Work *w, int *nfree, int *nenter, int *ileave, int *wrk, int updatd, int cnstnd,
                  int iter)
{
    const int n = w->n;
    int *index = w->index, *indx2 = w->indx2, *iwhere = w->iwhere;
    *nenter = 0; *ileave = n + 1;
    if (iter > 0 && cnstnd) {
        for (int i = 1; i <= *nfree; ++i) {
            int k = index[i - 1];
            if (iwhere[k - 1] > 0) { (*ileave)--; indx2[*ileave - 1] = k; }
        }
        for (int i = 1 + *nfree; i <= n; ++i) {
            int k = index[i - 1];
            if (iwhere[k - 1] <= 0) { (*nenter)++; indx2[*nenter - 1] = k; }
        }
    }
    *wrk = (*ileave < n + 1) || (*nenter > 0) || updatd;
    *nfree = 0;
    int iact = n + 1;
    for (int i = 1; i <= n; ++i) {
        if (iwhere[i - 1] <= 0) { (*nfree)++; index[*nfree - 1] = i; }
        else { iact--; index[iact - 1] = i; }
    }
}

/* ---- LEL^T factorisation of the indefinite K matrix of the subspace problem --------------- */
static int formk(Work *w, int nsub, int nenter, int ileave, int iupdat, int updatd, double theta,
                 int col, int head)
{
    const int n = w->n, m = w->m;
    const int *ind = w->index, *indx2 = w->indx2;
    int upcl;
    if (updatd) {
        if (iupdat > m) {
            for (int jy = 1; jy <= m - 1; ++jy) {
                int js = m + jy;
                for (int i = 0; i < m - jy; ++i) WN1(jy + i, jy) = WN1(jy + 1 + i, jy + 1);
                for (int i = 0; i < m - jy; ++i) WN1(js + i, js) = WN1(js + 1 + i, js + 1);
                for (int i = 0; i < m - 1; ++i) WN1(m + 1 + i, jy) = WN1(m + 2 + i, jy + 1);
            }
        }
        int pbegin = 1, pend = nsub, dbegin = nsub + 1, dend = n;
        int iy = col, is = m + col;
        int ipntr = head + col - 1;
        if (ipntr > m) ipntr -= m;
        int jpntr = head;
        for (int jy = 1; jy <= col; ++jy) {
            int js = m + jy;
            double temp1 = 0.0, temp2 = 0.0, temp3 = 0.0;
            for (int k = pbegin; k <= pend; ++k) { int k1 = ind[k - 1]; temp1 += WY(k1, ipntr) * WY(k1, jpntr); }
            for (int k = dbegin; k <= dend; ++k) {
                int k1 = ind[k - 1];
                temp2 += WS(k1, ipntr) * WS(k1, jpntr);
                temp3 += WS(k1, ipntr) * WY(k1, jpntr);
            }
            WN1(iy, jy) = temp1; WN1(is, js) = temp2; WN1(is, jy) = temp3;
            jpntr = jpntr % m + 1;
        }
        int jy = col;
        jpntr = head + col - 1;
        if (jpntr > m) jpntr -= m;
        ipntr = head;
        for (int i = 1; i <= col; ++i) {
            int is2 = m + i;
            double temp3 = 0.0;
            for (int k = pbegin; k <= pend; ++k) { int k1 = ind[k - 1]; temp3 += WS(k1, ipntr) * WY(k1, jpntr); }
            ipntr = ipntr % m + 1;
            WN1(is2, jy) = temp3;
        }
        upcl = col - 1;
    } else {
        upcl = col;
    }
    int ipntr = head;
    for (int iy = 1; iy <= upcl; ++iy) {
        int is = m + iy;
        int jpntr = head;
        for (int jy = 1; jy <= iy; ++jy) {
            int js = m + jy;
            double temp1 = 0.0, temp2 = 0.0, temp3 = 0.0, temp4 = 0.0;
            for (int k = 1; k <= nenter; ++k) {
                int k1 = indx2[k - 1];
                temp1 += WY(k1, ipntr) * WY(k1, jpntr);
                temp2 += WS(k1, ipntr) * WS(k1, jpntr);
            }
            for (int k = ileave; k <= n; ++k) {
                int k1 = indx2[k - 1];
                temp3 += WY(k1, ipntr) * WY(k1, jpntr);
                temp4 += WS(k1, ipntr) * WS(k1, jpntr);
            }
            WN1(iy, jy) += temp1 - temp3;
            WN1(is, js) += -temp2 + temp4;
            jpntr = jpntr % m + 1;
        }
        ipntr = ipntr % m + 1;
    }
    ipntr = head;
    for (int is = m + 1; is <= m + upcl; ++is) {
        int jpntr = head;
        for (int jy = 1; jy <= upcl; ++jy) {
            double temp1 = 0.0, temp3 = 0.0;
            for (int k = 1; k <= nenter; ++k) { int k1 = indx2[k - 1]; temp1 += WS(k1, ipntr) * WY(k1, jpntr); }
            for (int k = ileave; k <= n; ++k) { int k1 = indx2[k - 1]; temp3 += WS(k1, ipntr) * WY(k1, jpntr); }
            if (is <= jy + m) WN1(is, jy) += temp1 - temp3;
            else WN1(is, jy) += -temp1 + temp3;
            jpntr = jpntr % m + 1;
        }
        ipntr = ipntr % m + 1;
    }
    const int m2 = 2 * m;
    for (int iy = 1; iy <= col; ++iy) {
        int is = col + iy, is1 = m + iy;
        for (int jy = 1; jy <= iy; ++jy) {
            int js = col + jy, js1 = m + jy;
            WN(jy, iy) = WN1(iy, jy) / theta;
            WN(js, is) = WN1(is1, js1) * theta;
        }
        for (int jy = 1; jy <= iy - 1; ++jy) WN(jy, is) = -WN1(is1, jy);
        for (int jy = iy; jy <= col; ++jy) WN(jy, is) = WN1(is1, jy);
        WN(iy, iy) += SY(iy, iy);
    }
    if (chol_upper(w->wn, m2, col)) return -1;
    const int col2 = 2 * col;
    for (int js = col + 1; js <= col2; ++js)
        if (tri_solve_upper(w->wn, m2, col, &WN(1, js), 1)) return -1;
    for (int is = col + 1; is <= col2; ++is)
        for (int js = is; js <= col2; ++js) {
            double dot = 0.0;
            for (int k = 1; k <= col; ++k) dot += WN(k, is) * WN(k, js);
            WN(is, js) += dot;
        }
    if (chol_upper(&WN(col + 1, col + 1), m2, col)) return -2;
    return 0;
}

/* ---- T = theta*S'S + L D^-1 L' and its Cholesky factor ----------------------------------- */
static int formt(Work *w, int col, double theta)
{
    const int m = w->m;
    for (int j = 1; j <= col; ++j) WT(1, j) = theta * SS(1, j);
    for (int i = 2; i <= col; ++i)
        for (int j = i; j <= col; ++j) {
            int k1 = (i < j ? i : j) - 1;
            double ddum = 0.0;
            for (int k = 1; k <= k1; ++k) ddum += SY(i, k) * SY(j, k) / SY(k, k);
            WT(i, j) = ddum + theta * SS(i, j);
        }
    return chol_upper(w->wt, m, col) ? -3 : 0;
}

/* ---- r = -Z'B(xcp - xk) - Z'g ----------------------------------------------------------- */
static int cmprlb(Work *w, const double *x, const double *g, double theta, int col, int head,
                  int nfree, int cnstnd)
{
    const int n = w->n, m = w->m;
    double *r = w->r, *z = w->z, *wa = w->wa;
    if (!cnstnd && col > 0) {
        for (int i = 0; i < n; ++i) r[i] = -g[i];
        return 0;
    }
    for (int i = 1; i <= nfree; ++i) {
        int k = w->index[i - 1];
        r[i - 1] = -theta * (z[k - 1] - x[k - 1]) - g[k - 1];
    }
    if (bmv(w, col, wa + 2 * m, wa)) return -8;
    int pointr = head;
    for (int j = 1; j <= col; ++j) {
        double a1 = wa[j - 1], a2 = theta * wa[col + j - 1];
        for (int i = 1; i <= nfree; ++i) {
            int k = w->index[i - 1];
            r[i - 1] += WY(k, pointr) * a1 + WS(k, pointr) * a2;
        }
        pointr = pointr % m + 1;
    }
    return 0;
}

/* ---- subspace minimisation with the Morales-Nocedal projection step ------------------------ */
static int subsm(Work *w, int nsub, const double *l, const double *u, const int *nbd,
                 const double *xx, const double *gg, double theta, int col, int head, int *iword)
{
    const int n = w->n, m = w->m;
    const int *ind = w->index;
    double *x = w->z, *d = w->r, *xp = w->xp, *wv = w->wa;
    if (nsub <= 0) return 0;
    int pointr = head;
    for (int i = 1; i <= col; ++i) {
        double temp1 = 0.0, temp2 = 0.0;
        for (int j = 1; j <= nsub; ++j) {
            int k = ind[j - 1];
            temp1 += WY(k, pointr) * d[j - 1];
            temp2 += WS(k, pointr) * d[j - 1];
        }
        wv[i - 1] = temp1;
        wv[col + i - 1] = theta * temp2;
        pointr = pointr % m + 1;
    }
    const int m2 = 2 * m, col2 = 2 * col;
    if (tri_solve_upper(w->wn, m2, col2, wv, 1)) return 1;
    for (int i = 0; i < col; ++i) wv[i] = -wv[i];
    if (tri_solve_upper(w->wn, m2, col2, wv, 0)) return 1;
    pointr = head;
    for (int jy = 1; jy <= col; ++jy) {
        int js = col + jy;
        for (int i = 1; i <= nsub; ++i) {
            int k = ind[i - 1];
            d[i - 1] += WY(k, pointr) * wv[jy - 1] / theta + WS(k, pointr) * wv[js - 1];
        }
        pointr = pointr % m + 1;
    }
    for (int i = 0; i < nsub; ++i) d[i] *= 1.0 / theta;

    *iword = 0;
    memcpy(xp, x, sizeof(double) * n);
    for (int i = 1; i <= nsub; ++i) {
        int k = ind[i - 1];
        double dk = d[i - 1], xk = x[k - 1];
        if (nbd[k - 1] != 0) {
            if (nbd[k - 1] == 1) {
                x[k - 1] = fmax(l[k - 1], xk + dk);
                if (x[k - 1] == l[k - 1]) *iword = 1;
            } else if (nbd[k - 1] == 2) {
                xk = fmax(l[k - 1], xk + dk);
                x[k - 1] = fmin(u[k - 1], xk);
                if (x[k - 1] == l[k - 1] || x[k - 1] == u[k - 1]) *iword = 1;
            } else if (nbd[k - 1] == 3) {
                x[k - 1] = fmin(u[k - 1], xk + dk);
                if (x[k - 1] == u[k - 1]) *iword = 1;
            }
        } else {
            x[k - 1] = xk + dk;
        }
    }
    if (*iword == 0) return 0;
    double dd_p = 0.0;
    for (int i = 0; i < n; ++i) dd_p += (x[i] - xx[i]) * gg[i];
    if (dd_p > 0.0) {
        memcpy(x, xp, sizeof(double) * n);
        double alpha = 1.0, temp1 = alpha;
        int ibd = 0;
        for (int i = 1; i <= nsub; ++i) {
            int k = ind[i - 1];
            double dk = d[i - 1];
            if (nbd[k - 1] != 0) {
                if (dk < 0.0 && nbd[k - 1] <= 2) {
                    double temp2 = l[k - 1] - x[k - 1];
                    if (temp2 >= 0.0) temp1 = 0.0;
                    else if (dk * alpha < temp2) temp1 = temp2 / dk;
                } else if (dk > 0.0 && nbd[k - 1] >= 2) {
                    double temp2 = u[k - 1] - x[k - 1];
                    if (temp2 <= 0.0) temp1 = 0.0;
                    else if (dk * alpha > temp2) temp1 = temp2 / dk;
                }
                if (temp1 < alpha) { alpha = temp1; ibd = i; }
            }
        }
        if (alpha < 1.0) {
            double dk = d[ibd - 1];
            int k = ind[ibd - 1];
            if (dk > 0.0) { x[k - 1] = u[k - 1]; d[ibd - 1] = 0.0; }
            else if (dk < 0.0) { x[k - 1] = l[k - 1]; d[ibd - 1] = 0.0; }
        }
        for (int i = 1; i <= nsub; ++i) {
            int k = ind[i - 1];
            x[k - 1] += alpha * d[i - 1];
        }
    }
    return 0;
}

/* ---- update of the limited-memory matrices -------------------------------------------------- */
static void matupd(Work *w, int *itail, int iupdat, int *col, int *head, double *theta, double rr,
                   double dr, double stp, double dtd)
{
    const int n = w->n, m = w->m;
    const double *d = w->d, *r = w->r;
    if (iupdat <= m) {
        *col = iupdat;
        *itail = (*head + iupdat - 2) % m + 1;
    } else {
        *itail = *itail % m + 1;
        *head = *head % m + 1;
    }
    for (int i = 1; i <= n; ++i) { WS(i, *itail) = d[i - 1]; WY(i, *itail) = r[i - 1]; }
    *theta = rr / dr;
    const int c = *col;
    if (iupdat > m) {
        for (int j = 1; j <= c - 1; ++j) {
            for (int i = 0; i < j; ++i) SS(1 + i, j) = SS(2 + i, j + 1);
            for (int i = 0; i < c - j; ++i) SY(j + i, j) = SY(j + 1 + i, j + 1);
        }
    }
    int pointr = *head;
    for (int j = 1; j <= c - 1; ++j) {
        double a = 0.0, b = 0.0;
        for (int i = 1; i <= n; ++i) { a += d[i - 1] * WY(i, pointr); b += WS(i, pointr) * d[i - 1]; }
        SY(c, j) = a;
        SS(j, c) = b;
        pointr = pointr % m + 1;
    }
    if (stp == 1.0) SS(c, c) = dtd;
    else SS(c, c) = stp * stp * dtd;
    SY(c, c) = dr;
}

/* ---- projected gradient sup-norm ------------------------------------------------------------- */
static double projgr(int n, const double *l, const double *u, const int *nbd, const double *x,
                     const double *g)
{
    double sbgnrm = 0.0;
    for (int i = 0; i < n; ++i) {
        double gi = g[i];
        if (nbd[i] != 0) {
            if (gi < 0.0) { if (nbd[i] >= 2) gi = fmax(x[i] - u[i], gi); }
            else { if (nbd[i] <= 2) gi = fmin(x[i] - l[i], gi); }
        }
        sbgnrm = fmax(sbgnrm, fabs(gi));
    }
    return sbgnrm;
}

/* ---- More'-Thuente line search (MINPACK-2 dcsrch / dcstep) ----------------------------------- */
typedef struct {
    int brackt, stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
} LsState;

enum { LS_FG = 0, LS_CONV = 1, LS_WARN = 2, LS_ERROR = 3 };

static void dcstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy,
                   double *stp, double fp, double dp, int *brackt, double stpmin, double stpmax)
{
    const double p66 = 0.66;
    double gamma, p, q, r, s, sgnd, stpc, stpf, stpq, theta;
    sgnd = dp * (*dx / fabs(*dx));
    if (fp > *fx) {
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
        if (*stp < *stx) gamma = -gamma;
        p = (gamma - *dx) + theta;
        q = ((gamma - *dx) + gamma) + dp;
        r = p / q;
        stpc = *stx + r * (*stp - *stx);
        stpq = *stx + ((*dx / ((*fx - fp) / (*stp - *stx) + *dx)) / 2.0) * (*stp - *stx);
        if (fabs(stpc - *stx) < fabs(stpq - *stx)) stpf = stpc;
        else stpf = stpc + (stpq - stpc) / 2.0;
        *brackt = 1;
    } else if (sgnd < 0.0) {
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
        if (*stp > *stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + *dx;
        r = p / q;
        stpc = *stp + r * (*stx - *stp);
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (fabs(stpc - *stp) > fabs(stpq - *stp)) stpf = stpc;
        else stpf = stpq;
        *brackt = 1;
    } else if (fabs(dp) < fabs(*dx)) {
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = fmax(fabs(theta), fmax(fabs(*dx), fabs(dp)));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (*dx / s) * (dp / s)));
        if (*stp > *stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (*dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = *stp + r * (*stx - *stp);
        else if (*stp > *stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (*brackt) {
            if (fabs(stpc - *stp) < fabs(stpq - *stp)) stpf = stpc;
            else stpf = stpq;
            if (*stp > *stx) stpf = fmin(*stp + p66 * (*sty - *stp), stpf);
            else stpf = fmax(*stp + p66 * (*sty - *stp), stpf);
        } else {
            if (fabs(stpc - *stp) > fabs(stpq - *stp)) stpf = stpc;
            else stpf = stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {
        if (*brackt) {
            theta = 3.0 * (fp - *fy) / (*sty - *stp) + *dy + dp;
            s = fmax(fabs(theta), fmax(fabs(*dy), fabs(dp)));
            gamma = s * sqrt((theta / s) * (theta / s) - (*dy / s) * (dp / s));
            if (*stp > *sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + *dy;
            r = p / q;
            stpc = *stp + r * (*sty - *stp);
            stpf = stpc;
        } else if (*stp > *stx) stpf = stpmax;
        else stpf = stpmin;
    }
    if (fp > *fx) { *sty = *stp; *fy = fp; *dy = dp; }
    else {
        if (sgnd < 0.0) { *sty = *stx; *fy = *fx; *dy = *dx; }
        *stx = *stp; *fx = fp; *dx = dp;
    }
    *stp = stpf;
}

/* one call of the line-search state machine; start != 0 on the first call of a search */
static int dcsrch(LsState *s, int start, double f, double g, double *stp, double ftol, double gtol,
                  double xtol, double stpmin, double stpmax)
{
    const double p5 = 0.5, p66 = 0.66, xtrapl = 1.1, xtrapu = 4.0;
    if (start) {
        if (*stp < stpmin || *stp > stpmax || g >= 0.0 || ftol < 0.0 || gtol < 0.0 || xtol < 0.0 ||
            stpmin < 0.0 || stpmax < stpmin)
            return LS_ERROR;
        s->brackt = 0; s->stage = 1;
        s->finit = f; s->ginit = g; s->gtest = ftol * s->ginit;
        s->width = stpmax - stpmin; s->width1 = s->width / p5;
        s->stx = 0.0; s->fx = s->finit; s->gx = s->ginit;
        s->sty = 0.0; s->fy = s->finit; s->gy = s->ginit;
        s->stmin = 0.0; s->stmax = *stp + xtrapu * *stp;
        return LS_FG;
    }
    double ftest = s->finit + *stp * s->gtest;
    if (s->stage == 1 && f <= ftest && g >= 0.0) s->stage = 2;
    int task = LS_FG;
    if (s->brackt && (*stp <= s->stmin || *stp >= s->stmax)) task = LS_WARN;
    if (s->brackt && s->stmax - s->stmin <= xtol * s->stmax) task = LS_WARN;
    if (*stp == stpmax && f <= ftest && g <= s->gtest) task = LS_WARN;
    if (*stp == stpmin && (f > ftest || g >= s->gtest)) task = LS_WARN;
    if (f <= ftest && fabs(g) <= gtol * (-s->ginit)) task = LS_CONV;
    if (task != LS_FG) return task;
    if (s->stage == 1 && f <= s->fx && f > ftest) {
        double fm = f - *stp * s->gtest;
        double fxm = s->fx - s->stx * s->gtest;
        double fym = s->fy - s->sty * s->gtest;
        double gm = g - s->gtest;
        double gxm = s->gx - s->gtest;
        double gym = s->gy - s->gtest;
        dcstep(&s->stx, &fxm, &gxm, &s->sty, &fym, &gym, stp, fm, gm, &s->brackt, s->stmin, s->stmax);
        s->fx = fxm + s->stx * s->gtest;
        s->fy = fym + s->sty * s->gtest;
        s->gx = gxm + s->gtest;
        s->gy = gym + s->gtest;
    } else {
        dcstep(&s->stx, &s->fx, &s->gx, &s->sty, &s->fy, &s->gy, stp, f, g, &s->brackt, s->stmin, s->stmax);
    }
    if (s->brackt) {
        if (fabs(s->sty - s->stx) >= p66 * s->width1) *stp = s->stx + p5 * (s->sty - s->stx);
        s->width1 = s->width;
        s->width = fabs(s->sty - s->stx);
    }
    if (s->brackt) { s->stmin = fmin(s->stx, s->sty); s->stmax = fmax(s->stx, s->sty); }
    else { s->stmin = *stp + xtrapl * (*stp - s->stx); s->stmax = *stp + xtrapu * (*stp - s->stx); }
    *stp = fmax(*stp, stpmin);
    *stp = fmin(*stp, stpmax);
    if ((s->brackt && (*stp <= s->stmin || *stp >= s->stmax)) ||
        (s->brackt && s->stmax - s->stmin <= xtol * s->stmax))
        *stp = s->stx;
    return LS_FG;
}

/* ---- driver ---------------------------------------------------------------------------------- */
int orc_lbfgsb_minimize(int n, const double *l, const double *u, double *x, double *f_out,
                        orc_fg_fn fg, void *ctx, const OrcLbfgsbOpts *o, int *nit_out, int *nfev_out)
{
    const int m = o->m;
    Work wk, *w = &wk;
    w->n = n; w->m = m;
    size_t nd = (size_t)2 * m * n + 5 * (size_t)n + 11 * (size_t)m * m + 8 * (size_t)m + (size_t)n;
    double *buf = (double *)calloc(nd, sizeof(double));
    int *ibuf = (int *)calloc((size_t)4 * n, sizeof(int));
    if (!buf || !ibuf) { free(buf); free(ibuf); return 2; }
    double *q = buf;
    w->ws = q; q += (size_t)m * n;  w->wy = q; q += (size_t)m * n;
    w->sy = q; q += m * m;  w->ss = q; q += m * m;  w->wt = q; q += m * m;
    w->wn = q; q += 4 * m * m;  w->snd = q; q += 4 * m * m;
    w->z = q; q += n;  w->r = q; q += n;  w->d = q; q += n;  w->t = q; q += n;  w->xp = q; q += n;
    w->wa = q; q += 8 * m;
    double *g = q; q += n;
    w->index = ibuf; w->iwhere = ibuf + n; w->indx2 = ibuf + 2 * n;
    int *nbd = ibuf + 3 * n;

    int warnflag = 2;
    int n_iterations = 0, nfev = 0;
    double f = 0.0;
    const double tol = o->factr * EPSMCH;

    /* SciPy driver: bounds are all finite here -> nbd = 2; x0 clipped into the box */
    for (int i = 0; i < n; ++i) {
        nbd[i] = 2;
        if (x[i] < l[i]) x[i] = l[i];
        if (x[i] > u[i]) x[i] = u[i];
    }

    int col = 0, head = 1, itail = 0, iupdat = 0, updatd = 0, iback = 0, nfree = n, ifun = 0;
    int iter = 0, nenter = 0, ileave = 0, nseg = 0, info = 0, iword = 0, wrk = 0;
    double theta = 1.0, fold = 0.0, dnorm = 0.0, sbgnrm, stp = 0.0, gd = 0.0, gdold = 0.0, dtd = 0.0;
    double stpmx = 0.0;
    int prjctd = 0, cnstnd = 0, boxed = 1;
    (void)prjctd; (void)iword; (void)dnorm;

    /* initial projection / variable status */
    for (int i = 0; i < n; ++i) {
        if (nbd[i] > 0) {
            if (nbd[i] <= 2 && x[i] <= l[i]) { if (x[i] < l[i]) { prjctd = 1; x[i] = l[i]; } }
            else if (nbd[i] >= 2 && x[i] >= u[i]) { if (x[i] > u[i]) { prjctd = 1; x[i] = u[i]; } }
        }
    }
    for (int i = 0; i < n; ++i) {
        if (nbd[i] != 2) boxed = 0;
        if (nbd[i] == 0) w->iwhere[i] = -1;
        else {
            cnstnd = 1;
            if (nbd[i] == 2 && u[i] - l[i] <= 0.0) w->iwhere[i] = 3;
            else w->iwhere[i] = 0;
        }
    }

    if (fg(ctx, x, &f, g)) { warnflag = -1; goto done; }
    nfev = 1;
    sbgnrm = projgr(n, l, u, nbd, x, g);
    if (sbgnrm <= o->pgtol) { warnflag = 0; goto done; }

    for (;;) {
        /* ---------------- new iteration: Cauchy point ---------------- */
        iword = -1;
        if (!cnstnd && col > 0) {
            memcpy(w->z, x, sizeof(double) * n);
            wrk = updatd; nseg = 0;
        } else {
            if (cauchy(w, x, l, u, nbd, g, theta, col, head, sbgnrm, &nseg)) {
                info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
                continue;
            }
            freev(w, &nfree, &nenter, &ileave, &wrk, updatd, cnstnd, iter);
        }
        /* ---------------- subspace minimisation ---------------- */
        if (nfree != 0 && col != 0) {
            if (wrk) info = formk(w, nfree, nenter, ileave, iupdat, updatd, theta, col, head);
            else info = 0;
            if (info != 0) {
                info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
                continue;
            }
            info = cmprlb(w, x, g, theta, col, head, nfree, cnstnd);
            if (info == 0) info = subsm(w, nfree, l, u, nbd, x, g, theta, col, head, &iword);
            if (info != 0) {
                info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
                continue;
            }
        }
        /* ---------------- line search along d = z - x ---------------- */
        for (int i = 0; i < n; ++i) w->d[i] = w->z[i] - x[i];
        {
            LsState ls;
            int restart = 0, abnormal = 0;
            dtd = 0.0;
            for (int i = 0; i < n; ++i) dtd += w->d[i] * w->d[i];
            dnorm = sqrt(dtd);
            stpmx = 1e10;
            if (cnstnd) {
                if (iter == 0) stpmx = 1.0;
                else {
                    for (int i = 0; i < n; ++i) {
                        double a1 = w->d[i];
                        if (nbd[i] != 0) {
                            if (a1 < 0.0 && nbd[i] <= 2) {
                                double a2 = l[i] - x[i];
                                if (a2 >= 0.0) stpmx = 0.0;
                                else if (a1 * stpmx < a2) stpmx = a2 / a1;
                            } else if (a1 > 0.0 && nbd[i] >= 2) {
                                double a2 = u[i] - x[i];
                                if (a2 <= 0.0) stpmx = 0.0;
                                else if (a1 * stpmx > a2) stpmx = a2 / a1;
                            }
                        }
                    }
                }
            }
            if (iter == 0 && !boxed) stp = fmin(1.0 / dnorm, stpmx);
            else stp = 1.0;
            memcpy(w->t, x, sizeof(double) * n);
            memcpy(w->r, g, sizeof(double) * n);
            fold = f; ifun = 0; iback = 0;
            int first = 1;
            for (;;) {
                gd = 0.0;
                for (int i = 0; i < n; ++i) gd += g[i] * w->d[i];
                info = 0;
                int lstask;
                if (ifun == 0) {
                    gdold = gd;
                    if (gd >= 0.0) { info = -4; }
                }
                if (info == 0) {
                    double stp_in = stp;
                    lstask = dcsrch(&ls, first, f, gd, &stp, 1e-3, 0.9, 0.1, 0.0, stpmx);
                    if (getenv("ORC_TRACE_LNS"))
                        fprintf(stderr, "lns iter=%d ifun=%d first=%d f=%.17g gd=%.6e stp_in=%.17g stp_out=%.17g task=%d stpmx=%.6e col=%d\n",
                                iter, ifun, first, f, gd, stp_in, stp, lstask, stpmx, col);
                    first = 0;
                    if (lstask == LS_ERROR) lstask = LS_FG; /* treated like the reference: keeps going */
                    if (lstask == LS_FG) {
                        ifun++; nfev++; iback = ifun - 1;
                        if (stp == 1.0) memcpy(x, w->z, sizeof(double) * n);
                        else for (int i = 0; i < n; ++i) x[i] = stp * w->d[i] + w->t[i];
                    }
                } else {
                    lstask = LS_FG;
                }
                if (info != 0 || iback >= o->maxls) {
                    /* restore the previous iterate */
                    memcpy(x, w->t, sizeof(double) * n);
                    memcpy(g, w->r, sizeof(double) * n);
                    f = fold;
                    if (col == 0) {
                        if (info == 0) { info = -9; nfev--; ifun--; iback--; }
                        abnormal = 1;
                        iter++;
                    } else {
                        if (info == 0) nfev--;
                        info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
                        restart = 1;
                    }
                    break;
                }
                if (lstask == LS_FG) {
                    if (fg(ctx, x, &f, g)) { warnflag = -1; goto done; }
                    continue;
                }
                break; /* line search converged (or warned): new iterate accepted */
            }
            if (abnormal) { warnflag = 2; goto done; }
            if (restart) continue;
        }
        iter++;
        sbgnrm = projgr(n, l, u, nbd, x, g);
        /* SciPy driver, task NEW_X */
        n_iterations++;
        if (n_iterations >= o->maxiter) { warnflag = 1; goto done; }
        if (nfev > o->maxfun) { warnflag = 1; goto done; }
        /* termination tests */
        if (sbgnrm <= o->pgtol) { warnflag = 0; goto done; }
        {
            double ddum = fmax(fabs(fold), fmax(fabs(f), 1.0));
            if ((fold - f) <= tol * ddum) { warnflag = 0; goto done; }
        }
        /* ---------------- BFGS update ---------------- */
        {
            double rr = 0.0, dr, ddum;
            for (int i = 0; i < n; ++i) { w->r[i] = g[i] - w->r[i]; rr += w->r[i] * w->r[i]; }
            if (stp == 1.0) { dr = gd - gdold; ddum = -gdold; }
            else {
                dr = (gd - gdold) * stp;
                for (int i = 0; i < n; ++i) w->d[i] *= stp;
                ddum = -gdold * stp;
            }
            if (dr <= EPSMCH * ddum) {
                updatd = 0;
            } else {
                updatd = 1; iupdat++;
                matupd(w, &itail, iupdat, &col, &head, &theta, rr, dr, stp, dtd);
                if (formt(w, col, theta) != 0) {
                    info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
                }
            }
        }
    }
done:
    *f_out = f;
    if (nit_out) *nit_out = n_iterations;
    if (nfev_out) *nfev_out = nfev;
    free(buf); free(ibuf);
    return warnflag;
}

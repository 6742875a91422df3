/*
 * sl_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never on the product path).
 *
 * Plain-C restatement of the single-layer EBCM T-matrix algorithm that hydroscatt reaches through
 * the third-party package `pytmatrix` (un-vendored, version unpinned: reference README.md:9;
 * call sites scattering.py:27,478-525,555-585, scattering_rain.py:27-29,143-162,183-191,276-283).
 * pytmatrix wraps Mishchenko's published double-precision code `ampld.lp.f`
 * (Mishchenko & Travis, JQSRT 60, 309 (1998); Mishchenko, Appl. Opt. 39, 1026 (2000)) as the two
 * f2py entry points `calctmat` and `calcampl`.  The algorithm restated here is the published one,
 * frozen in SURVEY.md rows P2-P8 and Appendix A (A.1-A.8).
 *
 * PARITY PINNING: the reference tree holds no golden vectors.  This oracle is pinned to the one
 * upstream known-answer vector of pytmatrix's own test-suite (test_tmatrix.py::test_single,
 * radius=2, wavelength=6.5, m=1.5+0.5j, axis_ratio=1/0.6; see tests/golden/kat_single.json) and
 * to physics identities (Mie limit, reciprocity, optical theorem).  Everything else: "parity
 * unpinned" (see DESIGN.md).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.
 *
 * Layout of a stored T-matrix (shared with the CUDA engine, see include/hydrotm.h):
 *   for m = 0..nmax:  nmin = max(m,1), NM = nmax-nmin+1, four NM x NM complex blocks
 *   (T11,T12,T21,T22), row-major [n-nmin][n'-nmin], interleaved (re,im);
 *   block m starts at complex offset 4*sum_{m'<m} NM(m')^2.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>

#define HSO_NPN1 100   /* upstream ampld.par.f limit on NMAX */
#define HSO_NPNG1 500  /* upstream limit on NGAUSS */
#define HSO_MAXTRIALS 1024

typedef double complex cplx;

typedef struct {
    int nmax, ngauss, ntrials, status; /* status: 0 ok, 1 nmax limit, 2 ngauss limit (kept going), 3 singular */
    int trial_nmax[HSO_MAXTRIALS], trial_ngauss[HSO_MAXTRIALS];
    double trial_qsca[HSO_MAXTRIALS], trial_qext[HSO_MAXTRIALS];
    double *t; /* interleaved complex, layout above */
    long tlen; /* number of complex entries */
} hso_tmat;

/* ------------------------------------------------------------------ quadrature (GAUSS) */
/* Gauss-Legendre nodes/weights on [-1,1] by Newton iteration on P_n, ascending order. */
void hso_gauss(int n, double *z, double *w)
{
    int ind = n % 2, k = n / 2 + ind;
    double f = (double)n, x = 0.0;
    for (int i = 1; i <= k; i++) {
        int m = n + 1 - i;
        if (i == 1) x = 1.0 - 2.0 / ((f + 1.0) * f);
        if (i == 2) x = (z[n - 1] - 1.0) * 4.0 + z[n - 1];
        if (i == 3) x = (z[n - 2] - z[n - 1]) * 1.6 + z[n - 2];
        if (i > 3) x = (z[m] - z[m + 1]) * 3.0 + z[m + 2];
        if (i == k && ind == 1) x = 0.0;
        int niter = 0;
        double check = 1e-16, pa, pb, pc;
        for (;;) {
            pb = 1.0;
            niter++;
            if (niter > 100) check *= 10.0;
            pc = x;
            double dj = 1.0;
            for (int j = 2; j <= n; j++) {
                dj += 1.0;
                pa = pb;
                pb = pc;
                pc = x * pb + (x * pb - pa) * (dj - 1.0) / dj;
            }
            pa = 1.0 / ((pb - x * pc) * f);
            pb = pa * pc * (1.0 - x * x);
            x -= pb;
            if (!(fabs(pb) > check * fabs(x))) break;
        }
        z[m - 1] = x;
        w[m - 1] = 2.0 * pa * pa * (1.0 - x * x);
        if (!(i == k && ind == 1)) {
            z[i - 1] = -z[m - 1];
            w[i - 1] = w[m - 1];
        }
    }
}

/* ------------------------------------------------------------------ radial functions */
/* j_n(x), (1/x)d[x j_n]/dx, n=1..nmax by downward ratio recurrence started nnmax orders above. */
static void rjb(double x, double *y, double *u, int nmax, int nnmax, double *zw)
{
    int l = nmax + nnmax;
    double xx = 1.0 / x;
    zw[l] = 1.0 / ((double)(2 * l + 1) * xx);
    for (int i = 1; i <= l - 1; i++) {
        int i1 = l - i;
        zw[i1] = 1.0 / ((double)(2 * i1 + 1) * xx - zw[i1 + 1]);
    }
    double z0 = 1.0 / (xx - zw[1]);
    double y0 = z0 * cos(x) * xx;
    double y1 = y0 * zw[1];
    u[1] = y0 - y1 * xx;
    y[1] = y1;
    for (int i = 2; i <= nmax; i++) {
        double yi1 = y[i - 1];
        double yi = yi1 * zw[i];
        u[i] = yi1 - (double)i * yi * xx;
        y[i] = yi;
    }
}

/* y_n(x) and derivative form by upward recurrence. */
static void ryb(double x, double *y, double *v, int nmax)
{
    double c = cos(x), s = sin(x);
    double x1 = 1.0 / x, x2 = x1 * x1, x3 = x2 * x1;
    double y1 = -c * x2 - s * x1;
    y[1] = y1;
    if (nmax >= 2) y[2] = (-3.0 * x3 + x1) * c - 3.0 * x2 * s;
    for (int i = 2; i <= nmax - 1; i++)
        y[i + 1] = (double)(2 * i + 1) * x1 * y[i] - y[i - 1];
    v[1] = -x1 * (c + y1);
    for (int i = 2; i <= nmax; i++)
        v[i] = y[i - 1] - (double)i * x1 * y[i];
}

/* j_n(z) and derivative form for complex z = xr + i xi. */
static void cjb(double xr, double xi, double *yr, double *yi, double *ur, double *ui,
                int nmax, int nnmax, double *czr, double *czi)
{
    int l = nmax + nnmax;
    double xrxi = 1.0 / (xr * xr + xi * xi);
    double cxxr = xr * xrxi, cxxi = -xi * xrxi;
    double qf = 1.0 / (double)(2 * l + 1);
    czr[l] = xr * qf;
    czi[l] = xi * qf;
    for (int i = 1; i <= l - 1; i++) {
        int i1 = l - i;
        qf = (double)(2 * i1 + 1);
        double ar = qf * cxxr - czr[i1 + 1];
        double ai = qf * cxxi - czi[i1 + 1];
        double ari = 1.0 / (ar * ar + ai * ai);
        czr[i1] = ar * ari;
        czi[i1] = -ai * ari;
    }
    double ar = cxxr - czr[1], ai = cxxi - czi[1];
    double ari = 1.0 / (ar * ar + ai * ai);
    double cz0r = ar * ari, cz0i = -ai * ari;
    double cr = cos(xr) * cosh(xi), ci = -sin(xr) * sinh(xi);
    ar = cz0r * cr - cz0i * ci;
    ai = cz0i * cr + cz0r * ci;
    double cy0r = ar * cxxr - ai * cxxi, cy0i = ai * cxxr + ar * cxxi;
    double cy1r = cy0r * czr[1] - cy0i * czi[1], cy1i = cy0i * czr[1] + cy0r * czi[1];
    ar = cy1r * cxxr - cy1i * cxxi;
    ai = cy1i * cxxr + cy1r * cxxi;
    yr[1] = cy1r; yi[1] = cy1i;
    ur[1] = cy0r - ar; ui[1] = cy0i - ai;
    for (int i = 2; i <= nmax; i++) {
        double qi = (double)i;
        double cyi1r = yr[i - 1], cyi1i = yi[i - 1];
        double cyir = cyi1r * czr[i] - cyi1i * czi[i];
        double cyii = cyi1i * czr[i] + cyi1r * czi[i];
        ar = cyir * cxxr - cyii * cxxi;
        ai = cyii * cxxr + cyir * cxxi;
        yr[i] = cyir; yi[i] = cyii;
        ur[i] = cyi1r - qi * ar; ui[i] = cyi1i - qi * ai;
    }
}

/* ------------------------------------------------------------------ angular functions (VIG) */
/* d^n_{0m}(theta) and d/dtheta, n=1..nmax, x=cos(theta); entries n<m are zero. */
static void vig(double x, int nmax, int m, double *dv1, double *dv2)
{
    double a = 1.0, qs = sqrt(1.0 - x * x), qs1 = 1.0 / qs;
    for (int n = 1; n <= nmax; n++) { dv1[n] = 0.0; dv2[n] = 0.0; }
    if (m == 0) {
        double d1 = 1.0, d2 = x;
        for (int n = 1; n <= nmax; n++) {
            double qn = n, qn1 = n + 1, qn2 = 2 * n + 1;
            double d3 = (qn2 * x * d2 - qn * d1) / qn1;
            double der = qs1 * (qn1 * qn / qn2) * (-d1 + d3);
            dv1[n] = d2; dv2[n] = der;
            d1 = d2; d2 = d3;
        }
        return;
    }
    double qmm = (double)(m * m);
    for (int i = 1; i <= m; i++) a = a * sqrt((double)(2 * i - 1) / (double)(2 * i)) * qs;
    double d1 = 0.0, d2 = a;
    for (int n = m; n <= nmax; n++) {
        double qn = n, qn2 = 2 * n + 1, qn1 = n + 1;
        double qnm = sqrt(qn * qn - qmm), qnm1 = sqrt(qn1 * qn1 - qmm);
        double d3 = (qn2 * x * d2 - qnm * d1) / qnm1;
        double der = qs1 * (-qn1 * qnm * d1 + qn * qnm1 * d3) / qn2;
        dv1[n] = d2; dv2[n] = der;
        d1 = d2; d2 = d3;
    }
}

/* pi_n = d/sin and tau_n at a single direction, with the |x|->1 limits (VIGAMPL). */
static void vigampl(double x, int nmax, int m, double *dv1, double *dv2)
{
    for (int n = 1; n <= nmax; n++) { dv1[n] = 0.0; dv2[n] = 0.0; }
    if (fabs(1.0 - fabs(x)) <= 1e-10) {
        if (m != 1) return;
        for (int n = 1; n <= nmax; n++) {
            double dn = 0.5 * sqrt((double)(n * (n + 1)));
            if (x < 0.0) dn = dn * (((n + 1) % 2) ? -1.0 : 1.0);
            dv1[n] = dn;
            if (x < 0.0) dn = -dn;
            dv2[n] = dn;
        }
        return;
    }
    double a = 1.0, qs = sqrt(1.0 - x * x), qs1 = 1.0 / qs, dsi = qs1;
    if (m == 0) {
        double d1 = 1.0, d2 = x;
        for (int n = 1; n <= nmax; n++) {
            double qn = n, qn1 = n + 1, qn2 = 2 * n + 1;
            double d3 = (qn2 * x * d2 - qn * d1) / qn1;
            double der = qs1 * (qn1 * qn / qn2) * (-d1 + d3);
            dv1[n] = d2 * dsi; dv2[n] = der;
            d1 = d2; d2 = d3;
        }
        return;
    }
    double qmm = (double)(m * m);
    for (int i = 1; i <= m; i++) a = a * sqrt((double)(2 * i - 1) / (double)(2 * i)) * qs;
    double d1 = 0.0, d2 = a;
    for (int n = m; n <= nmax; n++) {
        double qn = n, qn2 = 2 * n + 1, qn1 = n + 1;
        double qnm = sqrt(qn * qn - qmm), qnm1 = sqrt(qn1 * qn1 - qmm);
        double d3 = (qn2 * x * d2 - qnm * d1) / qnm1;
        double der = qs1 * (-qn1 * qnm * d1 + qn * qnm1 * d3) / qn2;
        dv1[n] = d2 * dsi; dv2[n] = der;
        d1 = d2; d2 = d3;
    }
}

/* ------------------------------------------------------------------ equal-surface-area ratio (SAREA) */
static double sarea(double d)
{
    double e, r;
    if (d < 1.0) {
        e = sqrt(1.0 - d * d);
        r = 0.5 * (pow(d, 2.0 / 3.0) + pow(d, -1.0 / 3.0) * asin(e) / e);
    } else {
        e = sqrt(1.0 - 1.0 / (d * d));
        r = 0.25 * (2.0 * pow(d, 2.0 / 3.0) + pow(d, -4.0 / 3.0) * log((1.0 + e) / (1.0 - e)) / e);
    }
    return 1.0 / sqrt(r);
}

/* ------------------------------------------------------------------ complex LU inverse (TT: ZGETRF+ZGETRI) */
static int lu_invert(cplx *a, int n, cplx *inv, int *piv)
{
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = fabs(creal(a[k * n + k])) + fabs(cimag(a[k * n + k]));
        for (int i = k + 1; i < n; i++) {
            double v = fabs(creal(a[i * n + k])) + fabs(cimag(a[i * n + k]));
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (best == 0.0) return 1;
        if (p != k)
            for (int j = 0; j < n; j++) { cplx t = a[k * n + j]; a[k * n + j] = a[p * n + j]; a[p * n + j] = t; }
        cplx rp = 1.0 / a[k * n + k];
        for (int i = k + 1; i < n; i++) {
            cplx l = a[i * n + k] * rp;
            a[i * n + k] = l;
            for (int j = k + 1; j < n; j++) a[i * n + j] -= l * a[k * n + j];
        }
    }
    /* inverse: solve A X = I column by column with the permuted identity */
    for (int c = 0; c < n; c++) {
        cplx *x = inv + (long)c * n; /* column c stored contiguously, transposed below */
        for (int i = 0; i < n; i++) x[i] = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < n; k++) { int p = piv[k]; if (p != k) { cplx t = x[k]; x[k] = x[p]; x[p] = t; } }
        for (int i = 0; i < n; i++) { cplx s = x[i]; for (int j = 0; j < i; j++) s -= a[i * n + j] * x[j]; x[i] = s; }
        for (int i = n - 1; i >= 0; i--) { cplx s = x[i]; for (int j = i + 1; j < n; j++) s -= a[i * n + j] * x[j]; x[i] = s / a[i * n + i]; }
    }
    return 0; /* inv holds columns: inv[c*n + r] = (A^-1)[r][c] */
}

/* ------------------------------------------------------------------ per-trial context */
typedef struct {
    int nmax, ngauss, np;
    double *x, *w, *s, *ss;       /* [2*ngauss] nodes, weights, 1/sin, 1/sin^2 */
    double *r, *dr, *ddr, *drr, *dri; /* [2*ngauss] */
    double *j, *y, *dj, *dy, *jr, *ji, *djr, *dji; /* [ngauss][nmax+1] (first half of the nodes) */
    double ppi, pir, pii;
    double *an, *ann;             /* [nmax+1], [(nmax+1)^2] */
} trial_ctx;

static long tm_offset(int nmax, int m)
{
    long off = 0;
    for (int mm = 0; mm < m; mm++) { int nm = nmax - (mm > 1 ? mm : 1) + 1; off += 4L * nm * nm; }
    return off;
}
long hso_tm_offset(int nmax, int m) { return tm_offset(nmax, m); }

static void trial_free(trial_ctx *c)
{
    free(c->x); free(c->r); free(c->j); free(c->an);
}

/* CONST + VARY + BESS for one (nmax, ngauss) (A.2, A.3) */
static void trial_setup(trial_ctx *c, int nmax, int ngauss, double rev, double lam, double mrr,
                        double mri, double eps, int np)
{
    int ng = 2 * ngauss;
    c->nmax = nmax; c->ngauss = ngauss; c->np = np;
    c->x = (double *)malloc(sizeof(double) * 4 * ng);
    c->w = c->x + ng; c->s = c->w + ng; c->ss = c->s + ng;
    c->r = (double *)malloc(sizeof(double) * 5 * ng);
    c->dr = c->r + ng; c->ddr = c->dr + ng; c->drr = c->ddr + ng; c->dri = c->drr + ng;
    long tab = (long)ngauss * (nmax + 1);
    c->j = (double *)malloc(sizeof(double) * 8 * tab);
    c->y = c->j + tab; c->dj = c->y + tab; c->dy = c->dj + tab;
    c->jr = c->dy + tab; c->ji = c->jr + tab; c->djr = c->ji + tab; c->dji = c->djr + tab;
    c->an = (double *)malloc(sizeof(double) * ((nmax + 1) + (long)(nmax + 1) * (nmax + 1)));
    c->ann = c->an + (nmax + 1);

    double *dd = (double *)malloc(sizeof(double) * (nmax + 1));
    for (int n = 1; n <= nmax; n++) {
        int nn = n * (n + 1);
        c->an[n] = (double)nn;
        double d = sqrt((double)(2 * n + 1) / (double)nn);
        dd[n] = d;
        for (int n1 = 1; n1 <= n; n1++) {
            double ddd = d * dd[n1] * 0.5;
            c->ann[n * (nmax + 1) + n1] = ddd;
            c->ann[n1 * (nmax + 1) + n] = ddd;
        }
    }
    free(dd);
    if (np == -2) {
        /* cylinder: two Gauss segments split at the edge angle (CONST, NP=-2 branch) */
        int ng1 = (int)((double)ngauss / 2.0), ng2 = ngauss - ng1;
        double xx = -cos(atan(eps));
        double *x1 = (double *)malloc(sizeof(double) * 4 * (ngauss + 1));
        double *w1 = x1 + ngauss + 1, *x2 = w1 + ngauss + 1, *w2 = x2 + ngauss + 1;
        hso_gauss(ng1, x1, w1);
        hso_gauss(ng2, x2, w2);
        for (int i = 0; i < ng1; i++) { c->w[i] = 0.5 * (xx + 1.0) * w1[i]; c->x[i] = 0.5 * (xx + 1.0) * x1[i] + 0.5 * (xx - 1.0); }
        for (int i = 0; i < ng2; i++) { c->w[i + ng1] = -0.5 * xx * w2[i]; c->x[i + ng1] = -0.5 * xx * x2[i] + 0.5 * xx; }
        for (int i = 0; i < ngauss; i++) { c->w[ng - 1 - i] = c->w[i]; c->x[ng - 1 - i] = -c->x[i]; }
        free(x1);
    } else {
        hso_gauss(ng, c->x, c->w);
    }
    for (int i = 0; i < ngauss; i++) {
        double yv = c->x[i];
        yv = 1.0 / (1.0 - yv * yv);
        c->ss[i] = yv; c->ss[ng - 1 - i] = yv;
        yv = sqrt(yv);
        c->s[i] = yv; c->s[ng - 1 - i] = yv;
    }
    /* surface r^2(theta) and (dr/dtheta)/r */
    if (np == -1) {
        double a = rev * pow(eps, 1.0 / 3.0);
        double aa = a * a, ee = eps * eps, ee1 = ee - 1.0;
        for (int i = 0; i < ngauss; i++) {
            double cth = c->x[i], cc = cth * cth, ssq = 1.0 - cc, sth = sqrt(ssq);
            double rr = 1.0 / (ssq + ee * cc);
            c->r[i] = aa * rr; c->r[ng - 1 - i] = aa * rr;
            c->dr[i] = rr * cth * sth * ee1; c->dr[ng - 1 - i] = -rr * cth * sth * ee1;
        }
    } else { /* np == -2: finite circular cylinder, eps = diameter/length (RSP3) */
        double h = rev * pow(2.0 / (3.0 * eps * eps), 1.0 / 3.0);
        double a = h * eps;
        for (int i = 0; i < ngauss; i++) {
            double co = -c->x[i], si = sqrt(1.0 - co * co), rad, rthet;
            if (si / co > a / h) { rad = a / si; rthet = -a * co / (si * si); }
            else { rad = h / co; rthet = h * si / (co * co); }
            c->r[i] = rad * rad; c->r[ng - 1 - i] = c->r[i];
            c->dr[i] = -rthet / rad; c->dr[ng - 1 - i] = -c->dr[i];
        }
    }
    double pi_ = acos(-1.0) * 2.0 / lam;
    c->ppi = pi_ * pi_; c->pir = c->ppi * mrr; c->pii = c->ppi * mri;
    double v = 1.0 / (mrr * mrr + mri * mri);
    double prr = mrr * v, pri = -mri * v, ta = 0.0;
    double *z = (double *)malloc(sizeof(double) * 3 * ng), *zr = z + ng, *zi = zr + ng;
    for (int i = 0; i < ng; i++) {
        double vv = sqrt(c->r[i]);
        v = vv * pi_;
        if (v > ta) ta = v;
        vv = 1.0 / v;
        c->ddr[i] = vv; c->drr[i] = prr * vv; c->dri[i] = pri * vv;
        z[i] = v; zr[i] = v * mrr; zi[i] = v * mri;
    }
    double tb = ta * sqrt(mrr * mrr + mri * mri);
    if (tb < (double)nmax) tb = (double)nmax;
    double tmx = ta > (double)nmax ? ta : (double)nmax;
    int nnmax1 = (int)(1.2 * sqrt(tmx) + 3.0);
    int nnmax2 = (int)(tb + 4.0 * pow(tb, 0.33333) + 1.2 * sqrt(tb));
    nnmax2 = nnmax2 - nmax + 5;
    int lw = nmax + (nnmax1 > nnmax2 ? nnmax1 : nnmax2) + 2;
    double *wk = (double *)malloc(sizeof(double) * 2 * lw);
    for (int i = 0; i < ngauss; i++) { /* only the x<0 half is ever used (mirror symmetry) */
        long o = (long)i * (nmax + 1);
        rjb(z[i], c->j + o, c->dj + o, nmax, nnmax1, wk);
        ryb(z[i], c->y + o, c->dy + o, nmax);
        cjb(zr[i], zi[i], c->jr + o, c->ji + o, c->djr + o, c->dji + o, nmax, nnmax2, wk, wk + lw);
    }
    free(wk); free(z);
}

/* TMATR0 / TMATR + TT for azimuthal mode m (A.5, A.6).  Writes the four T blocks. */
static int tmatr(const trial_ctx *c, int m, double *tout)
{
    int nmax = c->nmax, ngauss = c->ngauss, ng = 2 * ngauss, n1s = nmax + 1;
    int mm1 = m > 1 ? m : 1, nm = nmax - mm1 + 1, nnm = 2 * nm;
    double qm = (double)m, qmm = qm * qm;
    double *d1 = (double *)calloc((size_t)2 * ngauss * n1s, sizeof(double)), *d2 = d1 + (long)ngauss * n1s;
    double *dv1 = (double *)malloc(sizeof(double) * 2 * n1s), *dv2 = dv1 + n1s;
    double *rr = (double *)malloc(sizeof(double) * 3 * ngauss), *ds = rr + ngauss, *dss = ds + ngauss;
    /* Wigner functions at the mirrored (x>0) node, reflected to the x<0 node i2 = ngauss-1-i */
    for (int i = 0; i < ngauss; i++) {
        int i1 = ngauss + i, i2 = ngauss - 1 - i;
        vig(c->x[i1], nmax, m, dv1, dv2);
        for (int n = mm1; n <= nmax; n++) {
            double si = ((n + m) % 2) ? -1.0 : 1.0;
            d1[(long)i2 * n1s + n] = dv1[n] * si;
            d2[(long)i2 * n1s + n] = -dv2[n] * si;
        }
    }
    for (int i = 0; i < ngauss; i++) {
        double wr = c->w[i] * c->r[i];
        ds[i] = c->s[i] * qm * wr;
        dss[i] = c->ss[i] * qmm;
        rr[i] = wr;
    }
    (void)ng;
    cplx *q = (cplx *)calloc((size_t)3 * nnm * nnm, sizeof(cplx)), *rgq = q + (long)nnm * nnm, *qinv = rgq + (long)nnm * nnm;
    for (int n1 = mm1; n1 <= nmax; n1++) {
        double an1 = c->an[n1];
        for (int n2 = mm1; n2 <= nmax; n2++) {
            double an2 = c->an[n2];
            double ar11 = 0, ai11 = 0, ar12 = 0, ai12 = 0, ar21 = 0, ai21 = 0, ar22 = 0, ai22 = 0;
            double gr11 = 0, gi11 = 0, gr12 = 0, gi12 = 0, gr21 = 0, gi21 = 0, gr22 = 0, gi22 = 0;
            int even = ((n1 + n2) % 2) == 0;
            if (!(m == 0 && !even)) {
                for (int i = 0; i < ngauss; i++) {
                    long o = (long)i * n1s;
                    double d1n1 = d1[o + n1], d2n1 = d2[o + n1], d1n2 = d1[o + n2], d2n2 = d2[o + n2];
                    double a11 = d1n1 * d1n2, a12 = d1n1 * d2n2, a21 = d2n1 * d1n2, a22 = d2n1 * d2n2;
                    double aa1 = a12 + a21, aa2 = a11 * dss[i] + a22;
                    double qj1 = c->j[o + n1], qy1 = c->y[o + n1];
                    double qjr2 = c->jr[o + n2], qji2 = c->ji[o + n2];
                    double qdjr2 = c->djr[o + n2], qdji2 = c->dji[o + n2];
                    double qdj1 = c->dj[o + n1], qdy1 = c->dy[o + n1];
                    double c1r = qjr2 * qj1, c1i = qji2 * qj1;
                    double b1r = c1r - qji2 * qy1, b1i = c1i + qjr2 * qy1;
                    double c2r = qjr2 * qdj1, c2i = qji2 * qdj1;
                    double b2r = c2r - qji2 * qdy1, b2i = c2i + qjr2 * qdy1;
                    double ddri = c->ddr[i];
                    double c3r = ddri * c1r, c3i = ddri * c1i, b3r = ddri * b1r, b3i = ddri * b1i;
                    double c4r = qdjr2 * qj1, c4i = qdji2 * qj1;
                    double b4r = c4r - qdji2 * qy1, b4i = c4i + qdjr2 * qy1;
                    double drri = c->drr[i], drii = c->dri[i];
                    double c5r = c1r * drri - c1i * drii, c5i = c1i * drri + c1r * drii;
                    double b5r = b1r * drri - b1i * drii, b5i = b1i * drri + b1r * drii;
                    double uri = c->dr[i], dsi = ds[i], rri = rr[i];
                    if (even) {
                        double f1 = rri * aa2, f2 = rri * uri * an1 * a12;
                        ar12 += f1 * b2r + f2 * b3r; ai12 += f1 * b2i + f2 * b3i;
                        gr12 += f1 * c2r + f2 * c3r; gi12 += f1 * c2i + f2 * c3i;
                        f2 = rri * uri * an2 * a21;
                        ar21 += f1 * b4r + f2 * b5r; ai21 += f1 * b4i + f2 * b5i;
                        gr21 += f1 * c4r + f2 * c5r; gi21 += f1 * c4i + f2 * c5i;
                    } else {
                        double c6r = qdjr2 * qdj1, c6i = qdji2 * qdj1;
                        double b6r = c6r - qdji2 * qdy1, b6i = c6i + qdjr2 * qdy1;
                        double c7r = c4r * ddri, c7i = c4i * ddri, b7r = b4r * ddri, b7i = b4i * ddri;
                        double c8r = c2r * drri - c2i * drii, c8i = c2i * drri + c2r * drii;
                        double b8r = b2r * drri - b2i * drii, b8i = b2i * drri + b2r * drii;
                        double e1 = dsi * aa1;
                        ar11 += e1 * b1r; ai11 += e1 * b1i; gr11 += e1 * c1r; gi11 += e1 * c1i;
                        double e2 = dsi * uri * a11, e3 = e2 * an2;
                        e2 = e2 * an1;
                        ar22 += e1 * b6r + e2 * b7r + e3 * b8r; ai22 += e1 * b6i + e2 * b7i + e3 * b8i;
                        gr22 += e1 * c6r + e2 * c7r + e3 * c8r; gi22 += e1 * c6i + e2 * c7i + e3 * c8i;
                    }
                }
            }
            double an12 = c->ann[n1 * n1s + n2] * 2.0; /* FACTOR=2: half range */
            double r11 = ar11 * an12, i11 = ai11 * an12, rg11 = gr11 * an12, ig11 = gi11 * an12;
            double r12 = ar12 * an12, i12 = ai12 * an12, rg12 = gr12 * an12, ig12 = gi12 * an12;
            double r21 = ar21 * an12, i21 = ai21 * an12, rg21 = gr21 * an12, ig21 = gi21 * an12;
            double r22 = ar22 * an12, i22 = ai22 * an12, rg22 = gr22 * an12, ig22 = gi22 * an12;
            int k1 = n1 - mm1, kk1 = k1 + nm, k2 = n2 - mm1, kk2 = k2 + nm;
            double tar11 = -r11, tai11 = -i11, tgr11 = -rg11, tgi11 = -ig11;
            double tar12 = i12, tai12 = -r12, tgr12 = ig12, tgi12 = -rg12;
            double tar21 = -i21, tai21 = r21, tgr21 = -ig21, tgi21 = rg21;
            double tar22 = -r22, tai22 = -i22, tgr22 = -rg22, tgi22 = -ig22;
            double tpir = c->pir, tpii = c->pii, tppi = c->ppi;
            q[k1 * nnm + k2] = (tpir * tar21 - tpii * tai21 + tppi * tar12) + I * (tpir * tai21 + tpii * tar21 + tppi * tai12);
            rgq[k1 * nnm + k2] = (tpir * tgr21 - tpii * tgi21 + tppi * tgr12) + I * (tpir * tgi21 + tpii * tgr21 + tppi * tgi12);
            q[k1 * nnm + kk2] = (tpir * tar11 - tpii * tai11 + tppi * tar22) + I * (tpir * tai11 + tpii * tar11 + tppi * tai22);
            rgq[k1 * nnm + kk2] = (tpir * tgr11 - tpii * tgi11 + tppi * tgr22) + I * (tpir * tgi11 + tpii * tgr11 + tppi * tgi22);
            q[kk1 * nnm + k2] = (tpir * tar22 - tpii * tai22 + tppi * tar11) + I * (tpir * tai22 + tpii * tar22 + tppi * tai11);
            rgq[kk1 * nnm + k2] = (tpir * tgr22 - tpii * tgi22 + tppi * tgr11) + I * (tpir * tgi22 + tpii * tgr22 + tppi * tgi11);
            q[kk1 * nnm + kk2] = (tpir * tar12 - tpii * tai12 + tppi * tar21) + I * (tpir * tai12 + tpii * tar12 + tppi * tai21);
            rgq[kk1 * nnm + kk2] = (tpir * tgr12 - tpii * tgi12 + tppi * tgr21) + I * (tpir * tgi12 + tpii * tgr12 + tppi * tgi21);
        }
    }
    int *piv = (int *)malloc(sizeof(int) * nnm);
    int sing = lu_invert(q, nnm, qinv, piv);
    /* T = -RgQ * Q^-1 ; qinv[c*n + r] = Qinv[r][c] */
    for (int i = 0; i < nnm; i++)
        for (int jc = 0; jc < nnm; jc++) {
            cplx t = 0;
            if (!sing) {
                const cplx *col = qinv + (long)jc * nnm;
                for (int k = 0; k < nnm; k++) t -= rgq[i * nnm + k] * col[k];
            } else t = NAN;
            int blk = (i >= nm ? 2 : 0) + (jc >= nm ? 1 : 0);
            int a = i % nm, b = jc % nm;
            double *dst = tout + 2 * ((long)blk * nm * nm + (long)a * nm + b);
            dst[0] = creal(t); dst[1] = cimag(t);
        }
    free(piv); free(q); free(rr); free(dv1); free(d1);
    return sing;
}

/* ------------------------------------------------------------------ calctmat (A.1) */
/* rat: 1 = axi is the equal-volume radius; other values = equal-surface-area radius (SAREA).
 * np: -1 spheroid, -2 cylinder.  rescale: apply Mishchenko's DDELT=0.1*DDELT.
 * Returns a heap object (free with hso_free); never touches global state. */
hso_tmat *hso_calctmat(double axi, double rat, double lam, double mrr, double mri, double eps, int np,
                       double ddelt, int ndgs, int rescale)
{
    hso_tmat *tm = (hso_tmat *)calloc(1, sizeof(hso_tmat));
    double p = acos(-1.0);
    if (fabs(rat - 1.0) > 1e-8 && np == -1) rat = sarea(eps);
    if (fabs(rat - 1.0) > 1e-8 && np == -2) {
        /* SAREAC */
        rat = pow(1.5 / eps, 1.0 / 3.0);
        rat = rat / sqrt((eps + 2.0) / (2.0 * eps));
    }
    if (rescale) ddelt = 0.1 * ddelt;
    double a = rat * axi;
    double xev = 2.0 * p * a / lam;
    int ixxx = (int)(xev + 4.05 * pow(xev, 0.333333));
    int inm1 = ixxx > 4 ? ixxx : 4;
    if (inm1 >= HSO_NPN1) { tm->status = 1; return tm; }
    double qext1 = 0.0, qsca1 = 0.0;
    int nmax = inm1, ngauss = 0, converged = 0;
    long tcap = tm_offset(HSO_NPN1, 1);
    double *t0 = (double *)malloc(sizeof(double) * 2 * 4 * (long)HSO_NPN1 * HSO_NPN1);
    (void)tcap;
    trial_ctx c;
    memset(&c, 0, sizeof(c));
    for (int nma = inm1; nma <= HSO_NPN1; nma++) {
        nmax = nma;
        ngauss = nmax * ndgs;
        if (ngauss > HSO_NPNG1) { tm->status = 2; free(t0); return tm; }
        if (c.x) trial_free(&c);
        trial_setup(&c, nmax, ngauss, a, lam, mrr, mri, eps, np);
        if (tmatr(&c, 0, t0)) { tm->status = 3; }
        double qext = 0.0, qsca = 0.0;
        long nn = (long)nmax * nmax;
        for (int n = 1; n <= nmax; n++) {
            long d = 2 * ((long)(n - 1) * nmax + (n - 1));
            double tr1 = t0[d], ti1 = t0[d + 1], tr2 = t0[2 * 3 * nn + d], ti2 = t0[2 * 3 * nn + d + 1];
            double dn1 = (double)(2 * n + 1);
            qsca += dn1 * (tr1 * tr1 + ti1 * ti1 + tr2 * tr2 + ti2 * ti2);
            qext += (tr1 + tr2) * dn1;
        }
        double dsca = fabs((qsca1 - qsca) / qsca), dext = fabs((qext1 - qext) / qext);
        qext1 = qext; qsca1 = qsca;
        if (tm->ntrials < HSO_MAXTRIALS) {
            tm->trial_nmax[tm->ntrials] = nmax; tm->trial_ngauss[tm->ntrials] = ngauss;
            tm->trial_qsca[tm->ntrials] = qsca; tm->trial_qext[tm->ntrials] = qext; tm->ntrials++;
        }
        if (dsca <= ddelt && dext <= ddelt) { converged = 1; break; }
        if (nma == HSO_NPN1) { tm->status = 1; trial_free(&c); free(t0); return tm; }
    }
    (void)converged;
    if (ngauss != HSO_NPNG1) {
        for (int ngaus = ngauss + 1; ngaus <= HSO_NPNG1; ngaus++) {
            ngauss = ngaus;
            trial_free(&c);
            trial_setup(&c, nmax, ngauss, a, lam, mrr, mri, eps, np);
            if (tmatr(&c, 0, t0)) { tm->status = 3; }
            double qext = 0.0, qsca = 0.0;
            long nn = (long)nmax * nmax;
            for (int n = 1; n <= nmax; n++) {
                long d = 2 * ((long)(n - 1) * nmax + (n - 1));
                double tr1 = t0[d], ti1 = t0[d + 1], tr2 = t0[2 * 3 * nn + d], ti2 = t0[2 * 3 * nn + d + 1];
                double dn1 = (double)(2 * n + 1);
                qsca += dn1 * (tr1 * tr1 + ti1 * ti1 + tr2 * tr2 + ti2 * ti2);
                qext += (tr1 + tr2) * dn1;
            }
            double dsca = fabs((qsca1 - qsca) / qsca), dext = fabs((qext1 - qext) / qext);
            qext1 = qext; qsca1 = qsca;
            if (tm->ntrials < HSO_MAXTRIALS) {
                tm->trial_nmax[tm->ntrials] = nmax; tm->trial_ngauss[tm->ntrials] = ngauss;
                tm->trial_qsca[tm->ntrials] = qsca; tm->trial_qext[tm->ntrials] = qext; tm->ntrials++;
            }
            if (dsca <= ddelt && dext <= ddelt) break;
            if (ngaus == HSO_NPNG1) tm->status = 2; /* upstream only prints a warning and goes on */
        }
    }
    tm->nmax = nmax; tm->ngauss = ngauss;
    tm->tlen = tm_offset(nmax, nmax + 1);
    tm->t = (double *)malloc(sizeof(double) * 2 * tm->tlen);
    memcpy(tm->t, t0, sizeof(double) * 2 * 4 * (long)nmax * nmax);
    free(t0);
    for (int m = 1; m <= nmax; m++)
        if (tmatr(&c, m, tm->t + 2 * tm_offset(nmax, m))) tm->status = 3;
    trial_free(&c);
    return tm;
}

void hso_free(hso_tmat *tm) { if (tm) { free(tm->t); free(tm); } }
int hso_nmax(const hso_tmat *tm) { return tm->nmax; }
int hso_ngauss(const hso_tmat *tm) { return tm->ngauss; }
int hso_ntrials(const hso_tmat *tm) { return tm->ntrials; }
int hso_status(const hso_tmat *tm) { return tm->status; }
long hso_tlen(const hso_tmat *tm) { return tm->tlen; }
const double *hso_tdata(const hso_tmat *tm) { return tm->t; }
void hso_trials(const hso_tmat *tm, int *nmax, int *ngauss, double *qsca, double *qext)
{
    for (int i = 0; i < tm->ntrials; i++) {
        nmax[i] = tm->trial_nmax[i]; ngauss[i] = tm->trial_ngauss[i];
        qsca[i] = tm->trial_qsca[i]; qext[i] = tm->trial_qext[i];
    }
}

/* ------------------------------------------------------------------ calcampl (A.7, A.8) */
/* T given in the packed layout; angles in degrees. S = [S11,S12,S21,S22] interleaved (re,im); Z row-major 4x4. */
void hso_ampl_from_t(const double *t, int nmax, double lam, double thet0, double thet, double phi0,
                     double phi, double alpha, double beta, double *S, double *Z)
{
    double pin = acos(-1.0), pin2 = pin * 0.5, pi_ = pin / 180.0;
    double alph = alpha * pi_, bet = beta * pi_;
    double thetl = thet0 * pi_, phil = phi0 * pi_, thetl1 = thet * pi_, phil1 = phi * pi_;
    double eps = 1e-7;
    if (thetl < pin2) thetl += eps;
    if (thetl > pin2) thetl -= eps;
    if (thetl1 < pin2) thetl1 += eps;
    if (thetl1 > pin2) thetl1 -= eps;
    if (phil < pin) phil += eps;
    if (phil > pin) phil -= eps;
    if (phil1 < pin) phil1 += eps;
    if (phil1 > pin) phil1 -= eps;
    if (bet <= pin2 && pin2 - bet <= eps) bet -= eps;
    if (bet > pin2 && bet - pin2 <= eps) bet += eps;

    double cb = cos(bet), sb = sin(bet), ct = cos(thetl), st = sin(thetl);
    double cp = cos(phil - alph), sp = sin(phil - alph);
    double ctp = ct * cb + st * sb * cp, thetp = acos(ctp);
    double cpp = cb * st * cp - sb * ct, spp = st * sp;
    double phip = atan(spp / cpp);
    if (phip > 0.0 && sp < 0.0) phip += pin;
    if (phip < 0.0 && sp > 0.0) phip += pin;
    if (phip < 0.0) phip += 2.0 * pin;

    double ct1 = cos(thetl1), st1 = sin(thetl1);
    double cp1 = cos(phil1 - alph), sp1 = sin(phil1 - alph);
    double ctp1 = ct1 * cb + st1 * sb * cp1, thetp1 = acos(ctp1);
    double cpp1 = cb * st1 * cp1 - sb * ct1, spp1 = st1 * sp1;
    double phip1 = atan(spp1 / cpp1);
    if (phip1 > 0.0 && sp1 < 0.0) phip1 += pin;
    if (phip1 < 0.0 && sp1 > 0.0) phip1 += pin;
    if (phip1 < 0.0) phip1 += 2.0 * pin;

    double ca = cos(alph), sa = sin(alph);
    double B[3][3] = {{ca * cb, sa * cb, -sb}, {-sa, ca, 0.0}, {ca * sb, sa * sb, cb}};
    cp = cos(phil); sp = sin(phil); cp1 = cos(phil1); sp1 = sin(phil1);
    double AL[3][2] = {{ct * cp, -sp}, {ct * sp, cp}, {-st, 0.0}};
    double AL1[3][2] = {{ct1 * cp1, -sp1}, {ct1 * sp1, cp1}, {-st1, 0.0}};
    ct = ctp; st = sin(thetp); cp = cos(phip); sp = sin(phip);
    ct1 = ctp1; st1 = sin(thetp1); cp1 = cos(phip1); sp1 = sin(phip1);
    double AP[2][3] = {{ct * cp, ct * sp, -st}, {-sp, cp, 0.0}};
    double AP1[2][3] = {{ct1 * cp1, ct1 * sp1, -st1}, {-sp1, cp1, 0.0}};
    double C[3][2], R[2][2], R1[2][2];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += B[i][k] * AL[k][j]; C[i][j] = x; }
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += AP[i][k] * C[k][j]; R[i][j] = x; }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += B[i][k] * AL1[k][j]; C[i][j] = x; }
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += AP1[i][k] * C[k][j]; R1[i][j] = x; }
    double d = 1.0 / (R1[0][0] * R1[1][1] - R1[0][1] * R1[1][0]);
    double x00 = R1[0][0];
    R1[0][0] = R1[1][1] * d; R1[0][1] = -R1[0][1] * d; R1[1][0] = -R1[1][0] * d; R1[1][1] = x00 * d;

    static const double cipow_re[4] = {1, 0, -1, 0}, cipow_im[4] = {0, 1, 0, -1};
    double dcth0 = ctp, dcth = ctp1, ph = phip1 - phip;
    cplx vv = 0, vh = 0, hv = 0, hh = 0;
    int n1s = nmax + 1;
    double *dv1 = (double *)malloc(sizeof(double) * 4 * n1s), *dv2 = dv1 + n1s, *dv01 = dv2 + n1s, *dv02 = dv01 + n1s;
    for (int m = 0; m <= nmax; m++) {
        int nmin = m > 1 ? m : 1, nm = nmax - nmin + 1;
        const double *tb = t + 2 * tm_offset(nmax, m);
        long bs = 2L * nm * nm;
        vigampl(dcth, nmax, m, dv1, dv2);
        vigampl(dcth0, nmax, m, dv01, dv02);
        double fc = 2.0 * cos(m * ph), fs = 2.0 * sin(m * ph);
        for (int nn = nmin; nn <= nmax; nn++) {
            double dv1nn = m * dv01[nn], dv2nn = dv02[nn];
            for (int n = nmin; n <= nmax; n++) {
                double dv1n = m * dv1[n], dv2n = dv2[n];
                long e = 2 * ((long)(n - nmin) * nm + (nn - nmin));
                cplx ct11 = tb[e] + I * tb[e + 1];
                cplx ct22 = tb[3 * bs + e] + I * tb[3 * bs + e + 1];
                int pw = ((nn - n - 1) % 4 + 4) % 4;
                double rn = sqrt((double)((2 * n + 1) * (2 * nn + 1)) / (double)(n * nn * (n + 1) * (nn + 1)));
                cplx cal = (cipow_re[pw] + I * cipow_im[pw]) * rn;
                if (m == 0) {
                    cplx cn = cal * dv2n * dv2nn;
                    vv += cn * ct22;
                    hh += cn * ct11;
                } else {
                    cplx ct12 = tb[bs + e] + I * tb[bs + e + 1];
                    cplx ct21 = tb[2 * bs + e] + I * tb[2 * bs + e + 1];
                    cplx cn1 = cal * fc, cn2 = cal * fs;
                    double d11 = dv1n * dv1nn, d12 = dv1n * dv2nn, d21 = dv2n * dv1nn, d22 = dv2n * dv2nn;
                    vv += (ct11 * d11 + ct21 * d21 + ct12 * d12 + ct22 * d22) * cn1;
                    vh += (ct11 * d12 + ct21 * d22 + ct12 * d11 + ct22 * d21) * cn2;
                    hv -= (ct11 * d21 + ct21 * d11 + ct12 * d22 + ct22 * d12) * cn2;
                    hh += (ct11 * d22 + ct21 * d12 + ct12 * d21 + ct22 * d11) * cn1;
                }
            }
        }
    }
    free(dv1);
    double dk = 2.0 * pin / lam;
    vv /= dk; vh /= dk; hv /= dk; hh /= dk;
    cplx cvv = vv * R[0][0] + vh * R[1][0], cvh = vv * R[0][1] + vh * R[1][1];
    cplx chv = hv * R[0][0] + hh * R[1][0], chh = hv * R[0][1] + hh * R[1][1];
    cplx s11 = R1[0][0] * cvv + R1[0][1] * chv, s12 = R1[0][0] * cvh + R1[0][1] * chh;
    cplx s21 = R1[1][0] * cvv + R1[1][1] * chv, s22 = R1[1][0] * cvh + R1[1][1] * chh;
    S[0] = creal(s11); S[1] = cimag(s11); S[2] = creal(s12); S[3] = cimag(s12);
    S[4] = creal(s21); S[5] = cimag(s21); S[6] = creal(s22); S[7] = cimag(s22);
#define CJ(z) conj(z)
    Z[0] = 0.5 * creal(s11 * CJ(s11) + s12 * CJ(s12) + s21 * CJ(s21) + s22 * CJ(s22));
    Z[1] = 0.5 * creal(s11 * CJ(s11) - s12 * CJ(s12) + s21 * CJ(s21) - s22 * CJ(s22));
    Z[2] = -creal(s11 * CJ(s12) + s22 * CJ(s21));
    Z[3] = -cimag(s11 * CJ(s12) - s22 * CJ(s21));
    Z[4] = 0.5 * creal(s11 * CJ(s11) + s12 * CJ(s12) - s21 * CJ(s21) - s22 * CJ(s22));
    Z[5] = 0.5 * creal(s11 * CJ(s11) - s12 * CJ(s12) - s21 * CJ(s21) + s22 * CJ(s22));
    Z[6] = -creal(s11 * CJ(s12) - s22 * CJ(s21));
    Z[7] = -cimag(s11 * CJ(s12) + s22 * CJ(s21));
    Z[8] = -creal(s11 * CJ(s21) + s22 * CJ(s12));
    Z[9] = -creal(s11 * CJ(s21) - s22 * CJ(s12));
    Z[10] = creal(s11 * CJ(s22) + s12 * CJ(s21));
    Z[11] = cimag(s11 * CJ(s22) + s21 * CJ(s12));
    Z[12] = -cimag(s21 * CJ(s11) + s22 * CJ(s12));
    Z[13] = -cimag(s21 * CJ(s11) - s22 * CJ(s12));
    Z[14] = cimag(s22 * CJ(s11) - s12 * CJ(s21));
    Z[15] = creal(s22 * CJ(s11) - s12 * CJ(s21));
#undef CJ
}

void hso_calcampl(const hso_tmat *tm, double lam, double thet0, double thet, double phi0, double phi,
                  double alpha, double beta, double *S, double *Z)
{
    hso_ampl_from_t(tm->t, tm->nmax, lam, thet0, thet, phi0, phi, alpha, beta, S, Z);
}

/* Batched convenience used by the CPU baseline: n particles, one worker; outputs nmax/ngauss/ntrials/status
 * and, when S != NULL, orientation-averaged S/Z for ngeom geometries with norient (alpha,beta,weight) triples. */
void hso_batch(int n, const double *rev, const double *lam, const double *mrr, const double *mri,
               const double *eps, int np, double ddelt, int ndgs, int rescale, int ngeom,
               const double *geom /*[ngeom][4] thet0,thet,phi0,phi*/, int norient, const double *alpha,
               const double *beta, const double *wgt, int *nmax, int *ngauss, int *ntrials, int *status,
               double *S /*[n][ngeom][8]*/, double *Z /*[n][ngeom][16]*/)
{
    for (int p = 0; p < n; p++) {
        hso_tmat *tm = hso_calctmat(rev[p], 1.0, lam[p], mrr[p], mri[p], eps[p], np, ddelt, ndgs, rescale);
        nmax[p] = tm->nmax; ngauss[p] = tm->ngauss; ntrials[p] = tm->ntrials; status[p] = tm->status;
        if (S && tm->t) {
            for (int g = 0; g < ngeom; g++) {
                double sa[8] = {0}, za[16] = {0}, s1[8], z1[16];
                for (int o = 0; o < norient; o++) {
                    hso_calcampl(tm, lam[p], geom[4 * g], geom[4 * g + 1], geom[4 * g + 2], geom[4 * g + 3],
                                 alpha[o], beta[o], s1, z1);
                    for (int k = 0; k < 8; k++) sa[k] += wgt[o] * s1[k];
                    for (int k = 0; k < 16; k++) za[k] += wgt[o] * z1[k];
                }
                memcpy(S + ((long)p * ngeom + g) * 8, sa, sizeof(sa));
                memcpy(Z + ((long)p * ngeom + g) * 16, za, sizeof(za));
            }
        }
        hso_free(tm);
    }
}

// hs_tmat.cuh -- single-layer EBCM T-matrix kernel (stages K1 special functions, K2 Q/RgQ assembly,
// K3 parity-decoupled complex solve, convergence controller) for one particle per thread group.
//
// Replaces pytmatrix's Fortran `calctmat` (Mishchenko ampld.lp.f: CONST, VARY, BESS, VIG, TMATR0, TMATR,
// TT), SURVEY.md rows P2-P7 / Appendix A.1-A.6.  Written B200-first: a particle's whole working set
// (radial-function tables, Wigner tables, the two parity-class linear systems) lives in shared memory
// for the entire convergence loop and the azimuthal-mode sweep; HBM sees 5 doubles in and the packed
// T-matrix out.
#pragma once
#include "hs_compat.cuh"
#include "../../include/hydrotm.h"

namespace hs {

struct TmArgs {
    const int *items;
    int n_items;
    int *queue;  // dynamic work queue counter
    const double *axi, *lam, *mr, *mi, *eps;
    double rat;
    int np;
    double ddelt;
    int ndgs;
    const double *gx, *gw;  // Gauss table: for rule 2g the g positive nodes ascending at g(g-1)/2
    int gmax;
    double *tarena;
    long long arena_cap;
    unsigned long long *arena_top;
    long long *toff;
    int *nmax, *ngauss, *ntrials, *status;
    int ws_cap;         // doubles available per group
    double *ws_global;  // GLOBAL_WS variant: per-CTA slices of ws_cap doubles
    double *ws_fb;      // fall-back workspace in global memory (L2-resident) for particles that outgrow the
    long long *prof;    // optional phase-cycle counters [8] (thread 0 of each group, clock64), null in production
    int tile_min;       // smallest block order NM assembled with the register-tiled contraction
    int debug;          // timing experiments only: bit0 skip solve, bit1 skip assembly, bit2 skip radial functions
    int fb_cap;         // shared-memory slice of their bucket: per-group slices of fb_cap doubles (or null / 0)
};

__host__ __device__ inline long long ws_need(int nmax, int g)
{
    int ld = nmax | 1;
    return ((13LL * g + 2LL * (nmax + 2) + 10LL * g * ld + 1) & ~1LL) + 8LL * nmax * nmax;  // sys is 16-byte aligned
}

__host__ __device__ inline long long tmat_size(int nmax)
{
    long long n = nmax;
    return 2 * (n * n + n * (n + 1) * (2 * n + 1) / 6);
}

struct Carve {
    double *x, *w, *s, *ss, *r, *dr, *ddr, *drr, *dri, *rr, *ds, *dss, *qsp, *sq, *rsq;
    double *bj, *by, *bdj, *bdy, *bjr, *bji, *bdjr, *bdji;
    double *d1, *d2;
    double2 *sys;
    int ld;
};

__device__ inline void carve(Carve &c, double *ws, int nmax, int g)
{
    int ld = nmax | 1;
    c.ld = ld;
    double *p = ws;
    c.x = p; p += g; c.w = p; p += g; c.s = p; p += g; c.ss = p; p += g;
    c.r = p; p += g; c.dr = p; p += g; c.ddr = p; p += g; c.drr = p; p += g;
    c.dri = p; p += g; c.rr = p; p += g; c.ds = p; p += g; c.dss = p; p += g; c.qsp = p; p += g;
    c.sq = p; p += nmax + 2; c.rsq = p; p += nmax + 2;
    long tab = (long)g * ld;
    c.bj = p; p += tab; c.by = p; p += tab; c.bdj = p; p += tab; c.bdy = p; p += tab;
    c.bjr = p; p += tab; c.bji = p; p += tab; c.bdjr = p; p += tab; c.bdji = p; p += tab;
    c.d1 = p; p += tab; c.d2 = p; p += tab;
    if ((p - ws) & 1) p++;  // double2 alignment
    c.sys = reinterpret_cast<double2 *>(p);
}

// per-group control block (shared memory)
struct Ctl {
    int item, nmax, g, phase, done, status, ntr, sing;
    double qsca1, qext1;
    unsigned long long toff;
    // storage for the solver's system descriptors / pivot records (SysDesc[8], SolveCtl)
    alignas(16) unsigned char sd[8 * 32];
    alignas(16) unsigned char sc[8 * 4 + 8 * 16 + 16];
};

template <bool WARP>
struct Grp {
    int tid, nt;
    __device__ __forceinline__ void sync() const
    {
        if (WARP) __syncwarp();
        else __syncthreads();
    }
};

// ---- K1: radial functions, one quadrature node per task (P4: RJB / RYB / CJB) ----
// j_n(x): downward ratio recurrence from order nmax+nnmax; ratios for n<=nmax are parked in y[] then
// turned into function values upward.  y,u hold n=1..nmax at index n-1.
__device__ inline void rjb(double x, double *y, double *u, int nmax, int nnmax)
{
    int l = nmax + nnmax;
    double xx = 1.0 / x;
    double zc = 1.0 / ((double)(2 * l + 1) * xx);
    for (int i1 = l - 1; i1 >= 1; i1--) {
        zc = 1.0 / ((double)(2 * i1 + 1) * xx - zc);
        if (i1 <= nmax) y[i1 - 1] = zc;
    }
    double z1 = y[0];
    double z0 = 1.0 / (xx - z1);
    double y0 = z0 * cos(x) * xx;
    double y1 = y0 * z1;
    u[0] = y0 - y1 * xx;
    y[0] = y1;
    for (int i = 2; i <= nmax; i++) {
        double yi1 = y[i - 2];
        double yi = yi1 * y[i - 1];
        u[i - 1] = yi1 - (double)i * yi * xx;
        y[i - 1] = yi;
    }
}

__device__ inline void ryb(double x, double *y, double *v, int nmax)
{
    double s, c;
    sincos(x, &s, &c);
    double x1 = 1.0 / x, x2 = x1 * x1, x3 = x2 * x1;
    double y1 = -c * x2 - s * x1;
    y[0] = y1;
    if (nmax >= 2) y[1] = (-3.0 * x3 + x1) * c - 3.0 * x2 * s;
    for (int i = 2; i <= nmax - 1; i++) y[i] = (double)(2 * i + 1) * x1 * y[i - 1] - y[i - 2];
    v[0] = -x1 * (c + y1);
    for (int i = 2; i <= nmax; i++) v[i - 1] = y[i - 2] - (double)i * x1 * y[i - 1];
}

__device__ inline void cjb(double xr, double xi, double *yr, double *yi, double *ur, double *ui,
                           int nmax, int nnmax)
{
    int l = nmax + nnmax;
    double xrxi = 1.0 / (xr * xr + xi * xi);
    double cxxr = xr * xrxi, cxxi = -xi * xrxi;
    double qf = 1.0 / (double)(2 * l + 1);
    double czr = xr * qf, czi = xi * qf;
    for (int i1 = l - 1; i1 >= 1; i1--) {
        qf = (double)(2 * i1 + 1);
        double ar = qf * cxxr - czr;
        double ai = qf * cxxi - czi;
        double ari = 1.0 / (ar * ar + ai * ai);
        czr = ar * ari;
        czi = -ai * ari;
        if (i1 <= nmax) { yr[i1 - 1] = czr; yi[i1 - 1] = czi; }
    }
    double z1r = yr[0], z1i = yi[0];
    double ar = cxxr - z1r, ai = cxxi - z1i;
    double ari = 1.0 / (ar * ar + ai * ai);
    double cz0r = ar * ari, cz0i = -ai * ari;
    double sxr, cxr;
    sincos(xr, &sxr, &cxr);
    double cr = cxr * cosh(xi), ci = -sxr * sinh(xi);
    ar = cz0r * cr - cz0i * ci;
    ai = cz0i * cr + cz0r * ci;
    double cy0r = ar * cxxr - ai * cxxi, cy0i = ai * cxxr + ar * cxxi;
    double cy1r = cy0r * z1r - cy0i * z1i, cy1i = cy0i * z1r + cy0r * z1i;
    ar = cy1r * cxxr - cy1i * cxxi;
    ai = cy1i * cxxr + cy1r * cxxi;
    yr[0] = cy1r; yi[0] = cy1i;
    ur[0] = cy0r - ar; ui[0] = cy0i - ai;
    for (int i = 2; i <= nmax; i++) {
        double qi = (double)i;
        double p1r = yr[i - 2], p1i = yi[i - 2];
        double zr = yr[i - 1], zi = yi[i - 1];
        double cyir = p1r * zr - p1i * zi;
        double cyii = p1i * zr + p1r * zi;
        ar = cyir * cxxr - cyii * cxxi;
        ai = cyii * cxxr + cyir * cxxi;
        yr[i - 1] = cyir; yi[i - 1] = cyii;
        ur[i - 1] = p1r - qi * ar; ui[i - 1] = p1i - qi * ai;
    }
}

// ---- K1: Wigner d^n_{0m} and d/dtheta at the mirrored node, reflected to x<0 (P5: VIG + TMATR sign rule) ----
__device__ inline void vig_reflect(double x, int nmax, int m, double *d1, double *d2)
{
    double qs = sqrt(1.0 - x * x), qs1 = 1.0 / qs;
    if (m == 0) {
        double a1 = 1.0, a2 = x;
        for (int n = 1; n <= nmax; n++) {
            double qn = n, qn1 = n + 1, qn2 = 2 * n + 1;
            double a3 = (qn2 * x * a2 - qn * a1) / qn1;
            double der = qs1 * (qn1 * qn / qn2) * (-a1 + a3);
            double si = (n & 1) ? -1.0 : 1.0;
            d1[n - 1] = a2 * si;
            d2[n - 1] = -der * si;
            a1 = a2; a2 = a3;
        }
        return;
    }
    double qmm = (double)(m * m), a = 1.0;
    for (int i = 1; i <= m; i++) a = a * sqrt((double)(2 * i - 1) / (double)(2 * i)) * qs;
    double a1 = 0.0, a2 = a;
    for (int n = m; n <= nmax; n++) {
        double qn = n, qn2 = 2 * n + 1, qn1 = n + 1;
        double qnm = sqrt(qn * qn - qmm), qnm1 = sqrt(qn1 * qn1 - qmm);
        double a3 = (qn2 * x * a2 - qnm * a1) / qnm1;
        double der = qs1 * (-qn1 * qnm * a1 + qn * qnm1 * a3) / qn2;
        double si = ((n + m) & 1) ? -1.0 : 1.0;
        d1[n - 1] = a2 * si;
        d2[n - 1] = -der * si;
        a1 = a2; a2 = a3;
    }
}

// CONST + VARY + BESS (+ VIG for m=0) for one (nmax, g) trial.
template <bool WARP>
__device__ void trial_setup(const Grp<WARP> &grp, const Carve &c, int nmax, int g, double a_rev, double lam,
                            double mrr, double mri, double eps, const double *gx, const double *gw)
{
    const double PI = 3.14159265358979323846;
    const double kw = PI * 2.0 / lam;
    const double A = a_rev * pow(eps, 1.0 / 3.0);
    const double aa = A * A, ee = eps * eps, ee1 = ee - 1.0;
    double v0 = 1.0 / (mrr * mrr + mri * mri);
    const double prr = mrr * v0, pri = -mri * v0;
    const long goff = (long)g * (g - 1) / 2;
    for (int i = grp.tid; i < g; i += grp.nt) {
        double xp = gx[goff + g - 1 - i], wv = gw[goff + g - 1 - i];
        double x = -xp;
        double yv = 1.0 / (1.0 - x * x);
        c.ss[i] = yv;
        c.s[i] = sqrt(yv);
        double cc = x * x, ssq = 1.0 - cc, sth = sqrt(ssq);
        double rq = 1.0 / (ssq + ee * cc);
        double r = aa * rq;
        c.r[i] = r;
        c.dr[i] = rq * x * sth * ee1;
        double v = sqrt(r) * kw;
        double vi = 1.0 / v;
        c.ddr[i] = vi;
        c.drr[i] = prr * vi;
        c.dri[i] = pri * vi;
        c.x[i] = xp;
        c.w[i] = wv;
        c.rr[i] = wv * r;
    }
    grp.sync();
    double ta = fmax(sqrt(c.r[0]) * kw, sqrt(c.r[g - 1]) * kw);
    double tb = ta * sqrt(mrr * mrr + mri * mri);
    tb = fmax(tb, (double)nmax);
    double tmx = fmax(ta, (double)nmax);
    int nnmax1 = (int)(1.2 * sqrt(tmx) + 3.0);
    int nnmax2 = (int)(tb + 4.0 * pow(tb, 0.33333) + 1.2 * sqrt(tb));
    nnmax2 = nnmax2 - nmax + 5;
    const int ld = c.ld;
    for (int t = grp.tid; t < 4 * g; t += grp.nt) {
        int kind = t / g, i = t - kind * g;
        long o = (long)i * ld;
        double v = sqrt(c.r[i]) * kw;
        if (kind == 2) cjb(v * mrr, v * mri, c.bjr + o, c.bji + o, c.bdjr + o, c.bdji + o, nmax, nnmax2);
        else if (kind == 0) rjb(v, c.bj + o, c.bdj + o, nmax, nnmax1);
        else if (kind == 1) ryb(v, c.by + o, c.bdy + o, nmax);
        else vig_reflect(c.x[i], nmax, 0, c.d1 + o, c.d2 + o);
    }
    grp.sync();
}

// Wigner tables + per-node weights for mode m >= 1.  Modes are visited in increasing m, so sin^m(theta) is carried
// in c.qsp[] and the square roots of the recurrence come from a per-mode table built by the whole group:
// a step is 4 multiplies-adds instead of 2 sqrt + 2 divisions.
template <bool WARP>
__device__ void mode_setup(const Grp<WARP> &grp, const Carve &c, int nmax, int g, int m, const double *cm,
                           const double *r2n1)
{
    const double qm = (double)m, qmm = qm * qm;
    for (int n = m + grp.tid; n <= nmax + 1; n += grp.nt) {
        double v = sqrt((double)n * (double)n - qmm);
        c.sq[n] = v;
        c.rsq[n] = (n > m) ? 1.0 / v : 0.0;
    }
    grp.sync();
    for (int i = grp.tid; i < g; i += grp.nt) {
        c.ds[i] = c.s[i] * qm * c.rr[i];
        c.dss[i] = c.ss[i] * qmm;
        const double x = c.x[i];
        const double qs = sqrt(1.0 - x * x), qs1 = 1.0 / qs;
        double qp = (m == 1) ? qs : c.qsp[i] * qs;
        c.qsp[i] = qp;
        double *d1 = c.d1 + (long)i * c.ld, *d2 = c.d2 + (long)i * c.ld;
        double a1 = 0.0, a2 = cm[m] * qp;
        for (int n = m; n <= nmax; n++) {
            const double qn = n, qn1 = n + 1, qn2 = 2 * n + 1;
            const double qnm = c.sq[n], qnm1 = c.sq[n + 1];
            const double a3 = (qn2 * x * a2 - qnm * a1) * c.rsq[n + 1];
            const double der = qs1 * (-qn1 * qnm * a1 + qn * qnm1 * a3) * r2n1[n];
            const double si = ((n + m) & 1) ? -1.0 : 1.0;
            d1[n - 1] = a2 * si;
            d2[n - 1] = -der * si;
            a1 = a2; a2 = a3;
        }
    }
    grp.sync();
}

// ---- K2: assembly of the two parity-class systems [Q^T | -RgQ^T] (P6: TMATR0 / TMATR) ----
__device__ __forceinline__ double2 combine(double pir, double pii, double ppi, double pr, double pi, double qr, double qi)
{
    return make_double2(pir * pr - pii * pi + ppi * qr, pir * pi + pii * pr + ppi * qi);
}

template <bool WARP>
__device__ void assemble(const Grp<WARP> &grp, const Carve &c, int nmax, int g, int m, double ppi,
                         double pir, double pii, const double *an, const double *dd)
{
    const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1;
    const int h0 = (NM + 1) >> 1, h1 = NM >> 1;
    const int nE = h0 * h0 + h1 * h1, nO = 2 * h0 * h1;
    const int ld = c.ld, ldc = 2 * NM;
    double2 *sys = c.sys;
    // even pairs: X12, X21
    for (int t = grp.tid; t < nE; t += grp.nt) {
        int k1, k2;
        if (t < h0 * h0) { int a = t / h0; k1 = 2 * a; k2 = 2 * (t - a * h0); }
        else { int u = t - h0 * h0; int a = u / h1; k1 = 2 * a + 1; k2 = 2 * (u - a * h1) + 1; }
        const int n1 = nmin + k1, n2 = nmin + k2;
        const double an1 = an[n1], an2 = an[n2];
        double ar12 = 0, ai12 = 0, gr12 = 0, gi12 = 0, ar21 = 0, ai21 = 0, gr21 = 0, gi21 = 0;
        const double *pj = c.bj + (n1 - 1), *py = c.by + (n1 - 1), *pdj = c.bdj + (n1 - 1), *pdy = c.bdy + (n1 - 1);
        const double *pjr = c.bjr + (n2 - 1), *pji = c.bji + (n2 - 1), *pdjr = c.bdjr + (n2 - 1), *pdji = c.bdji + (n2 - 1);
        const double *pd1a = c.d1 + (n1 - 1), *pd2a = c.d2 + (n1 - 1), *pd1b = c.d1 + (n2 - 1), *pd2b = c.d2 + (n2 - 1);
        for (int i = 0; i < g; i++) {
            const long o = (long)i * ld;
            double d1n1 = pd1a[o], d2n1 = pd2a[o], d1n2 = pd1b[o], d2n2 = pd2b[o];
            double a11 = d1n1 * d1n2, a12 = d1n1 * d2n2, a21 = d2n1 * d1n2, a22 = d2n1 * d2n2;
            double aa2 = (m == 0) ? a22 : a11 * c.dss[i] + a22;
            double qj1 = pj[o], qy1 = py[o], qdj1 = pdj[o], qdy1 = pdy[o];
            double qjr2 = pjr[o], qji2 = pji[o], qdjr2 = pdjr[o], qdji2 = pdji[o];
            double c1r = qjr2 * qj1, c1i = qji2 * qj1;
            double b1r = c1r - qji2 * qy1, b1i = c1i + qjr2 * qy1;
            double c2r = qjr2 * qdj1, c2i = qji2 * qdj1;
            double b2r = c2r - qji2 * qdy1, b2i = c2i + qjr2 * qdy1;
            double ddri = c.ddr[i];
            double c3r = ddri * c1r, c3i = ddri * c1i, b3r = ddri * b1r, b3i = ddri * b1i;
            double c4r = qdjr2 * qj1, c4i = qdji2 * qj1;
            double b4r = c4r - qdji2 * qy1, b4i = c4i + qdjr2 * qy1;
            double drri = c.drr[i], drii = c.dri[i];
            double c5r = c1r * drri - c1i * drii, c5i = c1i * drri + c1r * drii;
            double b5r = b1r * drri - b1i * drii, b5i = b1i * drri + b1r * drii;
            double uri = c.dr[i], rri = c.rr[i];
            double f1 = rri * aa2, f2 = rri * uri * an1 * a12;
            ar12 += f1 * b2r + f2 * b3r; ai12 += f1 * b2i + f2 * b3i;
            gr12 += f1 * c2r + f2 * c3r; gi12 += f1 * c2i + f2 * c3i;
            f2 = rri * uri * an2 * a21;
            ar21 += f1 * b4r + f2 * b5r; ai21 += f1 * b4i + f2 * b5i;
            gr21 += f1 * c4r + f2 * c5r; gi21 += f1 * c4i + f2 * c5i;
        }
        const double an12 = dd[n1] * dd[n2] * 0.5 * 2.0;
        double r12 = ar12 * an12, i12 = ai12 * an12, rg12 = gr12 * an12, ig12 = gi12 * an12;
        double r21 = ar21 * an12, i21 = ai21 * an12, rg21 = gr21 * an12, ig21 = gi21 * an12;
        // t12 = -i*x12 -> (i12, -r12); t21 = +i*x21 -> (-i21, r21)
        double2 qa = combine(pir, pii, ppi, -i21, r21, i12, -r12);
        double2 qb = combine(pir, pii, ppi, i12, -r12, -i21, r21);
        double2 ga = combine(pir, pii, ppi, -ig21, rg21, ig12, -rg12);
        double2 gb = combine(pir, pii, ppi, ig12, -rg12, -ig21, rg21);
        const int ca = (n1 + 1) & 1, cb = ca ^ 1;
        sys[(long)(ca * NM + k2) * ldc + k1] = qa;
        sys[(long)(ca * NM + k2) * ldc + NM + k1] = make_double2(-ga.x, -ga.y);
        sys[(long)(cb * NM + k2) * ldc + k1] = qb;
        sys[(long)(cb * NM + k2) * ldc + NM + k1] = make_double2(-gb.x, -gb.y);
    }
    // odd pairs: X11, X22 (identically zero for m = 0)
    for (int t = grp.tid; t < nO; t += grp.nt) {
        int k1, k2;
        if (t < h0 * h1) { int a = t / h1; k1 = 2 * a; k2 = 2 * (t - a * h1) + 1; }
        else { int u = t - h0 * h1; int a = u / h0; k1 = 2 * a + 1; k2 = 2 * (u - a * h0); }
        const int n1 = nmin + k1, n2 = nmin + k2;
        double ar11 = 0, ai11 = 0, gr11 = 0, gi11 = 0, ar22 = 0, ai22 = 0, gr22 = 0, gi22 = 0;
        if (m != 0) {
            const double an1 = an[n1], an2 = an[n2];
            const double *pj = c.bj + (n1 - 1), *py = c.by + (n1 - 1), *pdj = c.bdj + (n1 - 1), *pdy = c.bdy + (n1 - 1);
            const double *pjr = c.bjr + (n2 - 1), *pji = c.bji + (n2 - 1), *pdjr = c.bdjr + (n2 - 1), *pdji = c.bdji + (n2 - 1);
            const double *pd1a = c.d1 + (n1 - 1), *pd2a = c.d2 + (n1 - 1), *pd1b = c.d1 + (n2 - 1), *pd2b = c.d2 + (n2 - 1);
            for (int i = 0; i < g; i++) {
                const long o = (long)i * ld;
                double d1n1 = pd1a[o], d2n1 = pd2a[o], d1n2 = pd1b[o], d2n2 = pd2b[o];
                double a11 = d1n1 * d1n2, a12 = d1n1 * d2n2, a21 = d2n1 * d1n2;
                double aa1 = a12 + a21;
                double qj1 = pj[o], qy1 = py[o], qdj1 = pdj[o], qdy1 = pdy[o];
                double qjr2 = pjr[o], qji2 = pji[o], qdjr2 = pdjr[o], qdji2 = pdji[o];
                double c1r = qjr2 * qj1, c1i = qji2 * qj1;
                double b1r = c1r - qji2 * qy1, b1i = c1i + qjr2 * qy1;
                double c2r = qjr2 * qdj1, c2i = qji2 * qdj1;
                double b2r = c2r - qji2 * qdy1, b2i = c2i + qjr2 * qdy1;
                double ddri = c.ddr[i];
                double c4r = qdjr2 * qj1, c4i = qdji2 * qj1;
                double b4r = c4r - qdji2 * qy1, b4i = c4i + qdjr2 * qy1;
                double drri = c.drr[i], drii = c.dri[i];
                double c6r = qdjr2 * qdj1, c6i = qdji2 * qdj1;
                double b6r = c6r - qdji2 * qdy1, b6i = c6i + qdjr2 * qdy1;
                double c7r = c4r * ddri, c7i = c4i * ddri, b7r = b4r * ddri, b7i = b4i * ddri;
                double c8r = c2r * drri - c2i * drii, c8i = c2i * drri + c2r * drii;
                double b8r = b2r * drri - b2i * drii, b8i = b2i * drri + b2r * drii;
                double uri = c.dr[i], dsi = c.ds[i];
                double e1 = dsi * aa1;
                ar11 += e1 * b1r; ai11 += e1 * b1i; gr11 += e1 * c1r; gi11 += e1 * c1i;
                double e2 = dsi * uri * a11, e3 = e2 * an2;
                e2 = e2 * an1;
                ar22 += e1 * b6r + e2 * b7r + e3 * b8r; ai22 += e1 * b6i + e2 * b7i + e3 * b8i;
                gr22 += e1 * c6r + e2 * c7r + e3 * c8r; gi22 += e1 * c6i + e2 * c7i + e3 * c8i;
            }
        }
        const double an12 = dd[n1] * dd[n2] * 0.5 * 2.0;
        double r11 = ar11 * an12, i11 = ai11 * an12, rg11 = gr11 * an12, ig11 = gi11 * an12;
        double r22 = ar22 * an12, i22 = ai22 * an12, rg22 = gr22 * an12, ig22 = gi22 * an12;
        // t11 = -x11, t22 = -x22
        double2 qa = combine(pir, pii, ppi, -r11, -i11, -r22, -i22);
        double2 qb = combine(pir, pii, ppi, -r22, -i22, -r11, -i11);
        double2 ga = combine(pir, pii, ppi, -rg11, -ig11, -rg22, -ig22);
        double2 gb = combine(pir, pii, ppi, -rg22, -ig22, -rg11, -ig11);
        const int ca = (n1 + 1) & 1, cb = ca ^ 1;
        sys[(long)(ca * NM + k2) * ldc + k1] = qa;
        sys[(long)(ca * NM + k2) * ldc + NM + k1] = make_double2(-ga.x, -ga.y);
        sys[(long)(cb * NM + k2) * ldc + k1] = qb;
        sys[(long)(cb * NM + k2) * ldc + NM + k1] = make_double2(-gb.x, -gb.y);
    }
    grp.sync();
}

// ---- K2 (tiled): the same integrals as a register-tiled panel contraction ----
// Every integrand of TMATR factorises into a left factor that depends on (n1, node) and a right factor that depends on
// (n2, node):  with P1 = d1 dF, P2 = d2 dF, P3 = d1 F, P4 = d2 F  (F = j or y at n1; d1, d2 the Wigner functions) and
// jc, djc the complex-argument functions at n2,
//   X12_F = sum_i P1 (W dss d1 jc) + P2 (W d2 jc) + an1 P3 (W u ddr d2 jc)
//   X21_F = sum_i P3 (W dss d1 djc) + P4 (W d2 djc + an2 W u d1 jc/zeta)
//   X11_F = sum_i P3 (ds d2 jc) + P4 (ds d1 jc)
//   X22_F = sum_i P1 (ds d2 djc + an2 ds u d1 jc/zeta) + P2 (ds d1 djc) + an1 P3 (ds u ddr d1 djc)
// (W = w r^2, ds = m W / sin, dss = m^2 / sin^2, u = r'/r).  Q uses F = j + i y, RgQ uses F = j, so the thread keeps
// X_j and X_y.  That is 20 real FMAs per (pair, node) instead of ~60, and a TM x TN tile of same-parity pairs per
// thread re-uses each loaded row TN (TM) times: the per-pair loop above is shared-memory-bandwidth bound
// (18 LDS per pair-node), the tiled one is not.
__device__ __forceinline__ void store_even(double2 *sys, int NM, int nmin, int k1, int k2, double ppi, double pir,
                                           double pii, const double *dd, const double *x12j, const double *x12y,
                                           const double *x21j, const double *x21y)
{
    const int n1 = nmin + k1, n2 = nmin + k2, ldc = 2 * NM;
    const double an12 = dd[n1] * dd[n2] * 0.5 * 2.0;
    // h = j + i y:  A = X_j + i X_y ; G = X_j
    const double r12 = (x12j[0] - x12y[1]) * an12, i12 = (x12j[1] + x12y[0]) * an12;
    const double rg12 = x12j[0] * an12, ig12 = x12j[1] * an12;
    const double r21 = (x21j[0] - x21y[1]) * an12, i21 = (x21j[1] + x21y[0]) * an12;
    const double rg21 = x21j[0] * an12, ig21 = x21j[1] * an12;
    double2 qa = combine(pir, pii, ppi, -i21, r21, i12, -r12);
    double2 qb = combine(pir, pii, ppi, i12, -r12, -i21, r21);
    double2 ga = combine(pir, pii, ppi, -ig21, rg21, ig12, -rg12);
    double2 gb = combine(pir, pii, ppi, ig12, -rg12, -ig21, rg21);
    const int ca = (n1 + 1) & 1, cb = ca ^ 1;
    sys[(long)(ca * NM + k2) * ldc + k1] = qa;
    sys[(long)(ca * NM + k2) * ldc + NM + k1] = make_double2(-ga.x, -ga.y);
    sys[(long)(cb * NM + k2) * ldc + k1] = qb;
    sys[(long)(cb * NM + k2) * ldc + NM + k1] = make_double2(-gb.x, -gb.y);
}

__device__ __forceinline__ void store_odd(double2 *sys, int NM, int nmin, int k1, int k2, double ppi, double pir,
                                          double pii, const double *dd, const double *x11j, const double *x11y,
                                          const double *x22j, const double *x22y)
{
    const int n1 = nmin + k1, n2 = nmin + k2, ldc = 2 * NM;
    const double an12 = dd[n1] * dd[n2] * 0.5 * 2.0;
    const double r11 = (x11j[0] - x11y[1]) * an12, i11 = (x11j[1] + x11y[0]) * an12;
    const double rg11 = x11j[0] * an12, ig11 = x11j[1] * an12;
    const double r22 = (x22j[0] - x22y[1]) * an12, i22 = (x22j[1] + x22y[0]) * an12;
    const double rg22 = x22j[0] * an12, ig22 = x22j[1] * an12;
    double2 qa = combine(pir, pii, ppi, -r11, -i11, -r22, -i22);
    double2 qb = combine(pir, pii, ppi, -r22, -i22, -r11, -i11);
    double2 ga = combine(pir, pii, ppi, -rg11, -ig11, -rg22, -ig22);
    double2 gb = combine(pir, pii, ppi, -rg22, -ig22, -rg11, -ig11);
    const int ca = (n1 + 1) & 1, cb = ca ^ 1;
    sys[(long)(ca * NM + k2) * ldc + k1] = qa;
    sys[(long)(ca * NM + k2) * ldc + NM + k1] = make_double2(-ga.x, -ga.y);
    sys[(long)(cb * NM + k2) * ldc + k1] = qb;
    sys[(long)(cb * NM + k2) * ldc + NM + k1] = make_double2(-gb.x, -gb.y);
}

template <bool WARP, int TM, int TN>
__device__ void assemble_tiled(const Grp<WARP> &grp, const Carve &c, int nmax, int g, int m, double ppi, double pir,
                               double pii, const double *an, const double *dd)
{
    const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1;
    const int h0 = (NM + 1) >> 1, h1 = NM >> 1;  // number of even-k / odd-k indices
    const int rE = (h0 + TM - 1) / TM, rO = (h1 + TM - 1) / TM, cE = (h0 + TN - 1) / TN, cO = (h1 + TN - 1) / TN;
    const int nEE = rE * cE, nOO = rO * cO, nEO = rE * cO, nOE = rO * cE;
    const int nEven = nEE + nOO, nEvenPad = (nEven + 31) & ~31, nOdd = (m == 0) ? 0 : nEO + nOE;
    const int ld = c.ld;
    for (int t = grp.tid; t < nEvenPad + nOdd; t += grp.nt) {
        if (t >= nEven && t < nEvenPad) continue;
        const bool even = t < nEven;
        int u = even ? t : t - nEvenPad;
        int prow, pcol, a0, b0, hrow, hcol;
        if (even) {
            if (u < nEE) { prow = 0; pcol = 0; a0 = u / cE; b0 = u - a0 * cE; hrow = h0; hcol = h0; }
            else { u -= nEE; prow = 1; pcol = 1; a0 = u / cO; b0 = u - a0 * cO; hrow = h1; hcol = h1; }
        } else {
            if (u < nEO) { prow = 0; pcol = 1; a0 = u / cO; b0 = u - a0 * cO; hrow = h0; hcol = h1; }
            else { u -= nEO; prow = 1; pcol = 0; a0 = u / cE; b0 = u - a0 * cE; hrow = h1; hcol = h0; }
        }
        int k1[TM], k2[TN];
        bool v1[TM], v2[TN];
        const double *pl[TM], *pr_[TN];
        double an1[TM], an2[TN];
#pragma unroll
        for (int r = 0; r < TM; r++) {
            int a = a0 * TM + r;
            v1[r] = a < hrow;
            if (!v1[r]) a = hrow - 1;
            k1[r] = 2 * a + prow;
            pl[r] = c.bj + (nmin + k1[r] - 1);
            an1[r] = an[nmin + k1[r]];
        }
#pragma unroll
        for (int q = 0; q < TN; q++) {
            int b = b0 * TN + q;
            v2[q] = b < hcol;
            if (!v2[q]) b = hcol - 1;
            k2[q] = 2 * b + pcol;
            pr_[q] = c.bjr + (nmin + k2[q] - 1);
            an2[q] = an[nmin + k2[q]];
        }
        // offsets of the other tables relative to bj / bjr (all tables have the same g x ld shape)
        const long oY = c.by - c.bj, oDJ = c.bdj - c.bj, oDY = c.bdy - c.bj, oD1 = c.d1 - c.bj, oD2 = c.d2 - c.bj;
        const long oJI = c.bji - c.bjr, oDJR = c.bdjr - c.bjr, oDJI = c.bdji - c.bjr, oRD1 = c.d1 - c.bjr, oRD2 = c.d2 - c.bjr;
        double xa[TM][TN][4], xb[TM][TN][4];  // even: xa = X12 (j.re, j.im, y.re, y.im), xb = X21; odd: X11, X22
#pragma unroll
        for (int r = 0; r < TM; r++)
#pragma unroll
            for (int q = 0; q < TN; q++)
#pragma unroll
                for (int e = 0; e < 4; e++) { xa[r][q][e] = 0.0; xb[r][q][e] = 0.0; }
        for (int i = 0; i < g; i++) {
            const long o = (long)i * ld;
            const double uri = c.dr[i], ddri = c.ddr[i], drri = c.drr[i], drii = c.dri[i];
            double P1j[TM], P2j[TM], P3j[TM], P4j[TM], P3aj[TM], P1y[TM], P2y[TM], P3y[TM], P4y[TM], P3ay[TM];
#pragma unroll
            for (int r = 0; r < TM; r++) {
                const double *b = pl[r] + o;
                const double d1 = b[oD1], d2 = b[oD2], fj = b[0], fy = b[oY], fdj = b[oDJ], fdy = b[oDY];
                P1j[r] = d1 * fdj; P2j[r] = d2 * fdj; P3j[r] = d1 * fj; P4j[r] = d2 * fj; P3aj[r] = an1[r] * P3j[r];
                P1y[r] = d1 * fdy; P2y[r] = d2 * fdy; P3y[r] = d1 * fy; P4y[r] = d2 * fy; P3ay[r] = an1[r] * P3y[r];
            }
            if (even) {
                const double rri = c.rr[i];
                const double dssi = (m == 0) ? 0.0 : c.dss[i];
#pragma unroll
                for (int q = 0; q < TN; q++) {
                    const double *b = pr_[q] + o;
                    const double d1 = b[oRD1], d2 = b[oRD2], jr = b[0], ji = b[oJI], djr = b[oDJR], dji = b[oDJI];
                    const double w1 = rri * dssi * d1, w2 = rri * d2, w3 = w2 * uri * ddri, w6 = an2[q] * rri * uri * d1;
                    const double jzr = jr * drri - ji * drii, jzi = ji * drri + jr * drii;
                    const double q1r = w1 * jr, q1i = w1 * ji, q2r = w2 * jr, q2i = w2 * ji, q3r = w3 * jr, q3i = w3 * ji;
                    const double q4r = w1 * djr, q4i = w1 * dji;
                    const double q5r = w2 * djr + w6 * jzr, q5i = w2 * dji + w6 * jzi;  // Q5 + an2 Q6
#pragma unroll
                    for (int r = 0; r < TM; r++) {
                        double *A = xa[r][q], *B = xb[r][q];
                        A[0] += P1j[r] * q1r + P2j[r] * q2r + P3aj[r] * q3r;
                        A[1] += P1j[r] * q1i + P2j[r] * q2i + P3aj[r] * q3i;
                        A[2] += P1y[r] * q1r + P2y[r] * q2r + P3ay[r] * q3r;
                        A[3] += P1y[r] * q1i + P2y[r] * q2i + P3ay[r] * q3i;
                        B[0] += P3j[r] * q4r + P4j[r] * q5r;
                        B[1] += P3j[r] * q4i + P4j[r] * q5i;
                        B[2] += P3y[r] * q4r + P4y[r] * q5r;
                        B[3] += P3y[r] * q4i + P4y[r] * q5i;
                    }
                }
            } else {
                const double dsi = c.ds[i];
#pragma unroll
                for (int q = 0; q < TN; q++) {
                    const double *b = pr_[q] + o;
                    const double d1 = b[oRD1], d2 = b[oRD2], jr = b[0], ji = b[oJI], djr = b[oDJR], dji = b[oDJI];
                    const double v1_ = dsi * d1, v2_ = dsi * d2, v5 = v1_ * uri * ddri, v6 = an2[q] * v1_ * uri;
                    const double jzr = jr * drri - ji * drii, jzi = ji * drri + jr * drii;
                    const double r1r = v2_ * jr, r1i = v2_ * ji, r2r = v1_ * jr, r2i = v1_ * ji;
                    const double r3r = v2_ * djr + v6 * jzr, r3i = v2_ * dji + v6 * jzi;  // R3 + an2 R6
                    const double r4r = v1_ * djr, r4i = v1_ * dji, r5r = v5 * djr, r5i = v5 * dji;
#pragma unroll
                    for (int r = 0; r < TM; r++) {
                        double *A = xa[r][q], *B = xb[r][q];
                        A[0] += P3j[r] * r1r + P4j[r] * r2r;
                        A[1] += P3j[r] * r1i + P4j[r] * r2i;
                        A[2] += P3y[r] * r1r + P4y[r] * r2r;
                        A[3] += P3y[r] * r1i + P4y[r] * r2i;
                        B[0] += P1j[r] * r3r + P2j[r] * r4r + P3aj[r] * r5r;
                        B[1] += P1j[r] * r3i + P2j[r] * r4i + P3aj[r] * r5i;
                        B[2] += P1y[r] * r3r + P2y[r] * r4r + P3ay[r] * r5r;
                        B[3] += P1y[r] * r3i + P2y[r] * r4i + P3ay[r] * r5i;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < TM; r++)
#pragma unroll
            for (int q = 0; q < TN; q++) {
                if (!(v1[r] && v2[q])) continue;
                if (even) store_even(c.sys, NM, nmin, k1[r], k2[q], ppi, pir, pii, dd, xa[r][q], xa[r][q] + 2, xb[r][q], xb[r][q] + 2);
                else store_odd(c.sys, NM, nmin, k1[r], k2[q], ppi, pir, pii, dd, xa[r][q], xa[r][q] + 2, xb[r][q], xb[r][q] + 2);
            }
    }
    if (m == 0) {
        // odd pairs vanish identically for m = 0
        const int ldc = 2 * NM;
        for (int t = grp.tid; t < NM * NM; t += grp.nt) {
            int k1 = t / NM, k2 = t - k1 * NM;
            if (((k1 + k2) & 1) == 0) continue;
            const double2 z = make_double2(0.0, 0.0);
            for (int cl = 0; cl < 2; cl++) {
                c.sys[(long)(cl * NM + k2) * ldc + k1] = z;
                c.sys[(long)(cl * NM + k2) * ldc + NM + k1] = z;
            }
        }
    }
    grp.sync();
}

// dispatcher: tiled contraction for larger blocks, per-pair loop for the tiny ones (too few tiles to fill a group)
template <bool WARP, bool TILED>
__device__ void assemble_any(const Grp<WARP> &grp, const Carve &c, int nmax, int g, int m, double ppi, double pir,
                             double pii, const double *an, const double *dd, int tile_min)
{
    if constexpr (TILED) {
        const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1;
        if (NM >= tile_min) { assemble_tiled<WARP, 2, 2>(grp, c, nmax, g, m, ppi, pir, pii, an, dd); return; }
    }
    assemble(grp, c, nmax, g, m, ppi, pir, pii, an, dd);
}

// ---- K3: lock-step Gauss-Jordan with row pivoting over several independent systems (P7: TT, T = -RgQ Q^-1) ----
// Each system solves A X = B in place inside an augmented row-major block; element (r, j) of [A | B] lives at
//   base[(off + stride*r) * ld + (j < n ? off + stride*j : rhs + off + stride*(j - n))].
// A parity-class system of a mode is {n = NM, off = 0, stride = 1, rhs = NM}; for m = 0 each class splits further into
// two interleaved sub-systems (stride 2), because T12 = T21 = 0.  All systems advance through the elimination steps
// together, so the serial latency of a step (pivot search, barriers) is paid once for up to HS_MAXSYS systems.
#define HS_MAXSYS 8
struct SysDesc {
    double2 *base;
    int n, ld, off, stride, rhs;
};
struct SolveCtl {
    unsigned long long key[2][HS_MAXSYS];  // packed (|pivot candidate|, row) maxima, ping-pong per step
};
__device__ __forceinline__ long sys_at(const SysDesc &d, int r, int j)
{
    return (long)(d.off + d.stride * r) * d.ld + (j < d.n ? d.off + d.stride * j : d.rhs + d.off + d.stride * (j - d.n));
}
// order-preserving key: |re|+|im| as IEEE bits (non-negative doubles compare like unsigned integers), the low 8
// mantissa bits replaced by (255 - row) so that the smaller row wins among (near-)equal candidates
__device__ __forceinline__ unsigned long long pivot_key(double2 v, int r)
{
    double av = fabs(v.x) + fabs(v.y);
    unsigned long long b;
#ifdef HS_HOST_EMU
    memcpy(&b, &av, 8);
#else
    b = (unsigned long long)__double_as_longlong(av);
#endif
    return (b & ~0xFFull) | (unsigned long long)(255 - r);
}
#ifdef HS_HOST_EMU
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
#endif

// The pivot search is fused into the elimination: while a step updates the next pivot column, the updating threads
// publish their candidates with one shared-memory atomicMax each, so a step costs two barriers and no serial phase.
template <bool WARP>
__device__ void solve_multi(const Grp<WARP> &grp, const SysDesc *sd, int nsys, SolveCtl *sc, int *sing)
{
    int nmaxsys = 0, rows_tot = 0;
    for (int s = 0; s < nsys; s++) { nmaxsys = sd[s].n > nmaxsys ? sd[s].n : nmaxsys; rows_tot += sd[s].n; }
    // candidates of the first pivot column
    for (int t = grp.tid; t < 2 * HS_MAXSYS; t += grp.nt) sc->key[t / HS_MAXSYS][t % HS_MAXSYS] = 0ull;
    grp.sync();
    for (int t = grp.tid; t < rows_tot; t += grp.nt) {
        int s = 0, r = t;
        while (r >= sd[s].n) { r -= sd[s].n; s++; }
        atomicMax(&sc->key[0][s], pivot_key(sd[s].base[sys_at(sd[s], r, 0)], r));
    }
    grp.sync();
    for (int k = 0; k < nmaxsys; k++) {
        const unsigned long long *keyk = sc->key[k & 1];
        // total live columns over the active systems
        int cstart[HS_MAXSYS + 1];
        int tot = 0;
#pragma unroll
        for (int s = 0; s < HS_MAXSYS; s++) {
            cstart[s] = tot;
            if (s < nsys && k < sd[s].n) tot += 2 * sd[s].n - k - 1;
        }
        cstart[HS_MAXSYS] = tot;
        // (1) swap + scale the pivot rows (every thread derives pivot row and reciprocal itself: no serial phase)
        for (int t = grp.tid; t < tot; t += grp.nt) {
            int s = 0;
#pragma unroll
            for (int q = 1; q < HS_MAXSYS; q++) s += (t >= cstart[q]) ? 1 : 0;
            const SysDesc d = sd[s];
            const int j = k + 1 + (t - cstart[s]);
            const unsigned long long key = keyk[s];
            const int p = 255 - (int)(key & 0xFFull);
            if ((key >> 8) == 0ull && j == k + 1) *sing = 1;
            const double2 pv = d.base[sys_at(d, p, k)];
            const double den = pv.x * pv.x + pv.y * pv.y;
            const double2 inv = make_double2(pv.x / den, -pv.y / den);
            const long ik = sys_at(d, k, j), ip = sys_at(d, p, j);
            double2 a = d.base[ik], b = d.base[ip];
            d.base[ik] = make_double2(b.x * inv.x - b.y * inv.y, b.x * inv.y + b.y * inv.x);
            if (p != k) d.base[ip] = a;
        }
        for (int t = grp.tid; t < HS_MAXSYS; t += grp.nt) sc->key[(k + 1) & 1][t] = 0ull;
        grp.sync();
        // (2) rank-1 update of all other rows: threads tile [row group] x [live column of any system]; the threads that
        //     own column k+1 publish the next pivot candidates
        if (tot > 0) {
            const int w = tot < grp.nt ? tot : grp.nt;
            const int nrg = grp.nt / w, rg = grp.tid / w, c0 = grp.tid - rg * w;
            if (rg < nrg) {
                for (int t = c0; t < tot; t += w) {
                    int s = 0;
#pragma unroll
                    for (int q = 1; q < HS_MAXSYS; q++) s += (t >= cstart[q]) ? 1 : 0;
                    const SysDesc d = sd[s];
                    const int j = k + 1 + (t - cstart[s]);
                    const int p = 255 - (int)(keyk[s] & 0xFFull);
                    // element (r, j) = base[r * rstep + row0 + colj]: hoist all index arithmetic out of the row loop
                    const long rstep = (long)d.stride * d.ld, row0 = (long)d.off * d.ld;
                    const long colj = (j < d.n ? d.off + d.stride * j : d.rhs + d.off + d.stride * (j - d.n));
                    const long colk = d.off + d.stride * k;
                    double2 *bj_ = d.base + row0 + colj;
                    const double2 *bk_ = d.base + row0 + colk;
                    const double2 nk = bj_[k * rstep];
                    const double2 fkk = bk_[k * rstep];
                    const bool next_col = (j == k + 1) && (k + 1 < d.n);
                    unsigned long long best = 0ull;
                    for (int r = rg; r < d.n; r += nrg) {
                        if (r == k) continue;
                        // after the row swap, row p holds the old row k whose column-k entry is still at (k, k)
                        const double2 f = (r == p) ? fkk : bk_[r * rstep];
                        double2 v = bj_[r * rstep];
                        v.x -= f.x * nk.x - f.y * nk.y;
                        v.y -= f.x * nk.y + f.y * nk.x;
                        bj_[r * rstep] = v;
                        if (next_col && r > k) { unsigned long long ky = pivot_key(v, r); best = ky > best ? ky : best; }
                    }
                    if (next_col && best) atomicMax(&sc->key[(k + 1) & 1][s], best);
                }
            }
        }
        grp.sync();
    }
}

// Both parity-class systems of one mode stored as [2][NM][2NM] (used by the two-layer kernel).
template <bool WARP>
__device__ void solve(const Grp<WARP> &grp, const Carve &c, int NM, Ctl *ctl)
{
    SysDesc *sd = reinterpret_cast<SysDesc *>(ctl->sd);
    if (grp.tid == 0) {
        for (int cl = 0; cl < 2; cl++) {
            SysDesc d;
            d.base = c.sys + (long)cl * NM * 2 * NM; d.n = NM; d.ld = 2 * NM; d.off = 0; d.stride = 1; d.rhs = NM;
            sd[cl] = d;
        }
    }
    grp.sync();
    solve_multi(grp, sd, 2, reinterpret_cast<SolveCtl *>(ctl->sc), &ctl->sing);
}

#ifdef HS_HOST_EMU
#define HS_PROF(slot)
#else
#define HS_PROF(slot)                                                         \
    if (a.prof && grp.tid == 0) {                                             \
        long long now_ = clock64();                                           \
        atomicAdd((unsigned long long *)&a.prof[slot], (unsigned long long)(now_ - tprof)); \
        tprof = now_;                                                         \
    }
#endif
// ---- one particle: convergence loop (A.1) + mode sweep ----
template <bool WARP, bool TILED>
__device__ void calctmat_particle(const TmArgs &a, int pidx, double *ws, double *ws_fb, const Grp<WARP> &grp, Ctl *ctl,
                                  const double *an, const double *dd)
{
    const double PI = 3.14159265358979323846;
    double axi = a.axi[pidx], lam = a.lam[pidx], mrr = a.mr[pidx], mri = a.mi[pidx], eps = a.eps[pidx];
    double rat = a.rat;
    if (fabs(rat - 1.0) > 1e-8) {
        // SAREA: equal-surface-area radius -> equal-volume radius ratio
        double e, r;
        if (eps < 1.0) {
            e = sqrt(1.0 - eps * eps);
            r = 0.5 * (pow(eps, 2.0 / 3.0) + pow(eps, -1.0 / 3.0) * asin(e) / e);
        } else {
            e = sqrt(1.0 - 1.0 / (eps * eps));
            r = 0.25 * (2.0 * pow(eps, 2.0 / 3.0) + pow(eps, -4.0 / 3.0) * log((1.0 + e) / (1.0 - e)) / e);
        }
        rat = 1.0 / sqrt(r);
    }
    const double arev = rat * axi;
    const double xev = 2.0 * PI * arev / lam;
    int ixxx = (int)(xev + 4.05 * pow(xev, 0.333333));
    int nmax0 = ixxx > 4 ? ixxx : 4;
    const double kw = PI * 2.0 / lam;
    const double ppi = kw * kw, pir = ppi * mrr, pii = ppi * mri;
    const double ddelt = a.ddelt;

    if (grp.tid == 0) {
        ctl->nmax = nmax0; ctl->g = 0; ctl->phase = 0; ctl->done = 0; ctl->status = HS_ST_OK;
        ctl->ntr = 0; ctl->sing = 0; ctl->qsca1 = 0.0; ctl->qext1 = 0.0;
        bool bad = !(axi > 0.0) || !(lam > 0.0) || !(eps > 0.0) || !isfinite(mrr) || !isfinite(mri) || !(mrr * mrr + mri * mri > 0.0);
        if (bad) { ctl->done = 2; ctl->status = HS_ST_BAD_INPUT; }
        else if (nmax0 >= HS_NPN1) { ctl->done = 2; ctl->status = HS_ST_NMAX_LIMIT; }
    }
    grp.sync();
    Carve c;
#ifndef HS_HOST_EMU
    long long tprof = clock64();
#endif
    while (true) {
        if (ctl->done) break;
        int nmax = ctl->nmax, g;
        if (ctl->phase == 0) g = nmax * a.ndgs;
        else g = ctl->g + 1;
        grp.sync();  // everyone has read ctl before thread 0 rewrites it
        // limits and capacity
        int stop = 0;
        if (g > HS_NPNG1) stop = HS_ST_NGAUSS_LIMIT;
        else if (g > a.gmax) stop = HS_ST_INTERNAL_GAUSS;
        double *wsp = ws;
        if (ws_need(nmax, g) > (long long)a.ws_cap) {
            // outgrew the shared-memory slice: continue in the global-memory fall-back slice (slower, no re-launch)
            if (ws_fb && ws_need(nmax, g) <= (long long)a.fb_cap) wsp = ws_fb;
            else stop = HS_ST_INTERNAL_ESCALATE;
        }
        if (stop) {
            if (grp.tid == 0) { ctl->status = stop; ctl->done = 2; }
            grp.sync();
            break;
        }
        carve(c, wsp, nmax, g);
        HS_PROF(0)
        if (!(a.debug & 4) || ctl->ntr == 0) trial_setup(grp, c, nmax, g, arev, lam, mrr, mri, eps, a.gx, a.gw);
        HS_PROF(1)
        if (!(a.debug & 2)) assemble_any<WARP, TILED>(grp, c, nmax, g, 0, ppi, pir, pii, an, dd, a.tile_min);
        HS_PROF(2)
        if (!(a.debug & 1)) {
            // m = 0: T12 = T21 = 0, so each parity class splits into two interleaved half-size systems
            SysDesc *sd = reinterpret_cast<SysDesc *>(ctl->sd);
            if (grp.tid == 0) {
                int q = 0;
                for (int cl = 0; cl < 2; cl++)
                    for (int par = 0; par < 2; par++) {
                        SysDesc d;
                        d.base = c.sys + (long)cl * nmax * 2 * nmax; d.ld = 2 * nmax; d.off = par; d.stride = 2; d.rhs = nmax;
                        d.n = (nmax + 1 - par) >> 1;
                        if (d.n > 0) sd[q++] = d;
                    }
                ctl->item = q;
            }
            grp.sync();
            solve_multi(grp, sd, ctl->item, reinterpret_cast<SolveCtl *>(ctl->sc), &ctl->sing);
        }
        HS_PROF(3)
        if (grp.tid == 0) {
            const int NM = nmax, ldc = 2 * NM;
            double qext = 0.0, qsca = 0.0;
            for (int n = 1; n <= nmax; n++) {
                int k = n - 1;
                double2 t1 = c.sys[(long)(((n + 1) & 1) * NM + k) * ldc + NM + k];
                double2 t2 = c.sys[(long)((n & 1) * NM + k) * ldc + NM + k];
                double dn1 = (double)(2 * n + 1);
                qsca += dn1 * (t1.x * t1.x + t1.y * t1.y + t2.x * t2.x + t2.y * t2.y);
                qext += (t1.x + t2.x) * dn1;
            }
            double dsca = fabs((ctl->qsca1 - qsca) / qsca), dext = fabs((ctl->qext1 - qext) / qext);
            ctl->qext1 = qext; ctl->qsca1 = qsca;
            ctl->ntr++;
            ctl->g = g;
            bool ok = (dsca <= ddelt && dext <= ddelt);
            if (ctl->phase == 0) {
                if (ok) { ctl->phase = 1; if (g == HS_NPNG1) ctl->done = 1; }
                else if (nmax + 1 > HS_NPN1) { ctl->status = HS_ST_NMAX_LIMIT; ctl->done = 2; }
                else ctl->nmax = nmax + 1;
            } else {
                if (ok) ctl->done = 1;
                else if (g == HS_NPNG1) { ctl->status = HS_ST_NGAUSS_LIMIT; ctl->done = 1; }
            }
        }
        grp.sync();
    }
    const int nmax = ctl->nmax, g = ctl->g;
    if (ctl->done == 1) {
        // converged at (nmax, g); sys holds the m=0 solution.  Reserve arena space.
        if (grp.tid == 0) {
            unsigned long long sz = (unsigned long long)tmat_size(nmax);
            unsigned long long off = atomicAdd(a.arena_top, sz);
            if (off + sz > (unsigned long long)a.arena_cap) { ctl->status = HS_ST_ARENA_FULL; ctl->done = 2; }
            ctl->toff = off;
        }
        grp.sync();
        if (ctl->done == 1) {
            double2 *tg = reinterpret_cast<double2 *>(a.tarena) + ctl->toff;
            // m = 0 block: already solved by the last trial
            {
                const int NM = nmax, ldc = 2 * NM;
                for (int t = grp.tid; t < 2 * NM * NM; t += grp.nt) {
                    int row = t / NM, col = t - row * NM;  // row = c*NM + k'
                    tg[t] = c.sys[(long)row * ldc + NM + col];
                }
                tg += 2 * NM * NM;
                grp.sync();
            }
            // m >= 1: assemble as many modes as fit in the remaining workspace, then solve them in lock-step
            const long base_dbl = (13L * g + 2L * (nmax + 2) + 10L * g * c.ld + 1) & ~1L;
            const long cap_dbl = (c.x == ws ? (long)a.ws_cap : (long)a.fb_cap);
            const long sys_cap = (cap_dbl - base_dbl) / 2;  // complex entries available for systems
            double2 *sysbase = c.sys;
            SysDesc *sd = reinterpret_cast<SysDesc *>(ctl->sd);
            int m = 1;
            while (m <= nmax) {
                int M = 0;
                long used = 0;
                while (m + M <= nmax && M < HS_MAXSYS / 2) {
                    long nm = nmax - (m + M) + 1;
                    if (M > 0 && used + 4 * nm * nm > sys_cap) break;
                    used += 4 * nm * nm;
                    M++;
                }
                long off = 0;
                for (int i = 0; i < M; i++) {
                    const int mm = m + i, NM = nmax - mm + 1;
                    c.sys = sysbase + off;
                    HS_PROF(0)
                    mode_setup(grp, c, nmax, g, mm, an + (HS_NPN1 + 1) * 2, an + (HS_NPN1 + 1) * 3);
                    HS_PROF(4)
                    if (!(a.debug & 2)) assemble_any<WARP, TILED>(grp, c, nmax, g, mm, ppi, pir, pii, an, dd, a.tile_min);
                    HS_PROF(5)
                    if (grp.tid == 0) {
                        for (int cl = 0; cl < 2; cl++) {
                            SysDesc d;
                            d.base = c.sys + (long)cl * NM * 2 * NM; d.n = NM; d.ld = 2 * NM; d.off = 0; d.stride = 1; d.rhs = NM;
                            sd[2 * i + cl] = d;
                        }
                    }
                    off += 4L * NM * NM;
                }
                grp.sync();
                HS_PROF(0)
                if (!(a.debug & 1)) solve_multi(grp, sd, 2 * M, reinterpret_cast<SolveCtl *>(ctl->sc), &ctl->sing);
                HS_PROF(6)
                off = 0;
                for (int i = 0; i < M; i++) {
                    const int NM = nmax - (m + i) + 1, ldc = 2 * NM;
                    const double2 *sy = sysbase + off;
                    for (int t = grp.tid; t < 2 * NM * NM; t += grp.nt) {
                        int row = t / NM, col = t - row * NM;
                        tg[t] = sy[(long)row * ldc + NM + col];
                    }
                    tg += 2 * NM * NM;
                    off += 4L * NM * NM;
                }
                grp.sync();
                HS_PROF(7)
                m += M;
            }
            c.sys = sysbase;
        }
    }
    if (grp.tid == 0) {
        int st = ctl->status;
        if (st == HS_ST_OK && ctl->sing) st = HS_ST_SINGULAR;
        a.status[pidx] = st;
        a.nmax[pidx] = nmax;
        a.ngauss[pidx] = g;
        a.ntrials[pidx] = ctl->ntr;
        a.toff[pidx] = (ctl->done == 1) ? (long long)ctl->toff : -1;
    }
    grp.sync();
}

#define HS_MAX_GROUPS 8
#ifndef HS_HOST_EMU

template <bool WARP, bool GLOBAL_WS, bool TILED>
__global__ void __launch_bounds__(256) k_calctmat(TmArgs a)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ Ctl ctls[HS_MAX_GROUPS];
    __shared__ double s_tab[4 * (HS_NPN1 + 1)];  // an | dd | cm (prod sqrt((2i-1)/2i)) | 1/(2n+1)
    double *s_an = s_tab, *s_dd = s_tab + (HS_NPN1 + 1);
    for (int n = threadIdx.x; n <= HS_NPN1; n += blockDim.x) {
        if (n == 0) { s_an[0] = 0.0; s_dd[0] = 0.0; }
        else {
            int nn = n * (n + 1);
            s_an[n] = (double)nn;
            s_dd[n] = sqrt((double)(2 * n + 1) / (double)nn);
        }
        s_tab[3 * (HS_NPN1 + 1) + n] = 1.0 / (double)(2 * n + 1);
    }
    if (threadIdx.x == 0) {
        double cacc = 1.0;
        s_tab[2 * (HS_NPN1 + 1)] = 1.0;
        for (int i = 1; i <= HS_NPN1; i++) { cacc = cacc * sqrt((double)(2 * i - 1) / (double)(2 * i)); s_tab[2 * (HS_NPN1 + 1) + i] = cacc; }
    }
    __syncthreads();
    Grp<WARP> grp;
    int gidx;
    if (WARP) { grp.tid = threadIdx.x & 31; grp.nt = 32; gidx = threadIdx.x >> 5; }
    else { grp.tid = threadIdx.x; grp.nt = blockDim.x; gidx = 0; }
    double *ws = GLOBAL_WS ? (a.ws_global + (long long)blockIdx.x * a.ws_cap) : (smem + (long long)gidx * a.ws_cap);
    const int ngrp = WARP ? (blockDim.x >> 5) : 1;
    double *ws_fb = (GLOBAL_WS || !a.ws_fb) ? nullptr : a.ws_fb + ((long long)blockIdx.x * ngrp + gidx) * a.fb_cap;
    Ctl *ctl = &ctls[gidx];
    while (true) {
        if (grp.tid == 0) ctl->item = atomicAdd(a.queue, 1);
        grp.sync();
        int it = ctl->item;
        grp.sync();
        if (it >= a.n_items) break;
        calctmat_particle<WARP, TILED>(a, a.items[it], ws, ws_fb, grp, ctl, s_an, s_dd);
    }
}

#endif  // HS_HOST_EMU

}  // namespace hs

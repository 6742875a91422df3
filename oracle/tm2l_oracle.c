/*
 * tm2l_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's two-layer (coated spheroid) EBCM T-matrix program
 *   /root/reference/hydroscatt/f_2l_tmatrix/tmatrix_py.f  (Bringi & Seliga 1977 formulation)
 * as hydroscatt drives it (scattering.py:32-67 -> tm2l.tmatrix(2)).  Each function cites the Fortran lines it
 * follows.  Quirks kept on purpose:
 *   - cpi = 4.0*atan(1.0) evaluated in SINGLE precision (tmatrix_py.f:161) -> dtr, conk, sigma, epps(2)
 *   - aovrb**(1./3.) with a single-precision exponent literal (:617-618)
 *   - fixed nrank=17, nm=7 (m=0..6), nsect=2, ib=8, anginc=90, kgauss=5 (:549-554, 740)
 *   - 31-node Patterson rule on [0, cpi/2] in the routine's evaluation order (:1340-1595)
 *   - mirror-symmetry zeroing (:1074-1085, 1219), cmxnrm normalisation (:345-368, 387-399)
 *   - Gauss-Jordan with pivot = largest-modulus element of the current ROW, all-zero rows skipped (:2177-2280)
 *   - two scattering angles uang = 90 deg (forward) and 270 deg (back) from integer division 180/(nuang-1) (:175)
 *   - output order borrp borip borrpe boripe forrp forip forpe foripe (:2515-2520); the e15.7 rounding and the
 *     duplicate last particle in fort.8 belong to the file layer (tests / shim), not to this numerical core.
 *
 * PARITY PINNING: the Fortran cannot be compiled in the build container (no Fortran compiler) and the reference
 * ships no sample output, so this oracle is pinned to physics instead: coated-sphere Mie (Aden-Kerker), the
 * single-layer oracle for core == shell, and reciprocity-type identities (tests/test_tm2l_cpu.py).  "parity
 * unpinned" against tmatrix_py.f itself.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "patterson31.h"

#define NRANK 17
#define NRANKI 18
#define NM 7
#define NN (2 * NRANK)
#define NNODE 31

typedef double complex cplx;

/* the Fortran's real-pair helpers, same formulas (tmatrix_py.f:1672-1746, 2153-2175) */
static inline cplx cmul(cplx a, cplx b)
{
    return CMPLX(creal(a) * creal(b) - cimag(a) * cimag(b), cimag(a) * creal(b) + creal(a) * cimag(b));
}
static inline cplx cdiv(cplx a, cplx b)
{
    double e = creal(b) * creal(b) + cimag(b) * cimag(b);
    return CMPLX((creal(a) * creal(b) + cimag(a) * cimag(b)) / e, (cimag(a) * creal(b) - creal(a) * cimag(b)) / e);
}
static inline cplx crs(cplx a, double s) { return CMPLX(creal(a) * s, cimag(a) * s); }
static double cdabx(double a, double b)
{
    if (a != 0.0) {
        if (b != 0.0) {
            double e = a > b ? a : b, f = a > b ? b : a, g = f / e;
            return fabs(e) * sqrt(1.0 + g * g);
        }
        return fabs(a);
    }
    return b != 0.0 ? fabs(b) : 0.0;
}
static cplx croot(cplx z)
{
    double a = creal(z), b = cimag(z);
    double dmag = pow(a * a + b * b, 0.25), angle = 0.5 * atan2(b, a);
    return CMPLX(dmag * cos(angle), dmag * sin(angle));
}

/* ---- genlgp (:1597-1670): pn[k] = pnmllg(k+1) ~ P_k^m(cos)/sin, k = 0..nrank ---- */
static void genlgp(double theta, double sinth, double costh, int kmv, double prodm, double *pn)
{
    double twm = 2.0 * kmv, dtwm = twm + 1.0, pla, plb;
    int ibeg;
    if (sinth == 0.0 && kmv == 0) { for (int i = 0; i < NRANKI; i++) pn[i] = 0.0; return; }
    if (theta == 0.0) {
        if (kmv != 1) { for (int i = 0; i < NRANKI; i++) pn[i] = 0.0; return; }
        pn[0] = 0.0; pn[1] = 1.0; pla = 1.0;
        plb = dtwm * costh * pla; pn[kmv + 1] = plb; ibeg = kmv + 3;
    } else if (kmv <= 0) {
        pla = 1.0 / sinth; plb = costh * pla;
        pn[0] = pla; pn[1] = plb; ibeg = 3;
    } else {
        for (int i = 0; i < kmv; i++) pn[i] = 0.0;
        if (sinth == 0.0 && kmv == 1) pla = 0.0;
        else pla = prodm * pow(sinth, (double)(kmv - 1));
        pn[kmv] = pla;
        plb = dtwm * costh * pla; pn[kmv + 1] = plb; ibeg = kmv + 3;
    }
    double cnmul = ibeg + ibeg - 3, cnm = 2.0, cnmm = dtwm;
    for (int ilgr = ibeg; ilgr <= NRANKI; ilgr++) {
        double plc = (cnmul * costh * plb - cnmm * pla) / cnm;
        pn[ilgr - 1] = plc;
        pla = plb; plb = plc;
        cnmul += 2.0; cnm += 1.0; cnmm += 1.0;
    }
}

/* ---- bessel (:1845-1881): power series for j_n(x) ---- */
static int bessel_series(int n, double x, double *answr)
{
    double sum = 1.0, apr = 1.0, topr = -0.5 * x * x, ci = 1.0, cni = (double)(2 * n) + 3.0, acr;
    int ok = 0;
    for (int i = 1; i <= 100; i++) {
        acr = topr * apr / (ci * cni);
        sum += acr;
        if (fabs(acr / sum) - 1.0e-20 <= 0.0) { ok = 1; break; }
        apr = acr; ci += 1.0; cni += 2.0;
    }
    if (!ok) return 1;
    double prod = (double)(2 * n) + 1.0, fact = 1.0;
    for (int k = 1; k <= n; k++) { fact = fact * x / prod; prod -= 2.0; }
    *answr = fact * sum;
    return 0;
}

/* ---- genbsl (:1748-1843): j_k, y_k (k = 0..nrank) for real argument ---- */
static void genbsl(double pckr, double *bj, double *by)
{
    int nval = NRANK - 1, got = 0;
    double ansa = 0, ansb = 0, answr;
    for (int i = 1; i <= 4 && !got; i++) {
        if (bessel_series(nval, pckr, &answr) == 0) {
            ansa = answr;
            nval += 1;
            if (bessel_series(nval, pckr, &answr) == 0) { ansb = answr; got = 1; break; }
            nval -= 1;
        }
        nval += NRANK;
    }
    if (nval - NRANK > 0) {
        int iend = nval - NRANK;
        double conn = 2 * (nval - 1) + 1.0;
        for (int ip = 1; ip <= iend; ip++) {
            double ansc = conn * ansa / pckr - ansb;
            conn -= 2.0; ansb = ansa; ansa = ansc;
        }
    }
    bj[NRANKI - 1] = ansb;
    bj[NRANKI - 2] = ansa;
    double conn = (double)(NRANK + NRANK - 1);
    int ie = NRANKI - 2;
    for (int jb = 1; jb <= NRANKI - 2; jb++) {
        double ansc = conn * ansa / pckr - ansb;
        bj[ie - 1] = ansc;
        ansb = ansa; ansa = ansc; ie--; conn -= 2.0;
    }
    double cskrx = cos(pckr) / pckr, snkrx = sin(pckr) / pckr, cmuln = 3.0;
    double snsa = -cskrx, snsb = -cskrx / pckr - snkrx;
    by[0] = snsa; by[1] = snsb;
    for (int i = 3; i <= NRANKI; i++) {
        double snsc = cmuln * snsb / pckr - snsa;
        by[i - 1] = snsc;
        snsa = snsb; snsb = snsc; cmuln += 2.0;
    }
}

/* ---- genbkr (:1883-2151): j_k, y_k (k = 0..nrank) for complex argument, downward recurrence with renormalisation ---- */
static void genbkr(cplx z, cplx *bs, cplx *cn)
{
    const double alarge = 1.0e30;
    double pckrr = creal(z), pckri = cimag(z);
    double rm = cdabx(pckrr, pckri);
    int nval = 50;
    if (rm > 25.0 && rm <= 150.0) nval = (int)(2 * rm);
    if (rm > 150.0) nval = 300;
    cplx rj[303];
    cplx pstore[21];
    int ipntr[21];
    memset(ipntr, 0, sizeof(ipntr));
    rj[nval + 1] = 0.0;
    rj[nval] = 1.0;
    if (rm > 2.0) rj[nval] = 1.0e-10;
    if (rm > 10.0) rj[nval] = 1.0e-20;
    if (rm > 25.0) rj[nval] = 1.0e-30;
    int ie = nval + 2;
    double eposx = exp(pckri), enegx = exp(-pckri);
    double ex1 = (enegx + eposx) / 2.0, ex2 = (enegx - eposx) / 2.0;
    cplx cs = CMPLX(sin(pckrr) * ex1, -cos(pckrr) * ex2);
    cplx ar = cdiv(cs, z);
    int k = 0;
    for (int i = 2; i <= nval; i++) {
        int ij = ie - i;
        double f = (double)(2 * ij - 1);
        cplx cf = cdiv(CMPLX(f, 0.0), z);
        rj[ij - 1] = cmul(rj[ij], cf) - rj[ij + 1];
        double temr = fabs(creal(rj[ij - 1])), temi = fabs(cimag(rj[ij - 1]));
        if (temr <= alarge && temi <= alarge) continue;
        int ii = ij + 1;
        if (ii > NRANKI) {
            rj[ij] = cdiv(rj[ij], rj[ij - 1]);
            rj[ij - 1] = 1.0;
            continue;
        }
        k++;
        if (k > 20) break; /* Fortran STOPs here */
        pstore[k] = rj[ij - 1];
        ipntr[k] = ij + 1;
        rj[ij - 1] = 1.0;
        rj[ij] = cdiv(rj[ij], pstore[k]);
    }
    cplx px = cdiv(ar, rj[1]);
    for (int i = 1; i <= NRANKI; i++) {
        if (k >= 1 && ipntr[k] == i) {
            px = cdiv(px, pstore[k]);
            k--;
        }
        bs[i - 1] = cmul(rj[i], px);
    }
    cplx cc = CMPLX(cos(pckrr) * ex1, sin(pckrr) * ex2);
    cn[0] = -cdiv(cc, z);
    cn[1] = cdiv(cn[0], z) - ar;
    for (int i = 3; i <= NRANKI; i++) {
        double f = (double)(2 * i - 3);
        cplx cf = cdiv(CMPLX(f, 0.0), z);
        cn[i - 1] = cmul(cn[i - 2], cf) - cn[i - 3];
    }
}

/* ---- prcssm (:2177-2280): b <- a^-1 b by Gauss-Jordan, pivot = largest element of the current row ---- */
static void prcssm(cplx a[NN][NN], cplx b[NN][NN])
{
    const int n = NN;
    int perm[NN]; /* perm[col] = row whose pivot sits in that column (the Fortran punns this into a(1,jmax,1)) */
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int i = 0; i < n; i++) {
        cplx amax = a[i][0];
        int jmax = 0;
        for (int j = 1; j < n; j++) {
            if (cdabx(creal(a[i][j]), cimag(a[i][j])) <= cdabx(creal(amax), cimag(amax))) continue;
            amax = a[i][j];
            jmax = j;
        }
        if (!(cdabx(creal(amax), cimag(amax)) > 0.0)) { perm[i] = i; continue; }
        for (int j = 0; j < n; j++) {
            a[i][j] = cdiv(a[i][j], amax);
            b[i][j] = cdiv(b[i][j], amax);
        }
        for (int k = 0; k < n; k++) {
            if (k == i) continue;
            cplx arat = -a[k][jmax];
            for (int j = 0; j < n; j++) {
                if (cdabx(creal(a[i][j]), cimag(a[i][j])) <= 0.0) continue;
                a[k][j] = cmul(arat, a[i][j]) + a[k][j];
            }
            a[k][jmax] = 0.0;
            for (int j = 0; j < n; j++) {
                if (cdabx(creal(b[i][j]), cimag(b[i][j])) <= 0.0) continue;
                b[k][j] = cmul(arat, b[i][j]) + b[k][j];
            }
        }
        perm[jmax] = i;
    }
    /* row interchanges as recorded (:2251-2277) */
    for (int i = 0; i < n; i++) {
        int k = i;
        do { k = perm[k]; } while (k < i);
        if (k > i)
            for (int j = 0; j < n; j++) { cplx t = b[i][j]; b[i][j] = b[k][j]; b[k][j] = t; }
    }
}

typedef struct {
    double theta, sinth, costh, ckr, dckr, ckr2, dckr2;
    double bj[NRANKI], by[NRANKI];             /* bsslsp, cneumn at ckr */
    cplx jk[NRANKI], yk[NRANKI], hk[NRANKI];   /* bslkp, cneum, hankl at sqrt(eps1) ckr  (iswt=1) */
    cplx j2[NRANKI], y2[NRANKI], h2[NRANKI];   /* bkp2, cnp2, hank2 at sqrt(eps1) ckr2   (iswt=2) */
    cplx j3[NRANKI];                           /* bkprr2 at sqrt(eps2) ckr2               (iswt=3) */
    double rbessl[NRANK + 1];
    cplx rhankl[NRANK + 1], rbskpr[NRANK + 1], rhskpr[NRANK + 1], rbess2[NRANK + 1], rhank2[NRANK + 1], rbskp2[NRANK + 1];
    cplx ckpr2, dkp2;
} node_t;

typedef struct {
    double cpi, dtr, rtd, conk, conk2, aovrb, aovrb2, sigma;
    cplx eps1, eps2, sq1, qs1, sq2, rat12, sqrat, qsrat;
} part_t;

/* gener (:792-1338): the 4 x 6 integrand values at one node for (irow, icol, m) */
static void gener(const part_t *P, const node_t *N, const double *pn, int irow, int icol, int kmv, double qem,
                  cplx t[4][6])
{
    const double b89 = 2.0;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 6; j++) t[i][j] = 0.0;
    double cmv = kmv, cm2 = cmv * cmv;
    double crow = irow, ccol = icol, crow1 = crow + 1.0, ccol1 = ccol + 1.0, crowm = cmv + crow, ccolm = cmv + ccol;
    double sinth = N->sinth, costh = N->costh, srmsin = sinth, ckr = N->ckr, dckr = N->dckr;
    cplx dcn = P->eps1, rat12 = P->rat12;
    double br = N->rbessl[irow];
    cplx h = N->rhankl[irow], h2 = N->rhank2[irow], b2 = N->rbess2[irow];
    double crij = crow + ccol, crssij = crow * ccol;
    double cmcrco = cm2 - qem * crssij * costh * costh;
    double pnr0c0 = pn[irow - 1] * pn[icol - 1], pnr0c1 = pn[irow - 1] * pn[icol];
    double pnr1c0 = pn[irow] * pn[icol - 1], pnr1c1 = pn[irow] * pn[icol];
    double b1a = crow * costh * pnr1c1 - crowm * pnr0c1;
    double b1b = ccol * costh * pnr1c1 - ccolm * pnr1c0;
    cplx bk = N->rbskpr[icol], hkk = N->rhskpr[icol];
    cplx hrow = CMPLX(N->bj[irow], N->by[irow]);
    cplx hbkml = cmul(hrow, N->jk[icol]);
    cplx bbkml = crs(N->jk[icol], N->bj[irow]);
    cplx hhkml = cmul(hrow, N->hk[icol]);
    cplx bhkml = crs(N->hk[icol], N->bj[irow]);
    cplx heps = cmul(P->qs1, hbkml), beps = cmul(P->qs1, bbkml), hheps = cmul(P->qs1, hhkml), bbeps = cmul(P->qs1, bhkml);
    cplx bk2 = N->rbskp2[icol];
    cplx hb2ml = cmul(N->h2[irow], N->j3[icol]);
    cplx bb2ml = cmul(N->j2[irow], N->j3[icol]);
    cplx heps2 = cmul(P->qsrat, hb2ml), beps2 = cmul(P->qsrat, bb2ml);
    cplx ckp2 = N->ckpr2, dkp2 = N->dkp2;

    if (((irow + icol) & 1) == 1) {
        if (kmv == 0) return;
        /* I and L sub-blocks */
        double b1 = b1a + b1b;
        cplx htbk = cmul(h, bk), btbk = crs(bk, br), bthk = crs(hkk, br), hthk = cmul(h, hkk);
        double fac = b89 * cmv * srmsin;
        /* a: (rho = h, chi = bk) */
        cplx suma = crs(crs(bk, crow * crow1) + crs(h, ccol * ccol1) - crssij * (crij + 2.0) / ckr, dckr * sinth * pnr1c1);
        cplx sum = crs(crs(1.0 + htbk, ckr) - crs(h, ccol) - crs(bk, crow) + crssij / ckr, b1 * ckr) + suma;
        t[0][0] = crs(cmul(sum, hbkml), fac);
        cplx sumb = crs(crs(bk, crow * crow1) + ccol * ccol1 * br - crssij * (crij + 2.0) / ckr, dckr * sinth * pnr1c1);
        sum = crs(crs(1.0 + btbk, ckr) - ccol * br - crs(bk, crow) + crssij / ckr, b1 * ckr) + sumb;
        t[0][1] = crs(cmul(sum, bbkml), fac);
        cplx sumc = crs(crs(hkk, crow * crow1) + ccol * ccol1 * br - crssij * (crij + 2.0) / ckr, dckr * sinth * pnr1c1);
        sum = crs(crs(1.0 + bthk, ckr) - ccol * br - crs(hkk, crow) + crssij / ckr, b1 * ckr) + sumc;
        t[0][2] = crs(cmul(sum, bhkml), fac);
        cplx sumd = crs(crs(hkk, crow * crow1) + crs(h, ccol * ccol1) - crssij * (crij + 2.0) / ckr, dckr * sinth * pnr1c1);
        sum = crs(crs(1.0 + hthk, ckr) - crs(h, ccol) - crs(hkk, crow) + crssij / ckr, b1 * ckr) + sumd;
        t[0][3] = crs(cmul(sum, hhkml), fac);
        /* x, y (inner surface) */
        cplx htbk2 = cmul(h2, bk2), btbk2 = cmul(b2, bk2);
        cplx tinv = cdiv(1.0, ckp2);
        cplx sumx = crs(crs(bk2, crow * crow1) + crs(h2, ccol * ccol1) - crs(tinv, crssij * (crij + 2.0)), pnr1c1 * sinth);
        cplx sumg = cmul(sumx, dkp2);
        cplx s1 = crs(cmul(1.0 + htbk2, ckp2) - crs(h2, ccol) - crs(bk2, crow) + crs(tinv, crssij), b1);
        cplx srx = sumg + cmul(s1, ckp2);
        t[0][4] = crs(cmul(srx, hb2ml), fac);
        cplx sumy = crs(crs(bk2, crow * crow1) + crs(b2, ccol * ccol1) - crs(tinv, crssij * (crij + 2.0)), pnr1c1 * sinth);
        cplx sumh = cmul(sumy, dkp2);
        s1 = crs(cmul(1.0 + btbk2, ckp2) - crs(b2, ccol) - crs(bk2, crow) + crs(tinv, crssij), b1);
        srx = sumh + cmul(s1, ckp2);
        t[0][5] = crs(cmul(srx, bb2ml), fac);
        /* L */
        sum = crs(crs(dcn + htbk, ckr) - crs(h, ccol) - crs(bk, crow) + crssij / ckr, b1 * ckr) + suma;
        t[1][0] = -crs(cmul(sum, heps), fac);
        sum = crs(crs(dcn + btbk, ckr) - ccol * br - crs(bk, crow) + crssij / ckr, b1 * ckr) + sumb;
        t[1][1] = -crs(cmul(sum, beps), fac);
        sum = crs(crs(dcn + bthk, ckr) - ccol * br - crs(hkk, crow) + crssij / ckr, b1 * ckr) + sumc;
        t[1][2] = -crs(cmul(sum, bbeps), fac);
        sum = crs(crs(dcn + hthk, ckr) - crs(h, ccol) - crs(hkk, crow) + crssij / ckr, b1 * ckr) + sumd;
        t[1][3] = -crs(cmul(sum, hheps), fac);
        s1 = crs(cmul(rat12 + htbk2, ckp2) - crs(h2, ccol) - crs(bk2, crow) + crs(tinv, crssij), b1);
        srx = cmul(s1, ckp2) + sumg;
        t[1][4] = -crs(cmul(srx, heps2), fac);
        s1 = crs(cmul(rat12 + btbk2, ckp2) - crs(b2, ccol) - crs(bk2, crow) + crs(tinv, crssij), b1);
        srx = cmul(s1, ckp2) + sumh;
        t[1][5] = -crs(cmul(srx, beps2), fac);
        return;
    }
    /* J and K sub-blocks (irow + icol even) */
    double a12 = cmcrco * pnr1c1 + qem * (crow * ccolm * costh * pnr1c0 + ccol * crowm * costh * pnr0c1 - crowm * ccolm * pnr0c0);
    b1a = ccol * ccol1 * b1a;
    b1b = crow * crow1 * b1b;
    double b1 = (b1a - b1b) * sinth;
    double dd = -qem * dckr;
    double fac = b89 * srmsin;
    cplx tail = crs(b1a - crs(dcn, b1b), sinth * dd);
    cplx c_ = cmul(dcn, h);
    cplx sum = crs(crs(bk - c_, ckr) + crs(dcn, crow) - ccol, a12 * ckr) + tail;
    t[2][0] = crs(cmul(sum, heps), fac);
    c_ = crs(dcn, br);
    sum = crs(crs(bk - c_, ckr) + crs(dcn, crow) - ccol, a12 * ckr) + tail;
    t[2][1] = crs(cmul(sum, beps), fac);
    sum = crs(crs(hkk - c_, ckr) + crs(dcn, crow) - ccol, a12 * ckr) + tail;
    t[2][2] = crs(cmul(sum, bbeps), fac);
    c_ = cmul(dcn, h);
    sum = crs(crs(hkk - c_, ckr) + crs(dcn, crow) - ccol, a12 * ckr) + tail;
    t[2][3] = crs(cmul(sum, hheps), fac);
    cplx d2 = crs(dkp2, -qem);
    cplx c2 = cmul(rat12, h2);
    cplx s1 = crs(cmul(bk2 - c2, ckp2) + crs(rat12, crow) - ccol, a12);
    cplx sx = cmul(crs(b1a - crs(rat12, b1b), sinth), d2);
    cplx srx = cmul(s1, ckp2) + sx;
    t[2][4] = crs(cmul(srx, heps2), fac);
    c2 = cmul(rat12, b2);
    s1 = crs(cmul(bk2 - c2, ckp2) + crs(rat12, crow) - ccol, a12);
    srx = cmul(s1, ckp2) + sx;
    t[2][5] = crs(cmul(srx, beps2), fac);
    /* K */
    sum = crs(crs(bk - h, ckr) + crow - ccol, a12 * ckr) + b1 * dd;
    t[3][0] = crs(cmul(sum, hbkml), fac);
    sum = crs(crs(bk - br, ckr) + crow - ccol, a12 * ckr) + b1 * dd;
    t[3][1] = crs(cmul(sum, bbkml), fac);
    sum = crs(crs(hkk - br, ckr) + crow - ccol, a12 * ckr) + b1 * dd;
    t[3][2] = crs(cmul(sum, bhkml), fac);
    sum = crs(crs(hkk - h, ckr) + crow - ccol, a12 * ckr) + b1 * dd;
    t[3][3] = crs(cmul(sum, hhkml), fac);
    s1 = crs(cmul(bk2 - h2, ckp2) + crow - ccol, a12);
    cplx sxk = crs(d2, b1);
    srx = cmul(s1, ckp2) + sxk;
    t[3][4] = crs(cmul(srx, hb2ml), fac);
    s1 = crs(cmul(bk2 - b2, ckp2) + crow - ccol, a12);
    srx = cmul(s1, ckp2) + sxk;
    t[3][5] = crs(cmul(srx, bb2ml), fac);
}

static void setup_node(const part_t *P, node_t *N, double theta)
{
    N->theta = theta;
    N->costh = cos(theta);
    N->sinth = sin(theta);
    /* genkr / genkr2 (:2526-2584) */
    double bovra = 1.0 / P->aovrb;
    double qb = 1.0 / sqrt((bovra * N->costh) * (bovra * N->costh) + N->sinth * N->sinth);
    N->ckr = P->conk * qb;
    N->dckr = P->conk * N->costh * N->sinth * (bovra * bovra - 1.0) * qb * qb * qb;
    double bovra2 = 1.0 / P->aovrb2;
    double qb2 = 1.0 / sqrt((bovra2 * N->costh) * (bovra2 * N->costh) + N->sinth * N->sinth);
    N->ckr2 = P->conk2 * qb2;
    N->dckr2 = P->conk2 * N->costh * N->sinth * (bovra2 * bovra2 - 1.0) * qb2 * qb2 * qb2;
    genbsl(N->ckr, N->bj, N->by);
    genbkr(crs(P->sq1, N->ckr), N->jk, N->yk);
    N->ckpr2 = crs(P->sq1, N->ckr2);
    N->dkp2 = crs(P->sq1, N->dckr2);
    genbkr(N->ckpr2, N->j2, N->y2);
    cplx dummy[NRANKI];
    genbkr(crs(P->sq2, N->ckr2), N->j3, dummy);
    for (int k = 0; k < NRANKI; k++) {
        N->hk[k] = CMPLX(creal(N->jk[k]) - cimag(N->yk[k]), cimag(N->jk[k]) + creal(N->yk[k]));
        N->h2[k] = CMPLX(creal(N->j2[k]) - cimag(N->y2[k]), cimag(N->j2[k]) + creal(N->y2[k]));
    }
    for (int k = 1; k <= NRANK; k++) {
        N->rbessl[k] = N->bj[k - 1] / N->bj[k];
        N->rhankl[k] = cdiv(CMPLX(N->bj[k - 1], N->by[k - 1]), CMPLX(N->bj[k], N->by[k]));
        N->rbskpr[k] = cmul(P->sq1, cdiv(N->jk[k - 1], N->jk[k]));
        N->rhskpr[k] = cmul(P->sq1, cdiv(N->hk[k - 1], N->hk[k]));
        N->rbess2[k] = cdiv(N->j2[k - 1], N->j2[k]);
        N->rhank2[k] = cdiv(N->h2[k - 1], N->h2[k]);
        N->rbskp2[k] = cmul(P->sqrat, cdiv(N->j3[k - 1], N->j3[k]));
    }
}

/* One particle (driver :1-473 + rddata :476-790 + addprc :2282-2523).
 * in : dpart, dcore [cm], aovrb, aovrb2, eps_out (re,im), eps_in (re,im), wavelength [cm]
 * out: amp[8] = borrp borip borrpe boripe forrp forip forpe foripe  [m]  (full double precision, not yet e15.7) */
void tm2l_particle(double dpart, double dcore, double aovrb, double aovrb2, double eor, double eoi, double eir,
                   double eii, double alamd, double *amp)
{
    part_t P;
    P.cpi = (double)(4.0f * atanf(1.0f));
    P.dtr = P.cpi / 180.0;
    P.rtd = 180.0 / P.cpi;
    const double third = (double)(1.f / 3.f);
    P.aovrb = aovrb; P.aovrb2 = aovrb2;
    P.conk = P.cpi * dpart / (alamd * pow(aovrb, third));
    P.conk2 = P.cpi * dcore / (alamd * pow(aovrb2, third));
    P.sigma = 2.0 * alamd / (P.cpi * 100.0);
    P.eps1 = CMPLX(eor, eoi);
    P.eps2 = CMPLX(eir, eii);
    P.sq1 = croot(P.eps1);
    P.qs1 = cdiv(1.0, P.sq1);
    P.sq2 = croot(P.eps2);
    P.rat12 = cdiv(P.eps2, P.eps1);
    P.sqrat = croot(P.rat12);
    P.qsrat = cdiv(1.0, P.sqrat);
    const double anginc = 90.0, rtsfct = 8.0 / P.conk;
    double uang[2] = {anginc, anginc + 180.0};

    /* quadrature nodes on [0, cpi/2] in evaluation order: centre, then (+x_j, -x_j) pairs */
    node_t *ND = (node_t *)malloc(sizeof(node_t) * NNODE);
    double a0 = 0.0, b0 = P.cpi / 2.0, sum = (b0 + a0) / 2.0, diff = (b0 - a0) / 2.0;
    setup_node(&P, &ND[0], sum);
    for (int j = 0; j < 15; j++) {
        double x = PAT31_X[j] * diff;
        setup_node(&P, &ND[1 + 2 * j], sum + x);
        setup_node(&P, &ND[2 + 2 * j], sum - x);
    }
    cplx(*A)[NN] = malloc(sizeof(cplx) * NN * NN), (*B)[NN] = malloc(sizeof(cplx) * NN * NN);
    cplx(*C)[NN] = malloc(sizeof(cplx) * NN * NN), (*D)[NN] = malloc(sizeof(cplx) * NN * NN);
    cplx(*X)[NN] = malloc(sizeof(cplx) * NN * NN), (*Y)[NN] = malloc(sizeof(cplx) * NN * NN);
    cplx(*M[6])[NN] = {A, B, C, D, X, Y};
    cplx acans[2][2] = {{0, 0}, {0, 0}};
    double cmxnrm[NRANK + 1];
    double pnn[NNODE][NRANKI];

    for (int im = 1; im <= NM; im++) {
        int kmv = im - 1;
        double cmv = kmv, prodm = 1.0, em;
        if (kmv > 0) {
            em = 2.0;
            double quanm = cmv;
            for (int i = 1; i <= kmv; i++) { quanm += 1.0; prodm = quanm * prodm / 2.0; }
        } else em = 1.0;
        double qem = -2.0 / em;
        for (int q = 0; q < 6; q++) memset(M[q], 0, sizeof(cplx) * NN * NN);
        for (int t = 0; t < NNODE; t++) genlgp(ND[t].theta, ND[t].sinth, ND[t].costh, kmv, prodm, pnn[t]);
        for (int irow = 1; irow <= NRANK; irow++) {
            int irow1 = irow + NRANK;
            for (int icol = 1; icol <= NRANK; icol++) {
                int icol1 = icol + NRANK;
                cplx res[4][6], fz[4][6], f[4][6], tt[4][6], acum[4][6];
                gener(&P, &ND[0], pnn[0], irow, icol, kmv, qem, fz);
                memset(acum, 0, sizeof(acum));
                for (int j = 0; j < 15; j++) {
                    gener(&P, &ND[1 + 2 * j], pnn[1 + 2 * j], irow, icol, kmv, qem, f);
                    gener(&P, &ND[2 + 2 * j], pnn[2 + 2 * j], irow, icol, kmv, qem, tt);
                    for (int ii = 0; ii < 4; ii++)
                        for (int jj = 0; jj < 6; jj++) {
                            cplx fsum = f[ii][jj] + tt[ii][jj];
                            acum[ii][jj] = acum[ii][jj] + crs(fsum, PAT31_W[j]);
                        }
                }
                for (int ii = 0; ii < 4; ii++)
                    for (int jj = 0; jj < 6; jj++) res[ii][jj] = crs(acum[ii][jj] + crs(fz[ii][jj], PAT31_WC), diff);
                for (int jj = 0; jj < 6; jj++) {
                    M[jj][icol - 1][irow1 - 1] += res[0][jj];
                    M[jj][icol1 - 1][irow - 1] += res[1][jj];
                    M[jj][icol1 - 1][irow1 - 1] += res[2][jj];
                    M[jj][icol - 1][irow - 1] += res[3][jj];
                }
            }
            double ckrow = irow, fctki = 1.0;
            if (kmv > 0) {
                if (irow < kmv) { cmxnrm[irow] = 1.0; continue; }
                int ibfct = irow - kmv + 1, iefct = irow + kmv;
                double fprod = ibfct;
                for (int l = ibfct; l <= iefct; l++) { fctki *= fprod; fprod += 1.0; }
            }
            cmxnrm[irow] = 4.0 * ckrow * (ckrow + 1.0) * fctki / (em * (2.0 * ckrow + 1.0));
        }
        prcssm(X, Y); /* T2 = x^-1 y, stored in Y */
        for (int l = 0; l < NRANK; l++)
            for (int k = 0; k < NRANK; k++) {
                double scale = cmxnrm[l + 1] / cmxnrm[k + 1];
                Y[l][k] = crs(Y[l][k], scale);
                Y[l][k + NRANK] = crs(Y[l][k + NRANK], scale);
                Y[l + NRANK][k] = crs(Y[l + NRANK][k], scale);
                Y[l + NRANK][k + NRANK] = crs(Y[l + NRANK][k + NRANK], scale);
            }
        for (int l = 0; l < NN; l++)
            for (int k = 0; k < NN; k++) {
                cplx s = 0.0;
                for (int j = 0; j < NN; j++) s = cmul(Y[l][j], C[j][k]) + s;
                X[l][k] = B[l][k] - s;
            }
        for (int l = 0; l < NN; l++)
            for (int k = 0; k < NN; k++) {
                cplx s = 0.0;
                for (int j = 0; j < NN; j++) s = cmul(Y[l][j], D[j][k]) + s;
                B[l][k] = A[l][k] - s;
            }
        prcssm(B, X); /* T = b^-1 x, stored in X */

        /* addprc (:2282-2523) */
        double pn[NRANKI];
        double theta = P.dtr * anginc, sinth = sin(theta), costh = cos(theta);
        genlgp(theta, sinth, costh, kmv, prodm, pn);
        cplx ab1[NN], ab2[NN], fg1[NN], fg2[NN];
        static const cplx ipow[4] = {1.0, I, -1.0, -I};
        for (int n = 1; n <= NRANK; n++) {
            int np = n + NRANK;
            double cn = n;
            cplx ci1 = ipow[n & 3], ci2 = ipow[(n + 1) & 3];
            double p1 = cn * costh * pn[n] - (cn + cmv) * pn[n - 1];
            double p2 = cmv * pn[n];
            ab1[n - 1] = -crs(ci1, p2);
            ab1[np - 1] = crs(ci2, p1);
            ab2[n - 1] = crs(ci1, p1);
            ab2[np - 1] = -crs(ci2, p2);
        }
        for (int j = 0; j < NN; j++) {
            cplx s1 = 0.0, s2 = 0.0;
            for (int i = 0; i < NN; i++) {
                s1 = s1 + cmul(ab1[i], X[i][j]);
                s2 = s2 + cmul(ab2[i], X[i][j]);
            }
            fg1[j] = s1; fg2[j] = s2;
        }
        for (int iu = 0; iu < 2; iu++) {
            theta = P.dtr * uang[iu];
            sinth = sin(theta);
            costh = cos(theta);
            genlgp(theta, sinth, costh, kmv, prodm, pn);
            cplx fa1 = 0.0, fa2 = 0.0;
            for (int n = 1; n <= NRANK; n++) {
                int np = n + NRANK;
                double cn = n;
                double p1 = cn * costh * pn[n] - (cn + cmv) * pn[n - 1];
                double p2 = cmv * pn[n];
                cplx cim = ipow[(4 - ((n + 1) & 3)) & 3]; /* (-i)^(n+1) */
                cplx f1 = crs(fg1[n - 1], p2);
                cplx g1 = CMPLX(-cimag(fg1[np - 1]) * p1, creal(fg1[np - 1]) * p1);
                fa1 = fa1 + cmul(cim, f1 + g1) / cmxnrm[n];
                cplx f2 = crs(fg2[n - 1], p1);
                cplx g2 = CMPLX(-cimag(fg2[np - 1]) * p2, creal(fg2[np - 1]) * p2);
                fa2 = fa2 - cmul(cim, f2 + g2) / cmxnrm[n];
            }
            acans[iu][0] += crs(fa1, rtsfct);
            acans[iu][1] += crs(fa2, rtsfct);
        }
    }
    /* far-field amplitudes: index 0 = uang(1) (forward), 1 = uang(nuang) (back) */
    amp[0] = P.sigma * creal(acans[1][0]) / rtsfct; amp[1] = P.sigma * cimag(acans[1][0]) / rtsfct;
    amp[2] = P.sigma * creal(acans[1][1]) / rtsfct; amp[3] = P.sigma * cimag(acans[1][1]) / rtsfct;
    amp[4] = P.sigma * creal(acans[0][0]) / rtsfct; amp[5] = P.sigma * cimag(acans[0][0]) / rtsfct;
    amp[6] = P.sigma * creal(acans[0][1]) / rtsfct; amp[7] = P.sigma * cimag(acans[0][1]) / rtsfct;
    free(A); free(B); free(C); free(D); free(X); free(Y); free(ND);
}

void tm2l_batch(int n, const double *p /* [n][8] */, double alamd, double *amp /* [n][8] */)
{
    for (int i = 0; i < n; i++)
        tm2l_particle(p[8 * i], p[8 * i + 1], p[8 * i + 2], p[8 * i + 3], p[8 * i + 4], p[8 * i + 5], p[8 * i + 6],
                      p[8 * i + 7], alamd, amp + 8 * i);
}

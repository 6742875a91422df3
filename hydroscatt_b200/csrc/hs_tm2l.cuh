// hs_tm2l.cuh -- two-layer (coated spheroid) EBCM T-matrix, one particle per CTA, everything resident in shared memory.
//
// Replaces the reference's Fortran program hydroscatt/f_2l_tmatrix/tmatrix_py.f (f2py module tm2l; rows F1-F11 of
// SURVEY.md 8(a)): rddata (:476-790), quad (:1340-1595), gener (:792-1338), genlgp (:1597-1670), genbsl/bessel
// (:1748-1881), genbkr (:1883-2151), prcssm (:2177-2280), addprc (:2282-2523), genkr/genkr2 (:2526-2584).
// Same fixed truncation (nrank 17, modes m = 0..6, 31-node Patterson rule on [0, cpi/2]) and the same numerical
// quirks (single-precision pi, single-precision 1/3 exponent, uang = 90/270 deg).  B200-first restructuring:
//   * radial functions and their ratios are computed once per quadrature node (the Fortran recomputes the ratio
//     loops for each of the 289 (irow, icol) pairs) and staged in shared memory as SoA tables;
//   * mirror symmetry splits every 34 x 34 matrix into two independent 17 x 17 parity-class matrices, so the two
//     Gauss-Jordan solves and the two matrix products per mode run on order-17 systems;
//   * rows/columns with n < m (identically zero) are left out instead of being skipped inside the solver.
#pragma once
#include "hs_compat.cuh"
#include "hs_tmat.cuh"
#include "hs_patterson31.h"

namespace hs {

#define T2_NRANK 17
#define T2_NRANKI 18
#define T2_NM 7
#define T2_NNODE 31

struct cd {
    double x, y;
};
__device__ __forceinline__ cd mk(double x, double y) { cd r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ cd operator+(cd a, cd b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd operator-(cd a, cd b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd operator-(cd a) { return mk(-a.x, -a.y); }
__device__ __forceinline__ cd operator*(cd a, cd b) { return mk(a.x * b.x - a.y * b.y, a.y * b.x + a.x * b.y); }
__device__ __forceinline__ cd operator*(cd a, double s) { return mk(a.x * s, a.y * s); }
__device__ __forceinline__ cd operator*(double s, cd a) { return mk(a.x * s, a.y * s); }
__device__ __forceinline__ cd operator+(cd a, double s) { return mk(a.x + s, a.y); }
__device__ __forceinline__ cd operator-(cd a, double s) { return mk(a.x - s, a.y); }
__device__ __forceinline__ cd operator/(cd a, cd b)
{
    double e = b.x * b.x + b.y * b.y;
    return mk((a.x * b.x + a.y * b.y) / e, (a.y * b.x - a.x * b.y) / e);
}
__device__ __forceinline__ double cdabx(double a, double b)
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
__device__ inline cd croot(cd z)
{
    double dmag = pow(z.x * z.x + z.y * z.y, 0.25), angle = 0.5 * atan2(z.y, z.x);
    double s, c;
    sincos(angle, &s, &c);
    return mk(dmag * c, dmag * s);
}

// ---- bessel + genbsl: j_k, y_k (k = 0..17) for a real argument ----
__device__ inline int t2_bessel_series(int n, double x, double *answr)
{
    double sum = 1.0, apr = 1.0, topr = -0.5 * x * x, ci = 1.0, cni = (double)(2 * n) + 3.0, acr;
    bool ok = false;
    for (int i = 1; i <= 100; i++) {
        acr = topr * apr / (ci * cni);
        sum += acr;
        if (fabs(acr / sum) - 1.0e-20 <= 0.0) { ok = true; break; }
        apr = acr; ci += 1.0; cni += 2.0;
    }
    if (!ok) return 1;
    double prod = (double)(2 * n) + 1.0, fact = 1.0;
    for (int k = 1; k <= n; k++) { fact = fact * x / prod; prod -= 2.0; }
    *answr = fact * sum;
    return 0;
}

__device__ inline void t2_genbsl(double pckr, double *bj, double *by)
{
    int nval = T2_NRANK - 1;
    double ansa = 0, ansb = 0, answr = 0;
    for (int i = 1; i <= 4; i++) {
        if (t2_bessel_series(nval, pckr, &answr) == 0) {
            ansa = answr;
            nval += 1;
            if (t2_bessel_series(nval, pckr, &answr) == 0) { ansb = answr; break; }
            nval -= 1;
        }
        nval += T2_NRANK;
    }
    if (nval - T2_NRANK > 0) {
        int iend = nval - T2_NRANK;
        double conn = 2 * (nval - 1) + 1.0;
        for (int ip = 1; ip <= iend; ip++) {
            double ansc = conn * ansa / pckr - ansb;
            conn -= 2.0; ansb = ansa; ansa = ansc;
        }
    }
    bj[T2_NRANKI - 1] = ansb;
    bj[T2_NRANKI - 2] = ansa;
    double conn = (double)(T2_NRANK + T2_NRANK - 1);
    int ie = T2_NRANKI - 2;
    for (int jb = 1; jb <= T2_NRANKI - 2; jb++) {
        double ansc = conn * ansa / pckr - ansb;
        bj[ie - 1] = ansc;
        ansb = ansa; ansa = ansc; ie--; conn -= 2.0;
    }
    double sn, cs;
    sincos(pckr, &sn, &cs);
    double cskrx = cs / pckr, snkrx = sn / pckr, cmuln = 3.0;
    double snsa = -cskrx, snsb = -cskrx / pckr - snkrx;
    by[0] = snsa; by[1] = snsb;
    for (int i = 3; i <= T2_NRANKI; i++) {
        double snsc = cmuln * snsb / pckr - snsa;
        by[i - 1] = snsc;
        snsa = snsb; snsb = snsc; cmuln += 2.0;
    }
}

// ---- genbkr: j_k (and y_k when want_y) for a complex argument; only the low orders are ever stored ----
__device__ inline void t2_genbkr(cd z, cd *bs, cd *cn, bool want_y)
{
    const double alarge = 1.0e30;
    double rm = cdabx(z.x, z.y);
    int nval = 50;
    if (rm > 25.0 && rm <= 150.0) nval = (int)(2 * rm);
    if (rm > 150.0) nval = 300;
    cd rjs[T2_NRANKI + 2];  // rj(1..nranki+1)
    cd pst[21];
    int ipn[21];
    double seed = 1.0;
    if (rm > 2.0) seed = 1.0e-10;
    if (rm > 10.0) seed = 1.0e-20;
    if (rm > 25.0) seed = 1.0e-30;
    cd a = mk(0.0, 0.0), b = mk(seed, 0.0);  // rj(ij+1), rj(ij)
    double eposx = exp(z.y), enegx = exp(-z.y);
    double ex1 = (enegx + eposx) / 2.0, ex2 = (enegx - eposx) / 2.0;
    double sn, cs;
    sincos(z.x, &sn, &cs);
    cd ar = mk(sn * ex1, -cs * ex2) / z;
    int k = 0;
    for (int ij = nval; ij >= 2; ij--) {
        double f = (double)(2 * ij - 1);
        cd cf = mk(f, 0.0) / z;
        cd c = b * cf - a;  // rj(ij-1)
        if (!(fabs(c.x) <= alarge && fabs(c.y) <= alarge)) {
            if (ij + 1 > T2_NRANKI) {
                b = b / c;
                c = mk(1.0, 0.0);
            } else {
                k++;
                if (k > 20) k = 20;
                pst[k] = c;
                ipn[k] = ij + 1;
                c = mk(1.0, 0.0);
                b = b / pst[k];
            }
        }
        if (ij <= T2_NRANKI + 1) rjs[ij - 1] = b;  // (possibly rescaled) rj(ij)
        a = b;
        b = c;
    }
    rjs[0] = b;  // rj(1)
    cd px = ar / rjs[0];
    for (int i = 1; i <= T2_NRANKI; i++) {
        if (k >= 1 && ipn[k] == i) { px = px / pst[k]; k--; }
        bs[i - 1] = rjs[i - 1] * px;
    }
    if (!want_y) return;
    cd cc = mk(cs * ex1, sn * ex2);
    cd c0 = -(cc / z);
    cd c1 = c0 / z - ar;
    cn[0] = c0; cn[1] = c1;
    for (int i = 3; i <= T2_NRANKI; i++) {
        double f = (double)(2 * i - 3);
        cd cf = mk(f, 0.0) / z;
        cd c2 = c1 * cf - c0;
        cn[i - 1] = c2;
        c0 = c1; c1 = c2;
    }
}

// ---- genlgp: pn[k] = pnmllg(k+1), k = 0..17 ----
__device__ inline void t2_genlgp(double theta, double sinth, double costh, int kmv, double prodm, double *pn)
{
    double dtwm = 2.0 * kmv + 1.0, pla, plb;
    int ibeg;
    if ((sinth == 0.0 && kmv == 0) || (theta == 0.0 && kmv != 1)) {
        for (int i = 0; i < T2_NRANKI; i++) pn[i] = 0.0;
        return;
    }
    if (theta == 0.0) {
        pn[0] = 0.0; pn[1] = 1.0; pla = 1.0;
        plb = dtwm * costh * pla; pn[kmv + 1] = plb; ibeg = kmv + 3;
    } else if (kmv <= 0) {
        pla = 1.0 / sinth; plb = costh * pla;
        pn[0] = pla; pn[1] = plb; ibeg = 3;
    } else {
        for (int i = 0; i < kmv; i++) pn[i] = 0.0;
        if (sinth == 0.0 && kmv == 1) pla = 0.0;
        else {
            pla = prodm;
            for (int e = 0; e < kmv - 1; e++) pla *= sinth;
        }
        pn[kmv] = pla;
        plb = dtwm * costh * pla; pn[kmv + 1] = plb; ibeg = kmv + 3;
    }
    double cnmul = ibeg + ibeg - 3, cnm = 2.0, cnmm = dtwm;
    for (int ilgr = ibeg; ilgr <= T2_NRANKI; ilgr++) {
        double plc = (cnmul * costh * plb - cnmm * pla) / cnm;
        pn[ilgr - 1] = plc;
        pla = plb; plb = plc;
        cnmul += 2.0; cnm += 1.0; cnmm += 1.0;
    }
}

// shared-memory layout of one particle (offsets in doubles)
struct T2Smem {
    // node scalars [31]
    double *sinth, *costh, *ckr, *dckr, *ckp2r, *ckp2i, *dkp2r, *dkp2i, *theta;
    // row tables [31][17] (index n-1)
    double *bj, *by, *j2r, *j2i, *h2r, *h2i, *rb, *rhr, *rhi, *rb2r, *rb2i, *rh2r, *rh2i;
    // column tables [31][17]
    double *jkr, *jki, *hkr, *hki, *j3r, *j3i, *rbkr, *rbki, *rhkr, *rhki, *rbk2r, *rbk2i;
    double *pn;       // [31][18]
    double2 *A, *B, *C, *D;  // [2][17][17] parity-class matrices of a, b, c, d
    double2 *aug;     // [2][17][34]: [x | y] then [b - T2 c ... ] see below
    double2 *aug2;    // [2][17][34]
};
#define T2_TAB (T2_NNODE * T2_NRANK)
__host__ __device__ inline int t2_smem_doubles()
{
    return 9 * T2_NNODE + 25 * T2_TAB + T2_NNODE * T2_NRANKI + 2 * (4 * 2 * 289 + 2 * 2 * 17 * 34);
}
__device__ inline void t2_carve(T2Smem &s, double *p)
{
    s.sinth = p; p += T2_NNODE; s.costh = p; p += T2_NNODE; s.ckr = p; p += T2_NNODE; s.dckr = p; p += T2_NNODE;
    s.ckp2r = p; p += T2_NNODE; s.ckp2i = p; p += T2_NNODE; s.dkp2r = p; p += T2_NNODE; s.dkp2i = p; p += T2_NNODE;
    s.theta = p; p += T2_NNODE;
    double **tabs[25] = {&s.bj, &s.by, &s.j2r, &s.j2i, &s.h2r, &s.h2i, &s.rb, &s.rhr, &s.rhi, &s.rb2r, &s.rb2i,
                         &s.rh2r, &s.rh2i, &s.jkr, &s.jki, &s.hkr, &s.hki, &s.j3r, &s.j3i, &s.rbkr, &s.rbki,
                         &s.rhkr, &s.rhki, &s.rbk2r, &s.rbk2i};
    for (int i = 0; i < 25; i++) { *tabs[i] = p; p += T2_TAB; }
    s.pn = p; p += T2_NNODE * T2_NRANKI;
    s.A = reinterpret_cast<double2 *>(p); p += 2 * 2 * 289;
    s.B = reinterpret_cast<double2 *>(p); p += 2 * 2 * 289;
    s.C = reinterpret_cast<double2 *>(p); p += 2 * 2 * 289;
    s.D = reinterpret_cast<double2 *>(p); p += 2 * 2 * 289;
    s.aug = reinterpret_cast<double2 *>(p); p += 2 * 2 * 17 * 34;
    s.aug2 = reinterpret_cast<double2 *>(p);
}

struct T2Args {
    int n;
    const double *dpart, *dcore, *aovrb, *aovrb2, *eor, *eoi, *eir, *eii;
    double alamd;
    double *amp;   // [n][8]
    int *status;   // [n]
};

__device__ __forceinline__ cd ld2(const double *re, const double *im, int i) { return mk(re[i], im[i]); }

// One particle per CTA.
__device__ void tm2l_particle(const T2Args &a, int p, double *smem, Ctl *ctl, double *s_cmx, double *s_acc,
                              const int tid, const int nt)
{
    T2Smem s;
    t2_carve(s, smem);
    const double cpi = 3.1415927410125732;  // 4.0*atan(1.0) in single precision (tmatrix_py.f:161)
    const double dtr = cpi / 180.0;
    const double third = 0.3333333432674408;  // the single-precision literal 1./3.
    const double dpart = a.dpart[p], dcore = a.dcore[p], aovrb = a.aovrb[p], aovrb2 = a.aovrb2[p], alamd = a.alamd;
    const double conk = cpi * dpart / (alamd * pow(aovrb, third));
    const double conk2 = cpi * dcore / (alamd * pow(aovrb2, third));
    const double sigma = 2.0 * alamd / (cpi * 100.0);
    const cd eps1 = mk(a.eor[p], a.eoi[p]), eps2 = mk(a.eir[p], a.eii[p]);
    const cd sq1 = croot(eps1), qs1 = mk(1.0, 0.0) / sq1, sq2 = croot(eps2);
    const cd rat12 = eps2 / eps1, sqrat = croot(rat12), qsrat = mk(1.0, 0.0) / sqrat;
    const double rtsfct = 8.0 / conk;
    const double diff = (cpi / 2.0) / 2.0, sum0 = (cpi / 2.0) / 2.0;

    // ---- K1: node geometry + radial functions (4 task kinds x 31 nodes) ----
    for (int t = tid; t < T2_NNODE; t += nt) {
        double theta;
        if (t == 0) theta = sum0;
        else {
            int j = (t - 1) >> 1;
            double x = PAT31_X[j] * diff;
            theta = ((t - 1) & 1) ? sum0 - x : sum0 + x;
        }
        double sn, cs;
        sincos(theta, &sn, &cs);
        s.theta[t] = theta; s.sinth[t] = sn; s.costh[t] = cs;
        double bovra = 1.0 / aovrb;
        double qb = 1.0 / sqrt((bovra * cs) * (bovra * cs) + sn * sn);
        s.ckr[t] = conk * qb;
        s.dckr[t] = conk * cs * sn * (bovra * bovra - 1.0) * qb * qb * qb;
        double bovra2 = 1.0 / aovrb2;
        double qb2 = 1.0 / sqrt((bovra2 * cs) * (bovra2 * cs) + sn * sn);
        double ckr2 = conk2 * qb2, dckr2 = conk2 * cs * sn * (bovra2 * bovra2 - 1.0) * qb2 * qb2 * qb2;
        s.ckp2r[t] = sq1.x * ckr2; s.ckp2i[t] = sq1.y * ckr2;
        s.dkp2r[t] = sq1.x * dckr2; s.dkp2i[t] = sq1.y * dckr2;
        // park kr2 for the iswt=3 task
        s.pn[t] = ckr2;
    }
    __syncthreads();
    for (int task = tid; task < 4 * T2_NNODE; task += nt) {
        const int kind = task / T2_NNODE, t = task - kind * T2_NNODE;
        const int o = t * T2_NRANK;
        if (kind == 0) {
            double bj[T2_NRANKI], by[T2_NRANKI];
            t2_genbsl(s.ckr[t], bj, by);
            for (int k = 1; k <= T2_NRANK; k++) {
                s.bj[o + k - 1] = bj[k]; s.by[o + k - 1] = by[k];
                s.rb[o + k - 1] = bj[k - 1] / bj[k];
                cd r = mk(bj[k - 1], by[k - 1]) / mk(bj[k], by[k]);
                s.rhr[o + k - 1] = r.x; s.rhi[o + k - 1] = r.y;
            }
        } else {
            cd z = (kind == 1) ? sq1 * s.ckr[t] : (kind == 2) ? mk(s.ckp2r[t], s.ckp2i[t]) : sq2 * s.pn[t];
            cd bs[T2_NRANKI], cn[T2_NRANKI];
            t2_genbkr(z, bs, cn, kind != 3);
            for (int k = 1; k <= T2_NRANK; k++) {
                cd jr = bs[k - 1] / bs[k];
                if (kind == 3) {
                    s.j3r[o + k - 1] = bs[k].x; s.j3i[o + k - 1] = bs[k].y;
                    cd r = sqrat * jr;
                    s.rbk2r[o + k - 1] = r.x; s.rbk2i[o + k - 1] = r.y;
                } else {
                    cd h0 = mk(bs[k - 1].x - cn[k - 1].y, bs[k - 1].y + cn[k - 1].x);
                    cd h1 = mk(bs[k].x - cn[k].y, bs[k].y + cn[k].x);
                    cd hr = h0 / h1;
                    if (kind == 1) {
                        s.jkr[o + k - 1] = bs[k].x; s.jki[o + k - 1] = bs[k].y;
                        s.hkr[o + k - 1] = h1.x; s.hki[o + k - 1] = h1.y;
                        cd r = sq1 * jr;
                        s.rbkr[o + k - 1] = r.x; s.rbki[o + k - 1] = r.y;
                        r = sq1 * hr;
                        s.rhkr[o + k - 1] = r.x; s.rhki[o + k - 1] = r.y;
                    } else {
                        s.j2r[o + k - 1] = bs[k].x; s.j2i[o + k - 1] = bs[k].y;
                        s.h2r[o + k - 1] = h1.x; s.h2i[o + k - 1] = h1.y;
                        s.rb2r[o + k - 1] = jr.x; s.rb2i[o + k - 1] = jr.y;
                        s.rh2r[o + k - 1] = hr.x; s.rh2i[o + k - 1] = hr.y;
                    }
                }
            }
        }
    }
    for (int w = tid; w < 4; w += nt) s_acc[w * 2] = s_acc[w * 2 + 1] = 0.0;  // acans: [fwd pol1, fwd pol2, back pol1, back pol2]
    __syncthreads();

    Grp<false> grp;
    grp.tid = tid; grp.nt = nt;
    Carve cv;

    for (int kmv = 0; kmv < T2_NM; kmv++) {
        const double cmv = kmv;
        double prodm = 1.0, em = 1.0;
        if (kmv > 0) {
            em = 2.0;
            double quanm = cmv;
            for (int i = 1; i <= kmv; i++) { quanm += 1.0; prodm = quanm * prodm / 2.0; }
        }
        const double qem = -2.0 / em, cm2 = cmv * cmv;
        const int nmin = kmv > 1 ? kmv : 1, NMq = T2_NRANK - nmin + 1, ld = NMq, lda = 2 * NMq;
        for (int t = tid; t < T2_NNODE; t += nt)
            t2_genlgp(s.theta[t], s.sinth[t], s.costh[t], kmv, prodm, s.pn + t * T2_NRANKI);
        for (int n = tid + 1; n <= T2_NRANK; n += nt) {
            double ck = n, fctki = 1.0, v;
            if (kmv > 0 && n < kmv) v = 1.0;
            else {
                if (kmv > 0) {
                    double fprod = n - kmv + 1;
                    for (int l = n - kmv + 1; l <= n + kmv; l++) { fctki *= fprod; fprod += 1.0; }
                }
                v = 4.0 * ck * (ck + 1.0) * fctki / (em * (2.0 * ck + 1.0));
            }
            s_cmx[n] = v;
        }
        __syncthreads();

        // ---- K2: assembly.  pair (n = irow, n' = icol); even pairs first (J,K), then odd pairs (I,L) ----
        const int h0 = (NMq + 1) >> 1, h1 = NMq >> 1;
        const int nE = h0 * h0 + h1 * h1, nO = (kmv == 0) ? 0 : 2 * h0 * h1;
        // warp-aligned start of the odd pairs so that no warp mixes the two code paths
        const int nEpad = (nE + 31) & ~31;
        for (int item = tid; item < nEpad + nO; item += nt) {
            int k1, k2;  // k1 = n - nmin (irow), k2 = n' - nmin (icol)
            bool even;
            if (item < nE) {
                even = true;
                if (item < h0 * h0) { int q = item / h0; k1 = 2 * q; k2 = 2 * (item - q * h0); }
                else { int u = item - h0 * h0; int q = u / h1; k1 = 2 * q + 1; k2 = 2 * (u - q * h1) + 1; }
            } else if (item >= nEpad) {
                even = false;
                int u = item - nEpad;
                if (u < h0 * h1) { int q = u / h1; k1 = 2 * q; k2 = 2 * (u - q * h1) + 1; }
                else { u -= h0 * h1; int q = u / h0; k1 = 2 * q + 1; k2 = 2 * (u - q * h0); }
            } else continue;
            const int irow = nmin + k1, icol = nmin + k2;
            const double crow = irow, ccol = icol, crow1 = crow + 1.0, ccol1 = ccol + 1.0;
            const double crowm = cmv + crow, ccolm = cmv + ccol, crij = crow + ccol, crssij = crow * ccol;
            cd acc[12];
#pragma unroll
            for (int q = 0; q < 12; q++) acc[q] = mk(0.0, 0.0);
            cd fz[12];
            for (int step = 0; step < T2_NNODE; step++) {
                // evaluation order of quad: pairs j=1..15 (+x, -x) accumulate first, centre (node 0) last
                const int t = (step < 30) ? step + 1 : 0;
                const double w = (step < 30) ? PAT31_W[step >> 1] : PAT31_WC;
                const int ro = t * T2_NRANK + irow - 1, co = t * T2_NRANK + icol - 1;
                const double sinth = s.sinth[t], costh = s.costh[t], ckr = s.ckr[t], dckr = s.dckr[t];
                const double *pn = s.pn + t * T2_NRANKI;
                const double pnr0c0 = pn[irow - 1] * pn[icol - 1], pnr0c1 = pn[irow - 1] * pn[icol];
                const double pnr1c0 = pn[irow] * pn[icol - 1], pnr1c1 = pn[irow] * pn[icol];
                double b1a = crow * costh * pnr1c1 - crowm * pnr0c1;
                double b1b = ccol * costh * pnr1c1 - ccolm * pnr1c0;
                const double br = s.rb[ro];
                const cd h = ld2(s.rhr, s.rhi, ro), h2 = ld2(s.rh2r, s.rh2i, ro), b2 = ld2(s.rb2r, s.rb2i, ro);
                const cd bk = ld2(s.rbkr, s.rbki, co), hk = ld2(s.rhkr, s.rhki, co), bk2 = ld2(s.rbk2r, s.rbk2i, co);
                const cd hrow = mk(s.bj[ro], s.by[ro]);
                const cd jkc = ld2(s.jkr, s.jki, co), hkc = ld2(s.hkr, s.hki, co), j3c = ld2(s.j3r, s.j3i, co);
                const cd hbkml = hrow * jkc, bbkml = jkc * s.bj[ro], hhkml = hrow * hkc, bhkml = hkc * s.bj[ro];
                const cd hb2ml = ld2(s.h2r, s.h2i, ro) * j3c, bb2ml = ld2(s.j2r, s.j2i, ro) * j3c;
                const cd ckp2 = mk(s.ckp2r[t], s.ckp2i[t]), dkp2 = mk(s.dkp2r[t], s.dkp2i[t]);
                cd v[12];
                if (!even) {
                    const double b1 = b1a + b1b;
                    const cd htbk = h * bk, btbk = bk * br, bthk = hk * br, hthk = h * hk;
                    const double fac = 2.0 * cmv * sinth;
                    const double gs = dckr * sinth * pnr1c1, tl = crssij * (crij + 2.0) / ckr, t0 = crssij / ckr;
                    const cd suma = (bk * (crow * crow1) + h * (ccol * ccol1) - tl) * gs;
                    const cd sumb = (bk * (crow * crow1) + (ccol * ccol1 * br - tl)) * gs;
                    const cd sumc = (hk * (crow * crow1) + (ccol * ccol1 * br - tl)) * gs;
                    const cd sumd = (hk * (crow * crow1) + h * (ccol * ccol1) - tl) * gs;
                    const cd ca = ((htbk + 1.0) * ckr - h * ccol - bk * crow + t0);
                    const cd cb = ((btbk + 1.0) * ckr - (ccol * br) - bk * crow + t0);
                    const cd cc = ((bthk + 1.0) * ckr - (ccol * br) - hk * crow + t0);
                    const cd cdd = ((hthk + 1.0) * ckr - h * ccol - hk * crow + t0);
                    const double bc = b1 * ckr;
                    v[0] = ((ca * bc + suma) * hbkml) * fac;
                    v[1] = ((cb * bc + sumb) * bbkml) * fac;
                    v[2] = ((cc * bc + sumc) * bhkml) * fac;
                    v[3] = ((cdd * bc + sumd) * hhkml) * fac;
                    const cd htbk2 = h2 * bk2, btbk2 = b2 * bk2;
                    const cd tinv = mk(1.0, 0.0) / ckp2;
                    const cd sumg = ((bk2 * (crow * crow1) + h2 * (ccol * ccol1) - tinv * (crssij * (crij + 2.0))) * (pnr1c1 * sinth)) * dkp2;
                    const cd sumh = ((bk2 * (crow * crow1) + b2 * (ccol * ccol1) - tinv * (crssij * (crij + 2.0))) * (pnr1c1 * sinth)) * dkp2;
                    const cd rx = h2 * ccol + bk2 * crow - tinv * crssij;
                    const cd ry = b2 * ccol + bk2 * crow - tinv * crssij;
                    v[4] = ((sumg + (((htbk2 + 1.0) * ckp2 - rx) * b1) * ckp2) * hb2ml) * fac;
                    v[5] = ((sumh + (((btbk2 + 1.0) * ckp2 - ry) * b1) * ckp2) * bb2ml) * fac;
                    // L
                    const cd e1k = eps1 * ckr;
                    v[6] = -((((ca + e1k - ckr) * bc + suma) * (qs1 * hbkml)) * fac);
                    v[7] = -((((cb + e1k - ckr) * bc + sumb) * (qs1 * bbkml)) * fac);
                    v[8] = -((((cc + e1k - ckr) * bc + sumc) * (qs1 * bhkml)) * fac);
                    v[9] = -((((cdd + e1k - ckr) * bc + sumd) * (qs1 * hhkml)) * fac);
                    v[10] = -(((((((rat12 + htbk2) * ckp2 - rx) * b1) * ckp2) + sumg) * (qsrat * hb2ml)) * fac);
                    v[11] = -(((((((rat12 + btbk2) * ckp2 - ry) * b1) * ckp2) + sumh) * (qsrat * bb2ml)) * fac);
                } else {
                    const double cmcrco = cm2 - qem * crssij * costh * costh;
                    const double a12 = cmcrco * pnr1c1 + qem * (crow * ccolm * costh * pnr1c0 + ccol * crowm * costh * pnr0c1 - crowm * ccolm * pnr0c0);
                    b1a = ccol * ccol1 * b1a;
                    b1b = crow * crow1 * b1b;
                    const double b1 = (b1a - b1b) * sinth, dd = -qem * dckr, fac = 2.0 * sinth;
                    const double ak = a12 * ckr;
                    const cd tail = (mk(b1a, 0.0) - eps1 * b1b) * (sinth * dd);
                    const cd e1h = eps1 * h, e1b = eps1 * br, e1r = eps1 * crow - ccol;
                    // J: a, b, c, d
                    v[0] = ((((bk - e1h) * ckr + e1r) * ak + tail) * (qs1 * hbkml)) * fac;
                    v[1] = ((((bk - e1b) * ckr + e1r) * ak + tail) * (qs1 * bbkml)) * fac;
                    v[2] = ((((hk - e1b) * ckr + e1r) * ak + tail) * (qs1 * bhkml)) * fac;
                    v[3] = ((((hk - e1h) * ckr + e1r) * ak + tail) * (qs1 * hhkml)) * fac;
                    const cd d2 = dkp2 * (-qem);
                    const cd sx = ((mk(b1a, 0.0) - rat12 * b1b) * sinth) * d2;
                    const cd r12r = rat12 * crow - ccol;
                    v[4] = ((((((bk2 - rat12 * h2) * ckp2 + r12r) * a12) * ckp2) + sx) * (qsrat * hb2ml)) * fac;
                    v[5] = ((((((bk2 - rat12 * b2) * ckp2 + r12r) * a12) * ckp2) + sx) * (qsrat * bb2ml)) * fac;
                    // K
                    const double rc = crow - ccol, bd = b1 * dd;
                    v[6] = ((((bk - h) * ckr + rc) * ak + bd) * hbkml) * fac;
                    v[7] = ((((bk - br) * ckr + rc) * ak + bd) * bbkml) * fac;
                    v[8] = ((((hk - br) * ckr + rc) * ak + bd) * bhkml) * fac;
                    v[9] = ((((hk - h) * ckr + rc) * ak + bd) * hhkml) * fac;
                    const cd sxk = d2 * b1;
                    v[10] = (((((bk2 - h2) * ckp2 + rc) * a12) * ckp2 + sxk) * hb2ml) * fac;
                    v[11] = (((((bk2 - b2) * ckp2 + rc) * a12) * ckp2 + sxk) * bb2ml) * fac;
                }
                if (step < 30) {
                    if ((step & 1) == 0) {
#pragma unroll
                        for (int q = 0; q < 12; q++) fz[q] = v[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < 12; q++) acc[q] = acc[q] + (fz[q] + v[q]) * w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 12; q++) acc[q] = (acc[q] + v[q] * w) * diff;
                }
            }
            // scatter into the parity-class matrices: first six -> class ca, last six -> class cb
            // even: [0..5] = J -> (2,n')-(2,n): class n'&1 ; [6..11] = K -> (1,n')-(1,n): class (n'+1)&1
            // odd : [0..5] = I -> (1,n')-(2,n): class (n'+1)&1 ; [6..11] = L -> (2,n')-(1,n): class n'&1
            const int c_first = even ? (icol & 1) : ((icol + 1) & 1), c_second = c_first ^ 1;
            const long e1 = (long)(c_first * NMq + k2) * ld + k1, e2 = (long)(c_second * NMq + k2) * ld + k1;
            const long g1 = (long)(c_first * NMq + k2) * lda + k1, g2 = (long)(c_second * NMq + k2) * lda + k1;
            s.A[e1] = make_double2(acc[0].x, acc[0].y); s.B[e1] = make_double2(acc[1].x, acc[1].y);
            s.C[e1] = make_double2(acc[2].x, acc[2].y); s.D[e1] = make_double2(acc[3].x, acc[3].y);
            s.aug[g1] = make_double2(acc[4].x, acc[4].y); s.aug[g1 + NMq] = make_double2(acc[5].x, acc[5].y);
            s.A[e2] = make_double2(acc[6].x, acc[6].y); s.B[e2] = make_double2(acc[7].x, acc[7].y);
            s.C[e2] = make_double2(acc[8].x, acc[8].y); s.D[e2] = make_double2(acc[9].x, acc[9].y);
            s.aug[g2] = make_double2(acc[10].x, acc[10].y); s.aug[g2 + NMq] = make_double2(acc[11].x, acc[11].y);
        }
        if (kmv == 0) {
            // odd (I, L) entries vanish for m = 0: zero them explicitly
            for (int item = tid; item < 2 * h0 * h1; item += nt) {
                int k1, k2, u = item;
                if (u < h0 * h1) { int q = u / h1; k1 = 2 * q; k2 = 2 * (u - q * h1) + 1; }
                else { u -= h0 * h1; int q = u / h0; k1 = 2 * q + 1; k2 = 2 * (u - q * h0); }
                const double2 z = make_double2(0.0, 0.0);
                for (int c = 0; c < 2; c++) {
                    const long e = (long)(c * NMq + k2) * ld + k1, g = (long)(c * NMq + k2) * lda + k1;
                    s.A[e] = z; s.B[e] = z; s.C[e] = z; s.D[e] = z; s.aug[g] = z; s.aug[g + NMq] = z;
                }
            }
        }
        __syncthreads();

        // ---- K3: T2 = x^-1 y (class-wise), rescale, X = b - T2 c, Bn = a - T2 d, T = Bn^-1 X ----
        cv.sys = s.aug;
        solve(grp, cv, NMq, ctl);
        for (int e = tid; e < 2 * NMq * NMq; e += nt) {
            int c = e / (NMq * NMq), r = (e / NMq) % NMq, q = e % NMq;
            double scale = s_cmx[nmin + r] / s_cmx[nmin + q];
            double2 &y = s.aug[(long)(c * NMq + r) * lda + NMq + q];
            y.x *= scale; y.y *= scale;
        }
        __syncthreads();
        for (int e = tid; e < 2 * NMq * NMq; e += nt) {
            int c = e / (NMq * NMq), r = (e / NMq) % NMq, q = e % NMq;
            const double2 *t2 = s.aug + (long)(c * NMq + r) * lda + NMq;
            double xr = 0, xi = 0, br_ = 0, bi_ = 0;
            for (int j = 0; j < NMq; j++) {
                double2 y = t2[j];
                double2 cc = s.C[(long)(c * NMq + j) * ld + q], dd2 = s.D[(long)(c * NMq + j) * ld + q];
                xr += y.x * cc.x - y.y * cc.y; xi += y.y * cc.x + y.x * cc.y;
                br_ += y.x * dd2.x - y.y * dd2.y; bi_ += y.y * dd2.x + y.x * dd2.y;
            }
            double2 bb = s.B[(long)(c * NMq + r) * ld + q], aa = s.A[(long)(c * NMq + r) * ld + q];
            s.aug2[(long)(c * NMq + r) * lda + NMq + q] = make_double2(bb.x - xr, bb.y - xi);
            s.aug2[(long)(c * NMq + r) * lda + q] = make_double2(aa.x - br_, aa.y - bi_);
        }
        __syncthreads();
        cv.sys = s.aug2;
        solve(grp, cv, NMq, ctl);

        // ---- addprc: incident coefficients, scattered coefficients, far field at 90 deg (forward) and 270 deg (back) ----
        // T_c[r][q] = aug2[c][r][NMq+q];  fg(tau_j, n_j) = sum_i ab(i) T(i, j) inside each class.
        for (int w = tid; w < 4 * NMq; w += nt) {
            const int pol = w & 1, idx = w >> 1, c = idx / NMq, q = idx - c * NMq;
            double pnl[T2_NRANKI];
            const double theta = dtr * 90.0;
            double sn, cs;
            sincos(theta, &sn, &cs);
            t2_genlgp(theta, sn, cs, kmv, prodm, pnl);
            cd sacc = mk(0.0, 0.0);
            for (int r = 0; r < NMq; r++) {
                const int n = nmin + r;
                const int tau = (((n + 1) & 1) == c) ? 1 : 2;
                const double p1 = n * cs * pnl[n] - (n + cmv) * pnl[n - 1], p2 = cmv * pnl[n];
                const int pw = (tau == 1) ? (n & 3) : ((n + 1) & 3);
                double mag;
                if (pol == 0) mag = (tau == 1) ? -p2 : p1;   // ab1
                else mag = (tau == 1) ? p1 : -p2;            // ab2
                cd ab = (pw == 0) ? mk(mag, 0.0) : (pw == 1) ? mk(0.0, mag) : (pw == 2) ? mk(-mag, 0.0) : mk(0.0, -mag);
                double2 tt = s.aug2[(long)(c * NMq + r) * lda + NMq + q];
                sacc = sacc + ab * mk(tt.x, tt.y);
            }
            // fg[pol][c][q] parked in matrix A storage (free after the products)
            s.A[(pol * 2 + c) * NMq + q] = make_double2(sacc.x, sacc.y);
        }
        __syncthreads();
        for (int w = tid; w < 4; w += nt) {
            const int iu = w >> 1, pol = w & 1;
            const double theta = dtr * (iu == 0 ? 90.0 : 270.0);
            double sn, cs;
            sincos(theta, &sn, &cs);
            double pnl[T2_NRANKI];
            t2_genlgp(theta, sn, cs, kmv, prodm, pnl);
            cd fa = mk(0.0, 0.0);
            for (int n = nmin; n <= T2_NRANK; n++) {
                const int k = n - nmin;
                const double p1 = n * cs * pnl[n] - (n + cmv) * pnl[n - 1], p2 = cmv * pnl[n];
                const int pw = (4 - ((n + 1) & 3)) & 3;  // (-i)^(n+1)
                const cd cim = (pw == 0) ? mk(1, 0) : (pw == 1) ? mk(0, 1) : (pw == 2) ? mk(-1, 0) : mk(0, -1);
                double2 f1 = s.A[(pol * 2 + ((n + 1) & 1)) * NMq + k];   // fg(tau=1, n)
                double2 f2 = s.A[(pol * 2 + (n & 1)) * NMq + k];         // fg(tau=2, n)
                const double pa = pol == 0 ? p2 : p1, pb = pol == 0 ? p1 : p2;
                cd f = mk(f1.x * pa, f1.y * pa), g = mk(-f2.y * pb, f2.x * pb);
                cd term = (cim * (f + g)) * (1.0 / s_cmx[n]);
                fa = (pol == 0) ? fa + term : fa - term;
            }
            s_acc[w * 2] += rtsfct * fa.x;
            s_acc[w * 2 + 1] += rtsfct * fa.y;
        }
        __syncthreads();
    }
    for (int w = tid; w < 4; w += nt) {
        // s_acc: [iu][pol]; output order borrp borip borrpe boripe forrp forip forpe foripe
        const int iu = w >> 1, pol = w & 1;
        const int slot = (iu == 1 ? 0 : 4) + 2 * pol;
        a.amp[(long)p * 8 + slot] = sigma * s_acc[w * 2] / rtsfct;
        a.amp[(long)p * 8 + slot + 1] = sigma * s_acc[w * 2 + 1] / rtsfct;
        if (w == 0) a.status[p] = ctl->sing ? HS_ST_SINGULAR : HS_ST_OK;
    }
    __syncthreads();
}

#ifndef HS_HOST_EMU
__global__ void __launch_bounds__(256, 1) k_tm2l(T2Args a, int *queue)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ Ctl ctl;
    __shared__ double s_cmx[T2_NRANKI];
    __shared__ double s_acc[8];
    __shared__ int s_item;
    while (true) {
        if (threadIdx.x == 0) { s_item = atomicAdd(queue, 1); ctl.sing = 0; }
        __syncthreads();
        int it = s_item;
        __syncthreads();
        if (it >= a.n) break;
        tm2l_particle(a, it, smem, &ctl, s_cmx, s_acc, threadIdx.x, blockDim.x);
    }
}
#endif

}  // namespace hs

// hs_tm2l.cuh -- two-layer (coated spheroid) EBCM T-matrix, one particle per CTA, everything resident in shared memory.
// The CTA is warp-specialised: 8 assembly warps integrate the six matrices of azimuthal mode m + 1 while 4 solver warps run
// the latency-bound chain of mode m (T2 = x^-1 y, the two products, T = Bn^-1 X, far field) out of a second matrix buffer.
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
#ifndef HS_HOST_EMU
#include "hs_staged.cuh"   // solve_grp2s: the shared-space Gauss-Jordan of the single-layer pipeline
#endif

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
    const cd zi = mk(1.0, 0.0) / z;  // one reciprocal instead of a complex division per recurrence step
    for (int ij = nval; ij >= 2; ij--) {
        double f = (double)(2 * ij - 1);
        cd cf = zi * f;
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
        cd cf = zi * f;
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
struct T2Mat {
    double2 *A, *B, *C, *D;  // [2][17][17] parity-class matrices of a, b, c, d
    double2 *aug;            // [2][17][34]: [x | y] -> [. | T2] -> [Bn | X] -> [. | T]
};
struct T2Smem {
    // node scalars [31]
    double *sinth, *costh, *ckr, *dckr, *ckp2r, *ckp2i, *dkp2r, *dkp2i, *theta, *rckr, *tinvr, *tinvi;
    // row tables [31][17] (index n-1)
    double *bj, *by, *j2r, *j2i, *h2r, *h2i, *rb, *rhr, *rhi, *rb2r, *rb2i, *rh2r, *rh2i;
    // column tables [31][17]
    double *jkr, *jki, *hkr, *hki, *j3r, *j3i, *rbkr, *rbki, *rhkr, *rhki, *rbk2r, *rbk2i;
    double *pn;       // [31][18]
    T2Mat mat[2];     // matrix buffers of two consecutive modes (assembly of m + 1 overlaps the solve chain of m)
};
#define T2_TAB (T2_NNODE * T2_NRANK)
#define T2_MAT_DOUBLES (2 * (4 * 2 * 289 + 2 * 17 * 34))
__host__ __device__ inline int t2_smem_doubles(int nbuf = 2)
{
    return 12 * T2_NNODE + 25 * T2_TAB + T2_NNODE * T2_NRANKI + nbuf * T2_MAT_DOUBLES;
}
__device__ inline void t2_carve(T2Smem &s, double *p, int nbuf = 2)
{
    s.sinth = p; p += T2_NNODE; s.costh = p; p += T2_NNODE; s.ckr = p; p += T2_NNODE; s.dckr = p; p += T2_NNODE;
    s.ckp2r = p; p += T2_NNODE; s.ckp2i = p; p += T2_NNODE; s.dkp2r = p; p += T2_NNODE; s.dkp2i = p; p += T2_NNODE;
    s.theta = p; p += T2_NNODE; s.rckr = p; p += T2_NNODE; s.tinvr = p; p += T2_NNODE; s.tinvi = p; p += T2_NNODE;
    double **tabs[25] = {&s.bj, &s.by, &s.j2r, &s.j2i, &s.h2r, &s.h2i, &s.rb, &s.rhr, &s.rhi, &s.rb2r, &s.rb2i,
                         &s.rh2r, &s.rh2i, &s.jkr, &s.jki, &s.hkr, &s.hki, &s.j3r, &s.j3i, &s.rbkr, &s.rbki,
                         &s.rhkr, &s.rhki, &s.rbk2r, &s.rbk2i};
    for (int i = 0; i < 25; i++) { *tabs[i] = p; p += T2_TAB; }
    s.pn = p; p += T2_NNODE * T2_NRANKI;
    if ((12 * T2_NNODE + 25 * T2_TAB + T2_NNODE * T2_NRANKI) & 1) p++;  // double2 alignment (14105 doubles before the matrices)
    for (int b = 0; b < 2; b++) {
        T2Mat &m = s.mat[b];
        if (b >= nbuf) { m = s.mat[0]; continue; }
        m.A = reinterpret_cast<double2 *>(p); p += 2 * 2 * 289;
        m.B = reinterpret_cast<double2 *>(p); p += 2 * 2 * 289;
        m.C = reinterpret_cast<double2 *>(p); p += 2 * 2 * 289;
        m.D = reinterpret_cast<double2 *>(p); p += 2 * 2 * 289;
        m.aug = reinterpret_cast<double2 *>(p); p += 2 * 2 * 17 * 34;
    }
}

struct T2Args {
    int n;
    const double *dpart, *dcore, *aovrb, *aovrb2, *eor, *eoi, *eir, *eii;
    double alamd;
    double *amp;   // [n][8]
    int *status;   // [n]
    unsigned long long *prof;  // HYDROTM_PROF: cycles per phase, summed over particles (slots: T2P_*) or null
};
enum { T2P_TABLES = 0, T2P_ASM, T2P_ASM_WAIT, T2P_CHAIN, T2P_CHAIN_WAIT, T2P_SOLVE1, T2P_PRODUCTS, T2P_SOLVE2, T2P_ADDPRC, T2P_PARTICLES, T2P_N };
#ifdef HS_HOST_EMU
#define T2_PROF(cond, slot)
#define T2_PROF_T0
#else
#define T2_PROF_T0 long long tprof_ = clock64();
#define T2_PROF(cond, slot)                                                                              \
    if (prof_ && (cond)) {                                                                               \
        const long long now_ = clock64();                                                                \
        atomicAdd(&prof_[slot], (unsigned long long)(now_ - tprof_));                                    \
        tprof_ = now_;                                                                                   \
    }
#endif

// a group of threads that works on one phase: lane `tid` of `nt`, synchronised by named barrier `bar` (0 = whole CTA)
struct T2Grp {
    int tid, nt, bar;
};
__device__ __forceinline__ void t2_sync(const T2Grp &g)
{
#ifndef HS_HOST_EMU
    if (g.bar == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(g.bar), "r"(g.nt) : "memory");
#else
    (void)g;
#endif
}

// per-particle constants (rddata, tmatrix_py.f:580-620, 740), recomputed by every thread
struct T2Par {
    double cpi, dtr, aovrb, aovrb2, conk, conk2, sigma, rtsfct, diff;
    cd eps1, sq1, qs1, sq2, rat12, sqrat, qsrat;
};
__device__ inline T2Par t2_par(const T2Args &a, int p)
{
    T2Par P;
    P.cpi = 3.1415927410125732;  // 4.0*atan(1.0) in single precision (tmatrix_py.f:161)
    P.dtr = P.cpi / 180.0;
    const double third = 0.3333333432674408;  // the single-precision literal 1./3.
    const double dpart = a.dpart[p], dcore = a.dcore[p], alamd = a.alamd;
    P.aovrb = a.aovrb[p]; P.aovrb2 = a.aovrb2[p];
    P.conk = P.cpi * dpart / (alamd * pow(P.aovrb, third));
    P.conk2 = P.cpi * dcore / (alamd * pow(P.aovrb2, third));
    P.sigma = 2.0 * alamd / (P.cpi * 100.0);
    P.eps1 = mk(a.eor[p], a.eoi[p]);
    const cd eps2 = mk(a.eir[p], a.eii[p]);
    P.sq1 = croot(P.eps1); P.qs1 = mk(1.0, 0.0) / P.sq1; P.sq2 = croot(eps2);
    P.rat12 = eps2 / P.eps1; P.sqrat = croot(P.rat12); P.qsrat = mk(1.0, 0.0) / P.sqrat;
    P.rtsfct = 8.0 / P.conk;
    P.diff = (P.cpi / 2.0) / 2.0;
    return P;
}

__device__ __forceinline__ cd ld2(const double *re, const double *im, int i) { return mk(re[i], im[i]); }

// ---- K1: node geometry + radial functions (4 task kinds x 31 nodes); ends with a group barrier ----
__device__ void t2_tables(const T2Par &P, T2Smem &s, const T2Grp &g)
{
    const double sum0 = P.diff, diff = P.diff;
    const cd sq1 = P.sq1, sq2 = P.sq2, sqrat = P.sqrat;
    for (int t = g.tid; t < T2_NNODE; t += g.nt) {
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
        double bovra = 1.0 / P.aovrb;
        double qb = 1.0 / sqrt((bovra * cs) * (bovra * cs) + sn * sn);
        s.ckr[t] = P.conk * qb;
        s.dckr[t] = P.conk * cs * sn * (bovra * bovra - 1.0) * qb * qb * qb;
        double bovra2 = 1.0 / P.aovrb2;
        double qb2 = 1.0 / sqrt((bovra2 * cs) * (bovra2 * cs) + sn * sn);
        double ckr2 = P.conk2 * qb2, dckr2 = P.conk2 * cs * sn * (bovra2 * bovra2 - 1.0) * qb2 * qb2 * qb2;
        s.ckp2r[t] = sq1.x * ckr2; s.ckp2i[t] = sq1.y * ckr2;
        s.dkp2r[t] = sq1.x * dckr2; s.dkp2i[t] = sq1.y * dckr2;
        // reciprocals the integrand would otherwise form for every (pair, node): 1 / kr and 1 / (k' r2)
        s.rckr[t] = 1.0 / s.ckr[t];
        { const cd ti = mk(1.0, 0.0) / mk(s.ckp2r[t], s.ckp2i[t]); s.tinvr[t] = ti.x; s.tinvi[t] = ti.y; }
        // park kr2 for the iswt=3 task
        s.pn[t] = ckr2;
    }
    t2_sync(g);
    for (int task = g.tid; task < 4 * T2_NNODE; task += g.nt) {
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
    t2_sync(g);
}

__device__ __forceinline__ double t2_prodm(int kmv)
{
    double prodm = 1.0;
    if (kmv > 0) {
        double quanm = kmv;
        for (int i = 1; i <= kmv; i++) { quanm += 1.0; prodm = quanm * prodm / 2.0; }
    }
    return prodm;
}

// The 12 complex integrand values of one (irow, icol) pair (gener, tmatrix_py.f:1000-1338) summed over the quadrature nodes
// [step0, step1) of the 31-node Patterson rule (quad, :1340-1595: step < 30 -> node step + 1, weight PAT31_W[step / 2]; step 30
// -> the centre node).  EVEN: irow + icol even -> sub-blocks J (acc 0..5: a b c d x y) and K (acc 6..11); else I and L.
template <bool EVEN>
__device__ __forceinline__ void t2_pair_nodes(const T2Par &P, const T2Smem &s, int kmv, int irow, int icol, int step0, int step1, cd *acc)
{
    const double cmv = kmv, em = kmv > 0 ? 2.0 : 1.0, qem = -2.0 / em, cm2 = cmv * cmv;
    const cd eps1 = P.eps1, qs1 = P.qs1, rat12 = P.rat12, qsrat = P.qsrat;
    const double crow = irow, ccol = icol, crow1 = crow + 1.0, ccol1 = ccol + 1.0;
    const double crowm = cmv + crow, ccolm = cmv + ccol, crij = crow + ccol, crssij = crow * ccol;
    for (int step = step0; step < step1; step++) {
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
        if (!EVEN) {
            const double b1 = b1a + b1b;
            const cd htbk = h * bk, btbk = bk * br, bthk = hk * br, hthk = h * hk;
            const double fac = 2.0 * cmv * sinth;
            const double rckr = s.rckr[t];
            const double gs = dckr * sinth * pnr1c1, tl = crssij * (crij + 2.0) * rckr, t0 = crssij * rckr;
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
            const cd tinv = mk(s.tinvr[t], s.tinvi[t]);
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
#pragma unroll
        for (int q = 0; q < 12; q++) { acc[q].x = fma(v[q].x, w, acc[q].x); acc[q].y = fma(v[q].y, w, acc[q].y); }
    }
}

// ---- K2: the six matrices of mode kmv (outer-surface a, b, c, d, inner-surface x, y) into M, as two parity-class blocks each.
// pair (n = irow, n' = icol); even pairs first (J, K), then odd pairs (I, L).  `split` (1, 2 or 4, a power of two <= 32) lanes
// share a pair, each summing a contiguous range of the quadrature nodes; partial sums are combined by warp shuffles (the group
// must then consist of whole warps).  No barrier at the end.
__device__ void t2_assemble(const T2Par &P, T2Smem &s, const T2Mat &M, int kmv, const T2Grp &g, int split)
{
    const double prodm = t2_prodm(kmv);
    const int nmin = kmv > 1 ? kmv : 1, NMq = T2_NRANK - nmin + 1, ld = NMq, lda = 2 * NMq;
    for (int t = g.tid; t < T2_NNODE; t += g.nt)
        t2_genlgp(s.theta[t], s.sinth[t], s.costh[t], kmv, prodm, s.pn + t * T2_NRANKI);
    t2_sync(g);
    const int h0 = (NMq + 1) >> 1, h1 = NMq >> 1;
    const int nE = h0 * h0 + h1 * h1, nO = (kmv == 0) ? 0 : 2 * h0 * h1;
    // warp-aligned start of the odd pairs so that no warp mixes the two code paths
    const int nEs = nE * split, nEpad = (nEs + 31) & ~31, tot = nEpad + nO * split;
    const int nstep = (T2_NNODE + split - 1) / split;
    for (int base = 0; base < tot; base += g.nt) {
        const int item = base + g.tid;       // all lanes of a warp stay in the loop (shuffles below)
        const bool even = item < nEpad;
        const int u0 = even ? item : item - nEpad;
        const int pair = u0 / split, part = u0 - pair * split;
        const bool valid = even ? (item < nEs) : (item < tot);
        int k1 = 0, k2 = 0;  // k1 = n - nmin (irow), k2 = n' - nmin (icol)
        if (valid) {
            if (even) {
                if (pair < h0 * h0) { int q = pair / h0; k1 = 2 * q; k2 = 2 * (pair - q * h0); }
                else { int u = pair - h0 * h0; int q = u / h1; k1 = 2 * q + 1; k2 = 2 * (u - q * h1) + 1; }
            } else {
                if (pair < h0 * h1) { int q = pair / h1; k1 = 2 * q; k2 = 2 * (pair - q * h1) + 1; }
                else { int u = pair - h0 * h1; int q = u / h0; k1 = 2 * q + 1; k2 = 2 * (u - q * h0); }
            }
        }
        const int irow = nmin + k1, icol = nmin + k2;
        cd acc[12];
#pragma unroll
        for (int q = 0; q < 12; q++) acc[q] = mk(0.0, 0.0);
        if (valid) {
            const int s0 = part * nstep, s1 = (s0 + nstep < T2_NNODE) ? s0 + nstep : T2_NNODE;
            if (even) t2_pair_nodes<true>(P, s, kmv, irow, icol, s0, s1, acc);
            else t2_pair_nodes<false>(P, s, kmv, irow, icol, s0, s1, acc);
        }
        for (int o = 1; o < split; o <<= 1) {
#pragma unroll
            for (int q = 0; q < 12; q++) {
                acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o);
                acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
            }
        }
        if (!valid || part != 0) continue;
#pragma unroll
        for (int q = 0; q < 12; q++) acc[q] = acc[q] * P.diff;
        // scatter into the parity-class matrices: first six -> class ca, last six -> class cb
        // even: [0..5] = J -> (2,n')-(2,n): class n'&1 ; [6..11] = K -> (1,n')-(1,n): class (n'+1)&1
        // odd : [0..5] = I -> (1,n')-(2,n): class (n'+1)&1 ; [6..11] = L -> (2,n')-(1,n): class n'&1
        const int c_first = even ? (icol & 1) : ((icol + 1) & 1), c_second = c_first ^ 1;
        const int e1 = (c_first * NMq + k2) * ld + k1, e2 = (c_second * NMq + k2) * ld + k1;
        const int g1 = (c_first * NMq + k2) * lda + k1, g2 = (c_second * NMq + k2) * lda + k1;
        M.A[e1] = make_double2(acc[0].x, acc[0].y); M.B[e1] = make_double2(acc[1].x, acc[1].y);
        M.C[e1] = make_double2(acc[2].x, acc[2].y); M.D[e1] = make_double2(acc[3].x, acc[3].y);
        M.aug[g1] = make_double2(acc[4].x, acc[4].y); M.aug[g1 + NMq] = make_double2(acc[5].x, acc[5].y);
        M.A[e2] = make_double2(acc[6].x, acc[6].y); M.B[e2] = make_double2(acc[7].x, acc[7].y);
        M.C[e2] = make_double2(acc[8].x, acc[8].y); M.D[e2] = make_double2(acc[9].x, acc[9].y);
        M.aug[g2] = make_double2(acc[10].x, acc[10].y); M.aug[g2 + NMq] = make_double2(acc[11].x, acc[11].y);
    }
    if (kmv == 0) {
        // odd (I, L) entries vanish for m = 0: zero them explicitly
        for (int item = g.tid; item < 2 * h0 * h1; item += g.nt) {
            int k1, k2, u = item;
            if (u < h0 * h1) { int q = u / h1; k1 = 2 * q; k2 = 2 * (u - q * h1) + 1; }
            else { u -= h0 * h1; int q = u / h0; k1 = 2 * q + 1; k2 = 2 * (u - q * h0); }
            const double2 z = make_double2(0.0, 0.0);
            for (int c = 0; c < 2; c++) {
                const int e = (c * NMq + k2) * ld + k1, gg = (c * NMq + k2) * lda + k1;
                M.A[e] = z; M.B[e] = z; M.C[e] = z; M.D[e] = z; M.aug[gg] = z; M.aug[gg + NMq] = z;
            }
        }
    }
}

#ifndef HS_HOST_EMU
// One parity-class system  L X = R  (L, R: [NMq][NMq] row-major with row strides ldl, ldr) eliminated by TWO warps out of
// registers: warp `part` 0 owns L, warp 1 owns R; lane r < 17 keeps ROW r of its part in registers, padded to order 17 with
// identity rows / columns (a padded row is the only candidate of its own column, so the real rows are eliminated exactly as in the
// unpadded system).  A step: the L lanes publish their entries of pivot column k (one STS each, double-buffered, one 64-lane named
// barrier per step); every warp finds the pivot row with two warp reductions (same rule as prcssm / solve_multi: largest
// |re| + |im| among the rows not used yet, low 8 mantissa bits dropped, ties to the lowest row); the lane that holds the pivot row
// scales it and publishes it to its own warp; every lane updates its row.  Rows are never moved: `vp` is the position the row would
// have after the reference's row swaps, so the lane ends up holding solution row vp, which it writes to out[vp][.] (row stride ldo).
// No LDS -> DFMA -> STS round trip per row update, ~300 instructions per step and warp.
__device__ __forceinline__ void t2_solve_rows(const double2 *L, int ldl, const double2 *R, int ldr, double2 *out, int ldo, int NMq,
                                              int part, int lane, int barid, double2 *colbuf, double2 *rowbuf, int *sing,
                                              const double *rowscale = nullptr, const double *colscale = nullptr)
{
    const unsigned FULL = 0xffffffffu;
    double2 v[T2_NRANK];
    const bool realr = lane < NMq;
    const double2 *src = part ? R + lane * ldr : L + lane * ldl;
#pragma unroll
    for (int j = 0; j < T2_NRANK; j++) {
        double2 x = make_double2(0.0, 0.0);
        if (realr && j < NMq) x = src[j];
        else if (part == 0 && j == lane) x.x = 1.0;
        v[j] = x;
    }
    int vp = lane;
    double2 f = v[0];                             // this row's entry of the pivot column (L part)
    double2 sc = make_double2(1.0, 0.0);          // pending scale of this row: true row = sc * v (set when the row is the pivot)
    // NMq steps: the padded rows / columns (identity) never interact with the real ones
    for (int k = 0; k < NMq; k++) {
        double2 *cb = colbuf + (k & 1) * T2_NRANK;
        if (part == 0 && lane < T2_NRANK) cb[lane] = f;
        asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(64) : "memory");
        double2 c = make_double2(0.0, 0.0);
        if (lane < T2_NRANK) c = cb[lane];
        // pivot row: only rows not used yet are candidates, and their pending scale is still 1
        const bool cand = lane < T2_NRANK && vp >= k;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(c.x) + fabs(c.y));
        const unsigned hi = (unsigned)(bits >> 32) + 1u, lo = ((unsigned)bits & 0xFFFFFF00u) | (unsigned)(255 - vp);
        const unsigned m1 = __reduce_max_sync(FULL, cand ? hi : 0u);
        const bool top = cand && hi == m1;
        const unsigned m2 = __reduce_max_sync(FULL, top ? lo : 0u);
        const int P = __ffs(__ballot_sync(FULL, top && lo == m2)) - 1;
        if (m1 == 1u && (m2 >> 8) == 0u && part == 0 && lane == 0) *sing = 1;
        // the pivot lane hands its row to the warp as it is (no arithmetic in a one-lane branch: FP64 issue slots are charged per
        // warp instruction).  With W = sc * V the true rows, the step  W_r -= W_r[k] (W_P / pivot)  reads  V_r -= (V_r[k] / pivot) V_P
        // for every row r != P whatever its pending scale, and the pivot row only records sc_P = 1 / pivot.
        if (lane == P) {
#pragma unroll
            for (int j = 0; j < T2_NRANK; j++) rowbuf[j] = v[j];
        }
        const double pvx = __shfl_sync(FULL, c.x, P), pvy = __shfl_sync(FULL, c.y, P);
        const int vpP = __shfl_sync(FULL, vp, P);
        const double rden = hs_rcp(pvx * pvx + pvy * pvy);
        const double2 inv = make_double2(pvx * rden, -pvy * rden);
        double2 cs = make_double2(-(c.x * inv.x - c.y * inv.y), -(c.x * inv.y + c.y * inv.x));
        if (lane == P) { cs = make_double2(0.0, 0.0); sc = inv; }
        __syncwarp();
        double2 fn = f;
#pragma unroll
        for (int j0 = 0; j0 < T2_NRANK; j0 += 4) {
            double2 rw[4];
#pragma unroll
            for (int q = 0; q < 4; q++) if (j0 + q < T2_NRANK) rw[q] = rowbuf[j0 + q];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = j0 + q;
                if (j < T2_NRANK) {
                    double2 u = v[j];
                    u.x = fma(-cs.y, rw[q].y, fma(cs.x, rw[q].x, u.x));
                    u.y = fma(cs.y, rw[q].x, fma(cs.x, rw[q].y, u.y));
                    v[j] = u;
                    if (j == k + 1) fn = u;
                }
            }
        }
        f = fn;
        if (vp == k) vp = vpP;
        if (lane == P) vp = k;
        __syncwarp();  // rowbuf is rewritten by the next step's pivot lane
    }
    if (part == 1 && realr) {
        // solution row vp = sc * v, times the caller's row / column factors (cmxnrm rescaling of T2, tmatrix_py.f:387-399)
        const double rs = rowscale ? rowscale[vp] : 1.0;
        const double2 s2 = make_double2(sc.x * rs, sc.y * rs);
#pragma unroll
        for (int j = 0; j < T2_NRANK; j++)
            if (j < NMq) {
                const double cj = colscale ? colscale[j] : 1.0;
                out[vp * ldo + j] = make_double2((v[j].x * s2.x - v[j].y * s2.y) * cj, (v[j].x * s2.y + v[j].y * s2.x) * cj);
            }
    }
}
#endif

struct T2SolveWs {
    Ctl *ctl;            // generic solver (host emulation): descriptors + pivot keys
    double2 *colbuf;     // [2 classes][2][17] published pivot columns (device)
    double2 *rowbuf;     // [4 warps][17] published pivot rows (device)
    int *sing;
};
// both parity classes: M.aug right half <- L^-1 R with [L | R] = M.aug (which = 0: x, y) or L = M.A, R = M.B (which = 1: Bn, X)
// (prcssm, tmatrix_py.f:2177-2280).  Ends with a group barrier.
__device__ __forceinline__ void t2_solve(const T2Mat &M, int NMq, const T2Grp &g, const T2SolveWs &ws, int which,
                                         const double *rowscale = nullptr, const double *colscale = nullptr)
{
    const int lda = 2 * NMq;
#ifdef HS_HOST_EMU
    if (which) {
        for (int e = g.tid; e < 2 * NMq * NMq; e += g.nt) {
            int c = e / (NMq * NMq), r = (e / NMq) % NMq, q = e % NMq;
            M.aug[(c * NMq + r) * lda + q] = M.A[(c * NMq + r) * NMq + q];
            M.aug[(c * NMq + r) * lda + NMq + q] = M.B[(c * NMq + r) * NMq + q];
        }
        t2_sync(g);
    }
    Grp<false> grp;
    grp.tid = g.tid; grp.nt = g.nt;
    Carve cv;
    cv.sys = M.aug;
    solve(grp, cv, NMq, ws.ctl);
    if (ws.ctl->sing) *ws.sing = 1;
    if (rowscale) {
        for (int e = g.tid; e < 2 * NMq * NMq; e += g.nt) {
            int c = e / (NMq * NMq), r = (e / NMq) % NMq, q = e % NMq;
            double2 &y = M.aug[(c * NMq + r) * lda + NMq + q];
            const double sc = rowscale[r] * colscale[q];
            y.x *= sc; y.y *= sc;
        }
        t2_sync(g);
    }
#else
    // 64 lanes per class (two whole warps; named barriers 1 and 2)
    const int cl = g.tid >> 6, warp = g.tid >> 5;
    double2 *aug = M.aug + cl * NMq * lda;
    if (which) t2_solve_rows(M.A + cl * NMq * NMq, NMq, M.B + cl * NMq * NMq, NMq, aug + NMq, lda, NMq, warp & 1, g.tid & 31, 1 + cl,
                             ws.colbuf + cl * 2 * T2_NRANK, ws.rowbuf + warp * T2_NRANK, ws.sing);
    else t2_solve_rows(aug, lda, aug + NMq, lda, aug + NMq, lda, NMq, warp & 1, g.tid & 31, 1 + cl, ws.colbuf + cl * 2 * T2_NRANK,
                       ws.rowbuf + warp * T2_NRANK, ws.sing, rowscale, colscale);
    t2_sync(g);
#endif
}

// Per-mode constants of the solve chain: cmxnrm (:345-368) and its reciprocal, and for the two far-field angles (90 deg forward,
// 270 deg back, :175) the angular factors p1 = n cos pn(n) - (n + m) pn(n - 1), p2 = m pn(n) of addprc (:2282-2523).
struct T2Prep {
    double cmx[T2_NRANKI], icmx[T2_NRANKI];
    double2 pp[2][T2_NRANKI];
};
// items of t2_prep: n = 1..17 (cmx) and the two angles; any subset of lanes may run them (item = lane index, stride nl)
__device__ inline void t2_prep(const T2Par &P, int kmv, T2Prep &pr, int item)
{
    const double cmv = kmv, em = kmv > 0 ? 2.0 : 1.0;
    if (item >= 1 && item <= T2_NRANK) {
        const int n = item;
        double ck = n, fctki = 1.0, v;
        if (kmv > 0 && n < kmv) v = 1.0;
        else {
            if (kmv > 0) {
                double fprod = n - kmv + 1;
                for (int l = n - kmv + 1; l <= n + kmv; l++) { fctki *= fprod; fprod += 1.0; }
            }
            v = 4.0 * ck * (ck + 1.0) * fctki / (em * (2.0 * ck + 1.0));
        }
        pr.cmx[n] = v;
        pr.icmx[n] = 1.0 / v;
    } else if (item == T2_NRANK + 1 || item == T2_NRANK + 2) {
        const int q = item - (T2_NRANK + 1);
        const double theta = P.dtr * (q == 0 ? 90.0 : 270.0);
        double sn, cs, pnl[T2_NRANKI];
        sincos(theta, &sn, &cs);
        t2_genlgp(theta, sn, cs, kmv, t2_prodm(kmv), pnl);
        pr.pp[q][0] = make_double2(0.0, 0.0);
        for (int n = 1; n <= T2_NRANK; n++) pr.pp[q][n] = make_double2(n * cs * pnl[n] - (n + cmv) * pnl[n - 1], cmv * pnl[n]);
    }
}
#define T2_PREP_ITEMS (T2_NRANK + 3)

// ---- K3 + addprc of mode kmv on the matrices in M: T2 = x^-1 y (class-wise), rescaled by cmxnrm; X = b - T2 c, Bn = a - T2 d (in
// place of b and a); T = Bn^-1 X; incident / scattered coefficients and the far field at 90 deg (forward) and 270 deg (back)
// accumulated into s_acc.  pr = t2_prep of this mode.  Device: g = the 128 solver lanes.  Ends with a group barrier.
__device__ void t2_chain(const T2Par &P, const T2Mat &M, int kmv, const T2Grp &g, const T2Prep &pr, double *s_acc, const T2SolveWs &ws,
                         unsigned long long *prof_ = nullptr)
{
    T2_PROF_T0
    const int nmin = kmv > 1 ? kmv : 1, NMq = T2_NRANK - nmin + 1, ld = NMq, lda = 2 * NMq;
    t2_solve(M, NMq, g, ws, 0, pr.cmx + nmin, pr.icmx + nmin);  // T2, rescaled by cmxnrm(row) / cmxnrm(column) (:387-399)
    T2_PROF(g.tid == 0, T2P_SOLVE1)
    // in place: A <- a - T2 d (= Bn), B <- b - T2 c (= X); every (c, r, q) reads only its own entries of A and B
    for (int e = g.tid; e < 2 * NMq * NMq; e += g.nt) {
        int c = e / (NMq * NMq), r = (e / NMq) % NMq, q = e % NMq;
        const double2 *t2 = M.aug + (c * NMq + r) * lda + NMq;
        const double2 *pc = M.C + c * NMq * ld + q, *pd = M.D + c * NMq * ld + q;
        double xr = 0, xi = 0, br_ = 0, bi_ = 0;
        for (int j = 0; j < NMq; j++) {
            const double2 y = t2[j], cc = pc[j * ld], dd2 = pd[j * ld];
            xr += y.x * cc.x - y.y * cc.y; xi += y.y * cc.x + y.x * cc.y;
            br_ += y.x * dd2.x - y.y * dd2.y; bi_ += y.y * dd2.x + y.x * dd2.y;
        }
        double2 &bb = M.B[(c * NMq + r) * ld + q], &aa = M.A[(c * NMq + r) * ld + q];
        bb = make_double2(bb.x - xr, bb.y - xi);
        aa = make_double2(aa.x - br_, aa.y - bi_);
    }
    t2_sync(g);
    T2_PROF(g.tid == 0, T2P_PRODUCTS)
    t2_solve(M, NMq, g, ws, 1);
    T2_PROF(g.tid == 0, T2P_SOLVE2)

    // ---- addprc: T_c[r][q] = aug[c][r][NMq+q];  fg(tau_j, n_j) = sum_i ab(i) T(i, j) inside each class.
    for (int w = g.tid; w < 4 * NMq; w += g.nt) {
        const int pol = w & 1, idx = w >> 1, c = idx / NMq, q = idx - c * NMq;
        cd sacc = mk(0.0, 0.0);
        for (int r = 0; r < NMq; r++) {
            const int n = nmin + r;
            const int tau = (((n + 1) & 1) == c) ? 1 : 2;
            const double2 pq = pr.pp[0][n];                     // incidence at 90 deg
            const int pw = (tau == 1) ? (n & 3) : ((n + 1) & 3);
            double mag;
            if (pol == 0) mag = (tau == 1) ? -pq.y : pq.x;   // ab1
            else mag = (tau == 1) ? pq.x : -pq.y;            // ab2
            cd ab = (pw == 0) ? mk(mag, 0.0) : (pw == 1) ? mk(0.0, mag) : (pw == 2) ? mk(-mag, 0.0) : mk(0.0, -mag);
            double2 tt = M.aug[(c * NMq + r) * lda + NMq + q];
            sacc = sacc + ab * mk(tt.x, tt.y);
        }
        // fg[pol][c][q] parked in matrix C storage (free after the products)
        M.C[(pol * 2 + c) * NMq + q] = make_double2(sacc.x, sacc.y);
    }
    t2_sync(g);
    for (int w = g.tid; w < 4; w += g.nt) {
        const int iu = w >> 1, pol = w & 1;
        cd fa = mk(0.0, 0.0);
        for (int n = nmin; n <= T2_NRANK; n++) {
            const int k = n - nmin;
            const double2 pq = pr.pp[iu][n];
            const int pw = (4 - ((n + 1) & 3)) & 3;  // (-i)^(n+1)
            const cd cim = (pw == 0) ? mk(1, 0) : (pw == 1) ? mk(0, 1) : (pw == 2) ? mk(-1, 0) : mk(0, -1);
            double2 f1 = M.C[(pol * 2 + ((n + 1) & 1)) * NMq + k];   // fg(tau=1, n)
            double2 f2 = M.C[(pol * 2 + (n & 1)) * NMq + k];         // fg(tau=2, n)
            const double pa = pol == 0 ? pq.y : pq.x, pb = pol == 0 ? pq.x : pq.y;
            cd f = mk(f1.x * pa, f1.y * pa), gg = mk(-f2.y * pb, f2.x * pb);
            cd term = (cim * (f + gg)) * pr.icmx[n];
            fa = (pol == 0) ? fa + term : fa - term;
        }
        s_acc[w * 2] += P.rtsfct * fa.x;
        s_acc[w * 2 + 1] += P.rtsfct * fa.y;
    }
    t2_sync(g);
    T2_PROF(g.tid == 0, T2P_ADDPRC)
}

__device__ inline void t2_store(const T2Args &a, int p, const T2Par &P, const double *s_acc, int sing, int w)
{
    // s_acc: [iu][pol]; output order borrp borip borrpe boripe forrp forip forpe foripe
    const int iu = w >> 1, pol = w & 1;
    const int slot = (iu == 1 ? 0 : 4) + 2 * pol;
    a.amp[(long)p * 8 + slot] = P.sigma * s_acc[w * 2] / P.rtsfct;
    a.amp[(long)p * 8 + slot + 1] = P.sigma * s_acc[w * 2 + 1] / P.rtsfct;
    if (w == 0) a.status[p] = sing ? HS_ST_SINGULAR : HS_ST_OK;
}

// One particle by ONE group, phase after phase (the single-lane host emulation of tests/emu; smem: t2_smem_doubles(1)).
__device__ void tm2l_particle(const T2Args &a, int p, double *smem, Ctl *ctl, double *s_cmx, double *s_acc,
                              const int tid, const int nt)
{
    T2Smem s;
    t2_carve(s, smem, 1);
    const T2Par P = t2_par(a, p);
    T2Grp g;
    g.tid = tid; g.nt = nt; g.bar = 0;
    int sing = 0;
    T2Prep *pr = reinterpret_cast<T2Prep *>(s_cmx);   // caller's scratch: at least sizeof(T2Prep)
    T2SolveWs ws;
    ws.ctl = ctl; ws.colbuf = nullptr; ws.rowbuf = nullptr; ws.sing = &sing;
    t2_tables(P, s, g);
    for (int w = tid; w < 8; w += nt) s_acc[w] = 0.0;  // acans: [fwd pol1, fwd pol2, back pol1, back pol2]
    t2_sync(g);
    for (int kmv = 0; kmv < T2_NM; kmv++) {
        t2_assemble(P, s, s.mat[0], kmv, g, 1);
        for (int it = tid; it < T2_PREP_ITEMS; it += nt) t2_prep(P, kmv, *pr, it);
        t2_sync(g);
        t2_chain(P, s.mat[0], kmv, g, *pr, s_acc, ws);
    }
    for (int w = tid; w < 4; w += nt) t2_store(a, p, P, s_acc, sing, w);
    t2_sync(g);
}

#ifndef HS_HOST_EMU
#define T2_ASM_THREADS 256
#define T2_SOLVE_THREADS 128
#define T2_ASM_REGS 192
#define T2_SOLVE_REGS 112
// Persistent CTAs pull particles from a queue and run ONE continuous software pipeline over the (particle, mode) items they
// draw: at stage t the assembly warps (two warpgroups, 192 registers per thread after setmaxnreg) integrate item t into matrix
// buffer t & 1 -- building the node tables first when the item opens a new particle -- while the solver warps (one warpgroup, 112
// registers) run the chain of item t - 1 out of buffer (t - 1) & 1 (the chain never reads the node tables, so the last mode of
// particle p overlaps the tables and the first mode of particle p + 1); one CTA barrier per stage.
__global__ void __launch_bounds__(T2_ASM_THREADS + T2_SOLVE_THREADS, 1) k_tm2l(T2Args a, int *queue, int split)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ T2Prep s_prep[T2_NM];
    __shared__ double s_acc[8];
    __shared__ double2 s_colbuf[2 * 2 * T2_NRANK], s_rowbuf[4 * T2_NRANK];
    __shared__ int s_pid[2], s_sing;
    const int tid = threadIdx.x;
    T2Smem s;
    t2_carve(s, smem, 2);
    unsigned long long *prof_ = a.prof;
    if (tid >= T2_ASM_THREADS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T2_SOLVE_REGS));
        T2Grp gsol;
        gsol.tid = tid - T2_ASM_THREADS; gsol.nt = T2_SOLVE_THREADS; gsol.bar = 3;
        T2SolveWs ws;
        ws.ctl = nullptr; ws.colbuf = s_colbuf; ws.rowbuf = s_rowbuf; ws.sing = &s_sing;
        __syncthreads();                                       // stage 0: nothing to solve yet
        T2Par P;
        T2_PROF_T0
        for (int t = 1;; t++) {
            const int it = s_pid[(t - 1) & 1], kmv = (t - 1) % T2_NM;
            if (it < 0) break;
            if (kmv == 0) {
                // a new particle: constants of its seven chains, accumulators, singular flag
                P = t2_par(a, it);
                for (int w = gsol.tid; w < T2_NM * 32; w += T2_SOLVE_THREADS) {
                    const int m = w >> 5, item = w & 31;
                    if (item < T2_PREP_ITEMS) t2_prep(P, m, s_prep[m], item);
                }
                if (gsol.tid < 8) s_acc[gsol.tid] = 0.0;
                if (gsol.tid == 8) s_sing = 0;
                t2_sync(gsol);
            }
            t2_chain(P, s.mat[(t - 1) & 1], kmv, gsol, s_prep[kmv], s_acc, ws, a.prof);
            if (kmv == T2_NM - 1 && gsol.tid < 4) t2_store(a, it, P, s_acc, s_sing, gsol.tid);
            T2_PROF(gsol.tid == 0, T2P_CHAIN)
            __syncthreads();
            T2_PROF(gsol.tid == 0, T2P_CHAIN_WAIT)
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(T2_ASM_REGS));
        T2Grp gasm;
        gasm.tid = tid; gasm.nt = T2_ASM_THREADS; gasm.bar = 4;
        T2Par P;
        int it = -1;
        T2_PROF_T0
        for (int t = 0;; t++) {
            const int kmv = t % T2_NM;
            if (kmv == 0) {
                if (tid == 0) s_pid[t & 1] = atomicAdd(queue, 1);
                t2_sync(gasm);
                it = s_pid[t & 1];
                if (it >= a.n) it = -1;
                if (it >= 0) {
                    P = t2_par(a, it);
                    t2_tables(P, s, gasm);
                    T2_PROF(tid == 0, T2P_TABLES)
                }
            }
            if (tid == 0) s_pid[t & 1] = it;
            if (it >= 0) t2_assemble(P, s, s.mat[t & 1], kmv, gasm, split);
            T2_PROF(tid == 0, T2P_ASM)
            __syncthreads();
            T2_PROF(tid == 0, T2P_ASM_WAIT)
            if (it < 0) break;
            if (prof_ && tid == 0 && kmv == T2_NM - 1) atomicAdd(&prof_[T2P_PARTICLES], 1ull);
        }
    }
}
#endif

}  // namespace hs

// hs_ampl.cuh -- K4: amplitude matrix S (2x2 complex) and phase matrix Z (4x4) from a packed T-matrix,
// with fused fixed-quadrature orientation averaging.
//
// Replaces pytmatrix's Fortran `calcampl` (Mishchenko AMPL + VIGAMPL) and the Python loop of
// orientation.orient_averaged_fixed; SURVEY.md rows P8, P9 / Appendix A.7-A.8.
// One CTA per particle; one thread per (geometry, orientation) item.  Every thread walks the same
// T elements in the same order, so T is read as warp-wide broadcasts straight from L2/HBM, once per CTA.
#pragma once
#include "hs_compat.cuh"
#include "../../include/hydrotm.h"

namespace hs {

struct AmplArgs {
    int n;
    const double *tarena;
    const long long *toff;
    const int *nmax;
    const int *status;
    const double *lam;
    int n_geom;
    const double *geom;  // [n_geom][4] thet0, thet, phi0, phi (deg)
    int n_orient;
    const double *alpha, *beta, *weight;
    double scale;
    int reduce;
    double *S, *Z;
};

// pi_n = d^n_{0m}/sin(theta) and tau_n = d d^n_{0m}/d theta at one direction (VIGAMPL), n=1..nmax at [n]
__device__ inline void vigampl(double x, int nmax, int m, double *dv1, double *dv2)
{
    for (int n = 1; n <= nmax; n++) { dv1[n] = 0.0; dv2[n] = 0.0; }
    if (fabs(1.0 - fabs(x)) <= 1e-10) {
        if (m != 1) return;
        for (int n = 1; n <= nmax; n++) {
            double dn = 0.5 * sqrt((double)(n * (n + 1)));
            if (x < 0.0) dn = dn * (((n + 1) & 1) ? -1.0 : 1.0);
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

struct C2 { double x, y; };
__device__ __forceinline__ void cmac(double &ar, double &ai, double tr, double ti, double cr, double ci)
{
    // a += t * c
    ar += tr * cr - ti * ci;
    ai += tr * ci + ti * cr;
}

// One (particle, geometry, orientation) evaluation; out[0..7] = S11,S12,S21,S22 (re,im), out[8..23] = Z row-major
__device__ void ampl_one(const double2 *__restrict__ t, int nmax, double lam, double thet0, double thet,
                         double phi0, double phi, double alpha, double beta, const double *fn, double *out)
{
    const double pin = 3.14159265358979323846, pin2 = pin * 0.5, rad = pin / 180.0;
    double alph = alpha * rad, bet = beta * rad;
    double thetl = thet0 * rad, phil = phi0 * rad, thetl1 = thet * rad, phil1 = phi * rad;
    const double eps = 1e-7;
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
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += B[i][k] * AL[k][j]; C[i][j] = x; }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += AP[i][k] * C[k][j]; R[i][j] = x; }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += B[i][k] * AL1[k][j]; C[i][j] = x; }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += AP1[i][k] * C[k][j]; R1[i][j] = x; }
    double d = 1.0 / (R1[0][0] * R1[1][1] - R1[0][1] * R1[1][0]);
    double x00 = R1[0][0];
    R1[0][0] = R1[1][1] * d; R1[0][1] = -R1[0][1] * d; R1[1][0] = -R1[1][0] * d; R1[1][1] = x00 * d;

    const double dcth0 = ctp, dcth = ctp1, ph = phip1 - phip;
    double vvr = 0, vvi = 0, vhr = 0, vhi = 0, hvr = 0, hvi = 0, hhr = 0, hhi = 0;
    double dv1[HS_NPN1 + 1], dv2[HS_NPN1 + 1], dv01[HS_NPN1 + 1], dv02[HS_NPN1 + 1];
    for (int m = 0; m <= nmax; m++) {
        const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1;
        vigampl(dcth, nmax, m, dv1, dv2);
        vigampl(dcth0, nmax, m, dv01, dv02);
        double fc = 2.0 * cos(m * ph), fs = 2.0 * sin(m * ph);
        for (int nn = nmin; nn <= nmax; nn++) {
            const double dv1nn = m * dv01[nn], dv2nn = dv02[nn];
            const double2 *row0 = t + (long)(0 * NM + (nn - nmin)) * NM;
            const double2 *row1 = t + (long)(1 * NM + (nn - nmin)) * NM;
            for (int n = nmin; n <= nmax; n++) {
                const double dv1n = m * dv1[n], dv2n = dv2[n];
                // CAL(n,nn) = i^(nn-n-1) * sqrt((2n+1)(2nn+1)/(n(n+1)nn(nn+1)))
                const double rn = fn[n] * fn[nn];
                const int pw = (nn - n - 1) & 3;
                double calr = (pw == 0) ? rn : (pw == 2) ? -rn : 0.0;
                double cali = (pw == 1) ? rn : (pw == 3) ? -rn : 0.0;
                const int k = n - nmin;
                const double2 ta = ((n + 1) & 1) ? row1[k] : row0[k];  // tau = 1 row
                const double2 tb = (n & 1) ? row1[k] : row0[k];        // tau = 2 row
                const bool even = (((n + nn) & 1) == 0);
                if (m == 0) {
                    if (even) {
                        double f = dv2n * dv2nn;
                        double cnr = calr * f, cni = cali * f;
                        cmac(vvr, vvi, tb.x, tb.y, cnr, cni);  // T22
                        cmac(hhr, hhi, ta.x, ta.y, cnr, cni);  // T11
                    }
                } else {
                    double cn1r = calr * fc, cn1i = cali * fc, cn2r = calr * fs, cn2i = cali * fs;
                    double d11 = dv1n * dv1nn, d12 = dv1n * dv2nn, d21 = dv2n * dv1nn, d22 = dv2n * dv2nn;
                    double s1r, s1i, s2r, s2i, s3r, s3i, s4r, s4i;
                    if (even) {  // ta = T11, tb = T22
                        s1r = ta.x * d11 + tb.x * d22; s1i = ta.y * d11 + tb.y * d22;
                        s2r = ta.x * d12 + tb.x * d21; s2i = ta.y * d12 + tb.y * d21;
                        s3r = ta.x * d21 + tb.x * d12; s3i = ta.y * d21 + tb.y * d12;
                        s4r = ta.x * d22 + tb.x * d11; s4i = ta.y * d22 + tb.y * d11;
                    } else {     // ta = T12, tb = T21
                        s1r = tb.x * d21 + ta.x * d12; s1i = tb.y * d21 + ta.y * d12;
                        s2r = tb.x * d22 + ta.x * d11; s2i = tb.y * d22 + ta.y * d11;
                        s3r = tb.x * d11 + ta.x * d22; s3i = tb.y * d11 + ta.y * d22;
                        s4r = tb.x * d12 + ta.x * d21; s4i = tb.y * d12 + ta.y * d21;
                    }
                    cmac(vvr, vvi, s1r, s1i, cn1r, cn1i);
                    cmac(vhr, vhi, s2r, s2i, cn2r, cn2i);
                    cmac(hvr, hvi, -s3r, -s3i, cn2r, cn2i);
                    cmac(hhr, hhi, s4r, s4i, cn1r, cn1i);
                }
            }
        }
        t += 2 * NM * NM;
    }
    const double dk = 2.0 * pin / lam;
    vvr /= dk; vvi /= dk; vhr /= dk; vhi /= dk; hvr /= dk; hvi /= dk; hhr /= dk; hhi /= dk;
    double cvvr = vvr * R[0][0] + vhr * R[1][0], cvvi = vvi * R[0][0] + vhi * R[1][0];
    double cvhr = vvr * R[0][1] + vhr * R[1][1], cvhi = vvi * R[0][1] + vhi * R[1][1];
    double chvr = hvr * R[0][0] + hhr * R[1][0], chvi = hvi * R[0][0] + hhi * R[1][0];
    double chhr = hvr * R[0][1] + hhr * R[1][1], chhi = hvi * R[0][1] + hhi * R[1][1];
    double s11r = R1[0][0] * cvvr + R1[0][1] * chvr, s11i = R1[0][0] * cvvi + R1[0][1] * chvi;
    double s12r = R1[0][0] * cvhr + R1[0][1] * chhr, s12i = R1[0][0] * cvhi + R1[0][1] * chhi;
    double s21r = R1[1][0] * cvvr + R1[1][1] * chvr, s21i = R1[1][0] * cvvi + R1[1][1] * chvi;
    double s22r = R1[1][0] * cvhr + R1[1][1] * chhr, s22i = R1[1][0] * cvhi + R1[1][1] * chhi;
    out[0] = s11r; out[1] = s11i; out[2] = s12r; out[3] = s12i;
    out[4] = s21r; out[5] = s21i; out[6] = s22r; out[7] = s22i;
    // products a*conj(b): re = ar*br+ai*bi, im = ai*br-ar*bi
#define RE(a, b) (a##r * b##r + a##i * b##i)
#define IM(a, b) (a##i * b##r - a##r * b##i)
    double *Z = out + 8;
    Z[0] = 0.5 * (RE(s11, s11) + RE(s12, s12) + RE(s21, s21) + RE(s22, s22));
    Z[1] = 0.5 * (RE(s11, s11) - RE(s12, s12) + RE(s21, s21) - RE(s22, s22));
    Z[2] = -(RE(s11, s12) + RE(s22, s21));
    Z[3] = -(IM(s11, s12) - IM(s22, s21));
    Z[4] = 0.5 * (RE(s11, s11) + RE(s12, s12) - RE(s21, s21) - RE(s22, s22));
    Z[5] = 0.5 * (RE(s11, s11) - RE(s12, s12) - RE(s21, s21) + RE(s22, s22));
    Z[6] = -(RE(s11, s12) - RE(s22, s21));
    Z[7] = -(IM(s11, s12) + IM(s22, s21));
    Z[8] = -(RE(s11, s21) + RE(s22, s12));
    Z[9] = -(RE(s11, s21) - RE(s22, s12));
    Z[10] = RE(s11, s22) + RE(s12, s21);
    Z[11] = IM(s11, s22) + IM(s21, s12);
    Z[12] = -(IM(s21, s11) + IM(s22, s12));
    Z[13] = -(IM(s21, s11) - IM(s22, s12));
    Z[14] = IM(s22, s11) - IM(s12, s21);
    Z[15] = RE(s22, s11) - RE(s12, s21);
#undef RE
#undef IM
}

#ifndef HS_HOST_EMU
// grid.x = particles; threads over (geometry, orientation) items; dynamic smem = items_per_pass*24 doubles
__global__ void k_ampl(AmplArgs a)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_fn[HS_NPN1 + 1];
    for (int n = threadIdx.x; n <= HS_NPN1; n += blockDim.x)
        s_fn[n] = (n == 0) ? 0.0 : sqrt((double)(2 * n + 1) / (double)(n * (n + 1)));
    __syncthreads();
    const int p = blockIdx.x;
    const int n_items = a.n_geom * a.n_orient;
    const bool bad = (a.status[p] != HS_ST_OK && a.status[p] != HS_ST_NGAUSS_LIMIT && a.status[p] != HS_ST_SINGULAR) || a.toff[p] < 0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (bad) {
        long cnt = a.reduce ? (long)a.n_geom : (long)n_items;
        for (long t = threadIdx.x; t < cnt * 8; t += blockDim.x) a.S[(long)p * cnt * 8 + t] = nan;
        for (long t = threadIdx.x; t < cnt * 16; t += blockDim.x) a.Z[(long)p * cnt * 16 + t] = nan;
        return;
    }
    const double2 *t = reinterpret_cast<const double2 *>(a.tarena) + a.toff[p];
    const int nmax = a.nmax[p];
    const double lam = a.lam[p];
    double out[24];
    if (!a.reduce) {
        for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
            int g = it / a.n_orient, o = it - g * a.n_orient;
            ampl_one(t, nmax, lam, a.geom[4 * g], a.geom[4 * g + 1], a.geom[4 * g + 2], a.geom[4 * g + 3],
                     a.alpha[o], a.beta[o], s_fn, out);
            double *S = a.S + ((long)p * n_items + it) * 8, *Z = a.Z + ((long)p * n_items + it) * 16;
            for (int k = 0; k < 8; k++) S[k] = out[k];
            for (int k = 0; k < 16; k++) Z[k] = out[8 + k];
        }
        return;
    }
    // reduce: deterministic in-order weighted sum over orientations (same order as orient_averaged_fixed)
    for (int g = 0; g < a.n_geom; g++) {
        double acc = 0.0;  // thread k<24 owns component k
        for (int o0 = 0; o0 < a.n_orient; o0 += blockDim.x) {
            int o = o0 + threadIdx.x;
            if (o < a.n_orient) {
                ampl_one(t, nmax, lam, a.geom[4 * g], a.geom[4 * g + 1], a.geom[4 * g + 2], a.geom[4 * g + 3],
                         a.alpha[o], a.beta[o], s_fn, out);
                double w = a.weight[o];
                for (int k = 0; k < 24; k++) sm[threadIdx.x * 25 + k] = w * out[k];
            }
            __syncthreads();
            int cnt = min((int)blockDim.x, a.n_orient - o0);
            if (threadIdx.x < 24)
                for (int j = 0; j < cnt; j++) acc += sm[j * 25 + threadIdx.x];
            __syncthreads();
        }
        if (threadIdx.x < 8) a.S[((long)p * a.n_geom + g) * 8 + threadIdx.x] = acc * a.scale;
        else if (threadIdx.x < 24) a.Z[((long)p * a.n_geom + g) * 16 + threadIdx.x - 8] = acc * a.scale;
    }
}

#endif  // HS_HOST_EMU

}  // namespace hs

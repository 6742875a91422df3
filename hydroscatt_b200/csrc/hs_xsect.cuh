// hs_xsect.cuh -- K4: orientation-averaged scattering cross sections straight from the T-matrix.
//
// Replaces pytmatrix's scatter.sca_xsect (scipy dblquad over calcampl: thousands of amplitude evaluations per
// particle; reference call sites scattering.py:478-486 and PSDIntegrator.init_scatter_table(angular_integration
// =True), scattering_rain.py:282-283).  For one orientation the integral over the sphere of
// Z11 -/+ Z12 is the scattering cross section of an h / v polarised plane wave, which by orthogonality of the
// vector spherical harmonics is a sum of squared scattered-field expansion coefficients:
//   C = (4 pi / k^2) sum_m w_m sum_n [ u_t^2 (|G1|^2+|G2|^2) + u_p^2 (|G3|^2+|G4|^2) ],  w_0 = 1, w_m>0 = 2
//   G1 = sum_n' b_n' (T11 m pi0 + T12 tau0)   G2 = sum_n' b_n' (T21 m pi0 + T22 tau0)
//   G3 = sum_n' b_n' (T11 tau0 + T12 m pi0)   G4 = sum_n' b_n' (T21 tau0 + T22 m pi0)
// with b_n' = i^n' sqrt((2n'+1)/(n'(n'+1))), pi0/tau0 the VIGAMPL functions at the incident direction in the
// particle frame and (u_t,u_p) the incident polarisation rotated into that frame (same R matrix as AMPL).
#pragma once
#include "hs_compat.cuh"
#include "hs_ampl.cuh"

namespace hs {

struct XsArgs {
    int n;
    const double *tarena;
    const long long *toff;
    const int *nmax;
    const int *status;
    const double *lam;
    int n_inc;
    const double *inc;  // [n_inc][2] thet0, phi0 (deg)
    int n_orient;
    const double *alpha, *beta, *weight;
    double scale;
    double *out;  // [n][n_inc][2]  sca_xsect_h, sca_xsect_v
};

// out2[0] = C_sca (h-pol), out2[1] = C_sca (v-pol) for one orientation
__device__ void xsect_one(const double2 *__restrict__ t, int nmax, double lam, double thet0, double phi0,
                          double alpha, double beta, const double *fn, double *out2)
{
    const double pin = 3.14159265358979323846, pin2 = pin * 0.5, rad = pin / 180.0;
    double alph = alpha * rad, bet = beta * rad, thetl = thet0 * rad, phil = phi0 * rad;
    const double eps = 1e-7;
    if (thetl < pin2) thetl += eps;
    if (thetl > pin2) thetl -= eps;
    if (phil < pin) phil += eps;
    if (phil > pin) phil -= eps;
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
    double ca = cos(alph), sa = sin(alph);
    double B[3][3] = {{ca * cb, sa * cb, -sb}, {-sa, ca, 0.0}, {ca * sb, sa * sb, cb}};
    cp = cos(phil); sp = sin(phil);
    double AL[3][2] = {{ct * cp, -sp}, {ct * sp, cp}, {-st, 0.0}};
    double stp = sin(thetp), cpp_ = cos(phip), spp_ = sin(phip);
    double AP[2][3] = {{ctp * cpp_, ctp * spp_, -stp}, {-spp_, cpp_, 0.0}};
    double C[3][2], R[2][2];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += B[i][k] * AL[k][j]; C[i][j] = x; }
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += AP[i][k] * C[k][j]; R[i][j] = x; }
    // incident polarisation in the particle frame: column 1 = h, column 0 = v
    const double uth = R[0][1], uph = R[1][1], utv = R[0][0], upv = R[1][0];

    double dv01[HS_NPN1 + 1], dv02[HS_NPN1 + 1];
    double acc_t = 0.0, acc_p = 0.0;  // sum w_m (|G1|^2+|G2|^2), sum w_m (|G3|^2+|G4|^2)
    for (int m = 0; m <= nmax; m++) {
        const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1;
        vigampl(ctp, nmax, m, dv01, dv02);
        const double wm = (m == 0) ? 1.0 : 2.0;
        for (int n = nmin; n <= nmax; n++) {
            const int k = n - nmin;
            const int ca_ = (n + 1) & 1, cb_ = n & 1;
            double g1r = 0, g1i = 0, g2r = 0, g2i = 0, g3r = 0, g3i = 0, g4r = 0, g4i = 0;
            for (int nn = nmin; nn <= nmax; nn++) {
                const int kp = nn - nmin;
                const double2 ta = t[(long)(ca_ * NM + kp) * NM + k];
                const double2 tb = t[(long)(cb_ * NM + kp) * NM + k];
                const double mp = m * dv01[nn], ta0 = dv02[nn];
                const bool even = (((n + nn) & 1) == 0);
                const double e = even ? mp : ta0, o = even ? ta0 : mp;
                // b = i^nn * fn[nn]
                const double f = fn[nn];
                const int pw = nn & 3;
                const double br = (pw == 0) ? f : (pw == 2) ? -f : 0.0;
                const double bi = (pw == 1) ? f : (pw == 3) ? -f : 0.0;
                const double tar = ta.x * br - ta.y * bi, tai = ta.x * bi + ta.y * br;
                const double tbr = tb.x * br - tb.y * bi, tbi = tb.x * bi + tb.y * br;
                g1r += tar * e; g1i += tai * e;
                g3r += tar * o; g3i += tai * o;
                g2r += tbr * o; g2i += tbi * o;
                g4r += tbr * e; g4i += tbi * e;
            }
            acc_t += wm * (g1r * g1r + g1i * g1i + g2r * g2r + g2i * g2i);
            acc_p += wm * (g3r * g3r + g3i * g3i + g4r * g4r + g4i * g4i);
        }
        t += 2 * NM * NM;
    }
    const double kk = 2.0 * pin / lam;
    const double c0 = 4.0 * pin / (kk * kk);
    out2[0] = c0 * (uth * uth * acc_t + uph * uph * acc_p);
    out2[1] = c0 * (utv * utv * acc_t + upv * upv * acc_p);
}

#ifndef HS_HOST_EMU
// grid.x = particles; threads over orientations; in-order weighted reduction as in k_ampl
__global__ void k_xsect(XsArgs a)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_fn[HS_NPN1 + 1];
    for (int n = threadIdx.x; n <= HS_NPN1; n += blockDim.x)
        s_fn[n] = (n == 0) ? 0.0 : sqrt((double)(2 * n + 1) / (double)(n * (n + 1)));
    __syncthreads();
    const int p = blockIdx.x;
    const bool bad = (a.status[p] != HS_ST_OK && a.status[p] != HS_ST_NGAUSS_LIMIT && a.status[p] != HS_ST_SINGULAR) || a.toff[p] < 0;
    if (bad) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int t = threadIdx.x; t < a.n_inc * 2; t += blockDim.x) a.out[(long)p * a.n_inc * 2 + t] = nan;
        return;
    }
    const double2 *t = reinterpret_cast<const double2 *>(a.tarena) + a.toff[p];
    const int nmax = a.nmax[p];
    const double lam = a.lam[p];
    for (int g = 0; g < a.n_inc; g++) {
        double acc = 0.0;
        for (int o0 = 0; o0 < a.n_orient; o0 += blockDim.x) {
            int o = o0 + threadIdx.x;
            if (o < a.n_orient) {
                double r2[2];
                xsect_one(t, nmax, lam, a.inc[2 * g], a.inc[2 * g + 1], a.alpha[o], a.beta[o], s_fn, r2);
                double w = a.weight[o];
                sm[threadIdx.x * 2] = w * r2[0];
                sm[threadIdx.x * 2 + 1] = w * r2[1];
            }
            __syncthreads();
            int cnt = min((int)blockDim.x, a.n_orient - o0);
            if (threadIdx.x < 2)
                for (int j = 0; j < cnt; j++) acc += sm[j * 2 + threadIdx.x];
            __syncthreads();
        }
        if (threadIdx.x < 2) a.out[((long)p * a.n_inc + g) * 2 + threadIdx.x] = acc * a.scale;
    }
}
#endif

}  // namespace hs

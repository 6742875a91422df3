// hs_psd.cuh -- K4: PSD weights, PSD integration (small FP64 GEMM over diameters) and radar observables.
//
// Replaces pytmatrix PSDIntegrator.get_SZ / get_angular_integrated (numpy trapz per DSD), the
// PSD classes' N(D) evaluation (row P11), the rectangle-rule sums of scattering.py:312-448 (row S5) and the
// radar.* / scatter.* observable formulas as used by scattering.py:530-587 (rows P12, S6).
#pragma once
#include "hs_compat.cuh"
#include "../../include/hydrotm.h"

namespace hs {

#ifndef HS_HOST_EMU

// W[d][i] = wq[i] * N_d(D_i).  kind 0: normalised gamma (p0=D0, p1=Nw, p2=mu), 1: exponential (p0=N0,
// p1=Lambda), 2: unnormalised gamma (p0=N0, p1=Lambda, p2=mu).  N = 0 above dmax[d] (and at D=0 for kinds 0, 2).
__global__ void k_psd_weights(int kind, int n_dsd, const double *p0, const double *p1, const double *p2,
                              const double *dmax, int n_D, const double *D, const double *wq, double *W)
{
    long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_dsd * n_D) return;
    int d = (int)(idx / n_D), i = (int)(idx - (long)d * n_D);
    double Di = D[i], v;
    if (kind == 0) {
        double D0 = p0[d], Nw = p1[d], mu = p2[d];
        double nf = Nw * 6.0 / (3.67 * 3.67 * 3.67 * 3.67) * pow(3.67 + mu, mu + 4.0) / tgamma(mu + 4.0);
        double x = Di / D0;
        v = nf * exp(mu * log(x) - (3.67 + mu) * x);
        if (Di > dmax[d] || Di == 0.0) v = 0.0;
    } else if (kind == 1) {
        v = p0[d] * exp(-p1[d] * Di);
        if (Di > dmax[d]) v = 0.0;
    } else {
        v = p0[d] * exp(p2[d] * log(Di) - p1[d] * Di);
        if (Di > dmax[d] || Di == 0.0) v = 0.0;
    }
    W[idx] = wq[i] * v;
}

// out[d][c] = sum_i W[d][i] * Tt[i][c]   (W: [n_dsd][n_D], Tt: [n_D][ncol], out: [n_dsd][ncol]); i ascending.
// Block = 32 column lanes x 8 row groups; each thread owns an 8 x 2 register tile.
#define PSD_RB 64
#define PSD_CB 64
__global__ void __launch_bounds__(256) k_psd_gemm(int n_dsd, int n_D, int ncol, const double *__restrict__ W,
                                                   const double *__restrict__ Tt, double *__restrict__ out)
{
    __shared__ double sW[PSD_RB][33];
    __shared__ double sT[32][PSD_CB];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.y * PSD_RB, c0 = blockIdx.x * PSD_CB;
    double acc[8][2];
#pragma unroll
    for (int r = 0; r < 8; r++) { acc[r][0] = 0.0; acc[r][1] = 0.0; }
    for (int i0 = 0; i0 < n_D; i0 += 32) {
        // stage W tile [64 rows][32 i] and T tile [32 i][64 cols]
        for (int e = threadIdx.x; e < PSD_RB * 32; e += 256) {
            int r = e >> 5, ii = e & 31;
            int gr = r0 + r, gi = i0 + ii;
            sW[r][ii] = (gr < n_dsd && gi < n_D) ? W[(long)gr * n_D + gi] : 0.0;
        }
        for (int e = threadIdx.x; e < 32 * PSD_CB; e += 256) {
            int ii = e >> 6, cc = e & 63;
            int gi = i0 + ii, gc = c0 + cc;
            sT[ii][cc] = (gi < n_D && gc < ncol) ? Tt[(long)gi * ncol + gc] : 0.0;
        }
        __syncthreads();
        const int lim = min(32, n_D - i0);
        for (int ii = 0; ii < lim; ii++) {
            double t0 = sT[ii][tx], t1 = sT[ii][tx + 32];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                double w = sW[ty * 8 + r][ii];
                acc[r][0] += w * t0;
                acc[r][1] += w * t1;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        int gr = r0 + ty * 8 + r;
        if (gr >= n_dsd) continue;
        if (c0 + tx < ncol) out[(long)gr * ncol + c0 + tx] = acc[r][0];
        if (c0 + tx + 32 < ncol) out[(long)gr * ncol + c0 + tx + 32] = acc[r][1];
    }
}

// Radar observables of compute_scattering_psd_tm / compute_scattering_sp_tm (scattering.py:451-587).
// in: [n][stride] rows holding, at column offsets: back geometry S(8) Z(16) at ob, forward geometry S(8) Z(16) at of,
// optional ext_xsect_h/v at oe (oe < 0: use 2*lambda*Im S22 / Im S11 of the forward S), sca_xsect_h/v at os_ (or < 0).
// out: [n][15] = sca_xsect_h, sca_xsect_v [dB], refl_h, refl_v, zdr, ldr_h, ldr_v [dB], rho_hv, delta_hv [deg],
//                ext_xsect_h, ext_xsect_v [dB], kdp, A_h, A_v, Adp
__global__ void k_observables(int n, int stride, const double *in, int ob, int of, int oe, int os_,
                              const double *lam, int lam_stride, double kw_sqr, double *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double PI = 3.14159265358979323846;
    const double *row = in + (long)i * stride;
    const double *Zb = row + ob + 8, *Sf = row + of;
    const double wl = lam[(long)i * lam_stride];
    double xs_h = 2.0 * PI * (Zb[0] - Zb[1] - Zb[4] + Zb[5]);
    double xs_v = 2.0 * PI * (Zb[0] + Zb[1] + Zb[4] + Zb[5]);
    double pref = wl * wl * wl * wl / (PI * PI * PI * PI * PI * kw_sqr);
    double *o = out + (long)i * 15;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    o[0] = os_ >= 0 ? 10.0 * log10(row[os_] / 1e6) : nan;
    o[1] = os_ >= 0 ? 10.0 * log10(row[os_ + 1] / 1e6) : nan;
    o[2] = 10.0 * log10(pref * xs_h);
    o[3] = 10.0 * log10(pref * xs_v);
    o[4] = 10.0 * log10(xs_h / xs_v);
    o[5] = 10.0 * log10((Zb[0] - Zb[1] + Zb[4] - Zb[5]) / (Zb[0] - Zb[1] - Zb[4] + Zb[5]));
    o[6] = 10.0 * log10((Zb[0] + Zb[1] - Zb[4] - Zb[5]) / (Zb[0] + Zb[1] + Zb[4] + Zb[5]));
    double a = (Zb[10] + Zb[15]) * (Zb[10] + Zb[15]) + (Zb[14] - Zb[11]) * (Zb[14] - Zb[11]);
    double b = (Zb[0] - Zb[1] - Zb[4] + Zb[5]), c = (Zb[0] + Zb[1] + Zb[4] + Zb[5]);
    o[7] = sqrt(a / (b * c));
    o[8] = 180.0 / PI * atan2(Zb[11] - Zb[14], -Zb[10] - Zb[15]);
    double ext_h = oe >= 0 ? row[oe] : 2.0 * wl * Sf[7];
    double ext_v = oe >= 0 ? row[oe + 1] : 2.0 * wl * Sf[1];
    o[9] = 10.0 * log10(ext_h / 1e6);
    o[10] = 10.0 * log10(ext_v / 1e6);
    o[11] = 1e-3 * (180.0 / PI) * wl * (Sf[6] - Sf[0]);
    double Ah = 2.0 * 4.343e-3 * ext_h, Av = 2.0 * 4.343e-3 * ext_v;
    o[12] = Ah;
    o[13] = Av;
    o[14] = Ah - Av;
}

#endif  // HS_HOST_EMU

}  // namespace hs

// hs_cant.cuh -- canting-angle ("angular moments") observables: the batched equivalent of hydroscatt's
//   compute_scattering_canting_sp   (scattering.py:182-309): one row of observables per diameter,
//   compute_scattering_canting_psd  (scattering.py:312-448, row S5): rectangle-rule sums over the diameters for one DSD,
//                                    called 9 300x by scattering_rain.py:295-316 and 11 011x by scattering_hail.py:298-372,
//   compute_scattering_mixture      (scattering.py:590-673, row S7): incoherent two-species mixture.
//
// Per diameter the reference forms j0 = |fh180|^2, j1 = |fh180 - fv180|^2, j2 = conj(fh180)(fh180 - fv180),
// j3 = fh0 - fv0 (scattering.py:227-235) and, with the angular moments a1..a5, a7 of the canting distribution, the eight
// real "canting terms" below; every PSD observable is a function of the psd_vals*delta_d-weighted sums of those terms,
// so one [n_dsd x n_D] x [n_D x 8] FP64 product per table replaces the Python loop over the DSDs.  HBM-bound: the
// weights W (n_dsd x n_D doubles) are read once per table, the terms stay in L2 / shared memory.
//
// Observable columns (order of the reference's DataFrame / dict, scattering.py:277-307, 417-446):
//   0 sca_xsect_h  1 sca_xsect_v  2 ext_xsect_h  3 ext_xsect_v  4 ldr_h  5 ldr_v  6 refl_h  7 refl_v  8 zdr   [dB]
//   9 rho_hv  10 delta_hv [deg]  11 kdp [deg/km]  12 A_h  13 A_v  14 Adp [dB/km]
#pragma once
#include "hs_compat.cuh"
#include "../../include/hydrotm.h"

namespace hs {

#ifndef HS_HOST_EMU

#define CANT_NT 8    // canting terms per diameter
#define CANT_NX 4    // at most this many caller-supplied extra columns (e.g. D^3, D^3 v for lwc / rainfall rate)

// terms of one diameter: f = (fv180, fh180, fv0, fh0) as re, im pairs; a = (a1, a2, a3, a4, a5, a7)
__device__ __forceinline__ void cant_terms(const double *f, const double *a, double *t)
{
    const double vr = f[0], vi = f[1], hr = f[2], hi = f[3];
    const double j0 = hr * hr + hi * hi;
    const double dr = hr - vr, di = hi - vi;                  // fh180 - fv180
    const double j1 = dr * dr + di * di;
    const double j2r = hr * dr + hi * di, j2i = hr * di - hi * dr;  // conj(fh180) (fh180 - fv180)
    const double j3r = f[6] - f[4], j3i = f[7] - f[5];        // fh0 - fv0
    const double a1 = a[0], a2 = a[1], a3 = a[2], a4 = a[3], a5 = a[4], a7 = a[5];
    t[0] = j0 - 2.0 * j2r * a2 + j1 * a4;                     // sca_xsect_h / 4 pi   (scattering.py:242, 376)
    t[1] = j0 - 2.0 * j2r * a1 + j1 * a3;                     // sca_xsect_v / 4 pi   (:246, 380)
    t[2] = f[7] - j3i * a2;                                   // ext_xsect_h / (2 lambda / 1000)   (:248, 383)
    t[3] = f[7] - j3i * a3;                                   // ext_xsect_v          (:250, 386)
    t[4] = j0 + j1 * a5 - j2r * a1 - j2r * a2;                // Re delta_co          (:237, 371)
    t[5] = -j2i * a1 + j2i * a2;                              // Im delta_co  (conj(j2) a2 contributes +j2i a2)
    t[6] = j3r * a7;                                          // kdp / (180 / pi lambda)  (:256, 396)
    t[7] = j1 * a5;                                           // ldr numerator / 4 pi (:262, 403)
}

// the 15 observables from the (summed) terms
__device__ __forceinline__ void cant_obs(const double *s, double wl, double kw_sqr, double *o)
{
    const double PI = 3.14159265358979323846;
    const double sca_h = 4.0 * PI * s[0], sca_v = 4.0 * PI * s[1];
    const double ext_h = 2.0 * wl / 1000.0 * s[2], ext_v = 2.0 * wl / 1000.0 * s[3];
    const double pref = wl * wl * wl * wl * 1e6 / (PI * PI * PI * PI * PI * kw_sqr);
    o[0] = 10.0 * log10(sca_h);
    o[1] = 10.0 * log10(sca_v);
    o[2] = 10.0 * log10(ext_h);
    o[3] = 10.0 * log10(ext_v);
    o[4] = 10.0 * log10(4.0 * PI * s[7] / sca_h);
    o[5] = 10.0 * log10(4.0 * PI * s[7] / sca_v);
    o[6] = 10.0 * log10(pref * sca_h);
    o[7] = 10.0 * log10(pref * sca_v);
    o[8] = 10.0 * log10(sca_h / sca_v);
    o[9] = 4.0 * PI * sqrt(s[4] * s[4] + s[5] * s[5]) / sqrt(sca_h * sca_v);
    o[10] = -(180.0 / PI) * atan2(s[5], s[4]);
    o[11] = s[6] * 180.0 / PI * wl;
    const double ah = 2.0 * 4.34e3 * ext_h, av = 2.0 * 4.34e3 * ext_v;
    o[12] = ah;
    o[13] = av;
    o[14] = ah - av;
}

// Single-particle form (compute_scattering_canting_sp): thread = (table, diameter).
//   f [n_tab][n_D][8], mom [..][6] addressed as mom[t * ms_tab + i * ms_D], out [n_tab][n_D][15]
__global__ void __launch_bounds__(128) k_canting_sp(int n_tab, int n_D, const double *__restrict__ f, const double *__restrict__ mom,
                                                     long ms_tab, long ms_D, const double *__restrict__ lam_tab, double kw_sqr,
                                                     double *__restrict__ out)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_tab * n_D) return;
    const int t = (int)(idx / n_D), i = (int)(idx - (long)t * n_D);
    double tt[CANT_NT], o[15];
    cant_terms(f + idx * 8, mom + t * ms_tab + i * ms_D, tt);
    cant_obs(tt, lam_tab[t], kw_sqr, o);
#pragma unroll
    for (int k = 0; k < 15; k++) out[idx * 15 + k] = o[k];
}

// PSD form (compute_scattering_canting_psd for every (DSD, table)): CTA = 32 DSDs x one table, thread = (DSD, column),
// column = 8 canting terms + up to 4 extra columns + padding to 16.  Diameters in chunks of 32: the W chunk is read
// coalesced (one DSD row segment per warp) and staged transposed-free in shared memory, the term chunk is computed on the
// fly from f and the moments (8 + 6 doubles per diameter).  Sum over i ascending.
//   W [n_dsd][n_D] = psd_vals * delta_d;  extra [n_D][n_extra] (or null);  out [n_dsd][n_tab][15];
//   out_extra [n_dsd][n_extra] (table-independent, written by the blockIdx.y == 0 CTAs)
#define CANT_RB 32
#define CANT_NC 16
__global__ void __launch_bounds__(CANT_RB * CANT_NC) k_canting_psd(int n_dsd, int n_D, int n_tab, const double *__restrict__ W,
                                                                    const double *__restrict__ f, const double *__restrict__ mom,
                                                                    long ms_tab, long ms_D, const double *__restrict__ lam_tab,
                                                                    double kw_sqr, int n_extra, const double *__restrict__ extra,
                                                                    double *__restrict__ out, double *__restrict__ out_extra)
{
    __shared__ double sW[CANT_RB][33];
    __shared__ double sT[32][CANT_NC + 1];
    __shared__ double sS[CANT_RB][CANT_NC + 1];
    const int tid = threadIdx.x, c = tid & (CANT_NC - 1), r = tid / CANT_NC;
    const int r0 = blockIdx.x * CANT_RB, t = blockIdx.y;
    double acc = 0.0;
    for (int i0 = 0; i0 < n_D; i0 += 32) {
        for (int e = tid; e < CANT_RB * 32; e += CANT_RB * CANT_NC) {
            const int rr = e >> 5, ii = e & 31;
            sW[rr][ii] = (r0 + rr < n_dsd && i0 + ii < n_D) ? W[(long)(r0 + rr) * n_D + i0 + ii] : 0.0;
        }
        if (tid < 32) {
            const int i = i0 + tid;
            double tt[CANT_NT];
            if (i < n_D) cant_terms(f + ((long)t * n_D + i) * 8, mom + t * ms_tab + i * ms_D, tt);
#pragma unroll
            for (int k = 0; k < CANT_NT; k++) sT[tid][k] = i < n_D ? tt[k] : 0.0;
            for (int k = 0; k < CANT_NC - CANT_NT; k++) sT[tid][CANT_NT + k] = (i < n_D && k < n_extra) ? extra[(long)i * n_extra + k] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int ii = 0; ii < 32; ii++) acc += sW[r][ii] * sT[ii][c];
        __syncthreads();
    }
    sS[r][c] = acc;
    __syncthreads();
    if (r0 + r < n_dsd) {
        if (c == 0) {
            double s[CANT_NT], o[15];
#pragma unroll
            for (int k = 0; k < CANT_NT; k++) s[k] = sS[r][k];
            cant_obs(s, lam_tab[t], kw_sqr, o);
            double *op = out + ((long)(r0 + r) * n_tab + t) * 15;
#pragma unroll
            for (int k = 0; k < 15; k++) op[k] = o[k];
        } else if (t == 0 && out_extra && c >= CANT_NT && c - CANT_NT < n_extra) {
            out_extra[(long)(r0 + r) * n_extra + (c - CANT_NT)] = acc;
        }
    }
}

// compute_scattering_mixture (scattering.py:590-673) on rows of the 15 canting observables; the columns the reference's
// mixture does not define (cross sections) are NaN.
__global__ void __launch_bounds__(128) k_canting_mixture(long n, const double *__restrict__ A, const double *__restrict__ B,
                                                          double *__restrict__ out)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *a = A + i * 15, *b = B + i * 15;
    double *o = out + i * 15;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double ah = pow(10.0, 0.1 * a[6]), av = pow(10.0, 0.1 * a[7]), bh = pow(10.0, 0.1 * b[6]), bv = pow(10.0, 0.1 * b[7]);
    const double refl_h = 10.0 * log10(ah + bh), refl_v = 10.0 * log10(av + bv);
    o[0] = o[1] = o[2] = o[3] = nan;
    o[4] = 10.0 * log10(pow(10.0, 0.1 * a[4]) * ah + pow(10.0, 0.1 * b[4]) * bh) - refl_h;
    o[5] = 10.0 * log10(pow(10.0, 0.1 * a[5]) * av + pow(10.0, 0.1 * b[5]) * bv) - refl_v;
    o[6] = refl_h;
    o[7] = refl_v;
    o[8] = refl_h - refl_v;
    o[9] = (a[9] * sqrt(ah * av) + b[9] * sqrt(bh * bv)) / sqrt(pow(10.0, 0.1 * refl_h) * pow(10.0, 0.1 * refl_v));
    o[10] = a[10] + b[10];
    o[11] = a[11] + b[11];
    o[12] = a[12] + b[12];
    o[13] = a[13] + b[13];
    o[14] = a[14] + b[14];
}

#endif  // HS_HOST_EMU

}  // namespace hs

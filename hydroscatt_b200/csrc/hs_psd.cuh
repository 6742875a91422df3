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


// ---- fused PSD integration + observables ----
// The observables of compute_scattering_psd_tm (scattering.py:530-587) read 14 of the 56 integrated columns of a table
// record (8 phase-matrix elements of the back geometry, 4 amplitude components of the forward geometry, the two
// extinction cross sections; + the 2 scattering cross sections when asked for).  This kernel integrates exactly those
// over the diameters (same i-ascending sum as k_psd_gemm) for a tile of 64 DSDs x 4 tables at a time and turns them
// into the 15 observables on the spot, so the [n_dsd][n_tab][56] array (1 GB for the rain sweep) is never written:
// HBM sees the table records (L2-resident), the weights and the observables.
// rows: [n_tab][n_D][56] table records (tables.py layout); W: [n_dsd][n_D]; obs: [n_dsd][n_tab][15].
#define PSDF_RB 64
#define PSDF_TB 4
#define PSDF_NC 16
__constant__ int c_psdf_col[PSDF_NC] = {8, 9, 12, 13, 18, 19, 22, 23, 24, 25, 30, 31, 54, 55, 48, 49};

// caller-selected observable columns (hs_psd_observables_sel): n columns, in the caller's order
struct PsdSel { int n; signed char col[15]; };

// DMMA = true: the 64 x 64 x n_D tile product on the FP64 tensor pipe (mma.sync.m8n8k4.f64): warp w owns rows 8 w .. 8 w + 7 of all
// eight column tiles; the table tile is staged transposed ([column][n_D | 1]) so that A and B fragments are read with the same
// two-wavefront pattern.  Same flops as the FMA loop, 5x fewer issue slots and 2.6x fewer shared-memory instructions.
template <bool DMMA>
__global__ void __launch_bounds__(256) k_psd_obs_fused(int n_dsd, int n_D, int n_tab, int tabs_per_cta, const double *__restrict__ W,
                                                        const double *__restrict__ rows, const double *__restrict__ lam_tab,
                                                        double kw_sqr, int hpol_quirk, int want_sca, PsdSel sel, double *__restrict__ obs)
{
    extern __shared__ double psm[];
    const int ldw = n_D | 1;
    double *sW = psm;                                  // [64][ldw]
    // one area, three uses per tile, separated by barriers: [n_D][64] table tile (4 tables x 16 columns) -> [64][65]
    // integrated values -> [64][61] staged observables.  (Three separate areas held the kernel to two CTAs per SM.)
    double *sT = sW + PSDF_RB * ldw;
    double *sO = sT;
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * PSDF_RB;
    const int t_begin = blockIdx.y * tabs_per_cta, t_end = min(n_tab, t_begin + tabs_per_cta);
    for (int e = tid; e < PSDF_RB * n_D; e += 256) {
        const int r = e / n_D, i = e - r * n_D;
        sW[r * ldw + i] = (r0 + r < n_dsd) ? W[(long)(r0 + r) * n_D + i] : 0.0;
    }
    // thread tile: rows {rr, rr + 32}, 8 of the 64 (table, column) slots
    const int rr = tid & 31, cg = tid >> 5;
    const double PI = 3.14159265358979323846;
    for (int t0 = t_begin; t0 < t_end; t0 += PSDF_TB) {
        __syncthreads();  // previous observables phase has read sO; first pass: sW complete
        for (int e = tid; e < n_D * (PSDF_TB * PSDF_NC); e += 256) {
            const int i = e >> 6, sc = e & 63, tt = sc >> 4, c = sc & 15;
            const int t = t0 + tt;
            const double v = (t < t_end) ? rows[((long)t * n_D + i) * 56 + c_psdf_col[c]] : 0.0;
            if (DMMA) sT[sc * ldw + i] = v;
            else sT[e] = v;
        }
        __syncthreads();
        if (DMMA) {
            const int warp = tid >> 5, lane = tid & 31, fr = lane >> 2, fk = lane & 3;
            double c[8][2];
#pragma unroll
            for (int q = 0; q < 16; q++) c[q >> 1][q & 1] = 0.0;
            const double *ap = sW + (8 * warp + fr) * ldw + fk;
            const double *bp = sT + fr * ldw + fk;
            // every lane of the warp must reach each mma.sync: no lane-dependent branch around it -- the ragged last step
            // reads a clamped index and zeroes the fragment element by a select
            for (int k0 = 0; k0 < n_D; k0 += 4) {
                const bool kv = k0 + fk < n_D;
                const int kk = kv ? k0 : 0;
                double av = ap[kk];
                av = kv ? av : 0.0;
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    double bv = bp[nt * 8 * ldw + kk];
                    bv = kv ? bv : 0.0;
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[nt][0]), "+d"(c[nt][1]) : "d"(av), "d"(bv));
                }
            }
            __syncthreads();  // every warp is done with the table tile
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                sO[(8 * warp + fr) * 65 + 8 * nt + 2 * fk] = c[nt][0];
                sO[(8 * warp + fr) * 65 + 8 * nt + 2 * fk + 1] = c[nt][1];
            }
        } else {
            double acc[2][8];
#pragma unroll
            for (int q = 0; q < 16; q++) acc[q >> 3][q & 7] = 0.0;
            for (int i = 0; i < n_D; i++) {
                const double w0 = sW[rr * ldw + i], w1 = sW[(rr + 32) * ldw + i];
                const double2 *tp = reinterpret_cast<const double2 *>(sT + i * 64 + cg * 8);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double2 tv = tp[q];
                    acc[0][2 * q] += w0 * tv.x; acc[0][2 * q + 1] += w0 * tv.y;
                    acc[1][2 * q] += w1 * tv.x; acc[1][2 * q + 1] += w1 * tv.y;
                }
            }
            __syncthreads();  // every warp is done with the table tile
#pragma unroll
            for (int q = 0; q < 8; q++) { sO[rr * 65 + cg * 8 + q] = acc[0][q]; sO[(rr + 32) * 65 + cg * 8 + q] = acc[1][q]; }
        }
        __syncthreads();
        // observables: thread = (row, table)
        {
            const int r = tid & 63, tt = tid >> 6, t = t0 + tt;
            double v[16];
#pragma unroll
            for (int q = 0; q < 16; q++) v[q] = sO[r * 65 + tt * 16 + q];
            __syncthreads();  // the staged observables overwrite the integrated values
            if (r0 + r < n_dsd && t < t_end) {
                const double z0 = v[0], z1 = v[1], z4 = v[2], z5 = v[3], z10 = v[4], z11 = v[5], z14 = v[6], z15 = v[7];
                const double wl = lam_tab[t];
                const double xs_h = 2.0 * PI * (z0 - z1 - z4 + z5), xs_v = 2.0 * PI * (z0 + z1 + z4 + z5);
                const double pref = wl * wl * wl * wl / (PI * PI * PI * PI * PI * kw_sqr);
                double *o = sT + r * 61 + tt * 15;  // staged: [64 rows][4 tables x 15] (+1 pad), sT is free until the next tile
                const double nan = __longlong_as_double(0x7ff8000000000000LL);
                o[0] = want_sca ? 10.0 * log10(v[14] / 1e6) : nan;
                o[1] = want_sca ? 10.0 * log10(v[15] / 1e6) : nan;
                o[2] = 10.0 * log10(pref * xs_h);
                o[3] = 10.0 * log10(pref * xs_v);
                o[4] = 10.0 * log10(xs_h / xs_v);
                o[5] = 10.0 * log10((z0 - z1 + z4 - z5) / (z0 - z1 - z4 + z5));
                o[6] = 10.0 * log10((z0 + z1 - z4 - z5) / (z0 + z1 + z4 + z5));
                const double a = (z10 + z15) * (z10 + z15) + (z14 - z11) * (z14 - z11);
                const double b = (z0 - z1 - z4 + z5), c = (z0 + z1 + z4 + z5);
                o[7] = sqrt(a / (b * c));
                o[8] = 180.0 / PI * atan2(z11 - z14, -z10 - z15);
                const double ext_h = v[12], ext_v = hpol_quirk ? v[12] : v[13];
                o[9] = 10.0 * log10(ext_h / 1e6);
                o[10] = 10.0 * log10(ext_v / 1e6);
                o[11] = 1e-3 * (180.0 / PI) * wl * (v[10] - v[8]);
                const double Ah = 2.0 * 4.343e-3 * ext_h, Av = 2.0 * 4.343e-3 * ext_v;
                o[12] = Ah;
                o[13] = Av;
                o[14] = Ah - Av;
            }
        }
        __syncthreads();
        // coalesced store: the 4 tables x n_sel observables of a DSD are contiguous in obs
        {
            const int nt_ = min(PSDF_TB, t_end - t0), rowlen = nt_ * sel.n;
            if (sel.n == 15) {
                for (int e = tid; e < PSDF_RB * rowlen; e += 256) {
                    const int r = e / rowlen, k = e - r * rowlen;
                    if (r0 + r < n_dsd) obs[((long)(r0 + r) * n_tab + t0) * 15 + k] = sT[r * 61 + k];
                }
            } else {
                for (int e = tid; e < PSDF_RB * rowlen; e += 256) {
                    const int r = e / rowlen, k = e - r * rowlen, tt = k / sel.n, q = k - tt * sel.n;
                    if (r0 + r < n_dsd) obs[((long)(r0 + r) * n_tab + t0) * sel.n + k] = sT[r * 61 + tt * 15 + sel.col[q]];
                }
            }
        }
    }
}

// Table records (tables.py layout, 56 doubles) from the outputs of hs_scatter_batch: S [n][G][2][2] complex, Z [n][G][4][4],
// xs [n][G][2]; gb / gf = geometry index of the back / forward geometry, fb / ff = index of the geometry holding the forward
// amplitude at the back / forward geometry's incidence direction (optical theorem: ext_h = 2 lambda Im S22, ext_v = 2 lambda Im S11).
__global__ void k_pack_rows(int n, int G, const double *__restrict__ S, const double *__restrict__ Z, const double *__restrict__ xs,
                            const double *__restrict__ lam, int gb, int gf, int fb, int ff, double *__restrict__ rows)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n * 56) return;
    const int i = (int)(idx / 56), c = (int)(idx - (long)i * 56);
    double v;
    if (c < 48) {
        const int g = c < 24 ? gb : gf, k = c < 24 ? c : c - 24;
        v = k < 8 ? S[((long)i * G + g) * 8 + k] : Z[((long)i * G + g) * 16 + (k - 8)];
    } else if (c < 52) {
        const int g = c < 50 ? gb : gf;
        v = xs[((long)i * G + g) * 2 + (c & 1)];
    } else {
        const int f = c < 54 ? fb : ff;
        v = 2.0 * lam[i] * S[((long)i * G + f) * 8 + ((c & 1) ? 1 : 7)];
    }
    rows[idx] = v;
}

// dst[i][:] = src[index[i]][:]  (ncol doubles per record; index < 0: record left untouched).  Puts all-gathered shard records
// back into table order (hydroscatt_b200/dist.py).
__global__ void k_gather_rows(long n_out, int ncol, const double *__restrict__ src, const long long *__restrict__ index,
                              double *__restrict__ dst)
{
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_out * ncol) return;
    const long i = idx / ncol;
    const int c = (int)(idx - i * ncol);
    const long long j = index[i];
    if (j >= 0) dst[idx] = src[j * ncol + c];
}

#endif  // HS_HOST_EMU

}  // namespace hs

// hs_scatter.cuh -- K4 fused: orientation-averaged amplitude / phase matrices for several scattering geometries
// AND the h/v scattering cross sections, from one pass over the T-matrix per orientation.
//
// Replaces (batched) pytmatrix calcampl + orientation.orient_averaged_fixed + scatter.sca_xsect
// (reference call sites scattering.py:478-525 and
// PSDIntegrator.init_scatter_table(angular_integration=True) at scattering_rain.py:282-283).
//
// AMPL's double sum  sum_{n,n'} CAL(n,n') [..T_{nn'}..]  factorises into an incident-side vector
//   G1_n = sum_n' b_n' (T11 m pi0_n' + T12 tau0_n')   G2_n = sum_n' b_n' (T21 m pi0_n' + T22 tau0_n')
//   G3_n = sum_n' b_n' (T11 tau0_n' + T12 m pi0_n')   G4_n = sum_n' b_n' (T21 tau0_n' + T22 m pi0_n')
// (b_n' = i^n' f_n', O(NMAX^2) per mode, depends only on orientation + incident direction) and an O(NMAX) sum per
// scattered direction:
//   VV += fc a_n (m pi_n G1 + tau_n G2)    VH += fs a_n (m pi_n G3 + tau_n G4)
//   HV -= fs a_n (tau_n G1 + m pi_n G2)    HH += fc a_n (tau_n G3 + m pi_n G4)      a_n = i^(-n-1) f_n
// The same G's give the cross sections: C = 4 pi/k^2 sum_m w_m sum_n [u_t^2(|G1|^2+|G2|^2) + u_p^2(|G3|^2+|G4|^2)].
// So back-scatter, forward-scatter and sca_xsect of one orientation cost one T-matrix sweep instead of three.
#pragma once
#include "hs_compat.cuh"
#include "hs_ampl.cuh"

namespace hs {

#define SC_MAXG 4  // scattered directions per incident direction handled in one sweep

struct ScArgs {
    int n;
    const double *tarena;
    const long long *toff;
    const int *nmax;
    const int *status;
    const double *lam;
    int n_inc;
    const double *inc;      // [n_inc][2] thet0, phi0
    const int *g_begin;     // [n_inc+1] geometry range of each incident direction (geometries sorted by inc)
    const double *gsc;      // [n_geom][2] thet, phi of each (sorted) geometry
    const int *g_out;       // [n_geom] output slot of each sorted geometry
    int n_geom;
    int n_orient;
    const double *alpha, *beta, *weight;
    double scale;
    double *S, *Z;          // [n][n_geom][8], [n][n_geom][16]
    double *xs;             // [n][n_inc][2] or null
    int nmax_bound;         // upper bound of nmax over the batch (sizes the shared-memory T stage)
    const int *order;       // CTA -> particle, heaviest (largest NMAX) first, or null
    const double *frames;   // [n_inc][max blocks][n_orient][SC_FR]: particle-frame directions + polarisation rotations (k_scatter_frames)
    int frame_blocks;       // geometry blocks per incident direction in `frames`
    const double2 *wtab;    // Wigner table (k_scatter_wig): [(d, block)][1 + SCK_G directions][tri(m, n)][n_orient] pairs
    int wig_nb;             // largest n in `wtab`
};

// the one validity predicate of the scatter stage (k_scatter, k_order_by_nmax): a particle has a usable T-matrix when the
// controller finished it (converged, NGAUSS cap with the last T kept, or a flagged near-singular solve) and it owns an arena slot
__host__ __device__ __forceinline__ bool sc_has_tmatrix(int status, long long toff)
{
    return (status == HS_ST_OK || status == HS_ST_NGAUSS_LIMIT || status == HS_ST_SINGULAR) && toff >= 0;
}

// triangular index of (m, n >= m) in a Wigner table that goes up to n = nb
__host__ __device__ __forceinline__ long sc_tri(int m, int n, int nb) { return (long)m * (nb + 1) - (long)m * (m - 1) / 2 + (n - m); }
__host__ __device__ __forceinline__ long sc_tri_size(int nb) { return (long)(nb + 1) * (nb + 2) / 2; }

// Wigner pi_n = d/sin, tau_n = dd/dtheta generated in lock-step with n (VIGAMPL as a recurrence)
struct WigRec {
    double x, qs1, d1, d2;
    int m;
    bool pole;
    __device__ void init(double x_, int m_)
    {
        x = x_; m = m_;
        pole = fabs(1.0 - fabs(x)) <= 1e-10;
        if (pole) return;
        double qs = sqrt(1.0 - x * x);
        qs1 = 1.0 / qs;
        if (m == 0) { d1 = 1.0; d2 = x; }
        else {
            double a = 1.0;
            for (int i = 1; i <= m; i++) a = a * sqrt((double)(2 * i - 1) / (double)(2 * i)) * qs;
            d1 = 0.0; d2 = a;
        }
    }
    // call with n = max(m,1), max(m,1)+1, ... in order
    __device__ void next(int n, double &pi_n, double &tau_n)
    {
        if (pole) {
            if (m != 1) { pi_n = 0.0; tau_n = 0.0; return; }
            double dn = 0.5 * sqrt((double)(n * (n + 1)));
            if (x < 0.0) dn = dn * (((n + 1) & 1) ? -1.0 : 1.0);
            pi_n = dn;
            tau_n = (x < 0.0) ? -dn : dn;
            return;
        }
        double qn = n, qn1 = n + 1, qn2 = 2 * n + 1, d3, der;
        if (m == 0) {
            d3 = (qn2 * x * d2 - qn * d1) / qn1;
            der = qs1 * (qn1 * qn / qn2) * (-d1 + d3);
        } else {
            double qmm = (double)(m * m);
            double qnm = sqrt(qn * qn - qmm), qnm1 = sqrt(qn1 * qn1 - qmm);
            d3 = (qn2 * x * d2 - qnm * d1) / qnm1;
            der = qs1 * (-qn1 * qnm * d1 + qn * qnm1 * d3) / qn2;
        }
        pi_n = d2 * qs1;
        tau_n = der;
        d1 = d2; d2 = d3;
    }
};

// particle-frame direction (cos theta_p, phi_p) of a lab direction + the 2x2 polarisation rotation (AMPL eqs.)
__device__ inline void frame(double thet_deg, double phi_deg, double alph, double bet, double &ctp, double &phip,
                             double R[2][2])
{
    const double pin = 3.14159265358979323846, pin2 = pin * 0.5, rad = pin / 180.0, eps = 1e-7;
    double thetl = thet_deg * rad, phil = phi_deg * rad;
    if (thetl < pin2) thetl += eps;
    if (thetl > pin2) thetl -= eps;
    if (phil < pin) phil += eps;
    if (phil > pin) phil -= eps;
    double cb = cos(bet), sb = sin(bet), ct = cos(thetl), st = sin(thetl);
    double cp = cos(phil - alph), sp = sin(phil - alph);
    ctp = ct * cb + st * sb * cp;
    double thetp = acos(ctp);
    double cpp = cb * st * cp - sb * ct, spp = st * sp;
    phip = atan(spp / cpp);
    if (phip > 0.0 && sp < 0.0) phip += pin;
    if (phip < 0.0 && sp > 0.0) phip += pin;
    if (phip < 0.0) phip += 2.0 * pin;
    double ca = cos(alph), sa = sin(alph);
    double B[3][3] = {{ca * cb, sa * cb, -sb}, {-sa, ca, 0.0}, {ca * sb, sa * sb, cb}};
    cp = cos(phil); sp = sin(phil);
    double AL[3][2] = {{ct * cp, -sp}, {ct * sp, cp}, {-st, 0.0}};
    double stp = sin(thetp), cpp_ = cos(phip), spp_ = sin(phip);
    double AP[2][3] = {{ctp * cpp_, ctp * spp_, -stp}, {-spp_, cpp_, 0.0}};
    double C[3][2];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += B[i][k] * AL[k][j]; C[i][j] = x; }
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) { double x = 0; for (int k = 0; k < 3; k++) x += AP[i][k] * C[k][j]; R[i][j] = x; }
}

// One (particle, orientation, incident direction): ng scattered directions -> out[g*24..] (S 8, Z 16), xs2 (h, v)
__device__ void scatter_one(const double2 *__restrict__ t, int nmax, double lam, double thet0, double phi0, int ng,
                            const double *gsc, double alpha, double beta, const double *fn, double *out, double *xs2)
{
    const double pin = 3.14159265358979323846, pin2 = pin * 0.5, rad = pin / 180.0, eps = 1e-7;
    double alph = alpha * rad, bet = beta * rad;
    if (bet <= pin2 && pin2 - bet <= eps) bet -= eps;
    if (bet > pin2 && bet - pin2 <= eps) bet += eps;
    double ctp0, phip0, R[2][2];
    frame(thet0, phi0, alph, bet, ctp0, phip0, R);
    double ctp[SC_MAXG], dphi[SC_MAXG], R1[SC_MAXG][2][2];
    for (int g = 0; g < ng; g++) {
        double ph, Rg[2][2];
        frame(gsc[2 * g], gsc[2 * g + 1], alph, bet, ctp[g], ph, Rg);
        dphi[g] = ph - phip0;
        double d = 1.0 / (Rg[0][0] * Rg[1][1] - Rg[0][1] * Rg[1][0]);
        R1[g][0][0] = Rg[1][1] * d; R1[g][0][1] = -Rg[0][1] * d; R1[g][1][0] = -Rg[1][0] * d; R1[g][1][1] = Rg[0][0] * d;
    }
    double acc[SC_MAXG][8];
    for (int g = 0; g < ng; g++)
        for (int k = 0; k < 8; k++) acc[g][k] = 0.0;
    double acc_t = 0.0, acc_p = 0.0;
    double pr[HS_NPN1 + 1], pi_[HS_NPN1 + 1], qr[HS_NPN1 + 1], qi[HS_NPN1 + 1];  // P = b m pi0, Q = b tau0
    for (int m = 0; m <= nmax; m++) {
        const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1;
        {
            WigRec w0;
            w0.init(ctp0, m);
            for (int nn = nmin; nn <= nmax; nn++) {
                double p0, t0;
                w0.next(nn, p0, t0);
                const double f = fn[nn], mp = m * p0 * f, tq = t0 * f;
                const int pw = nn & 3;
                pr[nn] = (pw == 0) ? mp : (pw == 2) ? -mp : 0.0;
                pi_[nn] = (pw == 1) ? mp : (pw == 3) ? -mp : 0.0;
                qr[nn] = (pw == 0) ? tq : (pw == 2) ? -tq : 0.0;
                qi[nn] = (pw == 1) ? tq : (pw == 3) ? -tq : 0.0;
            }
        }
        WigRec ws[SC_MAXG];
        double fc[SC_MAXG], fs[SC_MAXG];
        for (int g = 0; g < ng; g++) {
            ws[g].init(ctp[g], m);
            if (m == 0) { fc[g] = 1.0; fs[g] = 0.0; }
            else { fc[g] = 2.0 * cos(m * dphi[g]); fs[g] = 2.0 * sin(m * dphi[g]); }
        }
        const double wm = (m == 0) ? 1.0 : 2.0;
        for (int n = nmin; n <= nmax; n++) {
            const int k = n - nmin;
            const double2 *ta = t + (long)(((n + 1) & 1) * NM) * NM + k;  // tau = 1 row of T, stride NM over n'
            const double2 *tb = t + (long)((n & 1) * NM) * NM + k;        // tau = 2 row
            double g1r = 0, g1i = 0, g2r = 0, g2i = 0, g3r = 0, g3i = 0, g4r = 0, g4i = 0;
            // n' with the parity of n: T11, T22 ; other parity: T12, T21
            for (int nn = nmin + ((n - nmin) & 1); nn <= nmax; nn += 2) {
                const long o = (long)(nn - nmin) * NM;
                const double2 a_ = ta[o], b_ = tb[o];
                const double Pr = pr[nn], Pi = pi_[nn], Qr = qr[nn], Qi = qi[nn];
                g1r += a_.x * Pr - a_.y * Pi; g1i += a_.x * Pi + a_.y * Pr;   // T11 * P
                g3r += a_.x * Qr - a_.y * Qi; g3i += a_.x * Qi + a_.y * Qr;   // T11 * Q
                g2r += b_.x * Qr - b_.y * Qi; g2i += b_.x * Qi + b_.y * Qr;   // T22 * Q
                g4r += b_.x * Pr - b_.y * Pi; g4i += b_.x * Pi + b_.y * Pr;   // T22 * P
            }
            for (int nn = nmin + ((n - nmin + 1) & 1); nn <= nmax; nn += 2) {
                const long o = (long)(nn - nmin) * NM;
                const double2 a_ = ta[o], b_ = tb[o];
                const double Pr = pr[nn], Pi = pi_[nn], Qr = qr[nn], Qi = qi[nn];
                g1r += a_.x * Qr - a_.y * Qi; g1i += a_.x * Qi + a_.y * Qr;   // T12 * Q
                g3r += a_.x * Pr - a_.y * Pi; g3i += a_.x * Pi + a_.y * Pr;   // T12 * P
                g2r += b_.x * Pr - b_.y * Pi; g2i += b_.x * Pi + b_.y * Pr;   // T21 * P
                g4r += b_.x * Qr - b_.y * Qi; g4i += b_.x * Qi + b_.y * Qr;   // T21 * Q
            }
            acc_t += wm * (g1r * g1r + g1i * g1i + g2r * g2r + g2i * g2i);
            acc_p += wm * (g3r * g3r + g3i * g3i + g4r * g4r + g4i * g4i);
            // a_n = i^(-n-1) f_n
            const double f = fn[n];
            const int pw = (-(n + 1)) & 3;
            const double ar = (pw == 0) ? f : (pw == 2) ? -f : 0.0, ai = (pw == 1) ? f : (pw == 3) ? -f : 0.0;
            for (int g = 0; g < ng; g++) {
                double p_, t_;
                ws[g].next(n, p_, t_);
                const double mp = m * p_;
                // u = m pi G1 + tau G2 ; v = m pi G3 + tau G4 ; w = tau G1 + m pi G2 ; z = tau G3 + m pi G4
                const double ur = mp * g1r + t_ * g2r, ui = mp * g1i + t_ * g2i;
                const double vr = mp * g3r + t_ * g4r, vi = mp * g3i + t_ * g4i;
                const double wr = t_ * g1r + mp * g2r, wi = t_ * g1i + mp * g2i;
                const double zr = t_ * g3r + mp * g4r, zi = t_ * g3i + mp * g4i;
                const double cr = ar * fc[g], ci = ai * fc[g], sr = ar * fs[g], si = ai * fs[g];
                acc[g][0] += ur * cr - ui * ci; acc[g][1] += ur * ci + ui * cr;   // VV
                acc[g][2] += vr * sr - vi * si; acc[g][3] += vr * si + vi * sr;   // VH
                acc[g][4] -= wr * sr - wi * si; acc[g][5] -= wr * si + wi * sr;   // HV
                acc[g][6] += zr * cr - zi * ci; acc[g][7] += zr * ci + zi * cr;   // HH
            }
        }
        t += 2 * NM * NM;
    }
    const double dk = 2.0 * pin / lam;
    if (xs2) {
        const double c0 = 4.0 * pin / (dk * dk);
        xs2[0] = c0 * (R[0][1] * R[0][1] * acc_t + R[1][1] * R[1][1] * acc_p);
        xs2[1] = c0 * (R[0][0] * R[0][0] * acc_t + R[1][0] * R[1][0] * acc_p);
    }
    for (int g = 0; g < ng; g++) {
        double vvr = acc[g][0] / dk, vvi = acc[g][1] / dk, vhr = acc[g][2] / dk, vhi = acc[g][3] / dk;
        double hvr = acc[g][4] / dk, hvi = acc[g][5] / dk, hhr = acc[g][6] / dk, hhi = acc[g][7] / dk;
        double cvvr = vvr * R[0][0] + vhr * R[1][0], cvvi = vvi * R[0][0] + vhi * R[1][0];
        double cvhr = vvr * R[0][1] + vhr * R[1][1], cvhi = vvi * R[0][1] + vhi * R[1][1];
        double chvr = hvr * R[0][0] + hhr * R[1][0], chvi = hvi * R[0][0] + hhi * R[1][0];
        double chhr = hvr * R[0][1] + hhr * R[1][1], chhi = hvi * R[0][1] + hhi * R[1][1];
        double s11r = R1[g][0][0] * cvvr + R1[g][0][1] * chvr, s11i = R1[g][0][0] * cvvi + R1[g][0][1] * chvi;
        double s12r = R1[g][0][0] * cvhr + R1[g][0][1] * chhr, s12i = R1[g][0][0] * cvhi + R1[g][0][1] * chhi;
        double s21r = R1[g][1][0] * cvvr + R1[g][1][1] * chvr, s21i = R1[g][1][0] * cvvi + R1[g][1][1] * chvi;
        double s22r = R1[g][1][0] * cvhr + R1[g][1][1] * chhr, s22i = R1[g][1][0] * cvhi + R1[g][1][1] * chhi;
        double *o = out + g * 24;
        o[0] = s11r; o[1] = s11i; o[2] = s12r; o[3] = s12i; o[4] = s21r; o[5] = s21i; o[6] = s22r; o[7] = s22i;
#define RE(a, b) (a##r * b##r + a##i * b##i)
#define IM(a, b) (a##i * b##r - a##r * b##i)
        double *Z = o + 8;
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
}

#ifndef HS_HOST_EMU
// ---- production kernel: same mathematics as scatter_one (which stays as the single-lane reference used by the
// host-emulation tests), restructured for the machine:
//   * one CTA per particle, one thread per orientation; every thread walks the same T elements, so the T block of
//     each azimuthal mode is staged ONCE in shared memory by the whole CTA (coalesced) and read back as broadcasts;
//   * sqrt(n^2-m^2), its reciprocal, 1/(n+1), 1/(2n+1) and prod sqrt((2i-1)/2i) come from per-CTA tables instead of
//     per-thread sqrt/div chains;
//   * the phases b_n' = i^n' f_n' are folded out of the inner loop (n' runs with stride 2, so i^n' = i^(n'&1) * sign):
//     the T . vector product is complex x real.
#define SCK_G 2  // scattered directions handled per sweep

struct WigTab {
    double x, qs, qs1, qspow, d1, d2;
    bool pole;
    __device__ __forceinline__ void setup(double x_)
    {
        x = x_;
        pole = fabs(1.0 - fabs(x)) <= 1e-10;
        qs = pole ? 0.0 : sqrt(1.0 - x * x);
        qs1 = pole ? 0.0 : 1.0 / qs;
        qspow = 1.0;
    }
    // mode m (call for m = 0, 1, 2, ... in order: qspow accumulates sin^m); cm = prod_{i<=m} sqrt((2i-1)/2i)
    __device__ __forceinline__ void start(int m, double cm)
    {
        if (m == 0) { d1 = 1.0; d2 = x; }
        else { qspow *= qs; d1 = 0.0; d2 = cm * qspow; }
    }
    // n = max(m,1), max(m,1)+1, ... in order
    __device__ __forceinline__ void next(int n, int m, double &pi_n, double &tau_n)
    {
        if (pole) {
            if (m != 1) { pi_n = 0.0; tau_n = 0.0; return; }
            double dn = 0.5 * sqrt((double)(n * (n + 1)));
            if (x < 0.0) dn = dn * (((n + 1) & 1) ? -1.0 : 1.0);
            pi_n = dn;
            tau_n = (x < 0.0) ? -dn : dn;
            return;
        }
        const double qn = n, qn1 = n + 1, qn2 = 2 * n + 1;
        const double r2n1 = 1.0 / qn2;
        double d3, der;
        if (m == 0) {
            d3 = (qn2 * x * d2 - qn * d1) * (1.0 / qn1);
            der = qs1 * (qn1 * qn * r2n1) * (-d1 + d3);
        } else {
            const double qmm = (double)(m * m);
            const double qnm = sqrt(qn * qn - qmm), qnm1 = sqrt(qn1 * qn1 - qmm);
            d3 = (qn2 * x * d2 - qnm * d1) * (1.0 / qnm1);
            der = qs1 * (-qn1 * qnm * d1 + qn * qnm1 * d3) * r2n1;
        }
        pi_n = d2 * qs1;
        tau_n = der;
        d1 = d2; d2 = d3;
    }
};

#define SC_FR (6 + 6 * SCK_G)
// frames[(d, block, o)]: ctp0, phip0, R (2x2) of the incident direction, then per scattered direction of the block
// cos(theta'), phi' - phi0', R1 = inverse polarisation rotation (what the orientation threads of k_scatter used to set up
// for every particle although it only depends on (orientation, geometry))
__global__ void k_scatter_frames(ScArgs a, double *frames)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y, b = blockIdx.z;
    if (o >= a.n_orient) return;
    const int gb = a.g_begin[d] + b * SCK_G, gend = a.g_begin[d + 1];
    const int ng = max(0, min(SCK_G, gend - gb));
    const double pin = 3.14159265358979323846, pin2 = pin * 0.5, rad = pin / 180.0, eps = 1e-7;
    double alph = a.alpha[o] * rad, bet = a.beta[o] * rad;
    if (bet <= pin2 && pin2 - bet <= eps) bet -= eps;
    if (bet > pin2 && bet - pin2 <= eps) bet += eps;
    double ctp0, phip0, R[2][2];
    frame(a.inc[2 * d], a.inc[2 * d + 1], alph, bet, ctp0, phip0, R);
    double *fr = frames + (((long)d * a.frame_blocks + b) * a.n_orient + o) * SC_FR;
    fr[0] = ctp0; fr[1] = phip0; fr[2] = R[0][0]; fr[3] = R[0][1]; fr[4] = R[1][0]; fr[5] = R[1][1];
    for (int g = 0; g < SCK_G; g++) {
        double *fg = fr + 6 + 6 * g;
        if (g < ng) {
            double ph, ct, Rg[2][2];
            frame(a.gsc[2 * (gb + g)], a.gsc[2 * (gb + g) + 1], alph, bet, ct, ph, Rg);
            const double dd = 1.0 / (Rg[0][0] * Rg[1][1] - Rg[0][1] * Rg[1][0]);
            fg[0] = ct; fg[1] = ph - phip0;
            fg[2] = Rg[1][1] * dd; fg[3] = -Rg[0][1] * dd; fg[4] = -Rg[1][0] * dd; fg[5] = Rg[0][0] * dd;
        } else {
            for (int k = 0; k < 6; k++) fg[k] = 0.0;
        }
    }
}

// Wigner table of a scatter call: the pi / tau functions of the particle-frame directions only depend on (orientation,
// geometry, m, n), not on the particle, so they are generated ONCE per call (same recurrence, same operation order as
// the per-thread WigTab that the orientation threads of k_scatter used to carry) instead of once per particle:
//   direction 0 (incident):       { P'_n, Q'_n } = { sgn f_n m pi0_n, sgn f_n tau0_n },  sgn = (-1)^(n >> 1)
//   direction 1 + g (scattered):  { f_n m pi_n, f_n tau_n }                              f_n = sqrt((2n+1) / (n(n+1)))
// grid (orientation blocks, m, (d, block, direction)); thread = orientation.
__global__ void __launch_bounds__(64) k_scatter_wig(ScArgs a, double2 *wtab)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y, nb = a.wig_nb;
    const int db = blockIdx.z / (1 + SCK_G), dir = blockIdx.z - db * (1 + SCK_G);
    const int d = db / a.frame_blocks, b = db - d * a.frame_blocks;
    if (o >= a.n_orient) return;
    const int gb = a.g_begin[d] + b * SCK_G, gend = a.g_begin[d + 1];
    const int ng = max(0, min(SCK_G, gend - gb));
    if (dir > ng || (b > 0 && ng == 0)) return;
    const double *fr = a.frames + ((long)db * a.n_orient + o) * SC_FR;
    WigTab w;
    w.setup(dir == 0 ? fr[0] : fr[6 + 6 * (dir - 1)]);
    double cm = 1.0;
    for (int i = 0; i <= m; i++) {
        if (i > 0) cm = cm * sqrt((double)(2 * i - 1) / (double)(2 * i));
        w.start(i, cm);
    }
    const int nmin = m > 1 ? m : 1;
    double2 *out = wtab + (((long)db * (1 + SCK_G) + dir) * sc_tri_size(nb) + sc_tri(m, nmin, nb)) * a.n_orient + o;
    for (int n = nmin; n <= nb; n++) {
        double pi_n, tau_n;
        w.next(n, m, pi_n, tau_n);
        const double f0 = sqrt((double)(2 * n + 1) / (double)(n * (n + 1)));
        const double f = (dir == 0 && (n & 2)) ? -f0 : f0;
        out[(long)(n - nmin) * a.n_orient] = make_double2(m * pi_n * f, tau_n * f);
    }
}

//   * SPLIT (1, 2 or 4) adjacent lanes share one orientation and take every SPLIT-th row n of a mode: more warps for the
//     same shared memory (the heavy NMAX classes are limited by their T stage), and the lanes of a group read the same
//     P' / Q' entries (fewer shared-memory wavefronts per FMA).
//   * NTO = orientations per pass as a compile-time constant (64, the usual case; 0 = blockDim.x / SPLIT): the [n'][thread]
//     strides of the P' / Q' vectors become immediates of the LDS instructions instead of two integer instructions per load.
template <int SPLIT, int NTO>
__global__ void __launch_bounds__(64 * SPLIT) k_scatter(ScArgs a)
{
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nto = NTO ? NTO : nt / SPLIT, oi = tid / SPLIT, h = tid % SPLIT;  // orientations per pass, this thread's one, its row slice
    const int p = a.order ? a.order[blockIdx.x] : blockIdx.x;
    // a particle outside its launch class's bounds (cannot happen with k_order_by_nmax's predicate; checked because the T stage
    // and the Wigner table are sized for the class) is reported as NaN instead of overrunning shared memory
    const bool bad = !sc_has_tmatrix(a.status[p], a.toff[p]) || a.nmax[p] > a.nmax_bound || a.nmax[p] > a.wig_nb;
    if (bad) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int t = tid; t < a.n_geom * 8; t += nt) a.S[(long)p * a.n_geom * 8 + t] = nan;
        for (int t = tid; t < a.n_geom * 16; t += nt) a.Z[(long)p * a.n_geom * 16 + t] = nan;
        if (a.xs) for (int t = tid; t < a.n_inc * 2; t += nt) a.xs[(long)p * a.n_inc * 2 + t] = nan;
        return;
    }
    const double2 *tg = reinterpret_cast<const double2 *>(a.tarena) + a.toff[p];
    const int nmax = a.nmax[p];
    const double lam = a.lam[p];
    // dynamic smem: [T stage: 2*nmax*nmax complex][reduction rows: nto * stride][accs][per-orientation amplitude sums: 8 SCK_G x nto]
    double2 *sT = reinterpret_cast<double2 *>(sm);
    const int stride = 24 + 3;  // one scattered direction (S 8, Z 16) + the two cross sections per reduction row
    double *srow = sm + 4L * a.nmax_bound * a.nmax_bound;
    const long vec_doubles = 2L * (a.nmax_bound + 2) * nto, row_doubles = (long)nto * stride;
    double *accs = srow + (vec_doubles > row_doubles ? vec_doubles : row_doubles);
    const double pin = 3.14159265358979323846;

    for (int d = 0; d < a.n_inc; d++) {
        const int gend = a.g_begin[d + 1];
        for (int gb = a.g_begin[d];; gb += SCK_G) {
            const int ng = max(0, min(SCK_G, gend - gb));
            const int ncomp = ng * 24 + 2;
            for (int c = tid; c < ncomp; c += nt) accs[c] = 0.0;
            for (int o0 = 0; o0 < a.n_orient; o0 += nto) {
                const int o = o0 + oi;
                const bool active = o < a.n_orient;
                const unsigned lanes = __ballot_sync(0xffffffffu, active);  // the SPLIT lanes of an orientation are in or out together
                // ---- per-thread setup ----
                double acc_t = 0.0, acc_p = 0.0;
                const long dblk = (long)d * a.frame_blocks + (gb - a.g_begin[d]) / SCK_G;
                // orientation / geometry quantities do not depend on the particle: they come from the table built once per
                // call by k_scatter_frames (AMPL's angle nudges and Euler rotations), read where they are needed
                const double *fr = a.frames + (dblk * a.n_orient + (active ? o : 0)) * SC_FR;
                // amplitude sums over the modes: per-thread slots in shared memory ([component][thread]), updated once per mode
                double *sacc = accs + (SCK_G * 24 + 2) + oi;
                if (h == 0) {
#pragma unroll
                    for (int k = 0; k < 8 * SCK_G; k++) sacc[k * nto] = 0.0;
                }
                // Wigner functions of this thread's directions: table of the call (k_scatter_wig), [tri(m, n)][orientation]
                const long tri = sc_tri_size(a.wig_nb);
                const double2 *wt0 = a.wtab + dblk * (1 + SCK_G) * tri * a.n_orient + (active ? o : 0);
                // P' = sgn f m pi0, Q' = sgn f tau0 of the current mode: per-thread vectors in shared memory ([n'][thread], aliased
                // with the reduction rows, which are only written after the mode sweep).  As thread-local arrays they lived in
                // L1-cached local memory, which the heavy class (3 CTAs x 55 KB of shared memory per SM) leaves too small.
                double *Pp = srow + oi, *Qp = srow + (long)(a.nmax_bound + 2) * nto + oi;
                long moff = 0;
                for (int m = 0; m <= nmax; m++) {
                    const int nmin = m > 1 ? m : 1, NM = nmax - nmin + 1;
                    __syncthreads();
                    for (int e = tid; e < 2 * NM * NM; e += nt) sT[e] = tg[moff + e];
                    __syncthreads();
                    moff += 2L * NM * NM;
                    if (!active) continue;
                    const double2 *wm0 = wt0 + sc_tri(m, nmin, a.wig_nb) * a.n_orient;
                    for (int nn = nmin + h; nn <= nmax; nn += SPLIT) {
                        const double2 v = wm0[(long)(nn - nmin) * a.n_orient];
                        Pp[nn * nto] = v.x;
                        Qp[nn * nto] = v.y;
                    }
                    if (SPLIT > 1) __syncwarp(lanes);
                    // per-mode sums over n of i^(-n-1) (f m pi +- f tau) (G1 +- G2), ... (see below)
                    double A[SCK_G][8];
#pragma unroll
                    for (int g = 0; g < SCK_G; g++)
#pragma unroll
                        for (int k = 0; k < 8; k++) A[g][k] = 0.0;
                    double xt = 0.0, xp = 0.0;
                    const double2 *wg = wt0 + (tri + sc_tri(m, nmin, a.wig_nb)) * a.n_orient;  // first scattered direction, n = nmin
                    for (int n = nmin + h; n <= nmax; n += SPLIT) {
                        const int k = n - nmin;
                        double2 wv[SCK_G];   // { f m pi_n, f tau_n } of the scattered directions: loaded ahead of the T sweep
#pragma unroll
                        for (int g = 0; g < SCK_G; g++)
                            if (g < ng) wv[g] = wg[((long)g * tri + k) * a.n_orient];
                        const double2 *ta = sT + (long)(((n + 1) & 1) * NM) * NM + k;  // tau = 1 row, stride NM over n'
                        const double2 *tb = sT + (long)((n & 1) * NM) * NM + k;        // tau = 2 row
                        // same-parity n': T11 (ta), T22 (tb); phase i^(n&1)
                        double e1r = 0, e1i = 0, e3r = 0, e3i = 0, e2r = 0, e2i = 0, e4r = 0, e4i = 0;
                        {   // running pointers: the row stride NM is a run-time value, one add per pointer and step
                            const int nn0 = nmin + (k & 1);
                            const double2 *pa = ta + (long)(k & 1) * NM, *pb = tb + (long)(k & 1) * NM;
                            const double *pP = Pp + nn0 * nto, *pQ = Qp + nn0 * nto;
                            for (int nn = nn0; nn <= nmax; nn += 2, pa += 2 * NM, pb += 2 * NM, pP += 2 * nto, pQ += 2 * nto) {
                                const double2 a_ = *pa, b_ = *pb;
                                const double P = *pP, Q = *pQ;
                                e1r += a_.x * P; e1i += a_.y * P; e3r += a_.x * Q; e3i += a_.y * Q;
                                e2r += b_.x * Q; e2i += b_.y * Q; e4r += b_.x * P; e4i += b_.y * P;
                            }
                        }
                        // other-parity n': T12 (ta), T21 (tb); phase i^((n+1)&1)
                        double o1r = 0, o1i = 0, o3r = 0, o3i = 0, o2r = 0, o2i = 0, o4r = 0, o4i = 0;
                        {
                            const int nn0 = nmin + ((k + 1) & 1);
                            const double2 *pa = ta + (long)((k + 1) & 1) * NM, *pb = tb + (long)((k + 1) & 1) * NM;
                            const double *pP = Pp + nn0 * nto, *pQ = Qp + nn0 * nto;
                            for (int nn = nn0; nn <= nmax; nn += 2, pa += 2 * NM, pb += 2 * NM, pP += 2 * nto, pQ += 2 * nto) {
                                const double2 a_ = *pa, b_ = *pb;
                                const double P = *pP, Q = *pQ;
                                o1r += a_.x * Q; o1i += a_.y * Q; o3r += a_.x * P; o3i += a_.y * P;
                                o2r += b_.x * P; o2i += b_.y * P; o4r += b_.x * Q; o4i += b_.y * Q;
                            }
                        }
                        double g1r, g1i, g2r, g2i, g3r, g3i, g4r, g4i;
                        if (n & 1) {  // same-parity phase = i, other-parity phase = 1
                            g1r = -e1i + o1r; g1i = e1r + o1i; g2r = -e2i + o2r; g2i = e2r + o2i;
                            g3r = -e3i + o3r; g3i = e3r + o3i; g4r = -e4i + o4r; g4i = e4r + o4i;
                        } else {      // same-parity phase = 1, other-parity phase = i
                            g1r = e1r - o1i; g1i = e1i + o1r; g2r = e2r - o2i; g2i = e2i + o2r;
                            g3r = e3r - o3i; g3i = e3i + o3r; g4r = e4r - o4i; g4i = e4i + o4r;
                        }
                        xt += g1r * g1r + g1i * g1i + g2r * g2r + g2i * g2i;
                        xp += g3r * g3r + g3i * g3i + g4r * g4r + g4i * g4i;
                        // With s = f m pi, t = f tau of a scattered direction (table), u = s G1 + t G2, w = t G1 + s G2,
                        // v = s G3 + t G4, z = t G3 + s G4:  u + w = (s + t)(G1 + G2), u - w = (s - t)(G1 - G2), same for v, z.
                        // A accumulates i^(-n-1) times those four products; u, v, w, z are recovered per mode (they are
                        // linear), which takes 8 FMAs per direction and n instead of 36.
                        const double p12r = g1r + g2r, p12i = g1i + g2i, m12r = g1r - g2r, m12i = g1i - g2i;
                        const double p34r = g3r + g4r, p34i = g3i + g4i, m34r = g3r - g4r, m34i = g3i - g4i;
                        const bool neg = ((-(n + 1)) & 2) != 0;   // i^(-n-1) = -1 or -i
#pragma unroll
                        for (int g = 0; g < SCK_G; g++) {
                            if (g >= ng) continue;
                            const double sp_ = neg ? -(wv[g].x + wv[g].y) : (wv[g].x + wv[g].y), sm_ = neg ? -(wv[g].x - wv[g].y) : (wv[g].x - wv[g].y);
                            if ((n + 1) & 1) {  // times i
                                A[g][0] -= sp_ * p12i; A[g][1] += sp_ * p12r; A[g][2] -= sm_ * m12i; A[g][3] += sm_ * m12r;
                                A[g][4] -= sp_ * p34i; A[g][5] += sp_ * p34r; A[g][6] -= sm_ * m34i; A[g][7] += sm_ * m34r;
                            } else {
                                A[g][0] += sp_ * p12r; A[g][1] += sp_ * p12i; A[g][2] += sm_ * m12r; A[g][3] += sm_ * m12i;
                                A[g][4] += sp_ * p34r; A[g][5] += sp_ * p34i; A[g][6] += sm_ * m34r; A[g][7] += sm_ * m34i;
                            }
                        }
                    }
                    const double wm = (m == 0) ? 1.0 : 2.0;
                    acc_t += wm * xt;
                    acc_p += wm * xp;
#pragma unroll
                    for (int g = 0; g < SCK_G; g++) {
                        if (g >= ng) continue;
                        double fc, fs;
                        if (m == 0) { fc = 0.5; fs = 0.0; }
                        else { double sn, cs; sincos(m * fr[6 + 6 * g + 1], &sn, &cs); fc = cs; fs = sn; }  // (2 cos, 2 sin) / 2
                        if (SPLIT > 1) {  // add up the row slices of this orientation's lanes
#pragma unroll
                            for (int k = 0; k < 8; k++)
#pragma unroll
                                for (int sh = 1; sh < SPLIT; sh <<= 1) A[g][k] += __shfl_xor_sync(lanes, A[g][k], sh);
                            if (h != 0) continue;
                        }
                        const double ur = A[g][0] + A[g][2], ui = A[g][1] + A[g][3], wr = A[g][0] - A[g][2], wi = A[g][1] - A[g][3];
                        const double vr = A[g][4] + A[g][6], vi = A[g][5] + A[g][7], zr = A[g][4] - A[g][6], zi = A[g][5] - A[g][7];
                        double *sa = sacc + 8 * g * nto;
                        sa[0 * nto] += fc * ur; sa[1 * nto] += fc * ui;   // VV
                        sa[2 * nto] += fs * vr; sa[3 * nto] += fs * vi;   // VH
                        sa[4 * nto] -= fs * wr; sa[5 * nto] -= fs * wi;   // HV
                        sa[6 * nto] += fc * zr; sa[7 * nto] += fc * zi;   // HH
                    }
                }
                if (SPLIT > 1 && active) {
#pragma unroll
                    for (int sh = 1; sh < SPLIT; sh <<= 1) {
                        acc_t += __shfl_xor_sync(lanes, acc_t, sh);
                        acc_p += __shfl_xor_sync(lanes, acc_p, sh);
                    }
                }
                __syncthreads();  // everybody is done with the P' / Q' vectors that share the reduction rows' memory
                // ---- per-thread epilogue: S, Z, cross sections -> weighted row in shared memory, one scattered direction at a
                // time (rows of 24 + 2 components; the reduction rows share their memory with the P' / Q' vectors) ----
                for (int g = 0; g < ng; g++) {
                    const int ncol = 24 + (g == 0 ? 2 : 0);
                    if (active && h == 0) {
                        const double dk = 2.0 * pin / lam, w = a.weight[o];
                        const double R[2][2] = {{fr[2], fr[3]}, {fr[4], fr[5]}};
                        double *row = srow + (long)oi * stride;
                        if (g == 0) {
                            const double c0 = 4.0 * pin / (dk * dk);
                            row[24] = w * c0 * (R[0][1] * R[0][1] * acc_t + R[1][1] * R[1][1] * acc_p);
                            row[25] = w * c0 * (R[0][0] * R[0][0] * acc_t + R[1][0] * R[1][0] * acc_p);
                        }
                        const double *sa = sacc + 8 * g * nto, *fg = fr + 6 + 6 * g;
                        const double R1[2][2] = {{fg[2], fg[3]}, {fg[4], fg[5]}};
                        const double rdk = 1.0 / dk;
                        double vvr = sa[0 * nto] * rdk, vvi = sa[1 * nto] * rdk, vhr = sa[2 * nto] * rdk, vhi = sa[3 * nto] * rdk;
                        double hvr = sa[4 * nto] * rdk, hvi = sa[5 * nto] * rdk, hhr = sa[6 * nto] * rdk, hhi = sa[7 * nto] * rdk;
                        double cvvr = vvr * R[0][0] + vhr * R[1][0], cvvi = vvi * R[0][0] + vhi * R[1][0];
                        double cvhr = vvr * R[0][1] + vhr * R[1][1], cvhi = vvi * R[0][1] + vhi * R[1][1];
                        double chvr = hvr * R[0][0] + hhr * R[1][0], chvi = hvi * R[0][0] + hhi * R[1][0];
                        double chhr = hvr * R[0][1] + hhr * R[1][1], chhi = hvi * R[0][1] + hhi * R[1][1];
                        double s11r = R1[0][0] * cvvr + R1[0][1] * chvr, s11i = R1[0][0] * cvvi + R1[0][1] * chvi;
                        double s12r = R1[0][0] * cvhr + R1[0][1] * chhr, s12i = R1[0][0] * cvhi + R1[0][1] * chhi;
                        double s21r = R1[1][0] * cvvr + R1[1][1] * chvr, s21i = R1[1][0] * cvvi + R1[1][1] * chvi;
                        double s22r = R1[1][0] * cvhr + R1[1][1] * chhr, s22i = R1[1][0] * cvhi + R1[1][1] * chhi;
                        double *o_ = row;
                        o_[0] = w * s11r; o_[1] = w * s11i; o_[2] = w * s12r; o_[3] = w * s12i;
                        o_[4] = w * s21r; o_[5] = w * s21i; o_[6] = w * s22r; o_[7] = w * s22i;
#define RE(a, b) (a##r * b##r + a##i * b##i)
#define IM(a, b) (a##i * b##r - a##r * b##i)
                        double *Z = o_ + 8;
                        Z[0] = w * 0.5 * (RE(s11, s11) + RE(s12, s12) + RE(s21, s21) + RE(s22, s22));
                        Z[1] = w * 0.5 * (RE(s11, s11) - RE(s12, s12) + RE(s21, s21) - RE(s22, s22));
                        Z[2] = w * -(RE(s11, s12) + RE(s22, s21));
                        Z[3] = w * -(IM(s11, s12) - IM(s22, s21));
                        Z[4] = w * 0.5 * (RE(s11, s11) + RE(s12, s12) - RE(s21, s21) - RE(s22, s22));
                        Z[5] = w * 0.5 * (RE(s11, s11) - RE(s12, s12) - RE(s21, s21) + RE(s22, s22));
                        Z[6] = w * -(RE(s11, s12) - RE(s22, s21));
                        Z[7] = w * -(IM(s11, s12) + IM(s22, s21));
                        Z[8] = w * -(RE(s11, s21) + RE(s22, s12));
                        Z[9] = w * -(RE(s11, s21) - RE(s22, s12));
                        Z[10] = w * (RE(s11, s22) + RE(s12, s21));
                        Z[11] = w * (IM(s11, s22) + IM(s21, s12));
                        Z[12] = w * -(IM(s21, s11) + IM(s22, s12));
                        Z[13] = w * -(IM(s21, s11) - IM(s22, s12));
                        Z[14] = w * (IM(s22, s11) - IM(s12, s21));
                        Z[15] = w * (RE(s22, s11) - RE(s12, s21));
#undef RE
#undef IM
                    }
                    __syncthreads();
                    const int cnt = min(nto, a.n_orient - o0);
                    for (int c = tid; c < ncol; c += nt) {
                        const int ci = c < 24 ? g * 24 + c : ng * 24 + (c - 24);
                        double acc_ = accs[ci];
                        for (int j = 0; j < cnt; j++) acc_ += srow[(long)j * stride + c];
                        accs[ci] = acc_;
                    }
                    __syncthreads();
                }
            }
            for (int c = tid; c < ncomp; c += nt) {
                const double v = accs[c] * a.scale;
                if (c < ng * 24) {
                    const int g = c / 24, k = c - g * 24, slot = a.g_out[gb + g];
                    if (k < 8) a.S[((long)p * a.n_geom + slot) * 8 + k] = v;
                    else a.Z[((long)p * a.n_geom + slot) * 16 + k - 8] = v;
                } else if (a.xs && gb == a.g_begin[d]) {
                    a.xs[((long)p * a.n_inc + d) * 2 + (c - ng * 24)] = v;
                }
            }
            __syncthreads();
            if (gb + SCK_G >= gend) break;
        }
    }
}
// CTA -> particle map with the largest NMAX first (one block; counting sort on NMAX): the heavy particles start first
// instead of forming the tail of the launch.
__global__ void __launch_bounds__(1024) k_order_by_nmax(int n, const int *nmax, const int *status, const long long *toff, int *order, int *hist_out)
{
    __shared__ int cnt[HS_NPN1 + 2];
    for (int t = threadIdx.x; t < HS_NPN1 + 2; t += blockDim.x) cnt[t] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int v = nmax[i];
        v = (sc_has_tmatrix(status[i], toff[i]) && v > 0 && v <= HS_NPN1) ? v : 0;
        atomicAdd(&cnt[v], 1);
    }
    __syncthreads();
    if (hist_out)
        for (int t = threadIdx.x; t < HS_NPN1 + 1; t += blockDim.x) hist_out[t] = cnt[t];  // pinned host memory: class sizes for the launcher
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int v = HS_NPN1; v >= 0; v--) { int c = cnt[v]; cnt[v] = acc; acc += c; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int v = nmax[i];
        v = (sc_has_tmatrix(status[i], toff[i]) && v > 0 && v <= HS_NPN1) ? v : 0;
        order[atomicAdd(&cnt[v], 1)] = i;
    }
}
#endif

}  // namespace hs

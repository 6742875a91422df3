"""Batched scatter-table build and PSD integration: the whole-table equivalent of what hydroscatt's scripts do one
diameter / one DSD at a time (scattering_rain.py:276-316: PSDIntegrator.init_scatter_table over 70 diameters and two
geometries, then compute_scattering_psd_tm for each of 9 300 DSDs).  Everything numerical is a C-ABI kernel; torch
is used for device memory and for re-packing rows.

Row layout of a table record (56 doubles per (table, diameter)), also the PSD-integrated layout:
   0.. 7  S back  (S11,S12,S21,S22 as re,im)      8..23  Z back  (row-major 4x4)
  24..31  S forw                                  32..47  Z forw
  48,49   sca_xsect h,v (back-geometry incidence) 50,51  sca_xsect h,v (forward-geometry incidence)
  52,53   ext_xsect h,v (back-geometry incidence) 54,55  ext_xsect h,v (forward-geometry incidence)
"""
import numpy as np
import torch

from . import engine as _engine
from .backend import trapz_weights, rect_weights

ROW = 56
OBS_COLUMNS = ("sca_xsect_h", "sca_xsect_v", "refl_h", "refl_v", "zdr", "ldr_h", "ldr_v", "rho_hv", "delta_hv",
               "ext_xsect_h", "ext_xsect_v", "kdp", "A_h", "A_v", "Adp")


def build_rows(eng, radius, wavelength, m_re, m_im, axis_ratio, geom_back, geom_forw, alpha, beta, weight, scale,
               ddelt=1e-3, ndgs=2, rescale_ddelt=True, rat=1.0):
    """T-matrices + orientation-averaged S/Z + cross sections for a flat particle list.

    Returns (rows [n, 56] float64 on the device, TMatrixBatch).  Two kernel families: k_calctmat (one launch per size bucket, concurrent streams)
    and the fused k_scatter (orientation-averaged S/Z of all geometries + cross sections in one T-matrix sweep).
    """
    gb, gf = tuple(geom_back[:4]), tuple(geom_forw[:4])
    geoms = [gb, gf]
    # forward amplitudes for the optical theorem at each geometry's incidence direction
    fw = [(g[0], g[0], g[2], g[2]) for g in geoms]
    allg = geoms + [f for f in dict.fromkeys(fw) if f not in geoms]
    tb, S, Z, xs = eng.calctmat_scatter(radius, wavelength, m_re, m_im, axis_ratio, allg, alpha, beta, weight=weight,
                                        scale=scale, want_xs=True, rat=rat, ddelt=ddelt, ndgs=ndgs, rescale_ddelt=rescale_ddelt)
    rows = eng.pack_rows(S, Z, xs, tb.lam, 0, 1, allg.index(fw[0]), allg.index(fw[1]))
    return rows, tb


def sp_observables(eng, rows, lam, kw_sqr=0.93):
    """Single-particle observables of compute_scattering_sp_tm (scattering.py:451-527) for every row: [n, 15]."""
    return eng.observables(rows, 0, 24, 54, 48, lam, kw_sqr)


def integrate_psd(eng, rows, n_tab, D, kind, p0, p1, p2, dmax, lam_tab, rule="trapz", kw_sqr=0.93,
                  hpol_quirk=True, want_integ=True, columns=None):
    """PSD-integrate n_tab tables (rows [n_tab*n_D, 56], table-major) over n_dsd size distributions and derive the
    observables of compute_scattering_psd_tm (scattering.py:530-587).

    Returns (obs [n_dsd, n_tab, 15], integ [n_dsd, n_tab, 56]) on the device.  want_integ=False runs the fused kernel
    (hs_psd_observables): only the integrals the observables read are formed, on chip, and integ is None; columns
    (names or indices of OBS_COLUMNS) then selects which observables are written: obs [n_dsd, n_tab, len(columns)].
    """
    D = np.asarray(D, dtype=np.float64)
    n_D = D.size
    wq = trapz_weights(D) if rule == "trapz" else rect_weights(D)
    W = eng.psd_weights(kind, p0, p1, p2, dmax, D, wq)                       # [n_dsd, n_D]
    if not want_integ:
        if columns is not None:
            columns = [OBS_COLUMNS.index(c) if isinstance(c, str) else int(c) for c in columns]
        return eng.psd_observables(W, rows, n_tab, lam_tab, kw_sqr, hpol_quirk, want_sca=False, columns=columns), None
    Tt = rows.reshape(n_tab, n_D, ROW).permute(1, 0, 2).reshape(n_D, n_tab * ROW).contiguous()
    integ = eng.psd_integrate(W, Tt)                                          # [n_dsd, n_tab*56]
    n_dsd = integ.shape[0]
    r2 = integ.reshape(n_dsd * n_tab, ROW)
    if hpol_quirk:
        # upstream scatter.ext_xsect ignores h_pol once an angular table exists: A_v == A_h
        r2[:, 55] = r2[:, 54]
        r2[:, 53] = r2[:, 52]
    lam_tab = eng._dev(lam_tab, n_tab)
    lam_rows = lam_tab.repeat(n_dsd)
    obs = eng.observables(r2, 0, 24, 54, -1, lam_rows, kw_sqr)
    return obs.reshape(n_dsd, n_tab, 15), integ.reshape(n_dsd, n_tab, ROW)


def fixed_orientation_nodes(std_deg, n_alpha=5, n_beta=10):
    """alpha[O], beta[O], weight[O], scale of pytmatrix's orient_averaged_fixed with gaussian_pdf(std)."""
    from . import install_shims
    install_shims()
    from pytmatrix import orientation, quadrature
    (bp, bw) = quadrature.get_points_and_weights(orientation.gaussian_pdf(std_deg), 0, 180, n_beta)
    ap = np.linspace(0, 360, n_alpha + 1)[:-1]
    return np.repeat(ap, n_beta), np.tile(bp, n_alpha), np.tile(bw, n_alpha), (1.0 / n_alpha) / bw.sum()

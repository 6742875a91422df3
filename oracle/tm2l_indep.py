"""Second, structurally independent two-layer oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Purpose ("pin parity harder"): oracle/tm2l_oracle.c and the CUDA kernel both restate the formulas of
hydroscatt/f_2l_tmatrix/tmatrix_py.f (`gener`, :792-1338), so a shared misreading of the Fortran would pass every test.
This file computes the same physical quantity -- far-field amplitudes of a coated (two-layer, coaxial) spheroid -- from
a DIFFERENT formulation with different numerics, sharing no formula and no code with the other two:

  * null-field / EBCM in Mishchenko's normalisation (normalised Wigner d functions, blocks
    X11..X22, Q built from t11..t22), generalised to an arbitrary pair (F outside, G inside) of radial functions;
    the layered T-matrix is composed as
        T2 = -Q2[j,j] Q2[h,j]^-1                      core in the shell medium  (surface 2, k = k0 m1 complex)
        T  = -(Q1[j,j] + Q1[j,h] T2)(Q1[h,j] + Q1[h,h] T2)^-1                    (surface 1, k = k0)
    (Peterson & Strom 1974) instead of the Fortran's six matrices a, b, c, d, x, y with cmxnrm rescaling;
  * scipy.special spherical Bessel functions (real and complex argument) instead of hand-written recurrences;
  * Gauss-Legendre quadrature in cos(theta) (any order) instead of the 31-node Patterson rule in theta;
  * LAPACK solves instead of row-pivoted Gauss-Jordan; far field from Appendix A.7 instead of `addprc`.

What it deliberately shares with the Fortran are its INPUT conventions and, when quirks=True, the two single-precision
constants that change the physical problem being solved (tmatrix_py.f:161 cpi = 4.0*atan(1.0) in REAL*4 -> k, the
integration limit cpi/2; :617-618 aovrb**(1./3.) with a REAL*4 exponent), and the fixed truncation nrank = 17, m <= 6.
With quirks=False it is the exact-pi solution (differs by ~1e-7).

`rayleigh_confocal` is the analytic electrostatic polarizability of a confocal coated spheroid (Bohren & Huffman 1983,
eq. 5.35): an anchor for ar != 1 with distinct layers that involves no T-matrix code at all.
"""
import numpy as np
from scipy.special import spherical_jn, spherical_yn

CPI32 = float(np.float32(4.0) * np.arctan(np.float32(1.0)))      # 3.1415927410125732
THIRD32 = float(np.float32(1.0) / np.float32(3.0))


def _radial(n_max, z):
    """f_n(z), n = 0..n_max for f in (j, y): arrays [n_max+1, len(z)] (complex z allowed)."""
    n = np.arange(n_max + 1)[:, None]
    return spherical_jn(n, z[None, :]), spherical_yn(n, z[None, :])


def _wigner(m, n_max, x):
    """d^n_{0m}(theta) and tau_n = d/dtheta d^n_{0m}, n = 0..n_max; x = cos(theta)."""
    s = np.sqrt(1.0 - x * x)
    d = np.zeros((n_max + 2, x.size))
    tau = np.zeros((n_max + 1, x.size))
    if m == 0:
        d[0], d[1] = 1.0, x
        for n in range(1, n_max + 1):
            d[n + 1] = ((2 * n + 1) * x * d[n] - n * d[n - 1]) / (n + 1)
            tau[n] = (n * (n + 1) / (2 * n + 1)) * (d[n + 1] - d[n - 1]) / s
    else:
        d[m] = s**m * np.prod(np.sqrt((2.0 * np.arange(1, m + 1) - 1.0) / (2.0 * np.arange(1, m + 1))))
        for n in range(m, n_max + 1):
            d[n + 1] = ((2 * n + 1) * x * d[n] - np.sqrt(n * n - m * m) * d[n - 1]) / np.sqrt((n + 1.0)**2 - m * m)
            tau[n] = (-(n + 1) * np.sqrt(n * n - m * m) * d[n - 1] + n * np.sqrt((n + 1.0)**2 - m * m) * d[n + 1]) / ((2 * n + 1) * s)
    return d[:n_max + 1], tau


def _q_matrix(m, n_max, k, mrel, r, u, x, w, outer, inner):
    """Generalised Q matrix of one surface for azimuthal mode m (Appendix A.5 / A.6 with the outside radial function
    F in {'h', 'j'} at z = k r and the inside function G in {'j', 'h'} at zeta = mrel k r).  Returns [2 NM, 2 NM]."""
    nmin = max(m, 1)
    ns = np.arange(nmin, n_max + 1)
    z = k * r
    zeta = mrel * z
    jz, yz = _radial(n_max, z)
    jq, yq = _radial(n_max, zeta)
    F = jz + 1j * yz if outer == "h" else jz.astype(complex)
    G = jq + 1j * yq if inner == "h" else jq.astype(complex)
    nn = ns[:, None]
    Fn, Ft = F[ns], F[ns - 1] - nn * F[ns] / z[None, :]
    Gn, Gt = G[ns], G[ns - 1] - nn * G[ns] / zeta[None, :]
    d, tau = _wigner(m, n_max, x)
    d, tau = d[ns], tau[ns]
    s = np.sqrt(1.0 - x * x)
    W = w * r * r
    cn = (ns * (ns + 1.0))
    # index order [n1, n2, node]
    A11 = d[:, None, :] * d[None, :, :]
    A12 = d[:, None, :] * tau[None, :, :]
    A21 = tau[:, None, :] * d[None, :, :]
    A22 = tau[:, None, :] * tau[None, :, :]
    b1 = Gn[None, :, :] * Fn[:, None, :]
    b2 = Gn[None, :, :] * Ft[:, None, :]
    b4 = Gt[None, :, :] * Fn[:, None, :]
    b6 = Gt[None, :, :] * Ft[:, None, :]
    b3, b5, b7, b8 = b1 / z, b1 / zeta, b4 / z, b2 / zeta
    c1, c2 = cn[:, None, None], cn[None, :, None]
    X11 = ((m * W / s) * (A12 + A21) * b1).sum(-1)
    X12 = (W * (m * m * A11 / (s * s) + A22) * b2 + W * u * c1 * A12 * b3).sum(-1)
    X21 = (W * (m * m * A11 / (s * s) + A22) * b4 + W * u * c2 * A21 * b5).sum(-1)
    X22 = ((m * W / s) * (A12 + A21) * b6 + (m * W * u / s) * A11 * (c1 * b7 + c2 * b8)).sum(-1)
    # mirror symmetry (half-range quadrature doubled): X11, X22 vanish for n1 + n2 even, X12, X21 for n1 + n2 odd
    odd = ((ns[:, None] + ns[None, :]) % 2 == 1)
    X11, X22, X12, X21 = X11 * odd, X22 * odd, X12 * ~odd, X21 * ~odd
    ann = 0.5 * np.sqrt((2 * ns[:, None] + 1.0) * (2 * ns[None, :] + 1.0) / (cn[:, None] * cn[None, :]))
    t11, t22, t12, t21 = -X11 * ann, -X22 * ann, -1j * X12 * ann, 1j * X21 * ann
    k2 = k * k
    return np.block([[k2 * mrel * t21 + k2 * t12, k2 * mrel * t11 + k2 * t22],
                     [k2 * mrel * t22 + k2 * t11, k2 * mrel * t12 + k2 * t21]])


def coated_spheroid_tmatrix(k0, a_h_out, ar_vh_out, a_h_in, ar_vh_in, m1, m2, n_max=17, m_max=6, n_gauss=96, x_lo=0.0):
    """List over m = 0..m_max of the [2 NM, 2 NM] T blocks of a coaxial two-layer spheroid.

    k0: vacuum wavenumber; a_h_*: horizontal (equatorial) semi-axes; ar_vh_* = vertical / horizontal semi-axis (the
    Fortran's aovrb, < 1 oblate); m1, m2: refractive indices of shell and core.  Mirror symmetry: the quadrature runs over
    cos(theta) in [x_lo, 1] and is doubled (x_lo = cos(cpi/2) reproduces the Fortran's integration limit)."""
    xg, wg = np.polynomial.legendre.leggauss(n_gauss)
    x = 0.5 * (1.0 - x_lo) * xg + 0.5 * (1.0 + x_lo)
    w = 2.0 * 0.5 * (1.0 - x_lo) * wg

    def rs(a_h, ar):
        s2 = 1.0 - x * x
        q = 1.0 / np.sqrt((x / ar)**2 + s2)                                  # genkr: kr = conk qb
        return a_h * q, x * np.sqrt(s2) * (1.0 / ar**2 - 1.0) * q * q         # r, (dr/dtheta) / r

    r1, u1 = rs(a_h_out, ar_vh_out)
    r2, u2 = rs(a_h_in, ar_vh_in)
    out = []
    for m in range(m_max + 1):
        q2 = {fg: _q_matrix(m, n_max, k0 * m1, m2 / m1, r2, u2, x, w, *fg) for fg in (("h", "j"), ("j", "j"))}
        T2 = -np.linalg.solve(q2[("h", "j")].T, q2[("j", "j")].T).T
        q1 = {fg: _q_matrix(m, n_max, k0, m1, r1, u1, x, w, *fg) for fg in (("h", "j"), ("j", "j"), ("h", "h"), ("j", "h"))}
        num = q1[("j", "j")] + q1[("j", "h")] @ T2
        den = q1[("h", "j")] + q1[("h", "h")] @ T2
        out.append(-np.linalg.solve(den.T, num.T).T)
    return out


def amplitudes_horizontal(Tm, k0, n_max, theta_inc=np.pi / 2):
    """(S_vv, S_hh) at back (phi - phi0 = pi) and forward (0) scattering for horizontal incidence, particle axis vertical
    (Appendix A.7 with alpha = beta = 0, no angle nudges).  Returns (Svv_back, Shh_back, Svv_fwd, Shh_fwd) in 1/k0 units."""
    x = np.array([np.cos(theta_inc)])
    res = []
    for dphi in (np.pi, 0.0):
        vv = hh = 0.0j
        for m, T in enumerate(Tm):
            nmin = max(m, 1)
            ns = np.arange(nmin, n_max + 1)
            nm = ns.size
            d, tau = _wigner(m, n_max, x)
            pi_n, tau_n = d[ns, 0] / np.sqrt(1.0 - x[0]**2), tau[ns, 0]
            n1, n2 = ns[:, None], ns[None, :]
            C = (1j)**((n2 - n1 - 1) % 4) * np.sqrt((2 * n1 + 1.0) * (2 * n2 + 1.0) / (n1 * (n1 + 1.0) * n2 * (n2 + 1.0)))
            T11, T12, T21, T22 = T[:nm, :nm], T[:nm, nm:], T[nm:, :nm], T[nm:, nm:]
            if m == 0:
                tt = tau_n[:, None] * tau_n[None, :]
                vv += (C * tt * T22).sum()
                hh += (C * tt * T11).sum()
            else:
                D11 = m * m * pi_n[:, None] * pi_n[None, :]
                D12 = m * pi_n[:, None] * tau_n[None, :]
                D21 = m * tau_n[:, None] * pi_n[None, :]
                D22 = tau_n[:, None] * tau_n[None, :]
                f = 2.0 * np.cos(m * dphi)
                vv += f * (C * (T11 * D11 + T21 * D21 + T12 * D12 + T22 * D22)).sum()
                hh += f * (C * (T11 * D22 + T21 * D12 + T12 * D21 + T22 * D11)).sum()
        res += [vv / k0, hh / k0]
    return tuple(res)


def tm2l(d_ext_cm, d_int_cm, ar_ext, ar_int, eps_ext, eps_int, wavelength_cm, quirks=True, n_max=17, m_max=6, n_gauss=96):
    """Same inputs and output layout as oracle.tm2l / tm2l.tmatrix (one particle): 8 doubles
    borrp borip borrpe boripe forrp forip forpe foripe, metres (sign convention: SURVEY.md Appendix C)."""
    pi = CPI32 if quirks else np.pi
    third = THIRD32 if quirks else 1.0 / 3.0
    lam_m = wavelength_cm / 100.0
    k0 = 2.0 * pi / lam_m
    a_out = (d_ext_cm / 200.0) / ar_ext**third          # horizontal semi-axis [m]: conk = cpi dpart / (alamd aovrb**(1./3.))
    a_in = (d_int_cm / 200.0) / ar_int**third
    m1, m2 = np.sqrt(complex(eps_ext)), np.sqrt(complex(eps_int))
    x_lo = float(np.cos(pi / 2.0)) if quirks else 0.0
    Tm = coated_spheroid_tmatrix(k0, a_out, ar_ext, a_in, ar_int, m1, m2, n_max, m_max, n_gauss, x_lo)
    vb, hb, vf, hf = amplitudes_horizontal(Tm, k0, n_max)
    return np.array([-vb.real, -vb.imag, -hb.real, -hb.imag, vf.real, vf.imag, hf.real, hf.imag])


# ---- analytic anchor ---------------------------------------------------------------------------------------------
def _depol_oblate_axis(c_over_a):
    """Depolarisation factor along the symmetry axis of an oblate spheroid (semi-axes a = b > c); the two equatorial
    factors are (1 - L) / 2 (Bohren & Huffman eq. 5.34)."""
    e2 = 1.0 - c_over_a**2
    e = np.sqrt(e2)
    return (1.0 / e2) * (1.0 - (np.sqrt(1.0 - e2) / e) * np.arcsin(e))


def rayleigh_confocal(a_h_out, ar_vh_out, a_h_in, eps1, eps2, k0):
    """Rayleigh-limit (S_vv, S_hh) of a CONFOCAL coated oblate spheroid, symmetry axis vertical, horizontal incidence:
    S = k0^2 alpha / (4 pi) with the coated-ellipsoid polarizability of Bohren & Huffman eq. 5.35.  Returns
    (S_vv, S_hh, ar_vh_in): the inner axis ratio is fixed by confocality (a^2 - c^2 equal for both surfaces)."""
    c_out = a_h_out * ar_vh_out
    f2 = a_h_out**2 - c_out**2
    c_in = np.sqrt(a_h_in**2 - f2)
    ar_in = c_in / a_h_in
    vol = 4.0 / 3.0 * np.pi * a_h_out**2 * c_out
    frac = (a_h_in**2 * c_in) / (a_h_out**2 * c_out)
    Lz_out, Lz_in = _depol_oblate_axis(ar_vh_out), _depol_oblate_axis(ar_in)
    res = []
    for L1, L2 in ((Lz_in, Lz_out), ((1.0 - Lz_in) / 2.0, (1.0 - Lz_out) / 2.0)):   # (core, outer) factors: vertical, horizontal
        ea = eps1 + (eps2 - eps1) * (L1 - frac * L2)
        num = (eps1 - 1.0) * ea + frac * eps1 * (eps2 - eps1)
        den = ea * (1.0 + (eps1 - 1.0) * L2) + frac * L2 * eps1 * (eps2 - eps1)
        res.append(k0**2 * vol * num / den / (4.0 * np.pi))
    return res[0], res[1], ar_in


# ---- single-layer use of the same blocks -------------------------------------------------------------------------
def spheroid_amplitudes(radius_mm, wavelength_mm, m, axis_ratio, n_max, n_gauss=96):
    """Homogeneous spheroid (pytmatrix conventions: equal-volume radius, axis_ratio = horizontal / vertical) through the
    generalised Q blocks above: T = -Q[j, j] Q[h, j]^-1 per azimuthal mode at a GIVEN truncation n_max, then the far field at
    horizontal incidence.  Returns (S_vv back, S_hh back, S_vv forward, S_hh forward) in mm, the units of pytmatrix's S --
    an independent check of the single-layer oracle's numerics at its own converged order; the
    convergence controller (row P2) is not involved."""
    k0 = 2.0 * np.pi / wavelength_mm
    ar_vh = 1.0 / axis_ratio
    a_h = radius_mm * axis_ratio ** (1.0 / 3.0)
    xg, wg = np.polynomial.legendre.leggauss(n_gauss)
    x = 0.5 * xg + 0.5
    w = wg  # half range, doubled
    s2 = 1.0 - x * x
    q = 1.0 / np.sqrt((x / ar_vh)**2 + s2)
    r, u = a_h * q, x * np.sqrt(s2) * (1.0 / ar_vh**2 - 1.0) * q * q
    Tm = []
    for mm in range(n_max + 1):
        qh = _q_matrix(mm, n_max, k0, complex(m), r, u, x, w, "h", "j")
        qj = _q_matrix(mm, n_max, k0, complex(m), r, u, x, w, "j", "j")
        Tm.append(-np.linalg.solve(qh.T, qj.T).T)
    return amplitudes_horizontal(Tm, k0, n_max)

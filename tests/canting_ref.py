"""Test infrastructure: compact numpy restatement of the reference's canting path, used as the checker of the CUDA kernels
where the reference tree is absent (GPU box).  tests/test_canting_cpu.py pins it to the UNMODIFIED reference functions
(scattering.compute_scattering_canting_sp / _psd / _mixture, scattering.py:182-448, 590-673) and to the committed golden
fixture tests/golden/rain_C10.json (written by the unmodified scattering_rain.py)."""
import numpy as np

COLS = ("sca_xsect_h", "sca_xsect_v", "ext_xsect_h", "ext_xsect_v", "ldr_h", "ldr_v", "refl_h", "refl_v", "zdr", "rho_hv",
        "delta_hv", "kdp", "A_h", "A_v", "Adp")


def moments_analytical(sigma_deg):
    """scattering.py:136-179"""
    s = np.asarray(sigma_deg, dtype=float) * np.pi / 180.0
    r = np.exp(-2.0 * s * s)
    p, q = 3 / 8 + r / 2 + r**4 / 8, 3 / 8 - r / 2 + r**4 / 8
    return dict(a1=(1 + r)**2 / 4, a2=(1 - r**2) / 4, a3=p * p, a4=q * p, a5=p * (1 - r**4) / 8, a7=r * (1 + r) / 2)


def _finish(sh, sv, eh, ev, dco, kd, ld, wl, kw):
    pref = wl**4 * 1e6 / (np.pi**5 * kw)
    sca_h, sca_v = 4 * np.pi * sh, 4 * np.pi * sv
    ext_h, ext_v = 2 * wl / 1000 * eh, 2 * wl / 1000 * ev
    with np.errstate(all="ignore"):
        out = [10 * np.log10(sca_h), 10 * np.log10(sca_v), 10 * np.log10(ext_h), 10 * np.log10(ext_v),
               10 * np.log10(4 * np.pi * ld / sca_h), 10 * np.log10(4 * np.pi * ld / sca_v),
               10 * np.log10(pref * sca_h), 10 * np.log10(pref * sca_v), 10 * np.log10(sca_h / sca_v),
               4 * np.pi * np.abs(dco) / np.sqrt(sca_h * sca_v), -np.angle(dco, deg=True), kd * 180 / np.pi * wl,
               2 * 4.34e3 * ext_h, 2 * 4.34e3 * ext_v, 2 * 4.34e3 * ext_h - 2 * 4.34e3 * ext_v]
    return np.stack(out, axis=-1)


def _terms(fv180, fh180, fv0, fh0, a):
    j0, j1 = np.abs(fh180)**2, np.abs(fh180 - fv180)**2
    j2, j3 = np.conj(fh180) * (fh180 - fv180), fh0 - fv0
    return (j0 - 2 * j2.real * a["a2"] + j1 * a["a4"], j0 - 2 * j2.real * a["a1"] + j1 * a["a3"],
            fh0.imag - j3.imag * a["a2"], fh0.imag - j3.imag * a["a3"],
            j0 + j1 * a["a5"] - j2 * a["a1"] - np.conj(j2) * a["a2"], j3.real * a["a7"], j1 * a["a5"])


def canting_sp(wl, fv180, fh180, fv0, fh0, a, kw=0.93):
    """[n_D, 15] (scattering.py:182-309)"""
    return _finish(*_terms(fv180, fh180, fv0, fh0, a), wl, kw)


def canting_psd(wl, fv180, fh180, fv0, fh0, a, delta_d, psd_vals, kw=0.93):
    """[n_dsd, 15] for psd_vals [n_dsd, n_D] (scattering.py:312-448)"""
    w = np.atleast_2d(psd_vals) * delta_d[None, :]
    return _finish(*[(w * t[None, :]).sum(1) for t in _terms(fv180, fh180, fv0, fh0, a)], wl, kw)


def mixture(A, B):
    """rows [..., 15] (scattering.py:590-673); cross-section columns NaN"""
    lin = lambda x: np.power(10.0, 0.1 * x)
    o = np.full(A.shape, np.nan)
    rh, rv = 10 * np.log10(lin(A[..., 6]) + lin(B[..., 6])), 10 * np.log10(lin(A[..., 7]) + lin(B[..., 7]))
    o[..., 6], o[..., 7], o[..., 8] = rh, rv, rh - rv
    o[..., 4] = 10 * np.log10(lin(A[..., 4]) * lin(A[..., 6]) + lin(B[..., 4]) * lin(B[..., 6])) - rh
    o[..., 5] = 10 * np.log10(lin(A[..., 5]) * lin(A[..., 7]) + lin(B[..., 5]) * lin(B[..., 7])) - rv
    o[..., 9] = ((A[..., 9] * np.sqrt(lin(A[..., 6]) * lin(A[..., 7])) + B[..., 9] * np.sqrt(lin(B[..., 6]) * lin(B[..., 7])))
                 / np.sqrt(lin(rh) * lin(rv)))
    for k in (10, 11, 12, 13, 14):
        o[..., k] = A[..., k] + B[..., k]
    return o


def rain_dsd_grid():
    """the 9 300 gamma DSDs of scattering_rain.py:236-267, in the script's order: (D0, Nw, mu)"""
    nw_aux = np.power(10, np.arange(1, 4 + 0.1, 0.1))
    mu_aux = np.arange(-1, 8 + 1, 1)
    d0_aux = np.arange(0.1, 3 + 0.1, 0.1)
    n_d0, n_nw, n_mu = d0_aux.size, nw_aux.size, mu_aux.size
    d0, nw, mu = (np.zeros(n_d0 * n_nw * n_mu) for _ in range(3))
    for i0, a in enumerate(d0_aux):
        for i1, b in enumerate(nw_aux):
            for i2, c in enumerate(mu_aux):
                k = i2 * n_d0 * n_nw + i1 * n_d0 + i0
                d0[k], nw[k], mu[k] = a, b, c
    return d0, nw, mu


def hail_dsd_grid():
    """the 11 011 exponential DSDs of scattering_hail.py:205-227 (Cheng et al. 1985): (N0, Lambda)"""
    lamb_aux = np.arange(0.1, 1.0 + 0.01, 0.01)
    c_aux = np.arange(60, 300 + 2, 2)
    n_l = lamb_aux.size
    n0, lamb = np.zeros(n_l * c_aux.size), np.zeros(n_l * c_aux.size)
    for il, l in enumerate(lamb_aux):
        for ic, c in enumerate(c_aux):
            n0[il + ic * n_l] = c * np.power(l, 4.11)
            lamb[il + ic * n_l] = l
    return n0, lamb


def golden_rain_amplitudes(fx):
    """(fv180, fh180, fv0, fh0) as scattering_rain.py:185-193 forms them from the fixture's S elements"""
    Sb, Sf = np.array(fx["S_back_vv_hh"]), np.array(fx["S_forw_vv_hh"])
    return (1e-3 * (Sb[:, 0] + 1j * Sb[:, 1]), -1e-3 * (Sb[:, 2] + 1j * Sb[:, 3]),
            1e-3 * (Sf[:, 0] + 1j * Sf[:, 1]), 1e-3 * (Sf[:, 2] + 1j * Sf[:, 3]))


def moments_mc(canting_angle, ele=0.0, n_part=500000, seed=0):
    """scattering.py:70-133 (compute_angular_moments): numpy default_rng stream, one alpha vector, one beta vector per sigma."""
    rng = np.random.default_rng(seed=seed)
    thet0 = np.pi / 2 - ele * np.pi / 180.0
    canting_angle = np.atleast_1d(np.asarray(canting_angle, dtype=float))
    out = {k: np.zeros(canting_angle.size) for k in ("a1", "a2", "a3", "a4", "a5", "a7")}
    alpha = rng.uniform(low=0.0, high=360.0, size=n_part) * np.pi / 180.0
    for ind, cant_angle in enumerate(canting_angle):
        beta = rng.normal(scale=cant_angle, size=n_part) * np.pi / 180.0
        a1 = np.power(np.cos(beta) * np.sin(thet0) - np.sin(beta) * np.cos(thet0) * np.cos(alpha), 2.0)
        a2 = np.sin(beta) * np.sin(beta) * np.sin(alpha) * np.sin(alpha)
        for k, v in (("a1", a1), ("a2", a2), ("a3", a1 * a1), ("a4", a2 * a2), ("a5", a1 * a2), ("a7", a1 - a2)):
            out[k][ind] = np.mean(v)
    return out

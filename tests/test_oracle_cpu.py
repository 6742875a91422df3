"""CPU suite: the single-layer oracle against the committed golden vector and physics identities."""
import json
import os

import numpy as np
import pytest
from scipy.special import spherical_jn, spherical_yn

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def test_kat_single():
    k = json.load(open(os.path.join(HERE, "golden", "kat_single.json")))
    s = k["scatterer"]
    tm = oracle.TMatrix(s["radius"], s["wavelength"], complex(*s["m"]), s["axis_ratio"], ddelt=s["ddelt"], ndgs=s["ndgs"])
    assert (tm.nmax, tm.ngauss, tm.ntrials) == (k["converged"]["nmax"], k["converged"]["ngauss"], k["converged"]["ntrials"])
    S, Z = tm.ampl(*k["geometry"])
    Sref = np.array(k["S"])[..., 0] + 1j * np.array(k["S"])[..., 1]
    Zref = np.array(k["Z"])
    tol = k["tolerance_rel"]
    for i in range(2):
        assert abs(S[i, i] - Sref[i, i]) / abs(Sref[i, i]) < tol
    # cross-polar terms are ~1e-24: compare absolutely against the co-polar scale, and relatively to 1e-3
    assert np.abs(S - Sref).max() / np.abs(Sref).max() < tol
    assert abs(S[0, 1] - Sref[0, 1]) / abs(Sref[0, 1]) < 1e-3
    big = np.abs(Zref) > 1e-6
    assert np.abs(Z[big] - Zref[big]).max() / np.abs(Zref).max() < tol


def test_ddelt_rescale_switch_same_nmax_on_kat():
    a = oracle.TMatrix(2.0, 6.5, 1.5 + 0.5j, 1 / 0.6, rescale_ddelt=True)
    b = oracle.TMatrix(2.0, 6.5, 1.5 + 0.5j, 1 / 0.6, rescale_ddelt=False)
    assert a.nmax == b.nmax == 7


def _mie_ab(x, m, nmax):
    n = np.arange(1, nmax + 1)
    jx, jmx = spherical_jn(n, x), spherical_jn(n, m * x)
    djx = spherical_jn(n, x, derivative=True)
    djmx = spherical_jn(n, m * x, derivative=True)
    hx = jx + 1j * spherical_yn(n, x)
    dhx = djx + 1j * spherical_yn(n, x, derivative=True)
    psi, dpsi = x * jx, jx + x * djx
    psim, dpsim = m * x * jmx, jmx + m * x * djmx
    xi, dxi = x * hx, hx + x * dhx
    a = (m * psim * dpsi - psi * dpsim) / (m * psim * dxi - xi * dpsim)
    b = (psim * dpsi - m * psi * dpsim) / (psim * dxi - m * xi * dpsim)
    return a, b


def test_sphere_limit_is_mie():
    """T11_nn = -b_n, T22_nn = -a_n for a (nearly) spherical particle."""
    r, lam, m = 1.3, 6.0, 1.7 + 0.3j
    tm = oracle.TMatrix(r, lam, m, 1.0000001)
    t = tm.tdata()
    nmax = tm.nmax
    blk = t[:4 * nmax * nmax].reshape(4, nmax, nmax)
    a, b = _mie_ab(2 * np.pi * r / lam, m, nmax)
    assert np.abs(np.diag(blk[0]) + b).max() < 2e-6
    assert np.abs(np.diag(blk[3]) + a).max() < 2e-6
    off = blk[0] - np.diag(np.diag(blk[0]))
    assert np.abs(off).max() < 1e-6


def test_reciprocity_and_optical_theorem():
    tm = oracle.TMatrix(1.8, 8.0, 1.6 + 0.0j, 1.5)
    # reciprocity: S(-n_inc, -n_sca) = [[S11, -S21], [-S12, S22]](n_sca, n_inc)
    t0, t1, p0, p1 = 40.0, 110.0, 20.0, 250.0
    S, _ = tm.ampl(t0, t1, p0, p1, 30.0, 25.0)
    Sr, _ = tm.ampl(180 - t1, 180 - t0, (p1 + 180) % 360, (p0 + 180) % 360, 30.0, 25.0)
    expect = np.array([[S[0, 0], -S[1, 0]], [-S[0, 1], S[1, 1]]])
    assert np.abs(Sr - expect).max() / np.abs(S).max() < 1e-4  # limited by the ddelt truncation
    # optical theorem for a lossless particle: C_ext = C_sca (h polarisation, brute-force quadrature)
    x, w = np.polynomial.legendre.leggauss(20)
    nph = 40
    sca = 0.0
    for xi, wi in zip(x, w):
        for p in (np.arange(nph) + 0.5) * 360.0 / nph:
            _, Z = tm.ampl(90.0, np.degrees(np.arccos(xi)), 0.0, p)
            sca += wi * (Z[0, 0] - Z[0, 1])
    sca *= 2 * np.pi / nph
    Sf, _ = tm.ampl(90.0, 90.0, 0.0, 0.0)
    ext = 2 * 8.0 * Sf[1, 1].imag
    assert abs(sca - ext) / ext < 1e-4


def test_rain_grid_convergence_table():
    """Converged NMAX / NGAUSS ranges and trial totals of the 70-drop rain grid (SURVEY.md section 6 probe)."""
    from tests.workloads import rain_grid
    expect = {"S": (5, 6, 11, 13, 240), "C": (5, 8, 11, 17, 254), "X": (5, 10, 11, 21, 295)}
    for band, e in expect.items():
        w = rain_grid(band)
        r = oracle.sl_batch(w["radius"], w["wavelength"], w["mr"], w["mi"], w["axis_ratio"])
        assert (r["nmax"].min(), r["nmax"].max(), r["ngauss"].min(), r["ngauss"].max(), r["ntrials"].sum()) == e
        assert (r["status"] == 0).all()
        assert (r["ngauss"] == 2 * r["nmax"] + 1).all()


def test_gauss_matches_numpy():
    for n in (4, 22, 34, 102):
        z, w = oracle.gauss(n)
        zn, wn = np.polynomial.legendre.leggauss(n)
        assert np.abs(z - zn).max() < 1e-14 and np.abs(w - wn).max() < 1e-14

"""GPU: fused scatter kernel, PSD integration, observables and the pytmatrix shim end-to-end against the oracle."""
import numpy as np
import torch
import pytest

import oracle
from tests.workloads import rain_grid

pytestmark = pytest.mark.gpu

GB = (90.0, 90.0, 0.0, 180.0)
GF = (90.0, 90.0, 0.0, 0.0)


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_fused_scatter_equals_ampl_and_xsect(engine):
    w = rain_grid("X")
    tb = engine.calctmat(w["radius"], w["wavelength"], w["mr"], w["mi"], w["axis_ratio"])
    al = np.repeat(np.linspace(0, 360, 6)[:-1], 10)
    be = np.tile(np.linspace(1.0, 62.0, 10), 5)
    wt = np.tile(np.linspace(0.3, 0.01, 10), 5)
    geoms = [GB, GF, (60.0, 120.0, 10.0, 200.0), (0.0, 180.0, 0.0, 0.0), (90.0, 35.0, 0.0, 77.0)]
    S1, Z1 = tb.ampl(geoms, al, be, weight=wt, scale=0.37, reduce=True)
    x1 = tb.xsect([(g[0], g[2]) for g in geoms], al, be, wt, 0.37)
    S2, Z2, x2 = tb.scatter(geoms, al, be, weight=wt, scale=0.37)
    assert _rel(S2.cpu().numpy(), S1.cpu().numpy()) < 1e-12
    assert _rel(Z2.cpu().numpy(), Z1.cpu().numpy()) < 1e-12
    assert _rel(x2.cpu().numpy(), x1.cpu().numpy()) < 1e-12
    # more orientations than one pass of the kernel holds (133 > 64), three scattered directions for one incidence (two
    # geometry blocks), a second incidence with one
    al = np.repeat(np.linspace(0, 360, 8)[:-1], 19)
    be = np.tile(np.linspace(0.5, 179.0, 19), 7)
    wt = np.tile(np.linspace(0.2, 0.02, 19), 7)
    geoms = [GB, GF, (90.0, 40.0, 0.0, 65.0), (30.0, 150.0, 20.0, 200.0)]
    S1, Z1 = tb.ampl(geoms, al, be, weight=wt, scale=0.11, reduce=True)
    x1 = tb.xsect([(g[0], g[2]) for g in geoms], al, be, wt, 0.11)
    S2, Z2, x2 = tb.scatter(geoms, al, be, weight=wt, scale=0.11)
    assert _rel(S2.cpu().numpy(), S1.cpu().numpy()) < 1e-12
    assert _rel(Z2.cpu().numpy(), Z1.cpu().numpy()) < 1e-12
    assert _rel(x2.cpu().numpy(), x1.cpu().numpy()) < 1e-12


def test_orientation_averaged_table_matches_oracle(engine):
    from hydroscatt_b200.tables import fixed_orientation_nodes
    al, be, wt, scale = fixed_orientation_nodes(10.0)
    w = rain_grid("C")
    sel = np.arange(0, 70, 9)
    tb = engine.calctmat(w["radius"][sel], w["wavelength"][sel], w["mr"][sel], w["mi"][sel], w["axis_ratio"][sel])
    S, Z, xs = tb.scatter([GB, GF], al, be, weight=wt, scale=scale)
    S, Z = S.cpu().numpy(), Z.cpu().numpy()
    ref = oracle.sl_batch(w["radius"][sel], w["wavelength"][sel], w["mr"][sel], w["mi"][sel], w["axis_ratio"][sel],
                          geoms=[GB[:4], GF[:4]], alpha=al, beta=be, wgt=wt)
    assert _rel(S, ref["S"] * scale) < 1e-6
    big = np.abs(ref["Z"]) > 1e-9 * np.abs(ref["Z"]).max()
    assert (np.abs(Z - ref["Z"] * scale)[big] / np.abs(ref["Z"] * scale)[big]).max() < 1e-6
    assert (xs.cpu().numpy() > 0).all()


def test_psd_integration_and_observables(engine):
    """PSD-integrated observables of compute_scattering_psd_tm against numpy trapz + the radar.* formulas."""
    from hydroscatt_b200 import tables
    al, be, wt, scale = tables.fixed_orientation_nodes(10.0)
    w = rain_grid("C")
    rows, tb = tables.build_rows(engine, w["radius"], w["wavelength"], w["mr"], w["mi"], w["axis_ratio"], GB, GF,
                                 al, be, wt, scale)
    D = w["D"]
    d0 = np.array([0.8, 1.5, 2.5, 1.5])
    nw = np.array([8000.0, 3000.0, 500.0, 8000.0])
    mu = np.array([0.0, 3.0, -1.0, 0.0])
    obs, integ = tables.integrate_psd(engine, rows, 1, D, "gamma", d0, nw, mu, np.full(4, 7.0), [53.5],
                                      hpol_quirk=False)
    obs = obs.cpu().numpy()[:, 0]
    R = rows.cpu().numpy()
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    from pytmatrix.psd import GammaPSD
    trapz = getattr(np, "trapezoid", None) or np.trapz
    for i in range(4):
        nd = GammaPSD(D0=d0[i], Nw=nw[i], mu=mu[i], D_max=7.0)(D)
        I = trapz(R * nd[:, None], D, axis=0)
        assert _rel(integ.cpu().numpy()[i, 0], I) < 1e-12
        Zb = I[8:24].reshape(4, 4)
        Sf = I[24:32]
        wl = 53.5
        xh = 2 * np.pi * (Zb[0, 0] - Zb[0, 1] - Zb[1, 0] + Zb[1, 1])
        xv = 2 * np.pi * (Zb[0, 0] + Zb[0, 1] + Zb[1, 0] + Zb[1, 1])
        refl_h = 10 * np.log10(wl**4 / (np.pi**5 * 0.93) * xh)
        zdr = 10 * np.log10(xh / xv)
        kdp = 1e-3 * 180 / np.pi * wl * (Sf[6] - Sf[0])
        rho = np.sqrt(((Zb[2, 2] + Zb[3, 3])**2 + (Zb[3, 2] - Zb[2, 3])**2) / ((xh / (2 * np.pi)) * (xv / (2 * np.pi))))
        Ah = 2 * 4.343e-3 * I[54]
        assert abs(obs[i, 2] - refl_h) < 1e-9 and abs(obs[i, 4] - zdr) < 1e-9
        assert abs(obs[i, 11] - kdp) < 1e-12 and abs(obs[i, 7] - rho) < 1e-12 and abs(obs[i, 12] - Ah) < 1e-12
    # regression anchor (restatement run): D0=1.5, Nw=8000, mu=0 -> refl_h = 40.4 dBZ on the
    # rectangle-rule canting path (asserted in tests/test_canting_gpu.py); the trapz / pytmatrix path with the 10 deg
    # Gaussian orientation average must land within a few tenths of a dB of it, with a rain-like Zdr
    assert abs(obs[3, 2] - 40.4) < 0.5, obs[3, 2]
    assert 0.3 < obs[3, 4] < 2.0 and obs[3, 11] > 0.0


def test_scatter_lane_splits_agree(engine, monkeypatch):
    """k_scatter<SPLIT>: 1, 2 or 4 lanes per orientation (row slices of a mode added up by shuffles) must give the same
    orientation-averaged S / Z / cross sections; more than 1024 particles so that the three NMAX classes are launched."""
    from hydroscatt_b200 import tables
    al, be, wt, scale = tables.fixed_orientation_nodes(10.0)
    ws = [rain_grid(b) for b in ("S", "C", "X", "Ku", "Ka", "W")]
    cat = lambda k: np.concatenate([w[k] for w in ws] * 3)
    tb = engine.calctmat(cat("radius"), cat("wavelength"), cat("mr"), cat("mi"), cat("axis_ratio"))
    assert tb.n >= 1024 and int(tb.nmax.max()) > 20 and int(tb.nmax.min()) <= 10

    def run(split):
        monkeypatch.setenv("HYDROTM_SC_SPLIT", split)
        S, Z, xs = tb.scatter([GB, GF], al, be, weight=wt, scale=scale, want_xs=True)
        monkeypatch.delenv("HYDROTM_SC_SPLIT")
        return [t.cpu().numpy().copy() for t in (torch.view_as_real(S), Z, xs)]

    ref = run("1,1,1")
    for split in ("2,2,2", "4,4,4", "4,2,1"):
        for a, b in zip(ref, run(split)):
            assert np.isfinite(b).all()
            assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max(), split


def test_fused_psd_observables_equal_unfused(engine):
    """hs_psd_observables (integration + observables in one kernel, nothing but the observables written) against the
    two-kernel path, several tables, ragged DSD / table counts, both settings of the h_pol quirk."""
    from hydroscatt_b200 import tables
    al, be, wt, scale = tables.fixed_orientation_nodes(10.0)
    ws = [rain_grid(b) for b in ("S", "C", "X", "Ku", "W")]
    cat = lambda k: np.concatenate([w[k] for w in ws])
    rows, tb = tables.build_rows(engine, cat("radius"), cat("wavelength"), cat("mr"), cat("mi"), cat("axis_ratio"), GB, GF,
                                 al, be, wt, scale)
    D = ws[0]["D"]
    rng = np.random.default_rng(5)
    nd = 131
    d0, nw, mu = rng.uniform(0.3, 3.0, nd), 10 ** rng.uniform(1, 4, nd), rng.integers(-1, 8, nd).astype(float)
    lam = [w["wavelength"][0] for w in ws]
    for quirk in (True, False):
        o1, _ = tables.integrate_psd(engine, rows, len(ws), D, "gamma", d0, nw, mu, np.full(nd, 7.0), lam, hpol_quirk=quirk)
        o2, none = tables.integrate_psd(engine, rows, len(ws), D, "gamma", d0, nw, mu, np.full(nd, 7.0), lam, hpol_quirk=quirk,
                                        want_integ=False)
        assert none is None
        o1, o2 = o1.cpu().numpy(), o2.cpu().numpy()
        assert np.isnan(o1[..., :2]).all() and np.isnan(o2[..., :2]).all()
        assert np.abs(o1[..., 2:] - o2[..., 2:]).max() < 1e-10  # dB / deg / deg km-1


def test_shim_end_to_end_on_gpu(engine):
    """pytmatrix shim on the CUDA back-end: the scattering_rain.py PSD flow for a few DSDs vs the oracle back-end."""
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    from pytmatrix import orientation, psd, radar, scatter, tmatrix_aux, refractive
    from pytmatrix import tmatrix as tmod
    from pytmatrix.tmatrix import Scatterer
    from tests.oracle_backend import OracleBackend

    def run(backend, D_max=3.0, npts=6):
        tmod.set_backend(backend)
        wl = tmatrix_aux.wl_C
        tm = Scatterer(wavelength=wl, m=refractive.m_w_10C[wl])
        tm.orient = orientation.orient_averaged_fixed
        tm.or_pdf = orientation.gaussian_pdf(10.0)
        out = {}
        tm.radius, tm.axis_ratio = 1.0, 1.0 / tmatrix_aux.dsr_thurai_2007(2.0)
        tm.set_geometry(tmatrix_aux.geom_horiz_back)
        out["sp_refl"] = radar.refl(tm)
        out["sp_zdr"] = radar.Zdr(tm)
        tm.set_geometry(tmatrix_aux.geom_horiz_forw)
        out["sp_kdp"] = radar.Kdp(tm)
        out["sp_ext"] = scatter.ext_xsect(tm)
        tm.psd_integrator = psd.PSDIntegrator()
        tm.psd_integrator.axis_ratio_func = lambda D: 1.0 / tmatrix_aux.dsr_thurai_2007(D)
        tm.psd_integrator.D_max = D_max
        tm.psd_integrator.num_points = npts
        tm.psd_integrator.geometries = (tmatrix_aux.geom_horiz_back, tmatrix_aux.geom_horiz_forw)
        tm.psd_integrator.init_scatter_table(tm, angular_integration=(backend is None))
        tm.psd = psd.GammaPSD(D0=1.2, Nw=5e3, mu=1, D_max=D_max)
        tm.set_geometry(tmatrix_aux.geom_horiz_back)
        out["refl"] = radar.refl(tm)
        out["zdr"] = radar.Zdr(tm)
        out["rho"] = radar.rho_hv(tm)
        out["ldr"] = radar.ldr(tm)
        tm.set_geometry(tmatrix_aux.geom_horiz_forw)
        out["kdp"] = radar.Kdp(tm)
        if backend is None:
            out["Ai"] = radar.Ai(tm)
            out["sca"] = scatter.sca_xsect(tm)
        tmod.set_backend(None)
        return out
    g = run(None)
    o = run(OracleBackend())
    for k in o:
        assert abs(g[k] - o[k]) <= 1e-6 * abs(o[k]), (k, g[k], o[k])
    assert g["Ai"] > 0 and g["sca"] > 0


def test_shim_cylinder_on_gpu(engine):
    """Scatterer(shape=SHAPE_CYLINDER) through the shim: per-call S / Z and a small orientation-averaged scatter table on
    the CUDA back-end against the oracle back-end."""
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    from pytmatrix import orientation, radar, tmatrix_aux
    from pytmatrix import tmatrix as tmod
    from pytmatrix.tmatrix import Scatterer
    from tests.oracle_backend import OracleBackend

    def run(backend):
        tmod.set_backend(backend)
        tm = Scatterer(radius=1.5, wavelength=tmatrix_aux.wl_X, m=complex(1.78, 0.002), axis_ratio=0.5,
                       shape=Scatterer.SHAPE_CYLINDER)
        tm.set_geometry(tmatrix_aux.geom_horiz_back)
        S, Z = tm.get_SZ()
        out = [np.asarray(S).ravel().view(float), np.asarray(Z).ravel()]
        tm.orient = orientation.orient_averaged_fixed
        tm.or_pdf = orientation.gaussian_pdf(20.0)
        out.append(np.array([radar.refl(tm), radar.Zdr(tm), radar.ldr(tm)]))
        tmod.set_backend(None)
        return out

    for a, b in zip(run(None), run(OracleBackend())):
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max()


def test_golden_rain_amplitudes_through_shim(engine):
    """The amplitudes the unmodified reference script feeds to its canting formulas (fixture generated by running
    scattering_rain.py on the oracle back-end, tools/gen_golden_rain.py) reproduced by the shim on the CUDA back-end."""
    import json
    import os
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    from pytmatrix import orientation, tmatrix_aux, refractive
    from pytmatrix import tmatrix as tmod
    from pytmatrix.tmatrix import Scatterer
    tmod.set_backend(None)
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rain_C10.json")))
    wl = tmatrix_aux.wl_C
    tm = Scatterer(wavelength=wl, m=refractive.m_w_10C[wl])
    tm.orient = orientation.orient_single
    for i in range(0, 70, 4):
        d = fx["D"][i]
        tm.radius = d / 2.0
        tm.axis_ratio = 1.0 / tmatrix_aux.dsr_thurai_2007(d)
        tm.set_geometry(tmatrix_aux.geom_horiz_back)
        S = tm.get_S()
        ref = fx["S_back_vv_hh"][i]
        assert abs(S[0, 0] - complex(ref[0], ref[1])) <= 1e-6 * abs(complex(ref[0], ref[1]))
        assert abs(S[1, 1] - complex(ref[2], ref[3])) <= 1e-6 * abs(complex(ref[2], ref[3]))
        tm.set_geometry(tmatrix_aux.geom_horiz_forw)
        S = tm.get_S()
        ref = fx["S_forw_vv_hh"][i]
        assert abs(S[0, 0] - complex(ref[0], ref[1])) <= 1e-6 * abs(complex(ref[0], ref[1]))
        assert abs(S[1, 1] - complex(ref[2], ref[3])) <= 1e-6 * abs(complex(ref[2], ref[3]))


def test_asym_product_rule_matches_brute_force(engine):
    """scatter.asym via the exact product rule against a much finer brute-force quadrature of the oracle's Z."""
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    from pytmatrix import scatter, tmatrix_aux, orientation
    from pytmatrix import tmatrix as tmod
    from pytmatrix.tmatrix import Scatterer
    tmod.set_backend(None)
    tm = Scatterer(radius=2.0, wavelength=6.5, m=complex(1.5, 0.5), axis_ratio=1.0 / 0.6)
    tm.orient = orientation.orient_single
    tm.set_geometry((60.0, 60.0, 10.0, 10.0, 30.0, 20.0))
    g_h, g_v = scatter.asym(tm, True), scatter.asym(tm, False)
    o = oracle.TMatrix(2.0, 6.5, 1.5 + 0.5j, 1 / 0.6)
    x, w = np.polynomial.legendre.leggauss(24)
    nph = 48
    num = np.zeros(2)
    den = np.zeros(2)
    t0, p0 = np.radians(60.0), np.radians(10.0)
    for xi, wi in zip(x, w):
        th = np.arccos(xi)
        for p in (np.arange(nph) + 0.5) * 2 * np.pi / nph:
            _, Z = o.ampl(60.0, np.degrees(th), 10.0, np.degrees(p), 30.0, 20.0)
            c = np.sin(t0) * np.sin(th) * np.cos(p - p0) + np.cos(t0) * np.cos(th)
            for k, s in enumerate((-1.0, 1.0)):
                I = Z[0, 0] + s * Z[0, 1]
                num[k] += wi * I * c
                den[k] += wi * I
    assert abs(g_h - num[0] / den[0]) < 1e-6 and abs(g_v - num[1] / den[1]) < 1e-6
    assert 0.0 < g_h < 1.0
    assert abs(scatter.ssa(tm) - scatter.sca_xsect(tm) / scatter.ext_xsect(tm)) < 1e-12


def test_calctmat_scatter_one_call_equals_two_calls(engine, monkeypatch):
    """hs_calctmat_scatter_batch (also with the scatter kernels overlapping the convergence rounds) gives the same
    tables as hs_calctmat_batch followed by hs_scatter_batch."""
    from hydroscatt_b200 import tables
    al, be, wt, scale = tables.fixed_orientation_nodes(10.0)
    ws = [rain_grid(b) for b in ("C", "Ka", "W")]
    cat = lambda k: np.concatenate([np.tile(w[k], 6) for w in ws])  # 1260 particles: the ordered / class-split scatter path
    args = [cat(k) for k in ("radius", "wavelength", "mr", "mi", "axis_ratio")]
    tb = engine.calctmat(*args)
    S0, Z0, x0 = tb.scatter([GB, GF], al, be, weight=wt, scale=scale, want_xs=True)
    for overlap in (False, True):
        if overlap:
            monkeypatch.setenv("HYDROTM_OVERLAP_SCATTER", "1")
        tb1, S1, Z1, x1 = engine.calctmat_scatter(*args, [GB[:4], GF[:4]], al, be, weight=wt, scale=scale, want_xs=True)
        monkeypatch.delenv("HYDROTM_OVERLAP_SCATTER", raising=False)
        assert (tb1.nmax == tb.nmax).all() and (tb1.ngauss == tb.ngauss).all() and (tb1.status == tb.status).all()
        for a, b in ((S0, S1), (Z0, Z1), (x0, x1)):
            a, b = a.cpu().numpy(), b.cpu().numpy()
            assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max(), overlap

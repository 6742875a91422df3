"""GPU: the batched canting-angle observables (hs_canting_observables / hs_canting_mixture) against
the numpy restatement of the reference functions (tests/canting_ref.py, pinned to the unmodified reference on the CPU suite) and
against the golden fixture of the unmodified scattering_rain.py; PSD weight kinds 1 / 2 and the rectangle rule (row P11, S5)."""
import json
import os

import numpy as np
import pytest
import torch

from tests import canting_ref as cr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FX = json.load(open(os.path.join(ROOT, "tests", "golden", "rain_C10.json")))


def _shim():
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    from pytmatrix import tmatrix as tmod
    tmod.set_backend(None)


def test_canting_kernels_match_reference_restatement_and_golden(engine):
    from hydroscatt_b200 import canting
    D = np.array(FX["D"])
    f = cr.golden_rain_amplitudes(FX)
    a = canting.compute_angular_moments_analytical(10.0)
    sp = canting.compute_scattering_canting_sp(53.5, *f, a, eng=engine).cpu().numpy()[0]
    assert np.allclose(sp, cr.canting_sp(53.5, *f, a), rtol=1e-12, atol=1e-12, equal_nan=True)
    assert np.abs(sp[FX["sp_rows"]] - np.array(FX["sp"])[:, :15]).max() < 1e-9       # dB, deg, deg/km: <= 1e-9
    # all 9 300 DSDs of scattering_rain.py:236-267 in one launch; weights from hs_psd_weights (GammaPSD)
    d0, nw, mu = cr.rain_dsd_grid()
    delta_d = D - np.append(0.0, D[:-1])
    N = engine.psd_weights("gamma", d0, nw, mu, np.full(d0.size, 7.0), D, np.ones_like(D))
    n0 = engine.launch_count
    obs, sums = canting.compute_scattering_canting_psd(53.5, *f, a, delta_d, N, extra=np.stack([D**3, D**2], 1), eng=engine)
    assert engine.launch_count - n0 == 1
    obs, sums, Nh = obs.cpu().numpy()[:, 0], sums.cpu().numpy(), N.cpu().numpy()
    ref = cr.canting_psd(53.5, *f, a, delta_d, Nh)
    ok = np.isfinite(ref)
    assert (np.isfinite(obs) == ok).all()
    assert np.abs(obs[ok] - ref[ok]).max() < 1e-9
    assert np.allclose(sums[:, 0], (Nh * delta_d * D**3).sum(1), rtol=1e-13) and np.allclose(sums[:, 1], (Nh * delta_d * D**2).sum(1), rtol=1e-13)
    gold = np.array(FX["psd"])
    rows = FX["psd_rows"]
    assert np.allclose(np.pi / 6.0 * sums[rows, 0], gold[:, 3], rtol=1e-9)               # lwc column (precip.compute_lwc)
    for k, name in enumerate(FX["psd_columns"][5:]):
        assert np.abs(obs[rows, cr.COLS.index(name)] - gold[:, 5 + k]).max() < 1e-9, name
    # SURVEY.md section 4 anchor: D0 = 1.5 mm, Nw = 8000, mu = 0 at C band, 10 degC -> ~40.4 dBZ on the rectangle-rule canting path
    Nmp = engine.psd_weights("gamma", [1.5], [8000.0], [0.0], [7.0], D, np.ones_like(D))
    mp = canting.compute_scattering_canting_psd(53.5, *f, a, delta_d, Nmp, eng=engine).cpu().numpy()[0, 0]
    assert abs(mp[6] - 40.4) < 0.1


def test_reference_call_sequence_on_cuda_amplitudes(engine):
    """scattering_rain.py:176-200 call sequence (orient_single, set_geometry back / forward, get_S, 1e-3 scaling and the sign
    flip on fh180) through the pytmatrix shim on the CUDA back-end, then the canting kernels: rows of the golden CSVs."""
    _shim()
    from pytmatrix import tmatrix_aux, refractive, orientation
    from pytmatrix.tmatrix import Scatterer
    from hydroscatt_b200 import canting
    wl = tmatrix_aux.wl_C
    sc = Scatterer(wavelength=wl, m=refractive.m_w_10C[wl])
    sc.orient = orientation.orient_single
    diam = np.arange(0.1, 7.0 + 0.1, 0.1)
    fv180, fh180, fv0, fh0 = (np.empty(diam.size, dtype=complex) for _ in range(4))
    for ind, d_part in enumerate(diam):
        sc.radius = d_part / 2.0
        sc.axis_ratio = 1.0 / tmatrix_aux.dsr_thurai_2007(d_part)
        sc.set_geometry(tmatrix_aux.geom_horiz_back)
        s = sc.get_S()
        fv180[ind], fh180[ind] = 1e-3 * s[0][0], -1e-3 * s[1][1]
        sc.set_geometry(tmatrix_aux.geom_horiz_forw)
        s = sc.get_S()
        fv0[ind], fh0[ind] = 1e-3 * s[0][0], 1e-3 * s[1][1]
    gf = cr.golden_rain_amplitudes(FX)
    for x, g in zip((fv180, fh180, fv0, fh0), gf):
        assert np.abs(x - g).max() <= 1e-6 * np.abs(g).max()
    a = canting.compute_angular_moments_analytical(10.0)
    sp = canting.compute_scattering_canting_sp(wl, fv180, fh180, fv0, fh0, a, eng=engine).cpu().numpy()[0]
    gold = np.array(FX["sp"])[:, :15]
    err = np.abs(sp[FX["sp_rows"]] - gold)
    # spheres (D < 0.7 mm, axis ratio 1): fh - fv is rounding noise, ldr = -300 dB carries no information
    noise = gold[:, 4] < -150.0
    assert noise.sum() == 2 and (sp[FX["sp_rows"]][noise, 4:6] < -150.0).all()
    err[noise, 4:6] = 0.0
    # north-star bar: observables within 0.01 dB / 0.01 deg/km; measured far inside it.  ldr (a 1e-6 ratio of the amplitudes'
    # DIFFERENCE) amplifies the 1e-7 amplitude error most.
    assert err.max() < 1e-3, err.max(0)
    d0, nw, mu = cr.rain_dsd_grid()
    rows = np.array(FX["psd_rows"])
    N = engine.psd_weights("gamma", d0[rows], nw[rows], mu[rows], np.full(rows.size, 7.0), diam, np.ones_like(diam))
    obs = canting.compute_scattering_canting_psd(wl, fv180, fh180, fv0, fh0, a, diam - np.append(0.0, diam[:-1]), N,
                                                 eng=engine).cpu().numpy()[:, 0]
    g = np.array(FX["psd"])
    for k, name in enumerate(FX["psd_columns"][5:]):
        assert np.abs(obs[:, cr.COLS.index(name)] - g[:, 5 + k]).max() < 1e-3, name


def test_hail_grid_exponential_psd_per_diameter_canting_and_mixture(engine):
    """The scattering_hail.py:205-372 shape of the problem: 11 011 exponential DSDs (hs_psd_weights kind 1 against the shim's
    ExponentialPSD), 800 particles with a size-dependent canting angle (moments per diameter), a second species on another
    diameter grid, several tables (bands) per launch, and the mixture of the two."""
    _shim()
    from pytmatrix.psd import ExponentialPSD, UnnormalizedGammaPSD
    from hydroscatt_b200 import canting
    rng = np.random.default_rng(11)
    n_D, n_tab = 800, 3
    d_init = np.linspace(0.1, 80.0, n_D)
    n0, lamb = cr.hail_dsd_grid()
    assert n0.size == 11011
    prob_break = np.clip(rng.uniform(-0.5, 0.6, n_D), 0.0, 1.0)
    delta_d = d_init - np.append(0.0, d_init[:-1])
    # kind 1 weights with wq = (1 - prob_break): psd_vals_hail = psd_no_break - psd_no_break * prob_break (scattering_hail.py:305-309)
    N = engine.psd_weights("exponential", n0, lamb, None, np.full(n0.size, d_init[-1]), d_init, 1.0 - prob_break)
    Nh = N.cpu().numpy()
    for d in (0, 5000, 11010):
        ref = ExponentialPSD(N0=n0[d], Lambda=lamb[d], D_max=d_init[-1])(d_init) * (1.0 - prob_break)
        assert np.allclose(Nh[d], ref, rtol=1e-13, atol=0)
    # default D_max = 11 / Lambda cut-off and the unnormalised gamma (kind 2), incl. D = 0
    Dz = np.concatenate([[0.0], d_init[:99]])
    lam2, mu2, n02 = rng.uniform(0.3, 3.0, 64), rng.uniform(-0.9, 6.0, 64), rng.uniform(1.0, 1e4, 64)
    W1 = engine.psd_weights("exponential", n02, lam2, None, 11.0 / lam2, Dz, np.ones_like(Dz)).cpu().numpy()
    W2 = engine.psd_weights("ungamma", n02, lam2, mu2, 11.0 / lam2, Dz, np.ones_like(Dz)).cpu().numpy()
    for d in range(64):
        assert np.allclose(W1[d], ExponentialPSD(N0=n02[d], Lambda=lam2[d])(Dz), rtol=1e-13, atol=0)
        assert np.allclose(W2[d], UnnormalizedGammaPSD(N0=n02[d], Lambda=lam2[d], mu=mu2[d])(Dz), rtol=1e-12, atol=0)
        assert W2[d, 0] == 0.0 and (W1[d, Dz > 11.0 / lam2[d]] == 0.0).all()
    # amplitudes: Rayleigh-like growth with size, distinct per table; extinction positive
    x = d_init / 30.0
    f = []
    for k in range(4):
        re = (1.0 + 0.2 * k) * x[None, :]**3 * (1.0 + 0.1 * rng.normal(size=(n_tab, n_D))) * 1e-3
        im = 0.3 * x[None, :]**3 * (0.5 + rng.uniform(size=(n_tab, n_D))) * 1e-3
        f.append(re + 1j * im)
    sig = 5.0 + 55.0 * rng.uniform(size=n_D)
    a = canting.compute_angular_moments_analytical(sig)
    lam_tab = np.array([53.5, 33.3, 111.0])
    obs = canting.compute_scattering_canting_psd(lam_tab, *f, a, delta_d, N, eng=engine)
    o = obs.cpu().numpy()
    for t in range(n_tab):
        ref = cr.canting_psd(lam_tab[t], *[x_[t] for x_ in f], a, delta_d, Nh)
        ok = np.isfinite(ref)
        assert (np.isfinite(o[:, t]) == ok).all() and ok.mean() > 0.99
        assert np.abs(o[:, t][ok] - ref[ok]).max() < 1e-9, t
    sp = canting.compute_scattering_canting_sp(lam_tab, *f, a, eng=engine).cpu().numpy()
    for t in range(n_tab):
        ref = cr.canting_sp(lam_tab[t], *[x_[t] for x_ in f], a)
        ok = np.isfinite(ref)
        assert np.abs(sp[t][ok] - ref[ok]).max() < 1e-9
    # second species: rain on its own grid with scalar moments; mixture of the two frames
    d_drop = np.arange(0.1, 8.0 + 0.1, 0.1)
    fr = [np.array(g_)[None, :] * np.ones((n_tab, 1)) for g_ in [np.interp(d_drop, np.array(FX["D"]), v.real) + 1j * np.interp(d_drop, np.array(FX["D"]), v.imag)
                                                           for v in cr.golden_rain_amplitudes(FX)]]
    Nr = rng.uniform(0.0, 50.0, (n0.size, d_drop.size))
    ar = canting.compute_angular_moments_analytical(10.0)
    obs_r = canting.compute_scattering_canting_psd(lam_tab, *fr, ar, d_drop - np.append(0.0, d_drop[:-1]), Nr, eng=engine)
    mix = canting.compute_scattering_mixture(obs, obs_r, eng=engine).cpu().numpy()
    ref = cr.mixture(o, obs_r.cpu().numpy())
    ok = np.isfinite(ref)
    assert (np.isfinite(mix) == ok).all()
    assert np.abs(mix[ok] - ref[ok]).max() < 1e-9
    assert np.isnan(mix[..., :4]).all()


def test_rect_rule_psd_integration_matches_numpy(engine):
    """integrate_psd(rule='rect') = the rectangle rule delta_d = D - D_prev of scattering_rain.py:271 through hs_psd_weights +
    hs_psd_integrate / hs_psd_observables, for every PSD kind, against numpy sums of the shim's PSD classes."""
    _shim()
    from pytmatrix.psd import GammaPSD, ExponentialPSD, UnnormalizedGammaPSD
    from hydroscatt_b200 import tables
    from tests.workloads import rain_grid
    al, be, wt, scale = tables.fixed_orientation_nodes(10.0)
    w = rain_grid("X")
    rows, _ = tables.build_rows(engine, w["radius"], w["wavelength"], w["mr"], w["mi"], w["axis_ratio"], (90.0, 90.0, 0.0, 180.0),
                                (90.0, 90.0, 0.0, 0.0), al, be, wt, scale)
    R = rows.cpu().numpy()
    D = w["D"]
    dd = D - np.append(0.0, D[:-1])
    trap = np.zeros_like(D)
    trap[:-1] += 0.5 * np.diff(D)
    trap[1:] += 0.5 * np.diff(D)
    cases = {"gamma": ([1.2, 2.0], [8000.0, 900.0], [0.0, 4.0], lambda a, b, c: GammaPSD(D0=a, Nw=b, mu=c, D_max=7.0)),
             "exponential": ([8000.0, 120.0], [2.2, 0.9], None, lambda a, b, c: ExponentialPSD(N0=a, Lambda=b, D_max=7.0)),
             "ungamma": ([5000.0, 30.0], [2.5, 1.5], [1.0, 3.0], lambda a, b, c: UnnormalizedGammaPSD(N0=a, Lambda=b, mu=c, D_max=7.0))}
    for kind, (p0, p1, p2, mk) in cases.items():
        for rule, wq in (("rect", dd), ("trapz", trap)):
            obs, integ = tables.integrate_psd(engine, rows, 1, D, kind, p0, p1, p2, [7.0, 7.0], [33.3], rule=rule, hpol_quirk=False)
            obs_f, _ = tables.integrate_psd(engine, rows, 1, D, kind, p0, p1, p2, [7.0, 7.0], [33.3], rule=rule, hpol_quirk=False,
                                            want_integ=False)
            integ, obs, obs_f = integ.cpu().numpy()[:, 0], obs.cpu().numpy()[:, 0], obs_f.cpu().numpy()[:, 0]
            for d in range(2):
                nd = mk(p0[d], p1[d], None if p2 is None else p2[d])(D)
                ref = (R * (nd * wq)[:, None]).sum(0)
                assert np.abs(integ[d] - ref).max() <= 1e-12 * np.abs(ref).max(), (kind, rule)
            assert np.abs(obs[:, 2:] - obs_f[:, 2:]).max() < 1e-9, (kind, rule)
    # rect and trapz are different rules on purpose: they must not coincide
    o_r, _ = tables.integrate_psd(engine, rows, 1, D, "gamma", [1.2], [8000.0], [0.0], [7.0], [33.3], rule="rect")
    o_t, _ = tables.integrate_psd(engine, rows, 1, D, "gamma", [1.2], [8000.0], [0.0], [7.0], [33.3], rule="trapz")
    assert 1e-4 < abs(float(o_r[0, 0, 2] - o_t[0, 0, 2])) < 0.5


def test_scatter_classes_keep_flagged_particles_in_bounds(engine):
    """on the n >= 1024 path of hs_scatter_batch a particle whose status is NGAUSS_LIMIT / SINGULAR (valid T-matrix,
    toff >= 0) must be launched in the NMAX class its order belongs to, not in class [0, 10]: same result as with status OK."""
    from hydroscatt_b200 import tables, engine as E
    from tests.workloads import rain_grid
    al, be, wt, scale = tables.fixed_orientation_nodes(10.0)
    ws = [rain_grid(b) for b in ("S", "C", "X", "Ku", "Ka", "W")]
    cat = lambda k: np.concatenate([w[k] for w in ws] * 3)
    tb = engine.calctmat(cat("radius"), cat("wavelength"), cat("mr"), cat("mi"), cat("axis_ratio"))
    assert tb.n >= 1024
    geoms = [(90.0, 90.0, 0.0, 180.0), (90.0, 90.0, 0.0, 0.0)]
    S0, Z0, x0 = (t.clone() for t in tb.scatter(geoms, al, be, weight=wt, scale=scale))
    heavy = int(torch.argmax(tb.nmax).item())
    mid = int(torch.nonzero((tb.nmax > 12) & (tb.nmax < 20))[0].item())
    assert int(tb.nmax[heavy]) > 20
    tb.status[heavy] = E.ST_NGAUSS_LIMIT
    tb.status[mid] = E.ST_SINGULAR
    S1, Z1, x1 = tb.scatter(geoms, al, be, weight=wt, scale=scale)
    assert torch.isfinite(torch.view_as_real(S1)).all()
    assert torch.equal(torch.view_as_real(S1), torch.view_as_real(S0)) and torch.equal(Z1, Z0) and torch.equal(x1, x0)
    # a particle without a T-matrix (toff < 0, e.g. NMAX * NDGS beyond NPNG1) is reported as NaN, never read
    tb.toff[heavy] = -1
    S2, _, _ = tb.scatter(geoms, al, be, weight=wt, scale=scale)
    assert torch.isnan(torch.view_as_real(S2[heavy])).all()
    keep = torch.ones(tb.n, dtype=torch.bool, device=S2.device)
    keep[heavy] = False
    assert torch.equal(torch.view_as_real(S2[keep]), torch.view_as_real(S0[keep]))


def test_mc_angular_moments_on_numpy_stream(engine):
    """hs_canting_moments_mc against the numpy restatement of scattering.compute_angular_moments
    (scattering.py:70-133): the device consumes the same PCG64 + ziggurat stream, so the moments agree at ROUNDING level
    (libm vs CUDA sin / cos, summation order), far below the Monte-Carlo noise (~1e-3 at 500 000 samples)."""
    import time
    from hydroscatt_b200 import canting
    sig = np.array([0.5, 1.0, 10.0, 25.0, 40.0, 63.0])
    for ele, n_part, seed in ((0.0, 20000, 0), (30.0, 50001, 7), (0.0, 500000, 0)):
        ref = cr.moments_mc(sig, ele=ele, n_part=n_part, seed=seed)
        got = canting.compute_angular_moments(sig, ele=ele, n_part=n_part, seed=seed, eng=engine)
        for k in canting.MOMENT_KEYS:
            assert np.abs(got[k] - ref[k]).max() < 1e-12, (ele, n_part, k, got[k], ref[k])
    # a different seed or a swapped sigma order changes the stream: the agreement above is not a coincidence of the means
    other = canting.compute_angular_moments(sig, n_part=20000, seed=1, eng=engine)
    ref0 = cr.moments_mc(sig, n_part=20000, seed=0)
    assert np.abs(other["a2"] - ref0["a2"]).max() > 1e-6
    # the hail scripts' size: 800 canting angles x 500 000 samples (scattering_hail.py) in one call; the stream is consumed in
    # order, so the first angles of the long run must equal a short reference run, and the late ones must still be sane
    sig800 = np.linspace(1.0, 60.0, 800)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    big = canting.compute_angular_moments(sig800, eng=engine)
    dt = time.perf_counter() - t0
    print("800 sigmas x 500 000 samples: %.1f ms" % (dt * 1e3))
    ref3 = cr.moments_mc(sig800[:3])
    for k in canting.MOMENT_KEYS:
        assert np.abs(big[k][:3] - ref3[k]).max() < 1e-12, k
        assert np.isfinite(big[k]).all()
    assert (np.diff(big["a1"]) < 2e-3).all() and 0.5 < big["a1"][-1] < big["a1"][0] <= 1.0   # <cos^2 beta>-like: decreasing in sigma
    with pytest.raises(Exception):
        canting.compute_angular_moments(sig, n_part=100, eng=engine)

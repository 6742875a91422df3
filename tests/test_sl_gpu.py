"""GPU parity of the single-layer path (calctmat + calcampl) against the CPU oracle, through the C-ABI."""
import numpy as np
import pytest

import oracle
from tests.layout import packed_to_full
from tests.workloads import rain_grid, WL

pytestmark = pytest.mark.gpu

GEOMS = [(90.0, 90.0, 0.0, 180.0), (90.0, 90.0, 0.0, 0.0)]


@pytest.fixture(params=["host-rounds", "device-rounds"], autouse=True)
def round_driver(request, monkeypatch):
    """Every test of this module runs with both round drivers of hs_calctmat_batch: the host-driven rounds
    (hs_staged_host.h) and the device-driven rounds (hs_staged_dev.h; the default from 256 particles per call)."""
    if request.param == "host-rounds":
        monkeypatch.setenv("HYDROTM_DEVCTL", "0")
    else:
        monkeypatch.setenv("HYDROTM_DEVCTL_MIN", "1")
    return request.param


def _rel(a, b):
    scale = np.abs(b).max()
    return np.abs(a - b).max() / scale


def test_kat_single(engine):
    """pytmatrix test_tmatrix.py::test_single known-answer vector."""
    tb = engine.calctmat(2.0, 6.5, 1.5, 0.5, 1.0 / 0.6)
    assert int(tb.nmax[0]) == 7 and int(tb.ngauss[0]) == 15 and int(tb.status[0]) == 0
    S, Z = tb.ampl([GEOMS[0]])
    S = S.cpu().numpy()[0, 0, 0]
    Z = Z.cpu().numpy()[0, 0, 0]
    S_ref = np.array([[3.89338755e-02 - 2.43467777e-01j, -1.11474042e-24 - 3.75103868e-24j],
                      [1.11461702e-24 + 3.75030914e-24j, -8.38637654e-02 + 3.10409912e-01j]])
    assert abs(S[0, 0] - S_ref[0, 0]) / abs(S_ref[0, 0]) < 1e-6
    assert abs(S[1, 1] - S_ref[1, 1]) / abs(S_ref[1, 1]) < 1e-6
    assert abs(Z[0, 0] - 8.20899248e-02) / 8.2e-2 < 1e-6


@pytest.mark.parametrize("band", ["S", "C", "X", "Ku", "Ka", "W"])
def test_rain_grid_matches_oracle(engine, band):
    w = rain_grid(band)
    tb = engine.calctmat(w["radius"], w["wavelength"], w["mr"], w["mi"], w["axis_ratio"])
    nmax = tb.nmax.cpu().numpy()
    ngauss = tb.ngauss.cpu().numpy()
    ntr = tb.ntrials.cpu().numpy()
    st = tb.status.cpu().numpy()
    S, Z = tb.ampl(GEOMS, alpha=[0.0, 72.0, 144.0], beta=[0.0, 13.4, 61.7])
    S = S.cpu().numpy()
    Z = Z.cpu().numpy()
    for i in range(len(w["D"])):
        tm = oracle.TMatrix(w["radius"][i], w["wavelength"][i], complex(w["mr"][i], w["mi"][i]), w["axis_ratio"][i])
        assert (nmax[i], ngauss[i], ntr[i], st[i]) == (tm.nmax, tm.ngauss, tm.ntrials, tm.status), (band, i)
        tf = packed_to_full(tb.packed(i), tm.nmax)
        assert _rel(tf, tm.tdata()) < 1e-6, (band, i)
        for g, geom in enumerate(GEOMS):
            for o, (a, b) in enumerate([(0.0, 0.0), (72.0, 13.4), (144.0, 61.7)]):
                So, Zo = tm.ampl(*geom, a, b)
                assert _rel(S[i, g, o], So) < 1e-6, (band, i, g, o)
                assert _rel(Z[i, g, o], Zo) < 1e-6, (band, i, g, o)


def test_orientation_reduce(engine):
    w = rain_grid("C")
    tb = engine.calctmat(w["radius"], w["wavelength"], w["mr"], w["mi"], w["axis_ratio"])
    al = np.repeat(np.linspace(0, 360, 6)[:-1], 10)
    be = np.tile(np.linspace(1.0, 62.0, 10), 5)
    wt = np.tile(np.linspace(0.3, 0.01, 10), 5)
    S1, Z1 = tb.ampl(GEOMS, al, be)
    S2, Z2 = tb.ampl(GEOMS, al, be, weight=wt, scale=0.2 / wt[:10].sum(), reduce=True)
    wd = np.asarray(wt)
    Sr = (S1.cpu().numpy() * wd[None, None, :, None, None]).sum(2) * 0.2 / wt[:10].sum()
    Zr = (Z1.cpu().numpy() * wd[None, None, :, None, None]).sum(2) * 0.2 / wt[:10].sum()
    assert _rel(S2.cpu().numpy(), Sr) < 1e-12
    assert _rel(Z2.cpu().numpy(), Zr) < 1e-12


def test_ice_crystal_like_ndgs15(engine):
    """scattering_ice_crystals.py regime: ndgs=15 (NGAUSS up to ~150, exercises the Gauss-table extension), extreme
    axis ratios, RADIUS_MAXIMUM converted to equal volume as Scatterer.equal_volume_from_maximum does."""
    cases = [(0.5, 8.0, 33.3, 1.5 + 0.002j), (1.0, 15.0, 33.3, 1.4 + 0.001j), (1.4, 29.5, 53.5, 1.35 + 0.001j),
             (2.5, 41.0, 33.3, 1.3 + 0.001j), (0.3, 0.5, 8.43, 1.78 + 0.002j)]
    L = np.array([c[0] for c in cases])
    ar = np.array([c[1] for c in cases])
    wl = np.array([c[2] for c in cases])
    m = np.array([c[3] for c in cases])
    req = np.where(ar > 1, (L / 2) / ar ** (1 / 3), (L / 2) / ar ** (2 / 3))
    tb = engine.calctmat(req, wl, m.real, m.imag, ar, ndgs=15)
    S, Z = tb.ampl(GEOMS, alpha=[0.0, 30.0], beta=[0.0, 15.0])
    S = S.cpu().numpy()
    for i in range(len(cases)):
        tm = oracle.TMatrix(req[i], wl[i], m[i], ar[i], ndgs=15)
        # extreme axis ratios are ill-conditioned in FP64 (SURVEY.md 7 hard part 3, section 6: eps >~ 38 "never
        # converges"): the oracle's own S moves by 1e-3 .. 100 % under FMA contraction there.  Such particles are
        # held to 10x the oracle's sensitivity instead of 1e-6, and where the oracle is pure rounding noise
        # (sensitivity > 1e-2) even its convergence decisions are noise: they are listed, not compared.
        sens = oracle.sl_sensitivity(req[i], wl[i], m[i], ar[i], ndgs=15)
        if sens > 1e-2:
            print("excluded (oracle is rounding noise): case", i, cases[i], "sensitivity", sens,
                  "gpu", int(tb.nmax[i]), int(tb.ngauss[i]), "oracle", tm.nmax, tm.ngauss)
            assert int(tb.nmax[i]) == tm.nmax
            continue
        assert (int(tb.nmax[i]), int(tb.ngauss[i]), int(tb.ntrials[i]), int(tb.status[i])) == \
            (tm.nmax, tm.ngauss, tm.ntrials, tm.status), i
        tol = max(1e-6, 10 * sens)
        for g, geom in enumerate(GEOMS):
            for o, (a, b) in enumerate([(0.0, 0.0), (30.0, 15.0)]):
                So, _ = tm.ampl(*geom, a, b)
                assert _rel(S[i, g, o], So) < tol, (i, g, o, sens)


def test_status_codes(engine):
    """bad input -> 5, starting NMAX beyond NPN1 -> 1 (upstream: print + STOP), others unaffected in the same batch."""
    tb = engine.calctmat([1.0, -1.0, 900.0, 1.0], [33.3, 33.3, 33.3, 33.3], [7.9, 7.9, 1.5, 7.9], [2.3, 2.3, 0.0, 2.3],
                         [1.2, 1.2, 1.1, 0.0])
    st = tb.status.cpu().numpy()
    assert st.tolist() == [0, 5, 1, 5]
    S, Z = tb.ampl(GEOMS)
    S = S.cpu().numpy()
    assert np.isfinite(S[0]).all() and np.isnan(S[1]).all() and np.isnan(S[2]).all()
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    from pytmatrix.tmatrix import Scatterer
    from pytmatrix import tmatrix as tmod
    from hydroscatt_b200.backend import ConvergenceError
    tmod.set_backend(None)
    with pytest.raises(ConvergenceError):
        Scatterer(radius=900.0, wavelength=33.3, m=1.5 + 0j, axis_ratio=1.1).get_SZ()


def test_equal_area_radius_and_large_batch_escalation(engine):
    """radius_type = equal surface area (SAREA) and a particle that outgrows every shared-memory bucket."""
    tb = engine.calctmat([2.0], [6.5], [1.5], [0.5], [1 / 0.6], rat=0.0)
    tm = oracle.TMatrix(2.0, 6.5, 1.5 + 0.5j, 1 / 0.6, rat=0.0)
    assert (int(tb.nmax[0]), int(tb.ngauss[0])) == (tm.nmax, tm.ngauss)
    S, _ = tb.ampl([GEOMS[0]])
    So, _ = tm.ampl(*GEOMS[0])
    assert _rel(S.cpu().numpy()[0, 0, 0], So) < 1e-6
    # hail-sized dry ice sphere-ish at W band: NMAX ~ 60 -> global-memory workspace bucket
    tb = engine.calctmat([20.0], [3.19], [1.78], [0.003], [1.05])
    tm = oracle.TMatrix(20.0, 3.19, 1.78 + 0.003j, 1.05)
    assert (int(tb.nmax[0]), int(tb.ngauss[0]), int(tb.status[0])) == (tm.nmax, tm.ngauss, tm.status)
    if tm.status == 0:
        S, _ = tb.ampl([GEOMS[1]])
        So, _ = tm.ampl(*GEOMS[1])
        assert _rel(S.cpu().numpy()[0, 0, 0], So) < 1e-5


def test_staged_pipeline_invariants(engine, monkeypatch):
    """The staged pipeline must not depend on how it is scheduled: speculation depth, number of particle groups,
    scratch budget (forces deferred particles) and an undersized arena (retry through HS_RC_ARENA_FULL) all give the
    same (NMAX, NGAUSS, trials, status) and T-matrices, also with several groups driven by one host thread; the fused one-CTA-per-particle kernel agrees too."""
    w = rain_grid("W")
    sel = slice(None, None, 3)
    args = [w[k][sel] for k in ("radius", "wavelength", "mr", "mi", "axis_ratio")]

    def run(**env):
        for k, v in env.items():
            monkeypatch.setenv(k, str(v))
        engine._arena_hint = {}
        tb = engine.calctmat(*args, arena_margin=env.get("_margin", 1.0))
        for k in env:
            monkeypatch.delenv(k, raising=False)
        meta = [t.cpu().numpy().copy() for t in (tb.nmax, tb.ngauss, tb.ntrials, tb.status)]
        return meta, [tb.packed(i) for i in (0, 7, len(meta[0]) - 1)]

    ref_meta, ref_T = run()
    assert (ref_meta[3] == 0).all()
    for env in (dict(HYDROTM_SPEC=1), dict(HYDROTM_SPEC=5), dict(HYDROTM_GROUPS=1), dict(HYDROTM_GROUPS=3, HYDROTM_SPEC=3),
                dict(HYDROTM_GROUPS=4, HYDROTM_THREADS=2), dict(HYDROTM_GROUPS=5, HYDROTM_THREADS=1, HYDROTM_SCRATCH_GB=0.002),
                dict(HYDROTM_SCRATCH_GB=0.002), dict(_margin=0.05), dict(HYDROTM_STAGED_MIN=1000), dict(HYDROTM_STAGED_MIN=12)):
        meta, T = run(**env)
        for a, b in zip(ref_meta, meta):
            assert (a == b).all(), env
        for a, b in zip(ref_T, T):
            # the fused kernel sums the quadrature in a different order than the DMMA contraction
            tol = 1e-9 if "HYDROTM_STAGED_MIN" in env else 1e-13
            assert np.abs(a - b).max() / np.abs(a).max() < tol, env


@pytest.mark.parametrize("rat", [1.0, 0.0])
def test_cylinders_match_oracle(engine, rat):
    """Finite circular cylinders (Scatterer.SHAPE_CYLINDER, Mishchenko NP = -2: two-segment quadrature split at the edge,
    RSP3 surface, SAREAC for equal-area radii) against the oracle: same (NMAX, NGAUSS, trials), T and S / Z to 1e-6."""
    cases = [(1.0, 6.283, 1.5 + 0.02j, 1.0), (2.0, 6.5, 1.5 + 0.5j, 0.5), (1.5, 8.0, 1.33 + 0.1j, 2.0),
             (0.8, 3.19, 1.78 + 0.002j, 1.5), (0.4, 8.43, 1.78 + 0.002j, 0.6), (2.2, 8.43, 1.78 + 0.0005j, 3.0),
             (3.0, 33.3, 7.35 + 2.78j, 0.8)]
    r, lam, m, eps = (np.array([c[k] for c in cases]) for k in range(4))
    tb = engine.calctmat(r, lam, m.real, m.imag, eps, rat=rat, shape=-2)
    nmax, ngauss, ntr, st = (t.cpu().numpy() for t in (tb.nmax, tb.ngauss, tb.ntrials, tb.status))
    oris = [(0.0, 0.0), (72.0, 13.4), (144.0, 61.7)]
    S, Z = tb.ampl(GEOMS, alpha=[o[0] for o in oris], beta=[o[1] for o in oris])
    S, Z = S.cpu().numpy(), Z.cpu().numpy()
    for i, c in enumerate(cases):
        tm = oracle.TMatrix(c[0], c[1], c[2], c[3], rat=rat, shape=-2)
        assert (nmax[i], ngauss[i], ntr[i], st[i]) == (tm.nmax, tm.ngauss, tm.ntrials, tm.status), (i, c)
        assert st[i] == 0
        assert _rel(packed_to_full(tb.packed(i), tm.nmax), tm.tdata()) < 1e-6, (i, c)
        for g, geom in enumerate(GEOMS):
            for o, (a, b) in enumerate(oris):
                So, Zo = tm.ampl(*geom, a, b)
                assert _rel(S[i, g, o], So) < 1e-6 and _rel(Z[i, g, o], Zo) < 1e-6, (i, g, o)
    # a sphere-like check of the shape itself: the cylinder with the volume of a sphere scatters differently from the sphere
    sph = engine.calctmat(r[:1], lam[:1], m.real[:1], m.imag[:1], eps[:1])
    assert _rel(packed_to_full(sph.packed(0), int(sph.nmax[0]))[:4], packed_to_full(tb.packed(0), int(nmax[0]))[:4]) > 1e-3


def test_device_rounds_equal_host_rounds(engine, monkeypatch):
    """Device-driven rounds (controller, planner and task lists on the GPU) against the host-driven rounds: same stage
    kernels, so (NMAX, NGAUSS, trials, status) are identical and every packed T-matrix is equal bit for bit -- whatever the
    number of groups, the speculation depth, the rounds enqueued per host visit, a scratch pool too small for one
    particle (grown on demand), an undersized arena (HS_RC_ARENA_FULL retry), NMAX > 32 (multi-block assembly + separate
    solve) and Gauss rules beyond the initial table (ndgs = 15)."""
    from tests.workloads import hail_grid, ice_grid
    from hydroscatt_b200.engine import tmat_size
    w = rain_grid("W")
    h = hail_grid(n=40)
    ic = ice_grid(seed=5, n=24)
    cases = [([np.concatenate([w[k], h[k]]) for k in ("radius", "wavelength", "mr", "mi", "axis_ratio")], dict()),
             ([ic[k] for k in ("radius", "wavelength", "mr", "mi", "axis_ratio")], dict(ndgs=15)),
             ([[1.0, -1.0, 900.0, 1.0, 2.0], [33.3] * 5, [7.9, 7.9, 1.5, 7.9, 7.9], [2.3, 2.3, 0.0, 2.3, 2.3], [1.2, 1.2, 1.1, 0.0, 1.3]], dict())]

    def run(args, kw, **env):
        for k, v in env.items():
            if not k.startswith("_"):
                monkeypatch.setenv(k, str(v))
        engine._arena_hint = {}
        tb = engine.calctmat(*args, arena_margin=env.get("_margin", 1.0), **kw)
        for k in env:
            monkeypatch.delenv(k, raising=False)
        meta = [t.cpu().numpy().copy() for t in (tb.nmax, tb.ngauss, tb.ntrials, tb.status)]
        toff = tb.toff.cpu().numpy()
        arena = tb.arena.cpu().numpy()
        T = [arena[2 * toff[i]:2 * (toff[i] + tmat_size(int(meta[0][i])))].copy() if toff[i] >= 0 else None for i in range(tb.n)]
        return meta, T

    for args, kw in cases:
        monkeypatch.delenv("HYDROTM_DEVCTL_MIN", raising=False)
        ref_meta, ref_T = run(args, kw, HYDROTM_DEVCTL=0)
        for env in (dict(), dict(HYDROTM_DEV_GROUPS=1), dict(HYDROTM_DEV_GROUPS=4, HYDROTM_SPEC=1), dict(HYDROTM_SPEC=5, HYDROTM_DEV_ROUNDS=1),
                    dict(HYDROTM_SCRATCH_GB=0.0005, HYDROTM_DEV_GROUPS=2), dict(_margin=0.05), dict(HYDROTM_ONE_LANE=1, HYDROTM_DEV_ROUNDS=2)):
            meta, T = run(args, kw, HYDROTM_DEVCTL=1, HYDROTM_DEVCTL_MIN=1, **env)
            for a, b in zip(ref_meta, meta):
                assert (a == b).all(), (env, np.nonzero(a != b)[0][:8], a[a != b][:8], b[a != b][:8])
            for a, b in zip(ref_T, T):
                assert (a is None) == (b is None), env
                if a is not None:
                    assert np.array_equal(a, b), env

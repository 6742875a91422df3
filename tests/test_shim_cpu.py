"""CPU suite: host logic of the pytmatrix / tm2l drop-in shims (oracle-backed back-end injected; no GPU)."""
import os
import sys

import numpy as np
import pytest

import hydroscatt_b200

hydroscatt_b200.install_shims()

from pytmatrix import orientation, psd, quadrature, radar, scatter, tmatrix_aux, refractive  # noqa: E402
from pytmatrix import tmatrix as tmod  # noqa: E402
from pytmatrix.tmatrix import Scatterer  # noqa: E402
from tests.oracle_backend import OracleBackend  # noqa: E402

REF = "/root/reference/hydroscatt"


@pytest.fixture(autouse=True)
def _backend():
    tmod.set_backend(OracleBackend())
    yield
    tmod.set_backend(None)


def test_quadrature_check_vector():
    """SURVEY.md Appendix A.9."""
    p, w = quadrature.get_points_and_weights(orientation.gaussian_pdf(10.0), 0, 180, 10)
    ref_p = [1.278174, 4.135278, 8.284836, 13.449503, 19.422382, 26.087092, 33.421177, 41.514338, 50.650619, 61.689073]
    ref_w = [2.705251e-02, 1.361438e-01, 2.780417e-01, 3.049208e-01, 1.847799e-01, 5.925695e-02, 9.200167e-03,
             5.930498e-04, 1.184094e-05, 3.534975e-08]
    assert np.allclose(p, ref_p, atol=1e-6)
    assert np.allclose(w, ref_w, rtol=1e-6)
    assert abs(w.sum() - 1.0000008) < 1e-7


def test_psd_classes():
    D = np.linspace(0.1, 7.0, 70)
    g = psd.GammaPSD(D0=1.5, Nw=8000.0, mu=0.0, D_max=7.0)
    v = g(D)
    assert abs(v[9] - 8000.0 * np.exp(-3.67 * 1.0 / 1.5)) < 1e-9  # f(0) = 1
    assert g(8.0) == 0.0 and g(0.0) == 0.0
    assert psd.GammaPSD(D0=1.5, Nw=8000.0, mu=0.0, D_max=7.0) == g and psd.GammaPSD(D0=1.6) != g
    e = psd.ExponentialPSD(N0=100.0, Lambda=2.0)
    assert e.D_max == 5.5 and e(6.0) == 0.0 and abs(e(1.0) - 100 * np.exp(-2)) < 1e-12
    u = psd.UnnormalizedGammaPSD(N0=2.0, Lambda=1.5, mu=2.0, D_max=5.0)
    assert abs(u(2.0) - 2.0 * 4.0 * np.exp(-3.0)) < 1e-12
    b = psd.BinnedPSD(np.array([0.0, 1.0, 2.0]), np.array([5.0, 7.0]))
    assert b(0.5) == 5.0 and b(1.5) == 7.0 and b(2.5) == 0.0 and list(b(np.array([0.5, 1.5]))) == [5.0, 7.0]


def test_aux_tables():
    assert tmatrix_aux.geom_horiz_back == (90.0, 90.0, 0.0, 180.0, 0.0, 0.0)
    assert tmatrix_aux.dsr_thurai_2007(0.5) == 1.0
    assert abs(tmatrix_aux.dsr_thurai_2007(4.0) - 0.7897008) < 1e-7
    assert refractive.m_w_10C[tmatrix_aux.wl_C] == complex(8.601, 1.687)


def test_scatterer_kat_and_caching():
    calls = {"t": 0, "a": 0}
    be = tmod.get_backend()
    ct, ca = be.calctmat, be.calcampl

    def cct(*a, **k):
        calls["t"] += 1
        return ct(*a, **k)

    def cca(*a, **k):
        calls["a"] += 1
        return ca(*a, **k)
    be.calctmat, be.calcampl = cct, cca
    tm = Scatterer(radius=2.0, wavelength=6.5, m=complex(1.5, 0.5), axis_ratio=1.0 / 0.6)
    S, Z = tm.get_SZ()
    assert tm.nmax == 7
    assert abs(S[0, 0] - complex(3.89338755e-02, -2.43467777e-01)) < 1e-7
    assert abs(Z[2, 3] - 8.33266362e-03) < 1e-8
    tm.get_S()
    tm.get_Z()
    radar.refl(tm)
    assert calls == {"t": 1, "a": 1}
    tm.set_geometry(tmatrix_aux.geom_horiz_forw)
    radar.Kdp(tm)
    assert calls == {"t": 1, "a": 2}
    tm.radius = 2.1
    tm.get_SZ()
    assert calls == {"t": 2, "a": 3}
    tm.set_geometry(tmatrix_aux.geom_horiz_back)
    with pytest.raises(ValueError):
        radar.Kdp(tm)


def test_orient_averaged_fixed_matches_manual_loop():
    tm = Scatterer(radius=1.5, wavelength=33.3, m=complex(7.942, 2.332), axis_ratio=1.0 / 0.85)
    tm.orient = orientation.orient_averaged_fixed
    tm.or_pdf = orientation.gaussian_pdf(10.0)
    S, Z = tm.get_SZ()
    # the upstream Python loop, spelled out
    S2 = np.zeros((2, 2), complex)
    Z2 = np.zeros((4, 4))
    ap = np.linspace(0, 360, tm.n_alpha + 1)[:-1]
    for a in ap:
        for b, w in zip(tm.beta_p, tm.beta_w):
            s, z = tm.get_SZ_single(alpha=a, beta=b)
            S2 += w * s
            Z2 += w * z
    S2 *= (1.0 / tm.n_alpha) / tm.beta_w.sum()
    Z2 *= (1.0 / tm.n_alpha) / tm.beta_w.sum()
    assert np.abs(S - S2).max() < 1e-14 and np.abs(Z - Z2).max() < 1e-14
    assert radar.Zdr(tm) > 1.0 and 0.9 < radar.rho_hv(tm) <= 1.0


def test_radius_maximum_conversion():
    tm = Scatterer(radius=1.0, wavelength=33.3, m=1.78 + 0.002j, axis_ratio=4.0, radius_type=Scatterer.RADIUS_MAXIMUM)
    assert abs(tm.equal_volume_from_maximum() - 1.0 / 4.0 ** (1 / 3)) < 1e-15
    tm.axis_ratio = 0.5
    assert abs(tm.equal_volume_from_maximum() - 1.0 / 0.5 ** (2 / 3)) < 1e-15


def test_psd_integrator_table_and_trapz():
    tm = Scatterer(wavelength=tmatrix_aux.wl_C, m=refractive.m_w_10C[tmatrix_aux.wl_C])
    tm.orient = orientation.orient_averaged_fixed
    tm.or_pdf = orientation.gaussian_pdf(10.0)
    tm.psd_integrator = psd.PSDIntegrator()
    tm.psd_integrator.axis_ratio_func = lambda D: 1.0 / tmatrix_aux.dsr_thurai_2007(D)
    tm.psd_integrator.D_max = 3.0
    tm.psd_integrator.num_points = 6
    tm.psd_integrator.geometries = (tmatrix_aux.geom_horiz_back, tmatrix_aux.geom_horiz_forw)
    with pytest.raises(AttributeError):
        tm.psd = psd.GammaPSD(D0=1.0, Nw=1e3, mu=2)
        tm.get_SZ()
    tm.psd_integrator.init_scatter_table(tm, angular_integration=False)
    st = tm.psd_integrator._S_table[tmatrix_aux.geom_horiz_back]
    assert st.shape == (2, 2, 6) and tm.psd_integrator._Z_table[tmatrix_aux.geom_horiz_forw].shape == (4, 4, 6)
    tm.psd = psd.GammaPSD(D0=1.0, Nw=1e3, mu=2, D_max=3.0)
    tm.set_geometry(tmatrix_aux.geom_horiz_back)
    S, Z = tm.get_SZ()
    D = tm.psd_integrator._psd_D
    trapz = getattr(np, "trapezoid", None) or np.trapz
    Zt = trapz(tm.psd_integrator._Z_table[tmatrix_aux.geom_horiz_back] * tm.psd(D), D)
    assert np.abs(Z - Zt).max() / np.abs(Zt).max() < 1e-13
    zh = 10 * np.log10(radar.refl(tm))
    assert 5 < zh < 60
    tm.set_geometry(tmatrix_aux.geom_horiz_forw)
    assert radar.Kdp(tm) > 0 and radar.Ai(tm) > 0


def test_scatter_table_save_load_round_trip(tmp_path):
    """PSDIntegrator.save_scatter_table / load_scatter_table: upstream's pickle layout
    (num_points, D_max, _psd_D, _S_table, _Z_table, _angular_table, _m_table, geometries) under "psd_scatter", and a loaded
    table gives the same PSD-integrated S / Z and angular-integrated quantities without any T-matrix computation."""
    import pickle
    tm = Scatterer(wavelength=tmatrix_aux.wl_X, m=refractive.m_w_10C[tmatrix_aux.wl_X])
    tm.orient = orientation.orient_averaged_fixed
    tm.or_pdf = orientation.gaussian_pdf(7.0)
    geoms = (tmatrix_aux.geom_horiz_back, tmatrix_aux.geom_horiz_forw)
    pi = psd.PSDIntegrator(num_points=5, D_max=2.5, geometries=geoms,
                           axis_ratio_func=lambda D: 1.0 / tmatrix_aux.dsr_thurai_2007(D))
    tm.psd_integrator = pi
    pi.init_scatter_table(tm, angular_integration=True)
    fn = str(tmp_path / "table.pkl")
    pi.save_scatter_table(fn, description="x band test")
    raw = pickle.load(open(fn, "rb"))
    assert set(raw) == {"description", "time", "psd_scatter", "version"} and len(raw["psd_scatter"]) == 8
    assert raw["psd_scatter"][0] == 5 and raw["psd_scatter"][1] == 2.5 and raw["psd_scatter"][7] == geoms
    tm.psd = psd.GammaPSD(D0=1.2, Nw=4e3, mu=1, D_max=2.5)
    ref = {}
    for g in geoms:
        tm.set_geometry(g)
        ref[g] = (tm.get_SZ(), scatter.sca_xsect(tm), scatter.ext_xsect(tm), scatter.asym(tm))

    class NoCompute(OracleBackend):
        def calctmat(self, *a, **k):
            raise AssertionError("a loaded table must not trigger a T-matrix computation")
        scatter_table = calctmat

    tmod.set_backend(NoCompute())
    tm2 = Scatterer(wavelength=tmatrix_aux.wl_X, m=refractive.m_w_10C[tmatrix_aux.wl_X])
    tm2.psd_integrator = psd.PSDIntegrator()
    (when, desc) = tm2.psd_integrator.load_scatter_table(fn)
    assert desc == "x band test" and tm2.psd_integrator.num_points == 5 and tm2.psd_integrator.geometries == geoms
    tm2.psd = psd.GammaPSD(D0=1.2, Nw=4e3, mu=1, D_max=2.5)
    for g in geoms:
        tm2.set_geometry(g)
        (S, Z) = tm2.get_SZ()
        assert np.array_equal(S, ref[g][0][0]) and np.array_equal(Z, ref[g][0][1])
        # (the CPU oracle back-end has no sca_xsect / asym: NaN tables; the round trip must preserve them bit for bit)
        got = (scatter.sca_xsect(tm2), scatter.ext_xsect(tm2), scatter.asym(tm2))
        assert np.array_equal(np.array(got), np.array(ref[g][1:]), equal_nan=True) and got[1] > 0


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_scattering_module_on_shim_regression_vectors():
    """The UNMODIFIED reference scattering.py on top of the shims: analytic-canting observables of single drops
    (scattering_rain.py:185-193 convention) against the restatement vectors of SURVEY.md section 4."""
    sys.path.insert(0, REF)
    try:
        import scattering as ref_sc
    finally:
        sys.path.remove(REF)
    wl = tmatrix_aux.wl_C
    tm = Scatterer(wavelength=wl, m=refractive.m_w_10C[wl])
    tm.orient = orientation.orient_single
    diam = np.array([1.0, 2.0, 4.0, 7.0])
    fv180 = np.zeros(4, complex)
    fh180 = np.zeros(4, complex)
    fv0 = np.zeros(4, complex)
    fh0 = np.zeros(4, complex)
    for i, d in enumerate(diam):
        tm.radius = d / 2.0
        tm.axis_ratio = 1.0 / tmatrix_aux.dsr_thurai_2007(d)
        tm.set_geometry(tmatrix_aux.geom_horiz_back)
        S = tm.get_S()
        fv180[i], fh180[i] = 1e-3 * S[0][0], -1e-3 * S[1][1]
        tm.set_geometry(tmatrix_aux.geom_horiz_forw)
        S = tm.get_S()
        fv0[i], fh0[i] = 1e-3 * S[0][0], 1e-3 * S[1][1]
    mom = ref_sc.compute_angular_moments_analytical(10.0)
    df = ref_sc.compute_scattering_canting_sp(wl, fv180, fh180, fv0, fh0, mom)
    assert np.allclose(df["refl_h"], [-0.014328, 18.031289, 35.510821, 56.712188], atol=2e-5)
    assert np.allclose(df["zdr"], [0.128963, 0.679145, 2.264351, 4.531446], atol=2e-5)
    assert np.allclose(df["rho_hv"], [0.999999, 0.999971, 0.999662, 0.997882], atol=2e-6)
    assert np.allclose(df["delta_hv"], [0.012956, 0.075485, -0.433305, 19.452988], atol=2e-5)
    assert np.allclose(df["kdp"], [7.6e-5, 0.003265, 0.099528, 0.623675], atol=1e-6)
    assert np.allclose(df["A_h"], [2.9e-5, 4.26e-4, 0.019682, 0.366805], atol=1e-6)


def test_tm2l_format_helpers(tmp_path):
    from f_2l_tmatrix import tm2l
    assert tm2l.fortran_e15_7(0.12345678) == "  0.1234568E+00"
    assert tm2l.fortran_e15_7(-1.5e-5) == " -0.1500000E-04"
    assert tm2l.fortran_e15_7(0.0) == "  0.0000000E+00"
    assert tm2l.fortran_e15_7(9.99999999e3) == "  0.1000000E+05"
    (tmp_path / "spongyice.inp").write_text("&ingest\npalamd=5.35\n/\n")
    assert tm2l.read_wavelength_cm(str(tmp_path / "spongyice.inp")) == 5.35
    (tmp_path / "fort.20").write_text("2\n1. 0.5 1.2 1.1 3.1 0.01 70. 20.\n2. 1.0 1.3 1.2 3.2 0.02 71. 21.\n")
    p = tm2l.read_particles(str(tmp_path / "fort.20"))
    assert p.shape == (2, 8) and p[1, 7] == 21.0

"""CPU suite: the two-layer oracle (restatement of tmatrix_py.f) against physics anchors and the
tm2l shim's file protocol with the oracle injected as back-end."""
import os

import numpy as np
import pytest
from scipy.special import spherical_jn, spherical_yn

import oracle


def test_core_equals_shell_matches_single_layer():
    """Identical core / shell permittivity == homogeneous spheroid; differs only by the Fortran's float32 pi."""
    m = 8.601 + 1.687j
    for (D_mm, ar) in [(1.0, 0.9707), (4.0, 0.79), (7.0, 0.5577)]:
        out = oracle.tm2l([D_mm / 10], [D_mm / 20], [ar], [ar], [m * m], [m * m], 5.35)[0]
        tm = oracle.TMatrix(D_mm / 2, 53.5, m, 1 / ar, ddelt=1e-8)  # converged well below the float32-pi effect
        Sb, _ = tm.ampl(90, 90, 0, 180)
        Sf, _ = tm.ampl(90, 90, 0, 0)
        ref = np.array([-1e-3 * Sb[0, 0], -1e-3 * Sb[1, 1], 1e-3 * Sf[0, 0], 1e-3 * Sf[1, 1]])
        got = out[0::2] + 1j * out[1::2]
        assert np.abs(got - ref).max() / np.abs(ref).max() < 3e-7


def _coated_mie(x, y, m1, m2, nmax):
    """Aden-Kerker coefficients (Bohren & Huffman 8.1): core (size x, index m1) in a shell (size y, index m2)."""
    n = np.arange(1, nmax + 1)

    def psi(z):
        return z * spherical_jn(n, z)

    def dpsi(z):
        return spherical_jn(n, z) + z * spherical_jn(n, z, derivative=True)

    def chi(z):
        return -z * spherical_yn(n, z)

    def dchi(z):
        return -spherical_yn(n, z) - z * spherical_yn(n, z, derivative=True)
    A = (m2 * psi(m2 * x) * dpsi(m1 * x) - m1 * dpsi(m2 * x) * psi(m1 * x)) / \
        (m2 * chi(m2 * x) * dpsi(m1 * x) - m1 * dchi(m2 * x) * psi(m1 * x))
    B = (m2 * psi(m1 * x) * dpsi(m2 * x) - m1 * psi(m2 * x) * dpsi(m1 * x)) / \
        (m2 * dchi(m2 * x) * psi(m1 * x) - m1 * dpsi(m1 * x) * chi(m2 * x))
    xi = psi(y) - 1j * chi(y)
    dxi = dpsi(y) - 1j * dchi(y)
    ta = dpsi(m2 * y) - A * dchi(m2 * y)
    tb = psi(m2 * y) - A * chi(m2 * y)
    a = (psi(y) * ta - m2 * dpsi(y) * tb) / (xi * ta - m2 * dxi * tb)
    ta = dpsi(m2 * y) - B * dchi(m2 * y)
    tb = psi(m2 * y) - B * chi(m2 * y)
    b = (m2 * psi(y) * ta - dpsi(y) * tb) / (m2 * xi * ta - dxi * tb)
    return n, a, b


def test_coated_sphere_matches_aden_kerker():
    """16 mm ice core (m = 1.78+0.002j) inside a 20 mm water shell at C band (SURVEY.md Appendix C probe)."""
    wl_cm = 5.35
    m_core, m_shell = 1.78 + 0.002j, 8.601 + 1.687j
    out = oracle.tm2l([2.0], [1.6], [1.0], [1.0], [m_shell**2], [m_core**2], wl_cm)[0]
    k = 2 * np.pi / (wl_cm * 1e-2)  # 1/m
    n, a, b = _coated_mie(np.pi * 1.6 / wl_cm, np.pi * 2.0 / wl_cm, m_core, m_shell, 25)
    s0 = 0.5 * np.sum((2 * n + 1) * (a + b))
    s180 = 0.5 * np.sum((2 * n + 1) * (-1.0) ** n * (b - a))
    fwd_v, fwd_h = out[4] + 1j * out[5], out[6] + 1j * out[7]
    bk_v, bk_h = out[0] + 1j * out[1], out[2] + 1j * out[3]
    assert abs(fwd_v - 1j * s0 / k) / abs(s0 / k) < 1e-6
    assert abs(fwd_h - 1j * s0 / k) / abs(s0 / k) < 1e-6
    assert abs(bk_h + bk_v) / abs(bk_h) < 1e-6           # sphere: borpe = -borp
    assert abs(abs(bk_h) - abs(s180 / k)) / abs(s180 / k) < 1e-6


def test_tm2l_shim_file_protocol(tmp_path, monkeypatch):
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    from f_2l_tmatrix import tm2l

    class OB:
        def tm2l(self, de, di, ae, ai, ee, ei, wl):
            return oracle.tm2l(de, di, ae, ai, ee, ei, wl)
    tm2l.set_backend(OB())
    try:
        monkeypatch.chdir(tmp_path)
        m = 8.601 + 1.687j
        e = m * m
        (tmp_path / "spongyice.inp").write_text("&ingest\npalamd=5.35\n/\n")
        lines = ["2"]
        for D in (0.2, 0.4):
            lines.append(f"{D} {D/2} 0.9 0.9 {e.real} {e.imag} {e.real} {e.imag}")
        (tmp_path / "fort.20").write_text("\n".join(lines) + "\n")
        tm2l.tmatrix(2)
        for f in ("spongyice.out", "pck_tmatrix2b.out", "fort.8"):
            assert (tmp_path / f).exists()
        rows = (tmp_path / "spongyice.out").read_text().splitlines()
        assert len(rows) == 2 and all(len(r) == 120 for r in rows)
        vals = np.array([[float(r[15 * i:15 * i + 15]) for i in range(8)] for r in rows])
        ref = oracle.tm2l([0.2, 0.4], [0.1, 0.2], 0.9, 0.9, e, e, 5.35)
        assert np.abs(vals - ref).max() / np.abs(ref).max() < 1e-6   # e15.7 = 7 significant digits
        assert (tmp_path / "fort.8").read_text().strip() == rows[-1].strip()   # duplicate last particle
        with pytest.raises(RuntimeError):
            tm2l.tmatrix(3)
    finally:
        tm2l.set_backend(None)


def test_reference_io_roundtrip_if_present(tmp_path, monkeypatch):
    """reference scattering_io writes the tm2l input, the shim solves it, reference read_scatt_double_layer parses it."""
    ref = "/root/reference/hydroscatt"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    import sys
    import hydroscatt_b200
    hydroscatt_b200.install_shims()
    sys.path.insert(0, ref)
    try:
        import scattering_io as sio
        import scattering as sc
    finally:
        sys.path.remove(ref)
    from f_2l_tmatrix import tm2l

    class OB:
        def tm2l(self, de, di, ae, ai, ee, ei, wl):
            return oracle.tm2l(de, di, ae, ai, ee, ei, wl)
    tm2l.set_backend(OB())
    try:
        work = tmp_path / "tm2l"
        work.mkdir()
        m = 8.601 + 1.687j
        d = np.array([2.0, 4.0])          # mm
        ar = np.array([0.93, 0.79])
        fname_model = sio.write_melting_hydro_scatt_model(
            d, d / 2, ar, ar, np.full(2, m), np.full(2, m), str(tmp_path / "model_scatt.txt"))
        fname_freq = sio.write_wavelength_file(53.5, str(tmp_path / "freq.inp"))
        out = sc.compute_tmatrix_2l(fname_freq, fname_model, str(work), str(tmp_path / "tmat.out"))
        df = sio.read_scatt_double_layer(out)
        assert len(df["fv180"]) == 2
        for i in range(2):
            tm = oracle.TMatrix(d[i] / 2, 53.5, m, 1 / ar[i])
            Sb, _ = tm.ampl(90, 90, 0, 180)
            Sf, _ = tm.ampl(90, 90, 0, 0)
            # scattering_rain.py:188-193 convention: fv180 = 1e-3 S_vv, fh180 = -1e-3 S_hh, fv0/fh0 = 1e-3 S
            assert abs(df["fv180"][i] - 1e-3 * Sb[0, 0]) / abs(1e-3 * Sb[0, 0]) < 2e-6
            assert abs(df["fh180"][i] + 1e-3 * Sb[1, 1]) / abs(1e-3 * Sb[1, 1]) < 2e-6
            assert abs(df["fv0"][i] - 1e-3 * Sf[0, 0]) / abs(1e-3 * Sf[0, 0]) < 2e-6
            assert abs(df["fh0"][i] - 1e-3 * Sf[1, 1]) / abs(1e-3 * Sf[1, 1]) < 2e-6
        assert not any((work / f).exists() for f in ("fort.8", "fort.20", "spongyice.out", "spongyice.inp"))
    finally:
        tm2l.set_backend(None)

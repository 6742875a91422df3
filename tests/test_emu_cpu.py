"""CPU suite: the CUDA device functions compiled as single-lane host code (tests/emu) against the oracle.  Checks
the kernel logic (convergence controller, parity-decoupled solve, packed layout, amplitude sums, closed-form cross
sections) without a GPU; the multi-lane execution itself is covered by the -m gpu tests."""
import numpy as np
import pytest

import oracle
from tests import emu
from tests.layout import packed_size, packed_to_full
from tests.workloads import rain_grid


@pytest.mark.parametrize("band", ["C", "Ka"])
def test_emu_calctmat_matches_oracle(band):
    w = rain_grid(band)
    sel = slice(None, None, 6)
    r = emu.calctmat(w["radius"][sel], w["wavelength"][sel], w["mr"][sel], w["mi"][sel], w["axis_ratio"][sel])
    for j, i in enumerate(range(70)[sel]):
        tm = oracle.TMatrix(w["radius"][i], w["wavelength"][i], complex(w["mr"][i], w["mi"][i]), w["axis_ratio"][i])
        assert (tm.nmax, tm.ngauss, tm.ntrials, tm.status) == (r["nmax"][j], r["ngauss"][j], r["ntrials"][j], r["status"][j])
        tp = r["arena"][r["toff"][j]:r["toff"][j] + packed_size(tm.nmax)]
        full = packed_to_full(tp, tm.nmax)
        assert np.abs(full - tm.tdata()).max() / np.abs(tm.tdata()).max() < 1e-9
        S, Z = emu.ampl(tp, tm.nmax, w["wavelength"][i], 90, 90, 0, 180, 72.0, 13.4)
        So, Zo = tm.ampl(90, 90, 0, 180, 72.0, 13.4)
        assert np.abs(S - So).max() / np.abs(So).max() < 1e-9
        assert np.abs(Z - Zo).max() / np.abs(Zo).max() < 1e-9


def test_emu_escalation_and_limits():
    # workspace too small -> internal escalate status (100); Gauss table too small -> 101
    r = emu.calctmat(2.0, 6.5, 1.5, 0.5, 1 / 0.6, ws_cap=500)
    assert r["status"][0] == 100
    r = emu.calctmat(2.0, 6.5, 1.5, 0.5, 1 / 0.6, gmax=13)
    assert r["status"][0] == 101
    r = emu.calctmat(-1.0, 6.5, 1.5, 0.5, 1.2)
    assert r["status"][0] == 5
    r = emu.calctmat(2.0, 6.5, 1.5, 0.5, 1 / 0.6, arena_cap=10)
    assert r["status"][0] == 4


def test_emu_xsect_equals_sphere_integral():
    r = emu.calctmat(2.0, 6.5, 1.5, 0.5, 1 / 0.6)
    nmax = int(r["nmax"][0])
    tp = r["arena"][:packed_size(nmax)]
    tm = oracle.TMatrix(2.0, 6.5, 1.5 + 0.5j, 1 / 0.6)
    x, w = np.polynomial.legendre.leggauss(18)
    nph = 36
    for (t0, p0, a, b) in [(90, 0, 0, 0), (60, 10, 72, 61.7)]:
        sh = sv = 0.0
        for xi, wi in zip(x, w):
            for p in (np.arange(nph) + 0.5) * 360.0 / nph:
                _, Z = tm.ampl(t0, np.degrees(np.arccos(xi)), p0, p, a, b)
                sh += wi * (Z[0, 0] - Z[0, 1])
                sv += wi * (Z[0, 0] + Z[0, 1])
        sh *= 2 * np.pi / nph
        sv *= 2 * np.pi / nph
        out = emu.xsect(tp, nmax, 6.5, t0, p0, a, b)
        assert abs(out[0] - sh) / sh < 1e-6 and abs(out[1] - sv) / sv < 1e-6


def test_gauss_tables_any_order():
    """The host-side generator of the device Gauss tables: n-point rules of any order (the two-segment quadrature of
    cylinders needs odd ones) against numpy, and the positive half used for spheroids."""
    for n in (1, 2, 3, 7, 8, 15, 31, 64, 125, 250):
        z, w = emu.gauss_full(n)
        zr, wr = np.polynomial.legendre.leggauss(n)
        assert np.abs(z - zr).max() < 5e-15 and np.abs(w - wr).max() < 5e-15, n
        assert (np.diff(z) > 0).all() and abs(w.sum() - 2.0) < 1e-13
        if n % 2 == 0:
            xp, wp = emu.gauss_positive(n)
            assert (xp == z[n // 2:]).all() and (wp == w[n // 2:]).all()

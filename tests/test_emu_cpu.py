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


def _start_orders(radius, wavelength, m, axis_ratio):
    """starting NMAX (upstream's formula) and the lower-bound guess of the converged order, as k_dv_hist computes them."""
    xev = 2.0 * np.pi * radius / wavelength
    n0 = max(4, int(xev + 4.05 * xev ** 0.333333))
    a = axis_ratio ** (1.0 / 3.0)
    mx = abs(m) * xev * max(a, a / axis_ratio)
    return n0, max(n0 + 1, int(np.floor(0.6035 * mx + 4.49)) - 2)


@pytest.mark.parametrize("band", ["C", "W"])
def test_device_round_controller_follows_reference_sequence(band):
    """The controller / planner of the device-driven rounds (hs_dvctl.h, compiled as host code) fed with the oracle's recorded
    trial sequence: whatever the speculation depth, it must consume exactly the reference's trials in the reference's order
    -- same (NMAX, NGAUSS, #trials, status) -- and never read a trial speculated beyond convergence (those carry poison)."""
    w = rain_grid(band)
    for i in range(0, 70, 9):
        m = complex(w["mr"][i], w["mi"][i])
        tm = oracle.TMatrix(w["radius"][i], w["wavelength"][i], m, w["axis_ratio"][i])
        assert tm.status == 0
        table = tm.trials()
        assert len(table[0]) == tm.ntrials
        s0, lb = _start_orders(w["radius"][i], w["wavelength"][i], m, w["axis_ratio"][i])
        assert s0 == table[0][0]
        seen = set()
        for spec, spec0, margin in [(0, 0, 1), (0, 0, 0), (0, 2, 2), (0, -2, 1), (1, 0, 1), (3, 0, 1), (7, 0, 1)]:
            r = emu.dv_rounds(s0, lb, table, spec=spec, spec0=spec0, margin=margin, ddelt=0.1 * 1e-3)
            assert (r["nmax"], r["ngauss"], r["ntrials"], r["status"], r["done"]) == (tm.nmax, tm.ngauss, tm.ntrials, 0, 1), (i, spec, r)
            assert r["unknown"] == 0 and r["planned"] >= tm.ntrials
            seen.add(r["rounds"])
        assert len(seen) > 1  # the speculation depth changes the number of rounds, not the answer


def test_device_round_controller_limits():
    """NMAX limit (upstream: print + STOP -> status 1), NGAUSS limit while searching NGAUSS (status 2, T-matrix kept: done = 1),
    NGAUSS beyond NPNG1 at the start of the NMAX search (status 2, no T-matrix: done = 2), bad input (status 5)."""
    n = np.arange(4, 101)
    # Qsca alternates by 50 %: never converges -> NMAX runs into NPN1 = 100
    r = emu.dv_rounds(4, 9, (n, 2 * n, np.where(n % 2, 1.0, 2.0), np.ones(n.size)))
    assert (r["status"], r["done"], r["nmax"]) == (1, 2, 100)
    # NMAX converges at 6 (two equal trials), then every NGAUSS trial differs until NGAUSS = 500
    tn = np.concatenate([[4, 5, 6], np.full(488, 6)])
    tg = np.concatenate([[8, 10, 12], np.arange(13, 501)])
    qs = np.concatenate([[1.0, 2.0, 2.0], 3.0 + (np.arange(488) % 2)])
    r = emu.dv_rounds(4, 5, (tn, tg, qs, np.ones(tn.size)), spec=3)
    assert (r["status"], r["done"], r["nmax"], r["ngauss"], r["ntrials"]) == (2, 1, 6, 500, 491)
    # ndgs = 15: NMAX 34 would need 510 > NPNG1 nodes
    n = np.arange(30, 40)
    r = emu.dv_rounds(30, 31, (n, 15 * n, np.where(n % 2, 1.0, 2.0), np.ones(n.size)), ndgs=15)
    assert (r["status"], r["done"], r["nmax"]) == (2, 2, 34)
    r = emu.dv_rounds(-5, 6, (np.array([4]), np.array([8]), np.ones(1), np.ones(1)))
    assert (r["status"], r["done"], r["nmax"], r["ntrials"]) == (5, 2, 4, 0)

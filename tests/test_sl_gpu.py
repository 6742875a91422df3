"""GPU parity of the single-layer path (calctmat + calcampl) against the CPU oracle, through the C-ABI."""
import numpy as np
import pytest

import oracle
from tests.layout import packed_to_full
from tests.workloads import rain_grid, WL

pytestmark = pytest.mark.gpu

GEOMS = [(90.0, 90.0, 0.0, 180.0), (90.0, 90.0, 0.0, 0.0)]


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

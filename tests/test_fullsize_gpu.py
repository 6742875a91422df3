"""GPU: size-independent properties at BASELINE full size (W2: 17 220 spheroids, 6 bands x 0..40 degC x 70 D) -- the
oracle is too slow to run all of it in a test, so the full grid is checked through identities, and a seeded sample
of it against the oracle."""
import numpy as np
import pytest

import bench
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def w2(engine):
    w = bench.workload_w2()
    tb = engine.calctmat(w["radius"], w["wavelength"], w["mr"], w["mi"], w["eps"])
    return w, tb


def test_w2_all_converged_and_sample_matches_oracle(engine, w2):
    w, tb = w2
    st = tb.status.cpu().numpy()
    nmax = tb.nmax.cpu().numpy()
    ngauss = tb.ngauss.cpu().numpy()
    ntr = tb.ntrials.cpu().numpy()
    assert (st == 0).all()
    assert nmax.min() >= 5 and nmax.max() <= 32
    # NGAUSS loop: every particle converges on its first NGAUSS increment
    assert (ngauss >= 2 * nmax + 1).all() and (ngauss == 2 * nmax + 1).mean() > 0.98
    # within one (band, T) table the largest drop needs at least the order of the smallest
    t = nmax.reshape(w["n_tab"], w["n_D"])
    assert (t[:, -1] >= t[:, 0]).all()
    idx = np.random.default_rng(11).choice(st.size, 160, replace=False)
    r = oracle.sl_batch(w["radius"][idx], w["wavelength"][idx], w["mr"][idx], w["mi"][idx], w["eps"][idx])
    assert np.array_equal(r["nmax"], nmax[idx]) and np.array_equal(r["ngauss"], ngauss[idx])
    assert np.array_equal(r["ntrials"], ntr[idx])


def test_w2_reciprocity_and_z_identities(engine, w2):
    """backscatter reciprocity S12 = -S21 and the S -> Z identities hold for every particle and orientation."""
    w, tb = w2
    S, Z = tb.ampl([(90.0, 90.0, 0.0, 180.0)], alpha=[0.0, 72.0, 200.0], beta=[0.0, 13.4, 41.5])
    S = S.cpu().numpy()[:, 0]
    Z = Z.cpu().numpy()[:, 0]
    scale = np.abs(S).max(axis=(2, 3))
    assert (np.abs(S[..., 0, 1] + S[..., 1, 0]) / scale).max() < 1e-6
    z11 = 0.5 * (np.abs(S) ** 2).sum(axis=(2, 3))
    assert (np.abs(Z[..., 0, 0] - z11) / z11).max() < 1e-13
    z33 = (S[..., 0, 0] * np.conj(S[..., 1, 1]) + S[..., 0, 1] * np.conj(S[..., 1, 0])).real
    assert (np.abs(Z[..., 2, 2] - z33) / z11).max() < 1e-13


def test_w2_lossless_optical_theorem(engine):
    """mi = 0: extinction (optical theorem on the forward amplitude) equals the closed-form scattering cross section."""
    w = bench.workload_w2()
    sel = np.arange(0, w["radius"].size, 3)
    tb = engine.calctmat(w["radius"][sel], w["wavelength"][sel], w["mr"][sel], 0.0 * w["mi"][sel], w["eps"][sel])
    assert (tb.status.cpu().numpy() == 0).all()
    al, be = [0.0, 144.0], [0.0, 26.1]
    for o in range(2):
        S, Z, xs = tb.scatter([(90.0, 90.0, 0.0, 0.0)], [al[o]], [be[o]])
        S = S.cpu().numpy()[:, 0]
        xs = xs.cpu().numpy()[:, 0]
        lam = w["wavelength"][sel]
        ext_h, ext_v = 2 * lam * S[:, 1, 1].imag, 2 * lam * S[:, 0, 0].imag
        # The truncation is only converged to ddelt = 1e-4 in the orientation-averaged Qsca/Qext, and for lossless
        # high-index drops the reference's convergence test can stop early next to a morphology-dependent resonance
        # (a handful of particles are off by per cent -- the oracle shows the same); so: quantiles, not the maximum.
        for e in (np.abs(ext_h - xs[:, 0]) / xs[:, 0], np.abs(ext_v - xs[:, 1]) / xs[:, 1]):
            assert np.median(e) < 1e-6 and np.quantile(e, 0.99) < 5e-3


def test_psd_integration_linearity_full_size(engine, w2):
    from hydroscatt_b200 import tables
    w, tb = w2
    rows = torch_rows = None
    al, be, wt, scale = tables.fixed_orientation_nodes(10.0)
    S, Z, xs = tb.scatter([bench.GEOM_BACK, bench.GEOM_FORW], al, be, weight=wt, scale=scale)
    import torch
    n = tb.n
    rows = torch.zeros((n, tables.ROW), dtype=torch.float64, device=S.device)
    rows[:, 0:8] = torch.view_as_real(S).reshape(n, 2, 8)[:, 0]
    rows[:, 8:24] = Z[:, 0].reshape(n, 16)
    rows[:, 24:32] = torch.view_as_real(S).reshape(n, 2, 8)[:, 1]
    rows[:, 32:48] = Z[:, 1].reshape(n, 16)
    d0, nw, mu = bench.dsd_grid()
    k = slice(0, 9300, 37)
    dm = np.full(d0[k].size, 7.0)
    _, I1 = tables.integrate_psd(engine, rows, w["n_tab"], w["D"], "gamma", d0[k], nw[k], mu[k], dm, w["lam_tab"])
    _, I2 = tables.integrate_psd(engine, rows, w["n_tab"], w["D"], "gamma", d0[k], 3.0 * nw[k], mu[k], dm, w["lam_tab"])
    a, b = I1.cpu().numpy()[..., :48], I2.cpu().numpy()[..., :48]
    assert np.abs(b - 3.0 * a).max() <= 1e-12 * np.abs(b).max()

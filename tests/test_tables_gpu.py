"""GPU: the table-record packing kernel (hs_pack_rows) and caller-selected observable columns (hs_psd_observables_sel)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _table(engine, band="X"):
    from hydroscatt_b200 import tables
    from tests.workloads import rain_grid
    al, be, wt, scale = tables.fixed_orientation_nodes(10.0)
    w = rain_grid(band)
    return w, (al, be, wt, scale)


def test_pack_rows_equals_layout_definition(engine):
    """rows[n, 56] from hs_pack_rows == the layout in tables.py's header built column by column on the host, for the radar
    geometries (forward amplitude = the forward geometry itself) and for a back geometry whose forward direction is a third
    geometry (vertical incidence)."""
    from hydroscatt_b200 import tables
    w, (al, be, wt, scale) = _table(engine)
    for gb, gf in (((90.0, 90.0, 0.0, 180.0), (90.0, 90.0, 0.0, 0.0)), ((0.0, 180.0, 0.0, 0.0), (180.0, 180.0, 0.0, 0.0))):
        fw = [(g[0], g[0], g[2], g[2]) for g in (gb, gf)]
        allg = [gb, gf] + [f for f in dict.fromkeys(fw) if f not in (gb, gf)]
        tb, S, Z, xs = engine.calctmat_scatter(w["radius"], w["wavelength"], w["mr"], w["mi"], w["axis_ratio"], allg, al, be,
                                               weight=wt, scale=scale, want_xs=True)
        rows, tb2 = tables.build_rows(engine, w["radius"], w["wavelength"], w["mr"], w["mi"], w["axis_ratio"], gb, gf, al, be, wt, scale)
        R = rows.cpu().numpy()
        Sh = torch.view_as_real(S).reshape(tb.n, len(allg), 8).cpu().numpy()
        Zh, xh, lam = Z.reshape(tb.n, len(allg), 16).cpu().numpy(), xs.cpu().numpy(), tb.lam.cpu().numpy()
        ref = np.empty((tb.n, 56))
        for gi in range(2):
            ref[:, 24 * gi:24 * gi + 8] = Sh[:, gi]
            ref[:, 24 * gi + 8:24 * gi + 24] = Zh[:, gi]
            ref[:, 48 + 2 * gi:50 + 2 * gi] = xh[:, gi]
            fi = allg.index(fw[gi])
            ref[:, 52 + 2 * gi] = 2.0 * lam * Sh[:, fi, 7]
            ref[:, 53 + 2 * gi] = 2.0 * lam * Sh[:, fi, 1]
        assert np.array_equal(R, ref)
    with pytest.raises(Exception):
        engine.pack_rows(S, Z, xs, tb.lam, 0, 7)


def test_selected_observable_columns(engine):
    """hs_psd_observables_sel: any subset / order of the 15 columns equals the same columns of the full frame, bit for bit."""
    from hydroscatt_b200 import tables
    w, (al, be, wt, scale) = _table(engine)
    ws = [w, _table(engine, "C")[0], _table(engine, "W")[0], _table(engine, "Ka")[0], _table(engine, "S")[0]]
    cat = lambda k: np.concatenate([x[k] for x in ws])
    rows, _ = tables.build_rows(engine, cat("radius"), cat("wavelength"), cat("mr"), cat("mi"), cat("axis_ratio"), (90.0, 90.0, 0.0, 180.0),
                                (90.0, 90.0, 0.0, 0.0), al, be, wt, scale)
    rng = np.random.default_rng(5)
    nd = 200
    d0, nw, mu = rng.uniform(0.3, 3.0, nd), 10 ** rng.uniform(1, 4, nd), rng.integers(-1, 8, nd).astype(float)
    lam_tab = [x["wavelength"][0] for x in ws]
    W = engine.psd_weights("gamma", d0, nw, mu, np.full(nd, 7.0), w["D"], np.ones_like(w["D"]))
    full = engine.psd_observables(W, rows, 5, lam_tab, want_sca=True).cpu().numpy()
    for cols in ([2, 4, 11, 7, 12, 14], [14], list(range(15)), [0, 1, 9], [13, 3, 8, 8, 5]):
        got = engine.psd_observables(W, rows, 5, lam_tab, columns=cols).cpu().numpy()
        assert got.shape == (nd, 5, len(cols))
        assert np.array_equal(got, full[:, :, cols], equal_nan=True), cols
    # by name through tables.integrate_psd; the cross sections are integrated only when selected
    names = ["refl_h", "zdr", "kdp", "rho_hv", "A_h", "Adp"]
    o, _ = tables.integrate_psd(engine, rows, 5, w["D"], "gamma", d0, nw, mu, np.full(nd, 7.0), lam_tab, want_integ=False, columns=names)
    o15, _ = tables.integrate_psd(engine, rows, 5, w["D"], "gamma", d0, nw, mu, np.full(nd, 7.0), lam_tab, want_integ=False)
    idx = [tables.OBS_COLUMNS.index(c) for c in names]
    assert np.array_equal(o.cpu().numpy(), o15.cpu().numpy()[:, :, idx])
    with pytest.raises(Exception):
        engine.psd_observables(W, rows, 5, lam_tab, columns=[15])


def test_sharded_build_equals_single_build_world1(engine):
    """dist.build_rows_sharded without a process group (world 1): same records, same order, same (NMAX, NGAUSS, status)."""
    from hydroscatt_b200 import tables, dist as hdist
    w, (al, be, wt, scale) = _table(engine, "Ka")
    args = (w["radius"], w["wavelength"], w["mr"], w["mi"], w["axis_ratio"], (90.0, 90.0, 0.0, 180.0), (90.0, 90.0, 0.0, 0.0),
            al, be, wt, scale)
    rows, meta = hdist.build_rows_sharded(engine, *args)
    ref, tb = tables.build_rows(engine, *args)
    assert torch.equal(rows, ref)
    assert torch.equal(meta[:, 0], tb.nmax) and torch.equal(meta[:, 1], tb.ngauss) and torch.equal(meta[:, 3], tb.status)
    # the reorder kernel against index_select, with a simulated 3-rank deal
    cost = hdist.predicted_cost(w["radius"], w["wavelength"], w["mr"], w["mi"], w["axis_ratio"])
    plans = [hdist.ShardPlan(cost, 3, r) for r in range(3)]
    nm = plans[0].n_max
    gathered = torch.zeros((3 * nm, 56), dtype=torch.float64, device=ref.device)
    for r, p in enumerate(plans):
        gathered[r * nm:r * nm + p.idx.size] = ref[torch.from_numpy(p.idx).to(ref.device)]
    assert torch.equal(engine.gather_rows(gathered, plans[0].inv_on(ref.device)), ref)


def test_sharded_build_world2_nccl():
    """Two ranks over NCCL when the box has two GPUs (the driver's GPU test box has one: skipped there; run with
    gpurun --gpus 2)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(root, "tests", "run_sharded.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert p.stdout.count("SHARDED_OK") == 2


def test_c_abi_error_path_reports_message(engine):
    """A rejected call returns a negative status and hs_last_error carries the reason (no exception crosses the ABI; the
    message is set under its own mutex -- a regression here deadlocked every error path)."""
    from hydroscatt_b200.engine import EngineError
    with pytest.raises(EngineError, match="ndgs"):
        engine.calctmat([1.0], 33.3, 7.9, 2.3, [1.2], ndgs=0)
    with pytest.raises(EngineError, match="shapes"):
        engine.calctmat([1.0], 33.3, 7.9, 2.3, [1.2], shape=3)
    tb = engine.calctmat([1.0], 33.3, 7.9, 2.3, [1.2])       # the context stays usable
    assert int(tb.status[0]) == 0

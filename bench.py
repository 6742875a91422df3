#!/usr/bin/env python
"""bench.py -- headline benchmark: T-matrix solves/s on the rain-sweep scatter-table build (BASELINE.json configs[1]).

One "step" = one complete scatter-table build of workload W2: 6 radar bands x T = 0..40 degC x 70
drop diameters = 17 220 spheroids -> converged T-matrix of every particle (all NMAX/NGAUSS trials + every azimuthal
mode), orientation-averaged S/Z for the back- and forward-scatter geometries (5 x 10 fixed quadrature, sigma = 10
deg), h/v scattering + extinction cross sections, PSD integration over the 9 300 gamma DSDs of
scattering_rain.py:236-267 for each of the 246 (band, T) tables and the radar observables of every DSD.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1: launched under torchrun, one rank per GPU.  The global table is N W2 sweeps (weak scaling: one sweep's worth of
particles per GPU; sweep 0 is W2, the others have jittered diameters); the GLOBAL particle list is dealt cost-sorted
round-robin over the ranks (hydroscatt_b200.dist.ShardPlan), every rank solves and scatters its shard with no data-path
collective, ONE all-gather over NCCL returns the complete table to every rank, and rank r integrates the PSDs of its
own 246 tables and copies only those results to the host.  `--scaling strong --replicas R` keeps the global table at R
sweeps for every N.  Sub-measurements in the same line: the two-layer path at W5's full size (108 150 coated spheroids,
split over the ranks) and, at N = 1, W1 x 4096.
`--impl reference` times the CPU restatement of the reference algorithm (oracle/, "port") on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BANDS = ("S", "C", "X", "Ku", "Ka", "W")
WL = dict(S=111.0, C=53.5, X=33.3, Ku=22.0, Ka=8.43, W=3.19)
GEOM_BACK = (90.0, 90.0, 0.0, 180.0)
GEOM_FORW = (90.0, 90.0, 0.0, 0.0)
CANTING = 10.0
DDELT, NDGS = 1e-3, 2
W2_NAME = ("W2 rain-sweep scatter-table build: 6 bands x 0..40 degC x 70 D = 17220 spheroids per GPU, "
           "2 geometries x 50 orientations, 9300 gamma DSDs x 246 tables")


# ------------------------------------------------------------------------------------------------ workload
def dsr_thurai_2007(D):
    lo = 1.173 - 0.5165 * D + 0.4698 * D**2 - 0.1317 * D**3 - 8.5e-3 * D**4
    hi = 1.065 - 6.25e-2 * D - 3.99e-3 * D**2 + 7.66e-4 * D**3 - 4.095e-5 * D**4
    return np.where(D < 0.7, 1.0, np.where(D < 1.5, lo, hi))


def workload_w2(rank=0):
    """Flat particle list, table-major: table t = (band, T), 70 diameters each.  rank > 0 jitters D by <= 1 %."""
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "water_refractive_index.json")))
    D = np.arange(0.1, 7.0 + 0.1, 0.1)
    if rank:
        D = D * (1.0 + np.random.default_rng(1234 + rank).uniform(-0.01, 0.01, D.size))
    eps = 1.0 / dsr_thurai_2007(D)
    lam_tab, m_tab = [], []
    for b in BANDS:
        for (mr, mi) in fx["m"][b]:
            lam_tab.append(WL[b])
            m_tab.append(complex(mr, mi))
    lam_tab = np.array(lam_tab)
    m_tab = np.array(m_tab)
    n_tab = lam_tab.size
    return dict(D=D, n_tab=n_tab, n_D=D.size, lam_tab=lam_tab,
                radius=np.tile(D / 2.0, n_tab), wavelength=np.repeat(lam_tab, D.size),
                mr=np.repeat(m_tab.real, D.size), mi=np.repeat(m_tab.imag, D.size), eps=np.tile(eps, n_tab))


def dsd_grid():
    """scattering_rain.py:236-267: 31 Nw x 10 mu x 30 D0 gamma DSDs, D_max = 7 mm."""
    nw = np.power(10, np.arange(1, 4 + 0.1, 0.1))
    mu = np.arange(-1, 8 + 1, 1)
    d0 = np.arange(0.1, 3 + 0.1, 0.1)
    n = d0.size * mu.size * nw.size
    d0v, nwv, muv = np.zeros(n), np.zeros(n), np.zeros(n)
    for i0, a in enumerate(d0):
        for i1, b in enumerate(nw):
            for i2, c in enumerate(mu):
                ind = i2 * d0.size * nw.size + i1 * d0.size + i0
                d0v[ind], nwv[ind], muv[ind] = a, b, c
    return d0v, nwv, muv


def orientation_nodes():
    from hydroscatt_b200.tables import fixed_orientation_nodes
    return fixed_orientation_nodes(CANTING)


# ------------------------------------------------------------------------------------------------ flop model
def algorithmic_flops(w, nmax, ngauss, ndgs=NDGS, n_orient=50, n_geom=2):
    """SURVEY.md 8(d) contract figure, per particle: (calctmat flops, amplitude flops)."""
    nmax = nmax.astype(np.float64)
    ngauss = ngauss.astype(np.float64)
    xev = 2.0 * np.pi * w["radius"] / w["wavelength"]
    n0 = np.maximum(4, np.floor(xev + 4.05 * xev**0.333333))
    mabs = np.hypot(w["mr"], w["mi"])
    a = w["radius"] * w["eps"] ** (1.0 / 3.0)
    kr = 2.0 * np.pi / w["wavelength"] * np.maximum(a, a / w["eps"])
    tot = np.zeros_like(nmax)

    def special(n, g):
        ta = np.maximum(kr, n)
        tb = np.maximum(kr * mabs, n)
        nn1 = np.floor(1.2 * np.sqrt(ta) + 3.0)
        nn2 = np.floor(tb + 4.0 * tb ** 0.33333 + 1.2 * np.sqrt(tb)) - n + 5
        return 60.0 * g * (2 * n + nn1 + nn2)

    def trial(n, g):
        return 40.0 * n * n * g + special(n, g) + 12.0 * g * n + (64.0 / 3.0) * n**3

    nmax_max = int(nmax.max())
    for n in range(4, nmax_max + 1):
        act = (n0 <= n) & (n <= nmax)
        tot += np.where(act, trial(float(n), float(n * ndgs)), 0.0)
    gmax = int((ngauss - nmax * ndgs).max())
    for k in range(1, gmax + 1):
        act = (nmax * ndgs + k) <= ngauss
        tot += np.where(act, trial(nmax, nmax * ndgs + k), 0.0)
    for m in range(1, nmax_max + 1):
        nm = nmax - m + 1
        act = nm >= 1
        tot += np.where(act, 113.0 * nm * nm * ngauss + 12.0 * ngauss * nmax + (64.0 / 3.0) * nm**3, 0.0)
    ampl = (96.0 * nmax * (nmax + 1) * (2 * nmax + 1) / 6.0 + 30.0 * nmax**2) * n_orient * n_geom
    return tot, ampl


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        self.enabled = enabled  # only rank 0 polls nvidia-smi: eight pollers on one host take cores away from the pipeline's host threads
        self.index, self.samples, self.reasons, self.stop = index, [], set(), False
        self.smax = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.smax = float(f[1])
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.enabled:
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.enabled:
            self.th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.smax,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_chunk(args):
    import oracle
    (rev, lam, mr, mi, eps, geoms, al, be, wt) = args
    r = oracle.sl_batch(rev, lam, mr, mi, eps, geoms=geoms, alpha=al, beta=be, wgt=wt, ddelt=DDELT, ndgs=NDGS)
    return r["nmax"].sum()


def cpu_reference_pass(w, stride, cores):
    """Oracle (restated reference algorithm) on every `stride`-th particle of W2: calctmat + orientation-averaged
    S/Z for the three amplitude geometries, one process per core.  Returns (particles, seconds)."""
    import multiprocessing as mp
    al, be, wt, _ = orientation_nodes()
    idx = np.arange(0, w["radius"].size, stride)
    # cost-sorted round-robin so every worker gets the same mix
    cost = (w["radius"] / w["wavelength"])[idx]
    idx = idx[np.argsort(-cost)]
    geoms = [GEOM_BACK, GEOM_FORW]
    nchunk = cores * 8
    jobs = []
    for c in range(nchunk):
        s = idx[c::nchunk]
        if s.size:
            jobs.append((w["radius"][s], w["wavelength"][s], w["mr"][s], w["mi"][s], w["eps"][s], geoms, al, be, wt))
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_chunk, jobs, chunksize=1)
    return idx.size, time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    w = workload_w2()
    cores = os.cpu_count() or 1
    stride = args.cpu_stride
    for _ in range(args.warmup):
        cpu_reference_pass(w, stride * 8, cores)
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        n, t = cpu_reference_pass(w, stride, cores)
        tot_n += n
        tot_t += t
    val = tot_n / tot_t
    sample = f"every {stride}th particle of W2 ({tot_n // max(args.steps, 1)} of 17220 per step): calctmat + 2x50 calcampl"
    line = {"impl": "reference", "metric": "T-matrix solves/s", "value": val, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": W2_NAME, "sample": sample, "parallelism": f"multiprocessing x{cores}",
                       "ddelt_rescale": True},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.cpu_tm2l_stride > 0:
        n2c, t2c = cpu_tm2l_pass(args.cpu_tm2l_stride, cores)
        line["tm2l"] = {"metric": "two-layer T-matrix particles/s", "value": n2c / t2c, "unit": "particles/s", "cores": cores, "kind": "port",
                        "sample": f"every {args.cpu_tm2l_stride}th particle of W5 ({n2c} of 108150): oracle/tm2l_oracle.c "
                                  f"(restated tmatrix_py.f), {t2c:.1f} s"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
OBS_NORTH_STAR = ("refl_h", "zdr", "kdp", "rho_hv", "A_h", "Adp")   # BASELINE.json north_star: Zh, Zdr, Kdp, rhohv, Ah, Adp


def workload_global(replicas):
    """`replicas` W2 sweeps (replica 0 = W2 itself, the others with diameters jittered by <= 1 %) as ONE flat particle list,
    table-major: replica r owns tables [246 r, 246 (r + 1))."""
    ws = [workload_w2(r) for r in range(replicas)]
    g = {k: np.concatenate([w[k] for w in ws]) for k in ("radius", "wavelength", "mr", "mi", "eps")}
    g.update(replicas=ws, n_tab=ws[0]["n_tab"], n_D=ws[0]["n_D"], lam_tab=ws[0]["lam_tab"])
    return g


def workload_w5():
    """W5: two-layer melting-hail profile, 350 D x 309 levels at C band = 108 150 coated spheroids."""
    nD, nL = 350, 309
    D = np.tile(np.linspace(0.02, 3.5, nD), nL)                                  # cm
    fmw = np.repeat(np.linspace(0.02, 0.98, nL), nD)
    rng = np.random.default_rng(7)
    return [D, D * (1.0 - 0.5 * fmw) ** (1.0 / 3.0), 1.0 - 0.25 * rng.uniform(0, 1, D.size),
            np.minimum(1.0, 0.8 + 0.25 * rng.uniform(0, 1, D.size)), np.full(D.size, (8.601 + 1.687j) ** 2),
            np.full(D.size, (1.78 + 0.002j) ** 2)]


def _cpu_tm2l_chunk(cols):
    import oracle
    return oracle.tm2l(*cols, 5.35).shape[0]


def cpu_tm2l_pass(stride, cores):
    """Two-layer oracle (restated tmatrix_py.f) on every `stride`-th particle of W5, one process per core."""
    import multiprocessing as mp
    w5 = workload_w5()
    idx = np.arange(0, w5[0].size, stride)
    nchunk = cores * 4
    jobs = [[c[idx[k::nchunk]] for c in w5] for k in range(nchunk) if idx[k::nchunk].size]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_tm2l_chunk, jobs, chunksize=1)
    return idx.size, time.perf_counter() - t0


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and os.environ.get("HS_BENCH_BIND", "1") != "0":
        from hydroscatt_b200.dist import bind_to_gpu_cpus
        bound = bind_to_gpu_cpus(local)  # before any pinned allocation
        if bound is not None:
            print("[bench] rank %d bound to CPUs %d-%d (%d)" % (rank, bound[0], bound[-1], len(bound)), file=sys.stderr)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from hydroscatt_b200.engine import Engine
    from hydroscatt_b200 import tables
    from hydroscatt_b200 import dist as hdist

    eng = Engine(local)
    dev = eng.device
    # weak scaling (default): the global table is `world` W2 sweeps, so every GPU owns one sweep's worth of particles;
    # strong: a fixed global table of --replicas sweeps at every N.  Either way the GLOBAL particle list is dealt
    # cost-sorted round-robin over the ranks (dist.ShardPlan), every rank solves and scatters its shard, ONE all-gather
    # returns the complete table to every rank, and rank r integrates the PSDs of its own tables.
    strong = args.scaling == "strong"
    R = args.replicas if strong else world
    assert R % world == 0, "--replicas must be a multiple of the number of GPUs"
    g = workload_global(R)
    n_glob = g["radius"].size
    plan = hdist.ShardPlan(hdist.predicted_cost(g["radius"], g["wavelength"], g["mr"], g["mi"], g["eps"]), world, rank)
    idx = plan.idx
    n = idx.size
    rep_per = R // world
    my_reps = list(range(rank * rep_per, (rank + 1) * rep_per))
    n_tab, n_D = g["n_tab"], g["n_D"]
    d0, nw, mu = dsd_grid()
    al, be, wt, scale = orientation_nodes()
    n_dsd = d0.size
    cols = None if os.environ.get("HS_BENCH_OBS", "north_star") == "all" else [tables.OBS_COLUMNS.index(c) for c in OBS_NORTH_STAR]
    n_obs = 15 if cols is None else len(cols)

    # pinned host inputs (e2e path) and device-resident copies (kernel path): this rank's shard of the global list
    host_in = {k: torch.from_numpy(np.ascontiguousarray(g[k][idx])).pin_memory() for k in ("radius", "wavelength", "mr", "mi", "eps")}
    host_dsd = {k: torch.from_numpy(v).pin_memory() for k, v in (("d0", d0), ("nw", nw), ("mu", mu))}
    IN_KEYS, DSD_KEYS = ("radius", "wavelength", "mr", "mi", "eps"), ("d0", "nw", "mu")
    host_in_all = torch.stack([host_in[k] for k in IN_KEYS]).pin_memory()
    host_dsd_all = torch.stack([host_dsd[k] for k in DSD_KEYS]).pin_memory()
    dev_in = {k: v.to(dev) for k, v in host_in.items()}
    dev_dsd = {k: v.to(dev) for k, v in host_dsd.items()}
    dmax = torch.full((n_dsd,), 7.0, dtype=torch.float64, device=dev)
    h2d = sum(v.numel() * 8 for v in host_in.values()) + sum(v.numel() * 8 for v in host_dsd.values())
    n_rows_mine = rep_per * n_tab * n_D

    def host_bufs():
        return dict(rows=torch.empty((n_rows_mine, tables.ROW), dtype=torch.float64).pin_memory(),
                    obs=torch.empty((rep_per, n_dsd, n_tab, n_obs), dtype=torch.float64).pin_memory(),
                    meta=torch.empty((4, n), dtype=torch.int32).pin_memory())

    host_out = [host_bufs(), host_bufs()]
    d2h = sum(t.numel() * t.element_size() for t in host_out[0].values())
    full_buf = torch.empty((n_glob, tables.ROW), dtype=torch.float64, device=dev) if world > 1 else None

    ev = lambda: torch.cuda.Event(enable_timing=True)
    stage_ms = {"calctmat": 0.0, "scatter": 0.0, "psd": 0.0}
    last = {}
    split = {"on": False}

    def step(inp, dsd, timed):
        e = [ev() for _ in range(6)]
        e[0].record()
        if args.unfused_scatter or split["on"]:
            tb = eng.calctmat(inp["radius"], inp["wavelength"], inp["mr"], inp["mi"], inp["eps"], ddelt=DDELT, ndgs=NDGS)
            e[1].record()
            S, Z, xs = tb.scatter([GEOM_BACK, GEOM_FORW], al, be, weight=wt, scale=scale, want_xs=True)
        else:
            # one library call: T-matrices + orientation-averaged S/Z + cross sections, the scatter kernels of finished
            # particles overlapping the convergence rounds of the others
            tb, S, Z, xs = eng.calctmat_scatter(inp["radius"], inp["wavelength"], inp["mr"], inp["mi"], inp["eps"],
                                                [GEOM_BACK, GEOM_FORW], al, be, weight=wt, scale=scale, want_xs=True,
                                                ddelt=DDELT, ndgs=NDGS)
            e[1].record()
        e[2].record()
        rows = eng.pack_rows(S, Z, xs, tb.lam, 0, 1)              # hs_pack_rows: the 56-double table records of the shard
        e[3].record()
        if world > 1:
            full = plan.gather(rows, eng=eng, out=full_buf)       # all_gather_into_tensor (NCCL) + hs_gather_rows
            mine = full[rank * n_rows_mine:(rank + 1) * n_rows_mine]
        else:
            mine = rows                                           # one rank: the shard is the table, already in table order
        e[4].record()
        obs = []
        for q, r in enumerate(my_reps):
            w = g["replicas"][r]
            o, _ = tables.integrate_psd(eng, mine[q * n_tab * n_D:(q + 1) * n_tab * n_D], n_tab, w["D"], "gamma", dsd["d0"], dsd["nw"],
                                        dsd["mu"], dmax, w["lam_tab"], rule="trapz", want_integ=False, columns=cols)
            obs.append(o)
        e[5].record()
        if timed:
            last["e"] = e
        last.update(tb=tb, rows=mine, obs=obs)
        return tb, mine, obs

    # End-to-end step: pinned host inputs -> device -> tables + observables -> pinned host results.  The D2H copies of
    # step k run on a copy stream and overlap the compute of step k+1 (two sets of host result buffers); every step
    # still moves all of its own inputs and results inside the timed region.
    copy_stream = torch.cuda.Stream(dev)
    copied = [None, None]

    def step_e2e(k=0):
        hb = host_out[k & 1]
        if copied[k & 1] is not None:
            copied[k & 1].synchronize()  # the consumer of step k-2's results is done with this buffer
        # one H2D copy per input block (the five particle columns, the three DSD parameter columns), not one per column
        inp_all = host_in_all.to(dev, non_blocking=True)
        dsd_all = host_dsd_all.to(dev, non_blocking=True)
        inp = {k_: inp_all[i_] for i_, k_ in enumerate(IN_KEYS)}
        dsd = {k_: dsd_all[i_] for i_, k_ in enumerate(DSD_KEYS)}
        tb, rows, obs = step(inp, dsd, False)
        meta_dev = torch.stack([tb.nmax, tb.ngauss, tb.ntrials, tb.status])  # one D2H copy for the four per-particle records
        ready = torch.cuda.Event()
        ready.record()
        copy_stream.wait_event(ready)
        with torch.cuda.stream(copy_stream):
            for t in [rows, meta_dev] + obs:
                t.record_stream(copy_stream)
            hb["rows"].copy_(rows, non_blocking=True)
            # the observables in pieces, so that the pipeline's small host<->device transfers of the next step are not
            # queued behind one multi-millisecond DMA
            nch = int(os.environ.get("HS_BENCH_D2H_CHUNKS", "64"))
            for q, o in enumerate(obs):
                for c in range(nch):
                    a_, b_ = c * n_dsd // nch, (c + 1) * n_dsd // nch
                    hb["obs"][q, a_:b_].copy_(o[a_:b_], non_blocking=True)
            hb["meta"].copy_(meta_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        copied[k & 1] = done

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_tf = eng.fp64_peak()
    peak_tensor_tf = eng.fp64_peak(tensor=True)
    for _ in range(max(args.warmup, 3)):
        step(dev_in, dev_dsd, False)
    l0 = eng.launch_count
    sync_all()
    with ClockSampler(local, enabled=(rank == 0)) as clk:
        t0, t1 = ev(), ev()
        t0.record()
        for _ in range(args.steps):
            step(dev_in, dev_dsd, True)
        t1.record()
        sync_all()
        ms_total = t0.elapsed_time(t1)
    launches = eng.launch_count - l0
    coll_ms = last["e"][3].elapsed_time(last["e"][4]) if world > 1 else 0.0
    # stage split (and the roofline of the dominant stage): a short pass with the T-matrix and scatter calls issued
    # separately, so that CUDA events on the launching stream bracket each stage
    split["on"] = True
    nsplit = max(3, min(args.steps, 10))
    step(dev_in, dev_dsd, True)
    sync_all()
    acc = {"calctmat": 0.0, "scatter": 0.0, "psd": 0.0}
    for _ in range(nsplit):
        step(dev_in, dev_dsd, True)
        sync_all()
        e = last["e"]
        for k, (a, b) in zip(("calctmat", "scatter", "psd"), ((0, 1), (1, 2), (4, 5))):
            acc[k] += e[a].elapsed_time(e[b]) / nsplit
    stage_ms.update(acc)
    split["on"] = False
    # e2e: host buffers in, host results out, same K
    for k in range(2):
        step_e2e(k)
    sync_all()
    tw0 = time.perf_counter()
    t0, t1 = ev(), ev()
    t0.record()
    for k in range(args.steps):
        step_e2e(k)
    t1.record()
    sync_all()
    ms_e2e = max(t0.elapsed_time(t1), 0.0)
    wall_e2e = (time.perf_counter() - tw0) * 1e3
    ms_e2e = max(ms_e2e, wall_e2e)  # host-side work counts for the end-to-end number

    # second half of the headline metric: the two-layer (tm2l) path at W5's full size (350 D x 309 levels =
    # 108 150 coated spheroids), the particle list split evenly over the ranks (independent particles, no collective)
    w5 = workload_w5()
    n5 = w5[0].size
    sl5 = slice(rank * n5 // world, (rank + 1) * n5 // world)
    t2_dev = [torch.from_numpy(np.ascontiguousarray(a[sl5])).to(dev) for a in w5]
    eng.tm2l(*t2_dev, 5.35)
    sync_all()
    q0, q1 = ev(), ev()
    q0.record()
    for _ in range(2):
        amp5, st5 = eng.tm2l(*t2_dev, 5.35)
    q1.record()
    sync_all()
    ms_tm2l = q0.elapsed_time(q1) / 2.0
    bad5 = int((st5 != 0).sum().item())

    # W1 x 4096 (the C-band 10 degC rain grid replicated 4096 times with jitter = 286 720 spheroids), one GPU only
    ms_w1, other = None, {}
    if world == 1 and not args.no_sublines:
        from tests.workloads import rain_grid
        Rw = 4096
        wg = rain_grid("C")
        rng = np.random.default_rng(1234)
        jd = 1.0 + rng.uniform(-0.02, 0.02, (Rw, 70))
        jm = 1.0 + rng.normal(0.0, 1e-3, (Rw, 70))
        w1 = [(wg["radius"] * jd).ravel(), np.tile(wg["wavelength"], Rw), (wg["mr"] * jm).ravel(), (wg["mi"] * jm).ravel(),
              np.tile(wg["axis_ratio"], Rw)]
        w1d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in w1]
        run_w1 = lambda: eng.calctmat_scatter(*w1d, [GEOM_BACK, GEOM_FORW], al, be, weight=wt, scale=scale, ddelt=DDELT, ndgs=NDGS)
        run_w1()
        torch.cuda.synchronize()
        q0, q1 = ev(), ev()
        q0.record()
        for _ in range(3):
            tb1 = run_w1()[0]
        q1.record()
        torch.cuda.synchronize()
        ms_w1 = q0.elapsed_time(q1) / 3.0
        bad1 = int((tb1.status != 0).sum().item())
        del w1d, tb1
        # W3 (ice crystals: plates / columns at X / Ka / W, ndgs 2 and 15) and W4 (single-layer hail, D 2..50 mm: NMAX up to ~55,
        # multi-block assembly + separate solve kernel): BASELINE configs[2], [3]; same call, same outputs
        from tests.workloads import ice_grid, hail_grid
        other = {}
        for name, wk, ndgs_ in (("w3_ice_ndgs2", ice_grid(seed=11, n=7680), 2), ("w3_ice_ndgs15", ice_grid(seed=12, n=1024), 15),
                                ("w4_hail_single_layer", hail_grid(seed=13, n=1600), 2)):
            wd = [torch.from_numpy(np.ascontiguousarray(wk[k])).to(dev) for k in ("radius", "wavelength", "mr", "mi", "axis_ratio")]
            run_o = lambda: eng.calctmat_scatter(*wd, [GEOM_BACK, GEOM_FORW], al, be, weight=wt, scale=scale, ddelt=DDELT, ndgs=ndgs_)
            run_o()
            torch.cuda.synchronize()
            q0, q1 = ev(), ev()
            q0.record()
            for _ in range(5):
                tbo = run_o()[0]
            q1.record()
            torch.cuda.synchronize()
            ms_o = q0.elapsed_time(q1) / 5.0
            nm_o, st_o = tbo.nmax.cpu().numpy(), tbo.status.cpu().numpy()
            other[name] = {"particles": int(st_o.size), "ms": ms_o, "value": st_o.size / (ms_o * 1e-3), "unit": "solves/s", "ndgs": ndgs_,
                           "nmax_range": [int(nm_o[st_o == 0].min()), int(nm_o[st_o == 0].max())],
                           "status_counts": {str(k): int((st_o == k).sum()) for k in np.unique(st_o)}}

    # Throughput with two table builds in flight (N = 1 only, a sub-line: the headline times one build at a time): two
    # contexts, two streams, two host threads; one build's latency-bound convergence rounds overlap the other's scatter / PSD kernels
    piped = None
    if world == 1 and rep_per == 1 and not args.no_sublines and (os.cpu_count() or 1) >= 12:  # (one sweep per GPU: the shard is one table set)
        import threading as _th
        engs2 = [eng, Engine(local)]
        streams2 = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        w0 = g["replicas"][0]

        def build_on(i, nsteps):
            with torch.cuda.stream(streams2[i]):
                for _ in range(nsteps):
                    tb_, S_, Z_, xs_ = engs2[i].calctmat_scatter(dev_in["radius"], dev_in["wavelength"], dev_in["mr"], dev_in["mi"], dev_in["eps"],
                                                                 [GEOM_BACK, GEOM_FORW], al, be, weight=wt, scale=scale, want_xs=True,
                                                                 ddelt=DDELT, ndgs=NDGS)
                    rows_ = engs2[i].pack_rows(S_, Z_, xs_, tb_.lam, 0, 1)
                    tables.integrate_psd(engs2[i], rows_, n_tab, w0["D"], "gamma", dev_dsd["d0"], dev_dsd["nw"], dev_dsd["mu"], dmax,
                                         w0["lam_tab"], rule="trapz", want_integ=False, columns=cols)

        torch.cuda.synchronize()
        for i in range(2):
            build_on(i, 2)
        torch.cuda.synchronize()
        kp = max(2, (args.steps // 2) * 2)
        tp0 = time.perf_counter()
        ths = [_th.Thread(target=build_on, args=(i, kp // 2)) for i in range(2)]
        for t_ in ths:
            t_.start()
        for t_ in ths:
            t_.join()
        torch.cuda.synchronize()
        dtp = time.perf_counter() - tp0
        piped = {"builds_in_flight": 2, "steps": kp, "ms_per_step": dtp * 1e3 / kp, "value": n * kp / dtp, "unit": "solves/s",
                 "timing": "host wall clock around both threads, device synchronised on both sides"}
        del engs2

    tms = torch.tensor([ms_total, ms_e2e, ms_tm2l, coll_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_tm2l, coll_ms = tms.tolist()

    tb = last["tb"]
    nmax = tb.nmax.cpu().numpy()
    ngauss = tb.ngauss.cpu().numpy()
    status = tb.status.cpu().numpy()
    wl = {"radius": g["radius"][idx], "wavelength": g["wavelength"][idx], "mr": g["mr"][idx], "mi": g["mi"][idx], "eps": g["eps"][idx]}
    fl_tm, fl_am = algorithmic_flops(wl, nmax, ngauss)
    if rank == 0:
        ms_step = ms_total / args.steps
        units = n_glob
        value = units / (ms_step * 1e-3)
        dom = max(stage_ms, key=stage_ms.get)
        dom_fl = {"calctmat": fl_tm.sum(), "scatter": fl_am.sum()}.get(dom, fl_tm.sum())
        dom_kernel = {"calctmat": "k_st_assemble (+k_st_radial, k_st_wigner, k_st_mode_small)", "scatter": "k_scatter", "psd": "k_psd_obs_fused"}[dom]
        ach = dom_fl / (stage_ms[dom] * 1e-3) / 1e12
        cores = os.cpu_count() or 1
        cpu = cpu2 = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            oracle.build()
            ncpu, tcpu = cpu_reference_pass(workload_w2(), args.cpu_stride, cores)
            cpu = {"value": ncpu / tcpu, "unit": "solves/s", "cores": cores, "kind": "port",
                   "sample": f"every {args.cpu_stride}th particle of W2 ({ncpu} of 17220): calctmat + 2x50 calcampl, "
                             f"no cross sections / PSD integration, {tcpu:.1f} s"}
            n2c, t2c = cpu_tm2l_pass(args.cpu_tm2l_stride, cores)
            cpu2 = {"value": n2c / t2c, "unit": "particles/s", "cores": cores, "kind": "port",
                    "sample": f"every {args.cpu_tm2l_stride}th particle of W5 ({n2c} of {n5}): oracle/tm2l_oracle.c (restated tmatrix_py.f), {t2c:.1f} s"}
        traffic, traffic_note = None, "no ncu capture committed"
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, traffic_note = tj.get("calctmat_stage_dram_bytes_per_step"), tj.get("note")
        line = {
            "metric": "T-matrix solves/s", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": W2_NAME if R == 1 else W2_NAME.replace("per GPU", "per sweep") + f"; global table = {R} sweeps "
                       f"(sweep 0 = W2, the others with diameters jittered <= 1 %) = {n_glob} spheroids, {R * n_tab} tables",
                       "l2": "working set per step (0.35 GB T arena, 0.3-0.7 GB stage scratch, 0.1-0.27 GB observables) exceeds the 126 MB L2",
                       "psd_output": f"{n_obs} observables per (DSD, table)" + (" (" + ", ".join(OBS_NORTH_STAR) + ": the set BASELINE north_star names; "
                                     "HS_BENCH_OBS=all writes the 15-column frame)" if cols is not None else "") +
                                     "; the integrated S/Z array is not materialised",
                       "parallelism": (f"global particle list dealt cost-sorted round-robin over {world} ranks (dist.ShardPlan), one "
                                       f"all_gather_into_tensor of the table records per step, PSD integration sharded by table")
                       if world > 1 else "single GPU",
                       "ddelt_rescale": True},
            "clocks": clk.summary(),
            "e2e": {"value": units / (ms_e2e / args.steps * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e / args.steps,
                    "note": "bytes summed over the ranks; every rank copies its shard's inputs in and its own tables + observables out"},
            "gpu_launches": int(launches),
            "stage_ms": stage_ms,
            "stage_ms_note": "stages timed in a separate pass with hs_calctmat_batch and hs_scatter_batch issued one after the other; "
                             "the headline step issues them as one hs_calctmat_scatter_batch call (scatter overlaps the convergence rounds)",
            # dominant stage of the step: the staged T-matrix pipeline (k_st_radial, k_st_wigner, k_st_assemble with the
            # DMMA contraction + fused solve, k_st_mode_small), timed live with CUDA events on the launching stream; its
            # kernels overlap on several streams, so the stage (not one launch) is the unit
            "roofline": {"bound": "tensor", "kernel": dom_kernel, "achieved": ach, "peak": peak_tensor_tf if dom == "calctmat" else peak_tf,
                         "unit": "TFLOP/s", "frac": ach / (peak_tensor_tf if dom == "calctmat" else peak_tf), "traffic": traffic,
                         "traffic_note": traffic_note,
                         "pipe": "FP64 tensor core (DMMA m8n8k4) + FP64 FMA" if dom == "calctmat" else "FP64 FMA",
                         "peak_source": "measured in this run (hs_fp64_tensor_peak: DMMA m8n8k4 loop %.1f TFLOP/s; hs_fp64_peak: DFMA loop "
                                        "%.1f TFLOP/s); MEASURED_PEAKS.json has no FP64 figure" % (peak_tensor_tf, peak_tf),
                         "algorithmic_gflop_per_step": {"calctmat": fl_tm.sum() / 1e9, "scatter": fl_am.sum() / 1e9},
                         "achieved_by_stage_tflops": {"calctmat": fl_tm.sum() / (stage_ms["calctmat"] * 1e-3) / 1e12,
                                                      "scatter": fl_am.sum() / (stage_ms["scatter"] * 1e-3) / 1e12},
                         "scatter_note": "the scatter figure counts SURVEY 8(d)'s per-call flops (one calcampl per orientation and "
                                         "geometry, plus the cross sections); the fused kernel shares one T-matrix sweep between "
                                         "them and executes ~12x fewer real flops, so it is a work-equivalent rate, not a pipe utilisation"},
            "tm2l": {"metric": "two-layer T-matrix particles/s", "workload": "W5 melting-hail profile 350 D x 309 levels, C band"
                     + (f", split over {world} ranks" if world > 1 else ""), "value": n5 / (ms_tm2l * 1e-3),
                     "particles": n5, "ms": ms_tm2l, "status_nonzero_rank0": bad5, "algorithmic_flop_per_particle": 4.8e7,
                     "achieved_tflops": 4.8e7 * n5 / (ms_tm2l * 1e-3) / 1e12,
                     "frac_of_fp64_peak": 4.8e7 * n5 / world / (ms_tm2l * 1e-3) / 1e12 / peak_tf},
            "status_counts": {str(k): int((status == k).sum()) for k in np.unique(status)},
            "nmax_range": [int(nmax.min()), int(nmax.max())],
        }
        if world > 1:
            line["collective"] = {"op": "all_gather_into_tensor (NCCL) + hs_gather_rows reorder", "bytes_per_rank": n * tables.ROW * 8,
                                  "ms": coll_ms}
        if ms_w1 is not None:
            line["w1x4096"] = {"workload": "W1 x 4096: C-band 10 degC rain grid, 70 D x 4096 jittered replicas", "particles": 70 * 4096,
                               "ms": ms_w1, "value": 70 * 4096 / (ms_w1 * 1e-3), "unit": "solves/s", "status_nonzero": bad1,
                               "what": "T-matrices + 2 geometries x 50 orientations + cross sections (hs_calctmat_scatter_batch)"}
        if ms_w1 is not None:
            line["other_workloads"] = other
        if piped is not None:
            line["pipelined_builds"] = piped
        if cpu:
            line["cpu_baseline"] = cpu
            line["tm2l"]["cpu_baseline"] = cpu2
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-stride", type=int, default=1, dest="cpu_stride")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sublines", action="store_true", dest="no_sublines", help="skip the W1 x 4096 sub-measurement")
    ap.add_argument("--cpu-tm2l-stride", type=int, default=8, dest="cpu_tm2l_stride")
    ap.add_argument("--scaling", default="weak", choices=("weak", "strong"),
                    help="weak (default): one W2 sweep per GPU; strong: a fixed global table of --replicas sweeps at every N")
    ap.add_argument("--replicas", type=int, default=8)
    ap.add_argument("--unfused-scatter", action="store_true", dest="unfused_scatter",
                    help="separate hs_calctmat_batch / hs_scatter_batch calls (per-stage timing) instead of hs_calctmat_scatter_batch")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

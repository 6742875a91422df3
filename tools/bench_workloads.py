"""Throughput on the other BASELINE workloads: W1xR replicated C-band rain, W3 ice crystals, W4 hail
(single layer), W5 two-layer melting-hail profile.  Timing only (parity: tests/test_workloads_gpu.py, tests/test_tm2l_gpu.py).

    python tools/bench_workloads.py            # prints one JSON line per workload
"""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from hydroscatt_b200.engine import Engine
from hydroscatt_b200 import tables
from tests.workloads import rain_grid, ice_grid, hail_grid

eng = Engine(0)
GB, GF = (90.0, 90.0, 0.0, 180.0), (90.0, 90.0, 0.0, 0.0)
al, be, wt, scale = tables.fixed_orientation_nodes(10.0)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


def sl(name, w, ndgs=2):
    dev = {k: torch.from_numpy(np.ascontiguousarray(w[k])).cuda() for k in ("radius", "wavelength", "mr", "mi", "axis_ratio")}
    run = lambda: eng.calctmat_scatter(dev["radius"], dev["wavelength"], dev["mr"], dev["mi"], dev["axis_ratio"], [GB, GF], al, be,
                                       weight=wt, scale=scale, ndgs=ndgs)
    dt, (tb, S, Z, xs) = timed(run)
    st = tb.status.cpu().numpy()
    nm = tb.nmax.cpu().numpy()
    print(json.dumps({"workload": name, "particles": int(st.size), "ms": dt * 1e3, "solves_per_s": st.size / dt,
                      "nmax_range": [int(nm[st == 0].min()), int(nm[st == 0].max())],
                      "status_counts": {str(k): int((st == k).sum()) for k in np.unique(st)},
                      "what": "T-matrices + 2 geometries x 50 orientations + cross sections (hs_calctmat_scatter_batch)"}), flush=True)


# W1xR: C-band 10 degC rain grid, R jittered replicas (numpy default_rng(1234), D (1 + U(-0.02, 0.02)), m (1 + N(0, 1e-3)))
import os
R = int(os.environ.get("HS_W1_REPLICAS", "1024"))  # SURVEY 8(d) states 4096 (286 720 particles): HS_W1_REPLICAS=4096
w = rain_grid("C")
rng = np.random.default_rng(1234)
jd = 1.0 + rng.uniform(-0.02, 0.02, (R, 70))
jm = 1.0 + rng.normal(0.0, 1e-3, (R, 70))
w1 = dict(radius=(w["radius"] * jd).ravel(), wavelength=np.tile(w["wavelength"], R), mr=(w["mr"] * jm).ravel(),
          mi=(w["mi"] * jm).ravel(), axis_ratio=np.tile(w["axis_ratio"], R))
sl("W1xR rain C-band 10 degC, 70 D x %d replicas" % R, w1)
ice = ice_grid(seed=11, n=7680)
sl("W3 ice crystals X/Ka/W, ndgs 2", ice)
ice15 = ice_grid(seed=12, n=1024)
sl("W3 ice crystals X/Ka/W, ndgs 15", ice15, ndgs=15)
sl("W4 hail single layer C/X/Ka, D 2..50 mm", hail_grid(seed=13, n=1600))

# W5: two-layer melting hail profile: 350 D x 309 levels at C band
nD, nL = 350, 309
D = np.tile(np.linspace(0.02, 3.5, nD), nL)                                  # cm
fmw = np.repeat(np.linspace(0.02, 0.98, nL), nD)
rng = np.random.default_rng(7)
t2 = [D, D * (1.0 - 0.5 * fmw) ** (1.0 / 3.0), 1.0 - 0.25 * rng.uniform(0, 1, D.size), np.minimum(1.0, 0.8 + 0.25 * rng.uniform(0, 1, D.size))]
dev2 = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in t2]
eo = torch.from_numpy(np.full(D.size, (8.601 + 1.687j) ** 2)).cuda()
ei = torch.from_numpy(np.full(D.size, (1.78 + 0.002j) ** 2)).cuda()
dt, (amp, st) = timed(lambda: eng.tm2l(*dev2, eo, ei, 5.35), reps=2)
st = st.cpu().numpy()
print(json.dumps({"workload": "W5 two-layer melting-hail profile, 350 D x 309 levels, C band", "particles": int(D.size), "ms": dt * 1e3,
                  "particles_per_s": D.size / dt, "status_counts": {str(k): int((st == k).sum()) for k in np.unique(st)},
                  "achieved_tflops_algorithmic": 4.8e7 * D.size / dt / 1e12}), flush=True)

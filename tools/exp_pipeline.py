"""Experiment: W2 table builds pipelined on P contexts / streams / host threads (python tools/exp_pipeline.py P steps)."""
import os, sys, time, threading
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hydroscatt_b200.engine import Engine
from hydroscatt_b200 import tables
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
w = bench.workload_w2()
d0, nw, mu = bench.dsd_grid()
al, be, wt, scale = bench.orientation_nodes()
cols = [tables.OBS_COLUMNS.index(c) for c in bench.OBS_NORTH_STAR]
engs = [Engine(0) for _ in range(P)]
streams = [torch.cuda.Stream() for _ in range(P)]
dev = {k: torch.from_numpy(np.ascontiguousarray(w[k])).cuda() for k in ("radius", "wavelength", "mr", "mi", "eps")}
dsd = [torch.from_numpy(a).cuda() for a in (d0, nw, mu)]
dmax = torch.full((d0.size,), 7.0, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()


def step(eng):
    tb, S, Z, xs = eng.calctmat_scatter(dev["radius"], dev["wavelength"], dev["mr"], dev["mi"], dev["eps"], [bench.GEOM_BACK, bench.GEOM_FORW],
                                        al, be, weight=wt, scale=scale, want_xs=True, ddelt=bench.DDELT, ndgs=bench.NDGS)
    rows = eng.pack_rows(S, Z, xs, tb.lam, 0, 1)
    obs, _ = tables.integrate_psd(eng, rows, w["n_tab"], w["D"], "gamma", *dsd, dmax, w["lam_tab"], rule="trapz", want_integ=False, columns=cols)
    return obs


def worker(i, n):
    with torch.cuda.stream(streams[i]):
        for _ in range(n):
            step(engs[i])


for i in range(P):
    with torch.cuda.stream(streams[i]):
        for _ in range(3):
            step(engs[i])
torch.cuda.synchronize()
t0 = time.perf_counter()
th = [threading.Thread(target=worker, args=(i, K // P)) for i in range(P)]
for t in th: t.start()
for t in th: t.join()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
n = (K // P) * P
print("P=%d: %d steps in %.1f ms -> %.3f ms per step, %.2f M solves/s" % (P, n, dt * 1e3, dt * 1e3 / n, 17220 * n / dt / 1e6), flush=True)

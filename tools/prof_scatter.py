"""Scatter stage alone on W2 (T-matrices built once): python tools/prof_scatter.py [reps]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from hydroscatt_b200.engine import Engine
w = bench.workload_w2()
eng = Engine(0)
al, be, wt, scale = bench.orientation_nodes()
args = [torch.from_numpy(w[k]).cuda() for k in ("radius", "wavelength", "mr", "mi", "eps")]
tb = eng.calctmat(*args, ddelt=bench.DDELT, ndgs=bench.NDGS)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    S, Z, xs = tb.scatter([bench.GEOM_BACK, bench.GEOM_FORW], al, be, weight=wt, scale=scale, want_xs=True)
    torch.cuda.synchronize()
    print("scatter W2: %.3f ms" % ((time.perf_counter() - t0) * 1e3))
print("checksum S %.15e Z %.15e xs %.15e" % (S.abs().sum().item(), Z.abs().sum().item(), xs.abs().sum().item()))

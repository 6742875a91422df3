"""Latency of single particles through hs_calctmat_batch (critical path of the heaviest items)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from hydroscatt_b200.engine import Engine
w = bench.workload_w2()
eng = Engine(0)
idx = [17219, 17219 - 35, 17219 - 70 * 20, 70 * 41 * 4 + 69, 70 * 41 + 69, 0]
for i in idx:
    a = [w[k][i:i + 1] for k in ("radius", "wavelength", "mr", "mi", "eps")]
    for _ in range(2):
        tb = eng.calctmat(*a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        tb = eng.calctmat(*a)
    torch.cuda.synchronize()
    print("particle %5d lam %.2f D %.1f: nmax %d ngauss %d trials %d  %.3f ms" % (
        i, w["wavelength"][i], 2 * w["radius"][i], int(tb.nmax[0]), int(tb.ngauss[0]), int(tb.ntrials[0]),
        (time.perf_counter() - t0) / 5 * 1e3))

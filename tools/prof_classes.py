"""Isolated timing of hs_calctmat_batch on NMAX classes of W2 (which bucket costs what; sum vs. combined run)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from hydroscatt_b200.engine import Engine
w = bench.workload_w2()
eng = Engine(0)
keys = ("radius", "wavelength", "mr", "mi", "eps")
tb = eng.calctmat(*[w[k] for k in keys])
nmax = tb.nmax.cpu().numpy()
ng = tb.ngauss.cpu().numpy()
fl, _ = bench.algorithmic_flops(w, nmax, ng)


def timeit(sel, reps=5):
    a = [torch.from_numpy(np.ascontiguousarray(w[k][sel])).cuda() for k in keys]
    for _ in range(2):
        eng.calctmat(*a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.calctmat(*a)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


tot = 0.0
for lo, hi in [(4, 6), (7, 8), (9, 12), (13, 16), (17, 21), (22, 24), (25, 30)]:
    sel = (nmax >= lo) & (nmax <= hi)
    if not sel.any():
        continue
    ms = timeit(sel)
    tot += ms
    print("nmax %2d-%2d: %5d particles %7.3f ms  %6.2f Gflop  %5.2f TF/s  %.1f us/particle-slot" % (
        lo, hi, sel.sum(), ms, fl[sel].sum() / 1e9, fl[sel].sum() / ms / 1e9, ms * 1e3 / max(1, sel.sum())))
print("sum of classes %.3f ms; all together %.3f ms" % (tot, timeit(np.ones_like(sel))))
for i in [17219, 17219 - 20, 17219 - 40, 70 * 41 * 4 + 69, 70 * 41 + 69, 30, 0]:
    s = np.zeros(nmax.size, bool)
    s[i] = True
    print("single particle %5d nmax %2d ngauss %2d: %.3f ms (%.2f Mflop)" % (i, nmax[i], ng[i], timeit(s), fl[i] / 1e6))

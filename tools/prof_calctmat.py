"""One W2 hs_calctmat_batch call with the phase counters on: HYDROTM_PROF=1 [HYDROTM_PROF_SEL=mask] python tools/prof_calctmat.py"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hydroscatt_b200.engine import Engine
eng = Engine(0)
w = bench.workload_w2()
dev = {k: torch.from_numpy(np.ascontiguousarray(w[k])).cuda() for k in ("radius", "wavelength", "mr", "mi", "eps")}
for r in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tb = eng.calctmat(dev["radius"], dev["wavelength"], dev["mr"], dev["mi"], dev["eps"], ddelt=bench.DDELT, ndgs=bench.NDGS)
    torch.cuda.synchronize()
    print("calctmat W2: %.3f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)

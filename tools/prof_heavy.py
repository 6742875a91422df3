"""Phase-cycle profile (HYDROTM_PROF=1) of the calctmat kernel on the heavy W-band part of W2."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from hydroscatt_b200.engine import Engine
w = bench.workload_w2()
sel = w["wavelength"] < 4.0 if len(sys.argv) < 2 else w["wavelength"] > 50.0
eng = Engine(0)
args = [w[k][sel] for k in ("radius", "wavelength", "mr", "mi", "eps")]
for _ in range(3):
    tb = eng.calctmat(*args)
torch.cuda.synchronize()
t0 = time.perf_counter()
tb = eng.calctmat(*args)
torch.cuda.synchronize()
print("particles", sel.sum(), "ms", (time.perf_counter() - t0) * 1e3, "nmax", int(tb.nmax.min()), int(tb.nmax.max()))

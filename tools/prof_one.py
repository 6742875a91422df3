"""One W2 hs_calctmat_batch call (for ncu captures of single launches): python tools/prof_one.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hydroscatt_b200.engine import Engine
eng = Engine(0)
w = bench.workload_w2()
dev = {k: torch.from_numpy(np.ascontiguousarray(w[k])).cuda() for k in ("radius", "wavelength", "mr", "mi", "eps")}
tb = eng.calctmat(dev["radius"], dev["wavelength"], dev["mr"], dev["mi"], dev["eps"], ddelt=bench.DDELT, ndgs=bench.NDGS)
torch.cuda.synchronize()
print("status ok:", int((tb.status == 0).sum()), "of", tb.n)

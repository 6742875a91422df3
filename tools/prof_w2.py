"""One hs_calctmat_batch pass over W2 (for ncu launch lists): python tools/prof_w2.py [reps]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from hydroscatt_b200.engine import Engine
w = bench.workload_w2()
eng = Engine(0)
args = [torch.from_numpy(w[k]).cuda() for k in ("radius", "wavelength", "mr", "mi", "eps")]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tb = eng.calctmat(*args)
    torch.cuda.synchronize()
    print("calctmat W2: %.3f ms, launches so far %d" % ((time.perf_counter() - t0) * 1e3, eng.launch_count))

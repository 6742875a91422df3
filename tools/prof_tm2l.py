"""One hs_tm2l_batch call on a slice of the W5 profile (for ncu): python tools/prof_tm2l.py [n_particles] [reps]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from hydroscatt_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = Engine(0)
rng = np.random.default_rng(7)
D = rng.uniform(0.02, 3.5, n)
fmw = rng.uniform(0.02, 0.98, n)
cols = [D, D * (1.0 - 0.5 * fmw) ** (1.0 / 3.0), 1.0 - 0.25 * rng.uniform(0, 1, n), np.minimum(1.0, 0.8 + 0.25 * rng.uniform(0, 1, n))]
dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in cols]
eo = torch.from_numpy(np.full(n, (8.601 + 1.687j) ** 2)).cuda()
ei = torch.from_numpy(np.full(n, (1.78 + 0.002j) ** 2)).cuda()
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    amp, st = eng.tm2l(*dev, eo, ei, 5.35)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("tm2l %d particles: %.3f ms, %.0f particles/s, status!=0: %d" % (n, dt * 1e3, n / dt, int((st != 0).sum())))

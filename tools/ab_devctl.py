"""A/B of the two round drivers of hs_calctmat_batch on one GPU: host-driven rounds (HYDROTM_DEVCTL=0, hs_staged_host.h)
against device-driven rounds (default, hs_staged_dev.h).  Same stage kernels, so (NMAX, NGAUSS, trials, status) must be
identical and the packed T-matrices equal bit for bit; prints the time of each.

    python tools/ab_devctl.py [w2|ice|hail|w1x] [repeats]
"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import workloads
from hydroscatt_b200.engine import Engine, tmat_size

which = sys.argv[1] if len(sys.argv) > 1 else "w2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if which == "w2":
    w = bench.workload_w2()
    inp = (w["radius"], w["wavelength"], w["mr"], w["mi"], w["eps"]); kw = dict(ddelt=bench.DDELT, ndgs=bench.NDGS)
elif which == "ice":
    g = workloads.ice_grid(n=4096)
    inp = (g["radius"], g["wavelength"], g["mr"], g["mi"], g["axis_ratio"]); kw = dict(ndgs=2)
elif which == "ice15":
    g = workloads.ice_grid(n=1024)
    inp = (g["radius"], g["wavelength"], g["mr"], g["mi"], g["axis_ratio"]); kw = dict(ndgs=15)
elif which == "hail":
    g = workloads.hail_grid(n=1600)
    inp = (g["radius"], g["wavelength"], g["mr"], g["mi"], g["axis_ratio"]); kw = dict(ndgs=2)
else:
    raise SystemExit("unknown workload")
eng = Engine(0)
dev = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda() for a in inp]


def run(devctl):
    os.environ["HYDROTM_DEVCTL"] = str(devctl)
    best = 1e9
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tb = eng.calctmat(*dev, **kw)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return tb, best


tb0, t0 = run(0)
tb1, t1 = run(1)
print("%s: %d particles  host-driven %.3f ms  device-driven %.3f ms" % (which, tb0.n, t0, t1), flush=True)
ok = True
for name in ("nmax", "ngauss", "ntrials", "status"):
    a, b = getattr(tb0, name).cpu().numpy(), getattr(tb1, name).cpu().numpy()
    same = np.array_equal(a, b)
    ok &= same
    print("  %-8s identical: %s" % (name, same) + ("" if same else "  (%d differ, first %s)" % ((a != b).sum(), np.nonzero(a != b)[0][:5])))
st = tb0.status.cpu().numpy()
print("  status histogram:", np.bincount(st, minlength=6).tolist(), " nmax range", int(tb0.nmax.min()), int(tb0.nmax.max()))
# packed T bit for bit (the arena slots differ between the runs: compare particle by particle through the offsets)
nm = tb0.nmax.cpu().numpy()
sz = np.array([tmat_size(int(v)) for v in nm])
o0, o1 = tb0.toff.cpu().numpy(), tb1.toff.cpu().numpy()
valid = (o0 >= 0) & (o1 >= 0)
ok &= bool(((o0 >= 0) == (o1 >= 0)).all())
a0, a1 = tb0.arena.cpu().numpy(), tb1.arena.cpu().numpy()
bad = 0
for i in np.nonzero(valid)[0]:
    x = a0[2 * o0[i]:2 * (o0[i] + sz[i])]; y = a1[2 * o1[i]:2 * (o1[i] + sz[i])]
    if not np.array_equal(x, y):
        bad += 1
        if bad <= 3:
            print("  T differs for particle", i, "nmax", nm[i], "max rel", np.abs(x - y).max() / np.abs(x).max())
print("  packed T bit-identical for %d of %d particles" % (valid.sum() - bad, valid.sum()))
ok &= bad == 0
print("AB", "OK" if ok else "MISMATCH")

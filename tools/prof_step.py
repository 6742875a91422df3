"""One W2 bench step between cudaProfilerStart / Stop, for ncu --profile-from-start off:
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/step.csv python tools/prof_step.py [split]
`split` issues hs_calctmat_batch and hs_scatter_batch separately (stage attribution)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hydroscatt_b200.engine import Engine
from hydroscatt_b200 import tables

split = len(sys.argv) > 1 and sys.argv[1] == "split"
eng = Engine(0)
w = bench.workload_w2()
d0, nw, mu = bench.dsd_grid()
al, be, wt, scale = bench.orientation_nodes()
dev = {k: torch.from_numpy(np.ascontiguousarray(w[k])).cuda() for k in ("radius", "wavelength", "mr", "mi", "eps")}
dsd = [torch.from_numpy(a).cuda() for a in (d0, nw, mu)]
dmax = torch.full((d0.size,), 7.0, dtype=torch.float64, device="cuda")
cols = [tables.OBS_COLUMNS.index(c) for c in bench.OBS_NORTH_STAR]


def step():
    if split:
        tb = eng.calctmat(dev["radius"], dev["wavelength"], dev["mr"], dev["mi"], dev["eps"], ddelt=bench.DDELT, ndgs=bench.NDGS)
        S, Z, xs = tb.scatter([bench.GEOM_BACK, bench.GEOM_FORW], al, be, weight=wt, scale=scale, want_xs=True)
    else:
        tb, S, Z, xs = eng.calctmat_scatter(dev["radius"], dev["wavelength"], dev["mr"], dev["mi"], dev["eps"], [bench.GEOM_BACK, bench.GEOM_FORW],
                                            al, be, weight=wt, scale=scale, want_xs=True, ddelt=bench.DDELT, ndgs=bench.NDGS)
    rows = eng.pack_rows(S, Z, xs, tb.lam, 0, 1)
    obs, _ = tables.integrate_psd(eng, rows, w["n_tab"], w["D"], "gamma", *dsd, dmax, w["lam_tab"], rule="trapz", want_integ=False, columns=cols)
    return obs


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("one step profiled; launches so far", eng.launch_count)

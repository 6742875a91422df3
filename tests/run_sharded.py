"""torchrun helper of tests/test_tables_gpu.py: dist.build_rows_sharded over NCCL (one rank per GPU) must reproduce the
single-GPU tables.build_rows bit for bit on every rank.  Prints 'SHARDED_OK <rank>' per rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from hydroscatt_b200.engine import Engine
    from hydroscatt_b200 import tables, dist as hdist
    from tests.workloads import rain_grid
    eng = Engine(local)
    al, be, wt, scale = tables.fixed_orientation_nodes(10.0)
    ws = [rain_grid(b) for b in ("S", "C", "X", "Ku", "Ka", "W")]
    cat = lambda k: np.concatenate([w[k] for w in ws])[:-3]          # 417 particles: uneven shards
    gb, gf = (90.0, 90.0, 0.0, 180.0), (90.0, 90.0, 0.0, 0.0)
    args = (cat("radius"), cat("wavelength"), cat("mr"), cat("mi"), cat("axis_ratio"), gb, gf, al, be, wt, scale)
    rows, meta = hdist.build_rows_sharded(eng, *args)
    ref, tb = tables.build_rows(eng, *args)
    assert torch.equal(rows, ref), float((rows - ref).abs().max())
    assert torch.equal(meta[:, 0], tb.nmax) and torch.equal(meta[:, 1], tb.ngauss) and torch.equal(meta[:, 3], tb.status)
    dist.barrier()
    print("SHARDED_OK", dist.get_rank(), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU suite: the multi-GPU sharding / gather logic with world_size 2 over gloo (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hydroscatt_b200 import dist as hdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        cost = rng.uniform(1, 100, n) ** 2
        idx = hdist.shard_indices(cost, world, rank)
        local = torch.from_numpy(np.stack([idx * 1.0, idx * 2.0 + 0.5, cost[idx]], axis=1))
        full = hdist.gather_rows(local, idx, n)
        ok = bool(torch.equal(full[:, 0], torch.arange(n, dtype=torch.float64))) and \
            bool(torch.allclose(full[:, 2], torch.from_numpy(cost)))
        # ShardPlan: same deal, one all_gather_into_tensor + reorder (uneven shards: n is odd, the short one is padded)
        plan = hdist.ShardPlan(cost)
        ok = ok and np.array_equal(plan.idx, idx) and plan.world == world and plan.rank == rank
        full2 = plan.gather(local)
        ok = ok and bool(torch.equal(full2, full))
        part = torch.full((4, 3), float(rank + 1), dtype=torch.float64)
        hdist.allreduce_sum(part)
        ok = ok and bool((part == 3.0).all())
        q.put((rank, ok, float(cost[idx].sum())))
    finally:
        dist.destroy_process_group()


def test_shard_is_a_partition_and_balanced():
    rng = np.random.default_rng(0)
    cost = rng.uniform(1, 30, 1001) ** 4
    parts = [hdist.shard_indices(cost, 4, r) for r in range(4)]
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(1001))
    loads = np.array([cost[p].sum() for p in parts])
    assert loads.max() / loads.min() < 1.05


def test_predicted_cost_orders_bands():
    c = hdist.predicted_cost([3.5, 3.5], [53.5, 3.19], [8.6, 3.1], [1.7, 1.7], [1.6, 1.6])
    assert c[1] > 50 * c[0]


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 257
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    loads = [l for _, _, l in res]
    assert max(loads) / min(loads) < 1.2


def test_gather_single_process():
    idx = np.array([2, 0, 1])
    rows = torch.tensor([[20.0], [0.0], [10.0]], dtype=torch.float64)
    full = hdist.gather_rows(rows, idx, 3)
    assert full[:, 0].tolist() == [0.0, 10.0, 20.0]


def test_bench_reference_arm_under_torchrun_world2():
    """bench.py --impl reference launched like the driver does for N = 2: rank 0 alone times the CPU port and prints ONE
    JSON line with the contract's keys, the other rank exits 0 without work (bounded sample: every 64th particle)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--cpu-stride", "64", "--cpu-tm2l-stride", "1024"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "T-matrix solves/s" and d["unit"] == "solves/s" and d["n_gpus"] == 2
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["tm2l"]["value"] > 0 and d["tm2l"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

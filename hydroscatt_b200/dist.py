"""Multi-GPU sharding of the particle grid: one process per GPU, torch.distributed for the plumbing (NCCL on the
GPUs; the same code runs over gloo on CPU tensors in the tests).

Particles are independent, so the solve needs no data-path collective: the flattened particle list
is dealt to the ranks by cost-sorted round-robin (every GPU gets the same mix of small and large NMAX), each rank
builds the table records of its shard, and ONE all-gather returns the complete table to every rank.  PSD partial
sums (when the integral is split over diameters) are combined with one all-reduce.
"""
import numpy as np
import torch
import torch.distributed as dist


def predicted_cost(radius, wavelength, m_re, m_im, axis_ratio):
    """~NMAX^4 work estimate per particle from the same predictor the engine uses for bucketing."""
    radius, wavelength, m_re, m_im, axis_ratio = [np.asarray(a, dtype=np.float64) for a in
                                                  (radius, wavelength, m_re, m_im, axis_ratio)]
    x = 2.0 * np.pi * radius / wavelength
    n0 = np.maximum(4.0, np.floor(x + 4.05 * x ** 0.333333))
    a = axis_ratio ** (1.0 / 3.0)
    mx = np.hypot(m_re, m_im) * x * np.maximum(a, a / axis_ratio)
    return np.maximum(n0 + 1.0, np.ceil(mx + 2.0)) ** 4


def shard_indices(cost, world_size, rank):
    """Indices of the particles rank `rank` owns: sort by descending cost, deal round-robin."""
    order = np.argsort(-np.asarray(cost, dtype=np.float64), kind="stable")
    return np.sort(order[rank::world_size])


def gather_rows(local_rows, local_idx, n_total, group=None):
    """All-gather per-particle records.  local_rows [n_local, C] (rows of the particles in local_idx) ->
    full [n_total, C] on every rank, in the original particle order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    C = local_rows.shape[1]
    out = torch.empty((n_total, C), dtype=local_rows.dtype, device=local_rows.device)
    if world == 1:
        out[torch.as_tensor(local_idx, device=local_rows.device)] = local_rows
        return out
    n_max = (n_total + world - 1) // world
    pad = torch.zeros((n_max, C + 1), dtype=local_rows.dtype, device=local_rows.device)
    pad[:, C] = -1.0
    pad[:local_rows.shape[0], :C] = local_rows
    pad[:local_rows.shape[0], C] = torch.as_tensor(local_idx, dtype=local_rows.dtype, device=local_rows.device)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    for b in bufs:
        idx = b[:, C]
        keep = idx >= 0
        out[idx[keep].to(torch.int64)] = b[keep][:, :C]
    return out


def allreduce_sum(t, group=None):
    """Sum PSD partial integrals over the ranks (in place)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class ShardPlan:
    """Cost-sorted round-robin deal of n particles over the ranks, and the permutation that puts the all-gathered shard records
    back into particle order.  Built once per particle list (host side, numpy); `gather` is then one collective
    (all_gather_into_tensor) plus one reorder kernel per step."""

    def __init__(self, cost, world=None, rank=None, group=None):
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        cost = np.asarray(cost, dtype=np.float64)
        order = np.argsort(-cost, kind="stable")
        self.world, self.rank, self.group, self.n = world, rank, group, cost.size
        self.parts = [np.sort(order[r::world]) for r in range(world)]
        self.idx = self.parts[rank]
        self.n_max = max(p.size for p in self.parts) if cost.size else 0
        inv = np.empty(cost.size, dtype=np.int64)
        for r, part in enumerate(self.parts):
            inv[part] = r * self.n_max + np.arange(part.size)
        self.inv = inv                    # particle i sits at record inv[i] of the gathered (rank-major) array
        self._inv_dev = {}

    def inv_on(self, device):
        key = str(device)
        if key not in self._inv_dev:
            self._inv_dev[key] = torch.from_numpy(self.inv).to(device)
        return self._inv_dev[key]

    def gather(self, local, eng=None, out=None):
        """local [n_local, C] records of this rank's particles (in the order of self.idx) -> [n, C] records of ALL particles
        in particle order, on every rank.  eng: hydroscatt_b200 Engine (the reorder runs as hs_gather_rows on its device);
        None: torch index_select (CPU tensors, gloo)."""
        C = local.shape[1]
        if self.world == 1:
            gathered = local
        else:
            if local.shape[0] != self.n_max:
                pad = torch.zeros((self.n_max, C), dtype=local.dtype, device=local.device)
                pad[:local.shape[0]] = local
                local = pad
            gathered = torch.empty((self.world * self.n_max, C), dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(gathered, local.contiguous(), group=self.group)
        inv = self.inv_on(local.device)
        if eng is not None:
            return eng.gather_rows(gathered, inv, out=out)
        res = gathered.index_select(0, inv)
        if out is not None:
            out.copy_(res)
            return out
        return res


def build_rows_sharded(eng, radius, wavelength, m_re, m_im, axis_ratio, geom_back, geom_forw, alpha, beta, weight,
                       scale, group=None, plan=None, **kw):
    """Sharded table build: every rank computes its cost-balanced shard (no collective inside the solve), then ONE all-gather
    gives every rank the complete table.  Returns (rows [n, 56], meta [n, 4] int32 = nmax, ngauss, trials,
    status) in particle order."""
    from . import tables
    arrs = [np.asarray(a, dtype=np.float64) for a in (radius, wavelength, m_re, m_im, axis_ratio)]
    n = max(a.size for a in arrs)
    arrs = [np.broadcast_to(a, (n,)) for a in arrs]
    if plan is None:
        plan = ShardPlan(predicted_cost(*arrs), group=group)
    idx = plan.idx
    rows, tb = tables.build_rows(eng, *[a[idx] for a in arrs], geom_back, geom_forw, alpha, beta, weight, scale, **kw)
    meta = torch.stack([tb.nmax, tb.ngauss, tb.ntrials, tb.status], dim=1).to(torch.float64)
    full = plan.gather(torch.cat([rows, meta], dim=1), eng=eng)
    return full[:, :tables.ROW].contiguous(), full[:, tables.ROW:].to(torch.int32)


def bind_to_gpu_cpus(device_index):
    """One process per GPU: run this process (and the pinned host buffers it allocates afterwards) on the CPU cores NVML
    reports as local to the GPU, so that result copies do not cross the socket interconnect.  Returns the CPU set, or
    None if NVML is unavailable / reports no affinity / the affinity is the whole machine (nothing to do)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None

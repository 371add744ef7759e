"""Multi-GPU: one process per GPU, candidates sharded contiguously, no data-path
collective; only the (min value, index) pair is reduced.  Works with
any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of `total` units owned by `rank`: GPU g owns
    [g*B/G, (g+1)*B/G) (integer arithmetic, covers every unit exactly once)."""
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    return lo, hi


def allreduce_argmin(local_val: float, local_idx: int, device=None):
    """Global which.min over ranks: lowest value, lowest GLOBAL index on ties, NaN/empty
    shards (idx < 0) never win.  NCCL has no MINLOC, so: all-reduce(MIN) on the value,
    then all-reduce(MIN) on `index if value == min else INT64_MAX`.  Returns (val, idx)
    identical on every rank and independent of the world size."""
    import torch
    import torch.distributed as dist
    big = float("inf")
    v = local_val if (local_idx >= 0 and local_val == local_val) else big
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return (local_val, local_idx) if v != big else (float("nan"), -1)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    tv = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(tv, op=dist.ReduceOp.MIN)
    gmin = float(tv.item())
    imax = np.iinfo(np.int64).max
    ti = torch.tensor([local_idx if (v == gmin and v != big) else imax], dtype=torch.int64, device=dev)
    dist.all_reduce(ti, op=dist.ReduceOp.MIN)
    gi = int(ti.item())
    if gi == imax:
        return float("nan"), -1
    return gmin, gi


def sharded_nll_argmin(engine, cand, family, sigma2, rank, world, **kw):
    """Each rank evaluates its slice of the candidate rows on its own GPU and the argmin is
    all-reduced.  cand: full (B x k) host matrix (every rank holds the same seeded matrix)."""
    lo, hi = shard_range(cand.shape[0], rank, world)
    if hi > lo:
        v, i = engine.nll_argmin(cand[lo:hi], family, sigma2, **kw)
        i = i + lo if i >= 0 else -1
    else:
        v, i = float("nan"), -1
    return allreduce_argmin(v, i)

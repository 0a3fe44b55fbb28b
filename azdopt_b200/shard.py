"""Sharding of roots over one process per GPU: contiguous blocks of roots per rank, no collective
inside a step; epoch-boundary collectives over torch.distributed (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank's roots; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_argmin(local_eval: float, local_parents: np.ndarray, local_permitted: np.ndarray, dist=None, device="cpu"):
    """All-gather of each rank's (best eval, state) and a first-minimum over ranks (optimizer/mod.rs:221: lowest
    index wins ties).  Returns (eval, rank, parents, permitted)."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_eval), 0, local_parents, local_permitted
    world = dist.get_world_size()
    rec = np.concatenate([np.array([local_eval], dtype=np.float32).view(np.uint8), local_parents.astype(np.uint8),
                          local_permitted.astype(np.uint32).view(np.uint8)])
    t = torch.from_numpy(rec.copy()).to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    recs = [o.cpu().numpy() for o in out]
    evals = [float(r[:4].view(np.float32)[0]) for r in recs]
    best = min(range(world), key=lambda r: (evals[r], r))
    r = recs[best]
    n = local_parents.size
    return evals[best], best, r[4:4 + n].copy(), r[4 + n:].view(np.uint32).copy()


def sum_counters(counters: dict, dist=None, device="cpu") -> dict:
    """Sum the workload counters of all ranks (whole-job simulations = sum of n_live)."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(counters)
    keys = sorted(counters)
    t = torch.tensor([counters[k] for k in keys], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return {k: int(v) for k, v in zip(keys, t.cpu().tolist())}

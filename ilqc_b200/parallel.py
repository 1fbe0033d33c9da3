"""Multi-GPU plumbing: the batch is the only parallel axis.

Problems are independent, so every rank (one process per GPU, torchrun / torch.distributed) solves a
contiguous slice of the batch with the identical kernels and no communication; the single collective of the
path is the final all-gather of the results (costs, controls, step sizes, status) over NCCL / NVLink.
"""
from dataclasses import fields

import torch
import torch.distributed as dist

from .engine import SolveResult


def shard_bounds(n_total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of rank `rank`; the first n_total % world ranks get one extra problem."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank: int, world: int):
    """Slice of a per-problem tensor ([B_total, ...]) owned by `rank` (None passes through)."""
    if t is None:
        return None
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def all_gather_result(res: SolveResult, n_total: int, group=None) -> SolveResult:
    """All-gather every per-problem field of a SolveResult (uneven shards are padded to the largest one)."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    out = {}
    for f in fields(res):
        t = getattr(res, f.name)
        if t is None:
            out[f.name] = None
            continue
        t = t.contiguous()
        if t.shape[0] < n_max:
            pad = torch.zeros((n_max - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat((t, pad))
        buf = torch.empty((world * n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t, group=group)
        parts = [buf[r * n_max: r * n_max + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
        out[f.name] = torch.cat(parts)
    return SolveResult(**out)


def solve_sharded(engine, env, algo, max_iter, x0, weights=None, cmd0=None, stepsize=None, group=None, **kw):
    """run_min_algo for a global batch split over the ranks of `group`; returns the gathered SolveResult on
    every rank.  x0 / weights / cmd0 are the GLOBAL per-problem tensors (host or device)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_total = x0.shape[0]
    st = shard(stepsize, rank, world) if torch.is_tensor(stepsize) else stepsize
    res = engine.solve(env, algo, max_iter, x0=shard(x0, rank, world), weights=shard(weights, rank, world),
                       cmd0=shard(cmd0, rank, world), stepsize=st, **kw)
    return all_gather_result(res, n_total, group)

"""Sharding of posterior draws / log-prob rows over the GPUs of one box (one process per GPU).

Every test row is independent given the context, so the rows are split evenly over the
ranks, the (small) context, weights and K/V caches are replicated, and there is NO collective on the data
path.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used only to gather the finished
draws / log-probs and to reduce accept counts.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first `total % world_size` ranks get one extra row."""
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_rows(local: torch.Tensor, total: int) -> torch.Tensor:
    """all_gather of ragged row shards (shard_bounds order) -> [total, ...] on every rank."""
    rank, world = _world()
    if world == 1:
        return local
    sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def reduce_counts(accepted: int, drawn: int, device=None) -> Tuple[int, int]:
    """all_reduce(SUM) of (accepted, drawn) -> global acceptance statistics."""
    rank, world = _world()
    if world == 1:
        return accepted, drawn
    t = torch.tensor([accepted, drawn], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t[0]), int(t[1])


def sample_sharded(posterior, num_samples: int, x: torch.Tensor, gather: bool = True, **sample_kwargs):
    """`posterior.sample((num_samples,), x)` with the draws split over the ranks.

    Each rank draws its share with a disjoint Philox row range and (by default) all ranks receive the
    concatenation in rank order.  Returns (samples, global acceptance rate)."""
    rank, world = _world()
    lo, hi = shard_bounds(num_samples, world, rank)
    posterior.rank_row_offset = rank << 40
    local = posterior.sample((hi - lo,), x, **sample_kwargs) if hi > lo else None
    with_lp = isinstance(local, tuple)
    acc = getattr(posterior, "last_acceptance_rate", 1.0) or 1.0
    dev = posterior.engine.device if world > 1 and dist.get_backend() == "nccl" else None
    n_acc, n_drawn = reduce_counts(hi - lo, int(round((hi - lo) / max(acc, 1e-12))), device=dev)
    rate = n_acc / max(n_drawn, 1)
    if not gather or world == 1:
        return local, rate
    if with_lp:
        s, lp = local
        if dev is not None:
            s, lp = s.to(dev), lp.to(dev)
        return (gather_rows(s, num_samples), gather_rows(lp, num_samples)), rate
    s = local.to(dev) if dev is not None else local
    return gather_rows(s, num_samples), rate


def log_prob_sharded(posterior, theta: torch.Tensor, x: torch.Tensor, **kw) -> torch.Tensor:
    rank, world = _world()
    lo, hi = shard_bounds(theta.shape[0], world, rank)
    local = posterior.log_prob(theta[lo:hi], x, **kw)
    if world == 1:
        return local
    dev = posterior.engine.device if dist.get_backend() == "nccl" else None
    return gather_rows(local.to(dev) if dev is not None else local, theta.shape[0])

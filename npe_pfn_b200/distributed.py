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


def _gather_device(posterior):
    """device on which the gather runs: the engine's GPU under NCCL, None (host tensors) under gloo"""
    rank, world = _world()
    return posterior.engine.device if world > 1 and dist.get_backend() == "nccl" else None


def sample_sharded(posterior, num_samples: int, x: torch.Tensor, gather: bool = True, device_result: bool = False,
                   **sample_kwargs):
    """`posterior.sample((num_samples,), x)` with the draws split over the ranks.

    Each rank draws its share with a disjoint Philox row range and (by default) all ranks receive the
    concatenation in rank order.  `device_result=True` keeps the (gathered) draws on the GPU instead of returning
    host tensors.  Returns (samples, global acceptance rate)."""
    rank, world = _world()
    lo, hi = shard_bounds(num_samples, world, rank)
    posterior.rank_row_offset = rank << 40
    if device_result:
        sample_kwargs = dict(sample_kwargs, return_device=True)
    local = posterior.sample((hi - lo,), x, **sample_kwargs) if hi > lo else None
    with_lp = isinstance(local, tuple)
    acc = getattr(posterior, "last_acceptance_rate", 1.0) or 1.0
    dev = _gather_device(posterior)
    n_acc, n_drawn = reduce_counts(hi - lo, int(round((hi - lo) / max(acc, 1e-12))), device=dev)
    rate = n_acc / max(n_drawn, 1)
    if not gather or world == 1:
        return local, rate

    def g(t):
        t = gather_rows(t.to(dev) if dev is not None else t, num_samples)
        return t if device_result or dev is None else t.cpu()

    if with_lp:
        return (g(local[0]), g(local[1])), rate
    return g(local), rate


def log_prob_sharded(posterior, theta: torch.Tensor, x: torch.Tensor, **kw) -> torch.Tensor:
    rank, world = _world()
    lo, hi = shard_bounds(theta.shape[0], world, rank)
    local = posterior.log_prob(theta[lo:hi], x, **kw)
    if world == 1:
        return local
    dev = _gather_device(posterior)
    out = gather_rows(local.to(dev) if dev is not None else local, theta.shape[0])
    return out.cpu() if dev is not None else out


def sample_batched_sharded(posterior, x: torch.Tensor, num_samples: int, gather: bool = True, **kw):
    """`posterior.sample_batched(x, (num_samples,))` with the OBSERVATIONS split over the ranks (BASELINE config 4: many
    observations against one shared context).  Every rank holds the same simulations, prefills the same per-dimension
    K/V caches and draws for its contiguous block of observations with a disjoint Philox row range; the per-observation
    results are gathered in observation order.  -> [num_obs, num_samples, dim_theta] (and log-probs if requested)."""
    rank, world = _world()
    num_obs = x.shape[0]
    lo, hi = shard_bounds(num_obs, world, rank)
    posterior.rank_row_offset = rank << 40
    local = posterior.sample_batched(x[lo:hi], (num_samples,), **kw) if hi > lo else None
    if not gather or world == 1:
        return local
    dev = _gather_device(posterior)
    with_lp = isinstance(local, tuple) or bool(kw.get("with_log_prob"))
    d_theta = posterior._theta_train.shape[1]

    def g(t, shape):
        if t is None:
            t = torch.empty((0,) + shape, dtype=torch.float32)
        out = gather_rows(t.to(dev) if dev is not None else t, num_obs)
        return out.cpu() if dev is not None else out

    if with_lp:
        s, lp = local if local is not None else (None, None)
        return g(s, (num_samples, d_theta)), g(lp, (num_samples,))
    return g(local, (num_samples, d_theta))

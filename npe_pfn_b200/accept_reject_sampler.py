"""Rejection loop of `/root/reference/npe_pfn/accept_reject_sampler.py:8-91`, same semantics:
first batch `min(num_samples, max_bs)`, then `min(max_bs, max(int(1.5 * remaining / acc), 100))`,
keep the FIRST `num_samples` accepted rows in proposal order, return a 3-tuple
`(samples, log_probs | None, acceptance_rate)`; when `max_iter_rejection` is exceeded the last
unfiltered candidate batch is appended (Appendix B.3 of SURVEY.md).

Works on whatever device the proposal returns (CUDA tensors from the B200 path stay on the device;
the only host synchronisation per round is the accepted count, as in the reference's `.sum().item()`).
`accept_reject_fn` may return either a boolean mask or, for the fused device path, a callable
attribute `compact(candidates, log_probs)` is used when present (support check + ordered compaction
in one kernel chain, `pfn_accept_compact`).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor
from tqdm import tqdm


@torch.no_grad()
def accept_reject_sample(
    proposal: Callable,
    accept_reject_fn: Callable,
    num_samples: int,
    show_progress_bars: bool = False,
    max_sampling_batch_size: int = 10_000,
    proposal_sampling_kwargs: Optional[Dict] = None,
    max_iter_rejection: int | None = None,
) -> Tuple[Tensor, Optional[Tensor], float]:
    if proposal_sampling_kwargs is None:
        proposal_sampling_kwargs = {}
    pbar = tqdm(disable=not show_progress_bars, total=num_samples, desc=f"Drawing {num_samples} posterior samples")

    accepted, accepted_log_probs = [], []
    num_remaining = num_samples
    num_sampled_total = 0
    num_accepted_total = 0
    sampling_batch_size = min(num_samples, max_sampling_batch_size)
    i = 0
    compact = getattr(accept_reject_fn, "compact", None)
    while num_remaining > 0:
        i += 1
        candidates, log_probs = proposal(sampling_batch_size, **proposal_sampling_kwargs)
        if compact is not None and candidates.is_cuda:
            kept, kept_lp, num_accepted = compact(candidates, log_probs)
            accepted.append(kept)
            if log_probs is not None:
                accepted_log_probs.append(kept_lp)
        else:
            are_accepted = accept_reject_fn(candidates)
            accepted.append(candidates[are_accepted])
            if log_probs is not None:
                accepted_log_probs.append(log_probs[are_accepted])
            num_accepted = int(are_accepted.sum().item())
        num_sampled_total += sampling_batch_size
        num_accepted_total += num_accepted
        num_remaining -= num_accepted
        pbar.update(num_accepted)

        acceptance_rate = num_accepted_total / num_sampled_total
        sampling_batch_size = min(max_sampling_batch_size,
                                  max(int(1.5 * num_remaining / max(acceptance_rate, 1e-12)), 100))
        if max_iter_rejection is not None and i > max_iter_rejection:
            accepted.append(candidates)
            if log_probs is not None:
                accepted_log_probs.append(log_probs)
            break
    pbar.close()

    samples = torch.cat(accepted, dim=0)[:num_samples]
    log_probs = torch.cat(accepted_log_probs, dim=0)[:num_samples] if accepted_log_probs else None
    final_acceptance_rate = len(samples) / max(num_sampled_total, 1)
    return samples, log_probs, final_acceptance_rate

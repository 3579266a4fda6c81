"""Prior-support rejection around the autoregressive sampler.

Semantics follow `/root/reference/npe_pfn/accept_reject_sampler.py:8-91`: the first round
proposes `min(num_samples, max_bs)` rows, later rounds `min(max_bs, max(int(1.5 * missing / rate), 100))`; the FIRST
`num_samples` accepted rows are returned in proposal order together with their log-probs (or None) and the overall
acceptance rate; once `max_iter_rejection` rounds have passed, the last round's UNFILTERED proposals are appended and the
loop stops.

The loop itself is organised around a small ledger object; tensors stay on the device the proposal returned them on.
When the acceptance function carries a `compact(candidates, log_probs)` attribute (the engine's support check +
ordered stream compaction, `pfn_accept_compact`), a round costs one kernel chain and one scalar read-back — the same
single host synchronisation per round the reference has in `.sum().item()`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import Tensor
from tqdm import tqdm

_MIN_ROUND = 100      # smallest follow-up round
_OVERDRAW = 1.5       # safety factor on the expected number of proposals still needed


def next_round_size(missing: int, proposed: int, kept: int, cap: int) -> int:
    """Rows to propose next, from the running acceptance rate (`accept_reject_sampler.py:64-72` of the reference)."""
    rate = max(kept / proposed, 1e-12)
    return min(cap, max(int(_OVERDRAW * missing / rate), _MIN_ROUND))


@dataclass
class _Ledger:
    wanted: int
    rows: List[Tensor] = field(default_factory=list)
    logps: List[Tensor] = field(default_factory=list)
    proposed: int = 0
    kept: int = 0

    @property
    def missing(self) -> int:
        return self.wanted - self.kept

    def book(self, rows: Tensor, logp: Optional[Tensor], n_proposed: int, n_kept: int) -> None:
        self.rows.append(rows)
        if logp is not None:
            self.logps.append(logp)
        self.proposed += n_proposed
        self.kept += n_kept

    def result(self) -> Tuple[Tensor, Optional[Tensor], float]:
        out = torch.cat(self.rows, dim=0)[:self.wanted]
        lp = torch.cat(self.logps, dim=0)[:self.wanted] if self.logps else None
        return out, lp, len(out) / max(self.proposed, 1)


def _filter_round(accept_reject_fn: Callable, cand: Tensor, logp: Optional[Tensor]):
    """-> (kept rows, kept log-probs | None, number kept) for one round of proposals."""
    fused = getattr(accept_reject_fn, "compact", None)
    if fused is not None and cand.is_cuda:
        return fused(cand, logp)
    mask = accept_reject_fn(cand)
    return cand[mask], (None if logp is None else logp[mask]), int(mask.sum().item())


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
    kwargs = proposal_sampling_kwargs or {}
    ledger = _Ledger(wanted=num_samples)
    bar = tqdm(disable=not show_progress_bars, total=num_samples, desc=f"Drawing {num_samples} posterior samples")
    round_size = min(num_samples, max_sampling_batch_size)
    rounds = 0
    while ledger.missing > 0:
        rounds += 1
        cand, logp = proposal(round_size, **kwargs)
        kept_rows, kept_logp, n_kept = _filter_round(accept_reject_fn, cand, logp)
        ledger.book(kept_rows, kept_logp, round_size, n_kept)
        bar.update(n_kept)
        round_size = next_round_size(ledger.missing, ledger.proposed, ledger.kept, max_sampling_batch_size)
        if max_iter_rejection is not None and rounds > max_iter_rejection:
            ledger.book(cand, logp, 0, 0)  # give up filtering: hand back the raw proposals of this round
            break
    bar.close()
    return ledger.result()

"""TSNPE truncated proposal ("posterior support") and the context filters, over the B200 posterior object.

Behavioural contract (what `/root/reference/npe_pfn/support_posterior.py` does):
  * `PosteriorSupport(prior, posterior, obs, ...)`: draw `num_samples_to_estimate_support` posterior samples, set
    `thr` = the `allowed_false_negatives`-quantile of their posterior log-prob (:42-52, :61-69);
  * `.sample((n,))`, rejection mode: propose from the prior (narrowed to the classifier's box when the posterior
    exposes one, :137-152), keep proposals whose posterior log-prob exceeds `thr`, at most `max_iter_rejection`
    rounds of `sampling_batch_size` proposals, then top up with raw prior draws (:133-174);
  * SIR mode (:184-258): groups of `oversample_sir` posterior draws, one survivor per group drawn with weights
    prior/posterior (draws under the adaptive log-prob quantile get weight zero);
  * context filters (:327-369) return `(theta, x)`; the default keeps the `context_size` simulations nearest to the
    observation in z-scored x, ordered by distance.
"""
from __future__ import annotations

import logging
from typing import Any, Callable, Dict, Mapping, Optional, Tuple

import torch
from torch import Tensor
from torch.distributions import Independent, Uniform
from tqdm.auto import tqdm

from .utils import BoxUniform

log = logging.getLogger(__name__)


class PosteriorSupport:
    def __init__(
        self,
        prior: Any,
        posterior: Any,
        obs: Tensor,
        num_samples_to_estimate_support: int = 10_000,
        batch_size_for_estimate_support: int = 10_000,
        allowed_false_negatives: float = 0.0,
        sampling_method: str = "rejection",
        max_iter_rejection: int = 1000,
        oversample_sir: int = 100,
        log_prob_kwargs: Mapping = {},
    ) -> None:
        self._prior, self._posterior, self._obs = prior, posterior, obs
        self._log_prob_kwargs = dict(log_prob_kwargs)
        self._posterior_thr = None
        self.sampling_method = sampling_method
        self.max_iter = max_iter_rejection
        self.oversample_sir = oversample_sir
        self.allowed_false_negatives = allowed_false_negatives
        if sampling_method == "rejection":
            # the same posterior draws serve for the quantile and (implicitly) for the support estimate
            draws = posterior.sample((num_samples_to_estimate_support,), obs,
                                     max_sampling_batch_size=batch_size_for_estimate_support)
            self.thr = self.tune_threshold(draws, allowed_false_negatives, batch_size=batch_size_for_estimate_support)

    # -- threshold ------------------------------------------------------------------------------------------
    def _posterior_log_prob(self, theta: Tensor, **extra) -> Tensor:
        return self._posterior.log_prob(theta, self._obs, **extra, **self._log_prob_kwargs)

    def tune_threshold(self, samples: Tensor, allowed_false_negatives: float = 0.0, batch_size: int = 10_000):
        return torch.quantile(self._posterior_log_prob(samples, max_sampling_batch_size=batch_size),
                              allowed_false_negatives)

    # -- dispatch -------------------------------------------------------------------------------------------
    def sample(self, sample_shape: torch.Size = torch.Size(), show_progress_bars: bool = True,
               sampling_batch_size: int = 10_000, return_acceptance_rate: bool = False, return_ess: bool = False):
        common = dict(sample_shape=sample_shape, show_progress_bars=show_progress_bars,
                      sampling_batch_size=sampling_batch_size)
        if self.sampling_method == "rejection":
            return self.sample_rejection(return_acceptance_rate=return_acceptance_rate, **common)
        if self.sampling_method == "sir":
            return self.sample_sir(return_ess=return_ess, **common)
        raise ValueError(f"Unknown sampling method: {self.sampling_method}")

    # -- rejection ------------------------------------------------------------------------------------------
    def _proposal_round(self, box: Tuple[Optional[Tensor], Optional[Tensor]], n: int):
        """One batch of prior proposals and their posterior log-prob.  Returns (candidates, log_probs, box,
        pre-acceptance rate); `box` is the classifier's padded bounding box once the posterior has one."""
        lo, hi = box
        if lo is None or hi is None:
            cand, pre_rate = self._prior.sample((n,)), None
        else:
            cand, pre_rate = prereject_with_bounds(self._prior, lo, hi, n)
        lp = self._posterior_log_prob(cand)
        new_lo, new_hi = self._posterior._get_classifier_bounds()
        if lo is not None and hi is not None:  # the box must not move between rounds
            assert torch.allclose(lo, new_lo) and torch.allclose(hi, new_hi)
        return cand, lp, (new_lo, new_hi), pre_rate

    #: run the proposal loop on the GPU when the prior is a uniform box and the posterior is engine-backed
    device_rejection = True

    def _sample_rejection_device(self, wanted: int, sampling_batch_size: int):
        """The truncated-prior proposal loop (support_posterior.py:133-160) without a host round trip per round: uniform
        proposals are drawn on the device (`pfn_uniform_box`; inside the classifier's padded box intersected with the
        prior box when the posterior has one, which is what `prereject_with_bounds` does for a uniform prior,
        :305-309), their posterior log-prob stays on the device (`NPE_PFN_Core._log_prob_device`), and
        `log_prob > thr` + ordered compaction append the accepted proposals at a device cursor (`pfn_accept_append`).
        Rounds are enqueued in batches sized by the running acceptance rate; the host reads the cursor once per batch.
        -> (accepted [<= wanted, d] on the device, log-prob acceptance rate, pre-acceptance rate, rounds used)"""
        post = self._posterior
        eng = post.engine
        dev = eng.device
        p_lo, p_hi = get_uniform_bounds(self._prior)
        kw = dict(self._log_prob_kwargs)
        mode = kw.pop("mode", "autoregressive")
        box_lo, box_hi, pre_rate = p_lo, p_hi, 1.0
        if mode == "ratio_based":
            post._ensure_ratio_classifier(post._validate_x(self._obs), **{k: v for k, v in kw.items() if k != "eps"})
            c_lo, c_hi = post._get_classifier_bounds()
            box_lo, box_hi = torch.max(c_lo, p_lo), torch.min(c_hi, p_hi)
            # volume fraction of the prior box inside the classifier's box (the reference estimates it from 10^6 raw draws)
            pre_rate = float(torch.clamp(box_hi - box_lo, min=0).div(p_hi - p_lo).prod())
        d = p_lo.numel()
        out = torch.empty(max(wanted, 1), d, dtype=torch.float32, device=dev)
        cursor = torch.zeros(2, dtype=torch.int64, device=dev)
        from .estimator import draw_seed
        seed = draw_seed()
        accepted = proposed = rounds = 0
        rate, first = 1.0, True
        self.last_sync_count = 0
        while accepted < wanted and rounds < self.max_iter:
            need = (wanted - accepted) / (rate * sampling_batch_size) * (1.0 if first else 1.5)
            n_rounds = int(min(self.max_iter - rounds, max(1, -(-need // 1))))
            for k in range(n_rounds):
                cand = eng.uniform_box(box_lo, box_hi, sampling_batch_size, seed, row0=proposed + k * sampling_batch_size)
                lp = post._log_prob_device(cand, self._obs, mode=mode, **kw)
                eng.accept_append(cand, out, cursor, score=lp, thr=self.thr)
            rounds += n_rounds
            proposed += n_rounds * sampling_batch_size
            accepted = int(cursor[0].item())  # one read-back per batch of rounds
            self.last_sync_count += 1
            rate = max(accepted / proposed, 1e-6)
            first = False
        return out[:min(accepted, wanted)], accepted / max(proposed, 1), pre_rate, rounds

    def _device_path_ok(self) -> bool:
        return (self.device_rejection and check_for_uniform(self._prior) and hasattr(self._posterior, "_log_prob_device")
                and self._log_prob_kwargs.get("mode", "autoregressive") in ("autoregressive", "ratio_based")
                and torch.cuda.is_available())

    def sample_rejection(self, sample_shape: torch.Size = torch.Size(), show_progress_bars: bool = True,
                         sampling_batch_size: int = 10_000, return_acceptance_rate: bool = False):
        shape = torch.Size(sample_shape)
        assert len(shape) == 1, "only 1-D sample shapes are supported"
        wanted = shape[0]
        if self._device_path_ok():
            good, lp_rate, pre_rate, _rounds = self._sample_rejection_device(wanted, sampling_batch_size)
            kept = [good.cpu()]
            overall = pre_rate * lp_rate
            log.info("pre-acceptance %.4g, log-prob acceptance %.4g, overall %.4g", pre_rate, lp_rate, overall)
            if good.shape[0] < wanted:  # iteration budget exhausted: fill with unrestricted prior draws (:171-174)
                kept.append(self._prior.sample((wanted - good.shape[0],)))
            out = torch.cat(kept)[:wanted]
            return (out, overall) if return_acceptance_rate else out
        bar = tqdm(disable=not show_progress_bars, total=wanted, desc=f"Drawing {wanted} restricted posterior samples")
        kept, n_kept, n_proposed = [], 0, 0
        box: Tuple[Optional[Tensor], Optional[Tensor]] = (None, None)
        pre_rate = 1.0
        rounds = 0
        while n_kept < wanted and rounds < self.max_iter:
            cand, lp, box, r = self._proposal_round(box, sampling_batch_size)
            if r is not None:
                pre_rate = r
            good = cand[(lp > self.thr).bool()]
            kept.append(good)
            n_kept += good.shape[0]
            n_proposed += sampling_batch_size
            rounds += 1
            bar.update(good.shape[0])
        bar.close()
        lp_rate = n_kept / max(n_proposed, 1)
        overall = pre_rate * lp_rate
        log.info("pre-acceptance %.4g, log-prob acceptance %.4g, overall %.4g", pre_rate, lp_rate, overall)
        if n_kept < wanted:  # iteration budget exhausted: fill with unrestricted prior draws
            missing = wanted - n_kept
            kept.append(self._prior.sample((missing,)))
            log.info("max_iter reached: added %d raw prior samples", missing)
        out = torch.cat(kept)[:wanted]
        assert out.shape[0] == wanted
        return (out, overall) if return_acceptance_rate else out

    # -- sampling importance resampling --------------------------------------------------------------------------
    def sample_sir(self, sample_shape: torch.Size = torch.Size(), show_progress_bars: bool = True,
                   sampling_batch_size: int = 10_000, return_ess: bool = False):
        shape = torch.Size(sample_shape)
        assert len(shape) == 1, "only 1-D sample shapes are supported"
        wanted = shape[0]
        k = self.oversample_sir
        assert sampling_batch_size % k == 0, "sampling_batch_size must be a multiple of oversample_sir"
        groups = sampling_batch_size // k
        bar = tqdm(disable=not show_progress_bars, total=wanted, desc=f"Drawing {wanted} restricted posterior samples")
        chosen, ess_parts, have = [], [], 0
        while have < wanted:
            theta, lp_post = self._posterior.sample((sampling_batch_size,), self._obs,
                                                    max_sampling_batch_size=sampling_batch_size, with_log_prob=True)
            lp_prior = self._prior.log_prob(theta)
            cut = torch.quantile(lp_post, self.allowed_false_negatives)  # adaptive threshold per batch
            lp_prior = torch.where(lp_post < cut, torch.full_like(lp_prior, -float("inf")), lp_prior)
            logw = torch.nan_to_num(lp_prior - lp_post, -float("inf")).reshape(groups, k)
            w = torch.softmax(logw, dim=1)
            ess_parts.append(1.0 / (w * w).sum(dim=1))
            pick = torch.distributions.Categorical(logits=logw).sample()
            chosen.append(theta.reshape(groups, k, -1)[torch.arange(groups), pick])
            have += groups
            bar.update(groups)
        bar.close()
        out = torch.cat(chosen)[:wanted]
        assert out.shape[0] == wanted
        ess = torch.cat(ess_parts)
        log.info("ESS mean %.3f min %.3f", ess.mean().item(), ess.min().item())
        return (out, ess) if return_ess else out


# ---- box pre-rejection ------------------------------------------------------------------------------------------
def check_for_uniform(proposal: Any) -> bool:
    return isinstance(proposal, BoxUniform) or (isinstance(proposal, Independent)
                                                and isinstance(proposal.base_dist, Uniform))


def get_uniform_bounds(proposal) -> Tuple[Tensor, Tensor]:
    return proposal.base_dist.low, proposal.base_dist.high


def prereject_with_bounds(proposal: Any, lower_bound: Tensor, upper_bound: Tensor, sampling_batch_size: int = 10_000,
                          pre_sampling_batch_size: int = 1_000_000):
    """`sampling_batch_size` proposal draws inside [lower, upper] and the fraction of raw draws that fell inside.

    A uniform proposal is intersected with the box analytically (one raw batch only estimates the rate);
    anything else is filtered batch by batch until enough draws survive (support_posterior.py:264-309)."""
    uniform = check_for_uniform(proposal)
    inside, n_inside, n_raw = [], 0, 0
    while True:
        raw = proposal.sample((pre_sampling_batch_size,))
        ok = ((raw >= lower_bound) & (raw <= upper_bound)).all(dim=1)
        inside.append(raw[ok])
        n_inside += int(ok.sum())
        n_raw += pre_sampling_batch_size
        if uniform or n_inside >= sampling_batch_size:
            break
    rate = n_inside / n_raw
    if uniform:
        p_lo, p_hi = get_uniform_bounds(proposal)
        return BoxUniform(torch.max(lower_bound, p_lo), torch.min(upper_bound, p_hi)).sample((sampling_batch_size,)), rate
    return torch.cat(inside)[:sampling_batch_size], rate


# ---- context filters: (obs, theta, x, context_size) -> (theta, x) -------------------------------------------------
def no_filtering(obs: Tensor, theta: Tensor, x: Tensor, context_size: int):
    return theta, x


def latest_filtering(obs: Tensor, theta: Tensor, x: Tensor, context_size: int):
    """the newest simulations are at the end"""
    return theta[-context_size:], x[-context_size:]


def random_filtering(obs: Tensor, theta: Tensor, x: Tensor, context_size: int):
    keep = torch.randperm(theta.shape[0])[:context_size]
    return theta[keep], x[keep]


def standardized_euclidean_filtering(obs: Tensor, theta: Tensor, x: Tensor, context_size: int):
    mu, sd = x.mean(dim=0), x.std(dim=0)
    dist = torch.norm((x - mu) / sd - (obs - mu) / sd, dim=1)
    nearest = torch.topk(dist, min(context_size, dist.shape[0]), largest=False).indices
    return theta[nearest], x[nearest]


_FILTERS: Dict[str, Callable] = {
    "no_filtering": no_filtering,
    "latest_filtering": latest_filtering,
    "random_filtering": random_filtering,
    "standardized_euclidean_filtering": standardized_euclidean_filtering,
}


def get_filtering_method(name):
    if isinstance(name, str) and name in _FILTERS:
        return _FILTERS[name]
    if callable(name):
        return name
    raise ValueError(f"Unknown filtering method: {name}")

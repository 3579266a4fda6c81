"""Round driver of TSNPE-PFN over the B200 posterior.

Contract of `/root/reference/npe_pfn/tsnpe_pfn.py:14-119`: split `num_simulations` over `num_rounds`; in every
round simulate from the current proposal (the prior in round 1), hand ALL simulations so far to
`append_simulations`, and — except after the last round — wrap the posterior in a `PosteriorSupport`
truncated prior that becomes the next proposal.  Returns the posterior object.
"""
from __future__ import annotations

import dataclasses
import logging
from typing import Callable, List, Mapping

import torch
from torch.distributions import Distribution

from .npe_pfn import TabPFN_Based_NPE_PFN
from .support_posterior import PosteriorSupport
from .utils import simulate_for_sbi

log = logging.getLogger(__name__)


@dataclasses.dataclass
class _Schedule:
    rounds: int
    sims_per_round: int
    sim_batch: int

    @staticmethod
    def plan(num_simulations: int, num_rounds: int, simulation_batch_size: int) -> "_Schedule":
        per_round = num_simulations if num_rounds == 1 else num_simulations // num_rounds
        if simulation_batch_size > per_round:
            log.warning("simulation_batch_size reduced to the %d simulations of one round", per_round)
            simulation_batch_size = per_round
        log.info("%s: %d simulations per round", "NPE-PFN" if num_rounds == 1 else "TSNPE-PFN", per_round)
        return _Schedule(num_rounds, per_round, simulation_batch_size)


def run_tsnpe_pfn(
    simulator: Callable,
    prior: Distribution,
    observation: torch.Tensor,
    num_simulations: int = 10_000,
    num_rounds: int = 10,
    proposal_batch_size: int = 1000,
    simulation_batch_size: int = 1000,
    num_samples_to_estimate_support: int = 10_000,
    allowed_false_negatives: float = 0.0001,
    context_size: int = 10_000,
    log_prob_mode: str = "ratio_based",
    sampling_method: str = "rejection",
    max_iter_rejection: int = 1000,
    oversample_sir: int = 100,
    filtering: str = "no_filtering",
    regressor_init_kwargs: Mapping = {},
    classifier_init_kwargs: Mapping = {},
):
    plan = _Schedule.plan(num_simulations, num_rounds, simulation_batch_size)
    posterior = TabPFN_Based_NPE_PFN(prior=prior, filter_type=filtering, filter_context_size=context_size,
                                     regressor_init_kwargs=regressor_init_kwargs,
                                     classifier_init_kwargs=classifier_init_kwargs)
    support_kwargs = dict(num_samples_to_estimate_support=num_samples_to_estimate_support,
                          batch_size_for_estimate_support=proposal_batch_size,
                          allowed_false_negatives=allowed_false_negatives, sampling_method=sampling_method,
                          max_iter_rejection=max_iter_rejection, oversample_sir=oversample_sir,
                          log_prob_kwargs={"mode": log_prob_mode})
    thetas: List[torch.Tensor] = []
    xs: List[torch.Tensor] = []
    proposal = prior
    for r in range(plan.rounds):
        log.info("round %d/%d: simulating from the %s", r + 1, plan.rounds, "prior" if r == 0 else "truncated prior")
        theta_r, x_r = simulate_for_sbi(simulator, proposal, num_simulations=plan.sims_per_round,
                                        simulation_batch_size=plan.sim_batch)
        thetas.append(theta_r)
        xs.append(x_r)
        posterior.append_simulations(torch.cat(thetas), torch.cat(xs))  # replaces: pass every round so far
        if r + 1 < plan.rounds:
            proposal = PosteriorSupport(prior, posterior, obs=observation, **support_kwargs)
    return posterior

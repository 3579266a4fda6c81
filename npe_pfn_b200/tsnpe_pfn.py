"""Round driver of TSNPE-PFN — same control flow as `/root/reference/npe_pfn/tsnpe_pfn.py:14-119`:
R rounds of {simulate from the current proposal, `append_simulations(all rounds so far)`, build a
`PosteriorSupport` truncated prior as the next proposal}."""
from __future__ import annotations

import logging
from typing import Callable, Mapping

import torch
from torch.distributions import Distribution

from .npe_pfn import TabPFN_Based_NPE_PFN
from .support_posterior import PosteriorSupport
from .utils import simulate_for_sbi

log = logging.getLogger(__name__)


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
    if num_rounds == 1:
        log.info("Running NPE_PFN")
        num_simulations_per_round = num_simulations
    else:
        log.info("Running TSNPE_PFN")
        num_simulations_per_round = num_simulations // num_rounds
    log.info(f"Number of simulations per round: {num_simulations_per_round}")
    if simulation_batch_size > num_simulations_per_round:
        simulation_batch_size = num_simulations_per_round
        log.warning("Reduced simulation_batch_size to num_simulation_per_round")

    tabpfn_posterior = TabPFN_Based_NPE_PFN(
        prior=prior,
        regressor_init_kwargs=regressor_init_kwargs,
        classifier_init_kwargs=classifier_init_kwargs,
        filter_type=filtering,
        filter_context_size=context_size,
    )
    proposal = prior
    theta_per_round, x_per_round = [], []
    posterior = tabpfn_posterior
    for round_num in range(num_rounds):
        log.info(f"Round {round_num + 1}/{num_rounds}")
        theta, x = simulate_for_sbi(simulator, proposal, num_simulations=num_simulations_per_round,
                                    simulation_batch_size=simulation_batch_size)
        theta_per_round.append(theta)
        x_per_round.append(x)
        posterior = tabpfn_posterior.append_simulations(torch.cat(theta_per_round, dim=0),
                                                        torch.cat(x_per_round, dim=0))
        if round_num == num_rounds - 1:
            break
        proposal = PosteriorSupport(
            prior,
            posterior,
            obs=observation,
            num_samples_to_estimate_support=num_samples_to_estimate_support,
            batch_size_for_estimate_support=proposal_batch_size,
            allowed_false_negatives=allowed_false_negatives,
            sampling_method=sampling_method,
            max_iter_rejection=max_iter_rejection,
            oversample_sir=oversample_sir,
            log_prob_kwargs={"mode": log_prob_mode},
        )
    return posterior

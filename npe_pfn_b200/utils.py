"""Small stand-ins for the two `sbi` helpers the reference's hot path touches
(`sbi.utils.BoxUniform` at support_posterior.py:5, `sbi.inference.simulate_for_sbi` at tsnpe_pfn.py:86);
`sbi` itself is not a dependency of this package."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
from torch import Tensor
from torch.distributions import Independent, Uniform, constraints


class BoxUniform(Independent):
    """Uniform on a box, event dim 1 (`Independent(Uniform(low, high), 1)`), exposing `.base_dist.low/.high`."""

    def __init__(self, low, high, reinterpreted_batch_ndims: int = 1):
        low = torch.as_tensor(low, dtype=torch.float32)
        high = torch.as_tensor(high, dtype=torch.float32)
        super().__init__(Uniform(low, high, validate_args=False), reinterpreted_batch_ndims, validate_args=False)


def simulate_for_sbi(simulator: Callable, proposal, num_simulations: int, simulation_batch_size: Optional[int] = None,
                     **_kw) -> Tuple[Tensor, Tensor]:
    theta = proposal.sample((num_simulations,))
    bs = simulation_batch_size or num_simulations
    xs = [simulator(theta[i:i + bs]) for i in range(0, num_simulations, bs)]
    return theta, torch.cat(xs, 0)


def box_bounds_of(prior) -> Optional[Tuple[Optional[Tensor], Optional[Tensor]]]:
    """(lo, hi) if the prior's support is an axis-aligned box (or all of R^d: (None, None)); None if the
    support has to be checked through `prior.support.check` / `prior.log_prob` on the host."""
    try:
        sup = prior.support
    except (NotImplementedError, AttributeError):
        return None
    while isinstance(sup, constraints.independent):
        sup = sup.base_constraint
    if isinstance(sup, constraints.interval) and type(sup) is constraints.interval:
        lo = torch.as_tensor(sup.lower_bound, dtype=torch.float32).reshape(-1)
        hi = torch.as_tensor(sup.upper_bound, dtype=torch.float32).reshape(-1)
        return lo, hi
    if sup is constraints.real or type(sup) is type(constraints.real) or sup is constraints.real_vector:
        return None, None
    return None

"""npe_pfn_b200 — B200-native (sm_100a) implementation of NPE-PFN's autoregressive posterior sampling and
log-prob hot path.  Same public names as the reference package (`/root/reference/npe_pfn/__init__.py:1-12`)."""
from .accept_reject_sampler import accept_reject_sample  # noqa: F401
from .npe_pfn import NPE_PFN_Core, TabPFN_Based_NPE_PFN  # noqa: F401
from .support_posterior import PosteriorSupport, get_filtering_method  # noqa: F401
from .tsnpe_pfn import run_tsnpe_pfn  # noqa: F401
from .uncond import TabPFN_Based_Uncond_Estimator  # noqa: F401
from .utils import BoxUniform, simulate_for_sbi  # noqa: F401

__all__ = [
    "TabPFN_Based_NPE_PFN",
    "TabPFN_Based_Uncond_Estimator",
    "NPE_PFN_Core",
    "run_tsnpe_pfn",
    "PosteriorSupport",
    "accept_reject_sample",
    "BoxUniform",
]

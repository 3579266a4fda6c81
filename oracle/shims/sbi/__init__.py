"""Minimal stand-in for the absent `sbi` package (only the names the reference
imports: support_posterior.py:5, tsnpe_pfn.py:5, restricted_prior.py:4).
Test infrastructure only."""
from . import inference, utils  # noqa: F401

"""CPU ORACLE — test infrastructure only.

The five-call estimator protocol of SURVEY.md §8b restated over
`oracle/tabpfn_oracle.py` (model) and `oracle/bar_head.c` (head):
`OracleTabPFNRegressor(**kw).fit(X, y).predict(X, output_type="full",
quantiles=[]) -> {"criterion", "logits"}`; `criterion.sample(logits)`;
`criterion(logits, y)`.  Call sites in the reference:
`/root/reference/npe_pfn/npe_pfn.py:48, 140, 143-146, 149-151`.

Single estimator, identity preprocessing, softmax temperature 0.9 (upstream
default).  PARITY UNPINNED w.r.t. real `tabpfn`.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from npe_pfn_b200.weights import PFNWeights

from . import bar_head
from . import tabpfn_oracle as model

_DEFAULT_WEIGHTS: Optional[PFNWeights] = None


def default_weights() -> PFNWeights:
    global _DEFAULT_WEIGHTS
    if _DEFAULT_WEIGHTS is None:
        _DEFAULT_WEIGHTS = PFNWeights.default()
    return _DEFAULT_WEIGHTS


def y_standardise(y: torch.Tensor):
    """(y_mean, y_std) as fp32 from fp64 accumulation.  Upstream `TabPFNRegressor.fit` standardises the target with
    numpy's POPULATION standard deviation plus 1e-20 (`np.std(y) + 1e-20`, SURVEY.md Appendix A.3 [U]); a constant
    target (std 0, where upstream would divide by 1e-20) keeps unit scale here."""
    yd = y.double()
    mean = yd.mean()
    std = yd.std(unbiased=False) if y.numel() > 1 else torch.tensor(0.0, dtype=torch.float64)
    mean32 = np.float32(mean.item())
    std32 = np.float32(std.item() + 1e-20) if std.item() > 0 else np.float32(0.0)
    if not np.isfinite(std32) or std32 == 0:
        std32 = np.float32(1.0)
    yz = (y.float() - torch.tensor(mean32)) / torch.tensor(std32)
    return float(mean32), float(std32), yz


class OracleCriterion:
    """Bar distribution in ORIGINAL theta units (upstream `renormalized_criterion_`)."""

    def __init__(self, borders_orig: torch.Tensor):
        self.borders = borders_orig

    def sample(self, logits: torch.Tensor, uniforms: Optional[torch.Tensor] = None) -> torch.Tensor:
        u = torch.rand(logits.shape[0]) if uniforms is None else uniforms
        theta, _idx, _u = bar_head.sample(logits, self.borders, uniforms=u)
        return theta

    def icdf_indices(self, logits, uniforms):
        return bar_head.sample(logits, self.borders, uniforms=uniforms)

    def __call__(self, logits: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return bar_head.nll(logits, self.borders, y)


class OracleTabPFNRegressor:
    def __init__(self, weights: Optional[PFNWeights] = None, softmax_temperature: float = 0.9,
                 n_estimators: int = 1, dtype=torch.float32, chunk: int = 2048, **_ignored):
        assert n_estimators == 1, "oracle restates a single estimator with identity preprocessing"
        self.w = weights or default_weights()
        self.temperature = float(softmax_temperature)
        self.dtype = dtype
        self.chunk = chunk
        self.cache = None

    def fit(self, X: torch.Tensor, y: torch.Tensor):
        X = torch.as_tensor(X, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        self.y_mean, self.y_std, yz = y_standardise(y)
        self.cache = model.prefill(self.w, X, yz, dtype=self.dtype)
        self.borders_orig = bar_head.renorm_borders(self.w.borders, self.y_mean, self.y_std)
        return self

    def raw_logits(self, X: torch.Tensor) -> torch.Tensor:
        X = torch.as_tensor(X, dtype=torch.float32)
        return model.forward_test(self.w, self.cache, X, dtype=self.dtype, chunk=self.chunk).float()

    def predict(self, X, output_type: str = "full", quantiles=None):
        assert output_type == "full"
        logits = self.raw_logits(X) / np.float32(self.temperature)
        return {"criterion": OracleCriterion(self.borders_orig), "logits": logits}

"""The five-call estimator protocol of the reference's hot path, served by the sm_100a engine.

The reference reaches its arithmetic through (`/root/reference/npe_pfn/npe_pfn.py`):

    TabPFNRegressor(**kw)                                   :48, :69
    model.fit(joint[:, :dx+d], joint[:, dx+d])              :140, :215, :502
    model.predict(X, output_type="full", quantiles=[])      :143-145, :217-219, :505-507
    pred["criterion"].sample(pred["logits"])                :146, :220
    pred["criterion"](pred["logits"], y)   (negative log density)   :149-151, :226-228, :510-512

`B200TabPFNRegressor` honours exactly that: constructible from kwargs alone, `fit` (re)builds the
context K/V cache in HBM (`pfn_prefill`), `predict` runs the test rows against it
(`pfn_forward_logits`), and the criterion maps to `pfn_head_sample` / `pfn_head_nll`.  Single
estimator, identity preprocessing, softmax temperature 0.9; the 8-member
sklearn-preprocessing ensemble of upstream `tabpfn` is a "next" row.
"""
from __future__ import annotations

from typing import Optional

import torch

from .engine import Engine, get_engine
from .weights import PFNWeights


def draw_seed() -> int:
    """Philox seed taken from torch's global CPU generator (so `torch.manual_seed` pins the draws,
    as it pins `torch.rand` inside upstream's `criterion.sample`)."""
    return int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())


class B200Criterion:
    """Bar distribution in original theta units for one fitted slot."""

    def __init__(self, engine: Engine, slot: int, output_device: str = "cpu", eps: float = 1e-15):
        self.engine = engine
        self.slot = slot
        self.output_device = output_device
        self._counter = 0

    @property
    def borders(self) -> torch.Tensor:
        return self.engine.slot_export(self.slot)["borders"]

    def _out(self, t: torch.Tensor) -> torch.Tensor:
        return t.cpu() if self.output_device == "cpu" else t

    def sample(self, logits: torch.Tensor, uniforms: Optional[torch.Tensor] = None, seed: Optional[int] = None,
               return_bins: bool = False):
        if uniforms is None and seed is None:
            seed = draw_seed()
        theta, bins, _ = self.engine.head_sample(self.slot, logits, uniforms=uniforms, seed=seed or 0,
                                                 return_bins=return_bins)
        if return_bins:
            return self._out(theta), self._out(bins)
        return self._out(theta)

    def __call__(self, logits: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self._out(self.engine.head_nll(self.slot, logits, y.reshape(-1)))


class B200TabPFNRegressor:
    def __init__(self, weights: Optional[PFNWeights] = None, device: Optional[int] = None,
                 softmax_temperature: float = 0.9, n_estimators: int = 1, output_device: str = "cpu",
                 slot: int = 0, engine: Optional[Engine] = None, **_ignored):
        if n_estimators != 1:
            raise NotImplementedError("npe_pfn_b200 implements a single estimator with identity preprocessing "
                                      "(the tabpfn ensemble is listed as a next row in DESIGN.md)")
        self.engine = engine or get_engine(device=device, weights=weights, softmax_temperature=softmax_temperature)
        self.slot = slot
        self.output_device = output_device
        self._fitted = False

    def fit(self, X: torch.Tensor, y: torch.Tensor):
        X = torch.as_tensor(X, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        assert X.ndim == 2 and X.shape[0] == y.shape[0], "fit expects X[N, F], y[N]"
        self.engine.__dict__.get("_slot_tags", {}).pop(self.slot, None)  # the slot no longer holds a posterior's cache
        self.engine.prefill(self.slot, X, y)
        self._fitted = True
        return self

    def predict(self, X, output_type: str = "full", quantiles=None):
        if not self._fitted:
            raise RuntimeError("predict called before fit")
        if output_type != "full":
            raise NotImplementedError("only output_type='full' is on the NPE-PFN hot path")
        X = torch.as_tensor(X, dtype=torch.float32)
        logits = self.engine.forward_logits(self.slot, X)
        return {"criterion": B200Criterion(self.engine, self.slot, self.output_device), "logits": logits}


_CLS_WEIGHTS = None


def default_classifier_weights() -> PFNWeights:
    """Seeded random init of the classifier architecture (no checkpoint offline), shared by engine and oracle."""
    global _CLS_WEIGHTS
    if _CLS_WEIGHTS is None:
        from .weights import classifier_config
        _CLS_WEIGHTS = PFNWeights.random_init(classifier_config())
    return _CLS_WEIGHTS


class B200TabPFNClassifier:
    """`TabPFNClassifier(**kw).fit(X, y in {0..C-1})` / `.predict_proba(X) -> ndarray[m, C]` as the reference's
    density-ratio wrapper uses it (`/root/reference/npe_pfn/npe_pfn.py:610, 661, 697`): the same per-feature
    transformer with the classifier's own weights, class indices fed unscaled to the y-encoder, a 10-way decoder of
    which the first `n_classes` logits are softmaxed (temperature 0.9).  Single estimator, identity preprocessing."""

    def __init__(self, weights: Optional[PFNWeights] = None, device: Optional[int] = None,
                 softmax_temperature: float = 0.9, n_estimators: int = 1, engine: Optional[Engine] = None, **_ignored):
        if n_estimators != 1:
            raise NotImplementedError("npe_pfn_b200 implements a single estimator with identity preprocessing")
        if engine is None:
            engine = get_engine(device=device, weights=weights or default_classifier_weights(),
                                softmax_temperature=softmax_temperature, max_slots=1)
            engine.set_option("standardize_y", 0)
        self.engine = engine
        self.n_classes = 0

    def fit(self, X, y):
        X = torch.as_tensor(X, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        assert X.ndim == 2 and X.shape[0] == y.shape[0], "fit expects X[N, F], y[N]"
        self.n_classes = int(y.max().item()) + 1
        assert 2 <= self.n_classes <= self.engine.cfg.num_buckets
        self.engine.prefill(0, X, y)
        return self

    def predict_proba(self, X):
        if not self.n_classes:
            raise RuntimeError("predict_proba called before fit")
        X = torch.as_tensor(X, dtype=torch.float32)
        logits = self.engine.forward_logits(0, X)[:, :self.n_classes]
        return torch.softmax(logits, dim=-1).cpu().numpy()

"""The five-call estimator protocol of the reference's hot path, served by the sm_100a engine.

The reference reaches its arithmetic through (`/root/reference/npe_pfn/npe_pfn.py`):

    TabPFNRegressor(**kw)                                   :48, :69
    model.fit(joint[:, :dx+d], joint[:, dx+d])              :140, :215, :502
    model.predict(X, output_type="full", quantiles=[])      :143-145, :217-219, :505-507
    pred["criterion"].sample(pred["logits"])                :146, :220
    pred["criterion"](pred["logits"], y)   (negative log density)   :149-151, :226-228, :510-512

`B200TabPFNRegressor` honours exactly that: constructible from kwargs alone, `fit` (re)builds the
context K/V cache in HBM (`pfn_prefill`), `predict` runs the test rows against it
(`pfn_forward_logits`), and the criterion maps to `pfn_head_sample` / `pfn_head_nll`.  Softmax
temperature 0.9.

DEFAULTS DIFFER FROM UPSTREAM in one place, on purpose and loudly: upstream `TabPFNRegressor()` /
`TabPFNClassifier()` average 8 / 4 preprocessing-ensemble members; here `n_estimators` defaults to 1 (a single
estimator with identity preprocessing) and a one-time `UserWarning` says so.  Pass `n_estimators=8` (regressor) /
`n_estimators=4` (classifier) for upstream's ensemble (`ensemble.py`: per-member feature pipelines,
target transforms, re-binning and probability averaging).  Keyword arguments this implementation has no use for
(`fit_mode`, `memory_saving_mode`, `inference_precision`, ...) are accepted and reported once in a warning instead of
being swallowed silently.  Every estimator object owns its fit: the engine slot it uses is tagged with the object's id
and re-built from the stored fit data if another estimator has used the slot in between.
"""
from __future__ import annotations

import itertools
import warnings
from typing import Optional

import torch

from .engine import Engine, get_engine
from .weights import PFNWeights


_EST_UID = itertools.count(1)
_WARNED = set()

#: upstream keyword arguments that are meaningful for upstream's runtime only (accepted, no effect here)
_UPSTREAM_RUNTIME_KWARGS = {"fit_mode", "memory_saving_mode", "inference_precision", "n_jobs", "ignore_pretraining_limits",
                            "model_path", "categorical_features_indices", "average_before_softmax",
                            "inference_config", "balance_probabilities"}


def _warn_once(key: str, msg: str):
    if key not in _WARNED:
        _WARNED.add(key)
        warnings.warn(msg, UserWarning, stacklevel=3)


def _resolve_device(device) -> Optional[int]:
    """upstream accepts device="auto" | "cuda" | "cuda:1" | torch.device; this engine runs on CUDA devices only"""
    if device is None or device == "auto":
        return None
    if isinstance(device, int):
        return device
    d = torch.device(device)
    if d.type != "cuda":
        raise ValueError(f"npe_pfn_b200 runs on CUDA devices (sm_100a) only, got device={device!r}; there is no CPU path")
    return d.index


def _check_kwargs(cls_name: str, n_estimators, upstream_default: int, ignored: dict) -> int:
    if n_estimators is None:
        _warn_once(cls_name + ".n_estimators",
                   f"{cls_name}: n_estimators defaults to 1 here (single estimator, identity preprocessing); upstream's "
                   f"default is {upstream_default} ensemble members. Pass n_estimators={upstream_default} for the "
                   f"upstream behaviour, or n_estimators=1 to silence this warning.")
        n_estimators = 1
    unknown = sorted(k for k in ignored if k not in _UPSTREAM_RUNTIME_KWARGS)
    if unknown:
        _warn_once(cls_name + ":" + ",".join(unknown), f"{cls_name}: ignoring unsupported keyword arguments {unknown}")
    return int(n_estimators)


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
    def __init__(self, weights: Optional[PFNWeights] = None, device=None,
                 softmax_temperature: float = 0.9, n_estimators: Optional[int] = None, output_device: str = "cpu",
                 slot: int = 0, engine: Optional[Engine] = None, random_state: int = 0,
                 fingerprint_feature: bool = True, svd_features: bool = True, **ignored):
        n_estimators = _check_kwargs("B200TabPFNRegressor", n_estimators, 8, ignored)
        if n_estimators < 1:
            raise ValueError("n_estimators must be >= 1")
        self.n_estimators = n_estimators
        self._engine = engine
        self._engine_kw = dict(device=_resolve_device(device), weights=weights, softmax_temperature=softmax_temperature,
                               **({} if n_estimators == 1 else {"max_slots": 16 * n_estimators}))
        self.member_specs = None
        if self.n_estimators > 1:
            from .ensemble import make_members
            self.member_specs = make_members(self.n_estimators, random_state, fingerprint_feature, svd_features)
        self.slot = slot
        self.output_device = output_device
        self._uid = next(_EST_UID)
        self._fit_data = None
        self._ens = None

    @property
    def engine(self) -> Engine:
        """The engine is created on first use (constructing an estimator needs no GPU, using one does: there is no
        CPU path, `get_engine` raises without a CUDA device)."""
        if self._engine is None:
            self._engine = get_engine(**self._engine_kw)
        if self.n_estimators > 1:
            assert self._engine.max_slots >= self.n_estimators, "engine has fewer slots than ensemble members"
        return self._engine

    def _tag(self):
        return ("estimator", self._uid)

    def _build(self):
        """(re)build this estimator's slot(s) from its fit data and tag them as its own"""
        X, y = self._fit_data
        eng = self.engine
        tags = eng.__dict__.setdefault("_slot_tags", {})
        if self.n_estimators == 1:
            eng.prefill(self.slot, X, y)
            tags[self.slot] = self._tag()
        else:
            from .ensemble import EnsembleDim
            slot0 = self.slot * self.n_estimators
            dev = eng.device
            self._ens = EnsembleDim(eng, self.member_specs, slot0).fit(X.to(dev), y.to(dev))
            for e in range(self.n_estimators):
                tags[slot0 + e] = self._tag()

    def _ensure_current(self):
        tags = self.engine.__dict__.setdefault("_slot_tags", {})
        slot0 = self.slot * self.n_estimators
        if any(tags.get(slot0 + e) != self._tag() for e in range(self.n_estimators)):
            self._build()  # another estimator / posterior used the slot since our fit: our state is ours, rebuild it

    def fit(self, X: torch.Tensor, y: torch.Tensor):
        X = torch.as_tensor(X, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        assert X.ndim == 2 and X.shape[0] == y.shape[0], "fit expects X[N, F], y[N]"
        self._fit_data = (X, y)
        self._build()
        return self

    def predict(self, X, output_type: str = "full", quantiles=None):
        if self._fit_data is None:
            raise RuntimeError("predict called before fit")
        if output_type != "full":
            raise NotImplementedError("only output_type='full' is on the NPE-PFN hot path")
        self._ensure_current()
        X = torch.as_tensor(X, dtype=torch.float32)
        if self._ens is not None:
            logits = self._ens.logits(X.to(self.engine.device))
            return {"criterion": B200Criterion(self.engine, self._ens.head_slot, self.output_device), "logits": logits}
        logits = self.engine.forward_logits(self.slot, X)
        return {"criterion": B200Criterion(self.engine, self.slot, self.output_device), "logits": logits}


_CLS_WEIGHTS = None


def default_classifier_weights() -> PFNWeights:
    """Classifier weights shared by the engine and the oracle: `$NPE_PFN_B200_CLASSIFIER_CKPT` if set, otherwise - only
    with the explicit random-init opt-in of `PFNWeights.default` - a seeded random init of the classifier architecture."""
    global _CLS_WEIGHTS
    if _CLS_WEIGHTS is None:
        from .weights import classifier_config
        _CLS_WEIGHTS = PFNWeights.default(classifier_config(), env="NPE_PFN_B200_CLASSIFIER_CKPT")
    return _CLS_WEIGHTS


class B200TabPFNClassifier:
    """`TabPFNClassifier(**kw).fit(X, y in {0..C-1})` / `.predict_proba(X) -> ndarray[m, C]` as the reference's
    density-ratio wrapper uses it (`/root/reference/npe_pfn/npe_pfn.py:610, 661, 697`): the same per-feature
    transformer with the classifier's own weights, class indices fed unscaled to the y-encoder, a 10-way decoder of
    which the first `n_classes` logits are softmaxed (temperature 0.9).  `n_estimators=1`: single estimator, identity
    preprocessing; `n_estimators > 1`: member feature pipelines + per-member class permutation, probabilities averaged
    (`ensemble.py`)."""

    def __init__(self, weights: Optional[PFNWeights] = None, device=None,
                 softmax_temperature: float = 0.9, n_estimators: Optional[int] = None, engine: Optional[Engine] = None,
                 random_state: int = 0, fingerprint_feature: bool = True, svd_features: bool = True, **ignored):
        n_estimators = _check_kwargs("B200TabPFNClassifier", n_estimators, 4, ignored)
        if n_estimators < 1:
            raise ValueError("n_estimators must be >= 1")
        self.n_estimators = n_estimators
        self._engine = engine
        self._engine_kw = dict(device=_resolve_device(device), weights=weights, softmax_temperature=softmax_temperature)
        self.n_classes = 0
        self.random_state = random_state
        self._ens = None
        self._uid = next(_EST_UID)
        self._fit_data = None
        if self.n_estimators > 1:
            from .ensemble import make_classifier_members
            self.member_specs = make_classifier_members(self.n_estimators, random_state, fingerprint_feature, svd_features)

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            kw = dict(self._engine_kw)
            kw["weights"] = kw["weights"] or default_classifier_weights()
            self._engine = get_engine(max_slots=max(self.n_estimators, 4), **kw)
            self._engine.set_option("standardize_y", 0)
        return self._engine

    def _tag(self):
        return ("classifier", self._uid)

    def _build(self):
        X, y = self._fit_data
        eng = self.engine
        tags = eng.__dict__.setdefault("_slot_tags", {})
        if self.n_estimators > 1:
            from .ensemble import EnsembleDim
            dev = eng.device
            self._ens = EnsembleDim(eng, self.member_specs, 0).fit(X.to(dev), y.to(dev), n_classes=self.n_classes,
                                                                   class_seed=self.random_state)
        else:
            eng.prefill(0, X, y)
        for e in range(self.n_estimators):
            tags[e] = self._tag()

    def fit(self, X, y):
        X = torch.as_tensor(X, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        assert X.ndim == 2 and X.shape[0] == y.shape[0], "fit expects X[N, F], y[N]"
        self.n_classes = int(y.max().item()) + 1
        assert 2 <= self.n_classes <= self.engine.cfg.num_buckets
        self._fit_data = (X, y)
        self._build()
        return self

    def predict_proba(self, X, return_device: bool = False):
        """class probabilities [m, n_classes]: a numpy array like upstream's, or (`return_device=True`, not in upstream) a
        CUDA tensor left on the device"""
        if not self.n_classes:
            raise RuntimeError("predict_proba called before fit")
        tags = self.engine.__dict__.setdefault("_slot_tags", {})
        if any(tags.get(e) != self._tag() for e in range(self.n_estimators)):
            self._build()  # another classifier object fitted into the shared engine since: rebuild OUR context
        X = torch.as_tensor(X, dtype=torch.float32)
        if self._ens is not None:
            probs = self._ens.class_probabilities(X.to(self.engine.device))
        else:
            probs = torch.softmax(self.engine.forward_logits(0, X)[:, :self.n_classes], dim=-1)
        return probs if return_device else probs.cpu().numpy()

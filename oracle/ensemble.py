"""CPU ORACLE — test infrastructure only (imported by tests/, smoke() and bench.py's CPU legs, never by the product).

The `n_estimators > 1` path of upstream `tabpfn`'s regressor as the reference reaches it through its default
constructor call (`/root/reference/npe_pfn/npe_pfn.py:48`: `TabPFNRegressor(**regressor_init_kwargs)`, upstream
default `n_estimators=8`), restated from SURVEY.md Appendix A.5:

  per member: constant-feature removal, one of two feature pipelines ("quantile_uni" with the original columns
  appended plus truncated-SVD components; "safepower" = standardise -> Yeo-Johnson -> standardise), a fingerprint
  feature, a seeded feature shuffle, and one of two target transforms (none; Yeo-Johnson on the standardised target);
  the member's bar-distribution borders are mapped back through the inverse target transform, its probabilities
  (softmax of logits / temperature) are re-binned onto the common borders by CDF interpolation under the member's
  piecewise-uniform density, probabilities are averaged over members and `log` gives the returned logits.

Here the sklearn transformers themselves (`QuantileTransformer`, `PowerTransformer`, the estimator classes `tabpfn`
calls) are the oracle for the preprocessing arithmetic; the product (npe_pfn_b200/ensemble.py + the CUDA kernels
`member_transform_kernel` / `ensemble_combine_kernel`) implements the same arithmetic on its own and is compared with
this file.  PARITY UNPINNED w.r.t. real `tabpfn` 2.2.1 (absent offline): member composition, seeds, the fingerprint
hash and the SVD solver follow this file's specification, not upstream's bytes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from npe_pfn_b200.weights import PFNWeights

from . import bar_head
from .estimator import OracleCriterion, default_weights, y_standardise
from . import tabpfn_oracle as model


@dataclass
class MemberSpec:
    x_kind: str          # "quantile" | "safepower" | "none"
    y_kind: str          # "none" | "safepower"
    perm_seed: int
    fingerprint: bool = True
    svd: bool = True


def make_members(n: int, random_state: int = 0, fingerprint: bool = True, svd: bool = True) -> List[MemberSpec]:
    """Member composition: the first ceil(n/2) members take the quantile pipeline, the rest safepower; target
    transforms alternate none / safepower inside each half (so member 0 always has the identity target transform);
    feature-shuffle seeds are a seeded permutation of `start .. start + n`."""
    rng = np.random.default_rng(random_state)
    start = int(rng.integers(0, 1000))
    shifts = rng.permutation(np.arange(start, start + n))
    half = (n + 1) // 2
    out = []
    for i in range(n):
        j = i if i < half else i - half
        out.append(MemberSpec("quantile" if i < half else "safepower", "none" if j % 2 == 0 else "safepower",
                              int(shifts[i]), fingerprint, svd))
    return out


def make_classifier_members(n: int, random_state: int = 0, fingerprint: bool = True, svd: bool = True) -> List[MemberSpec]:
    """Classifier ensemble: quantile pipeline and untouched features alternate; class indices are permuted per member."""
    rng = np.random.default_rng(random_state)
    start = int(rng.integers(0, 1000))
    shifts = rng.permutation(np.arange(start, start + n))
    return [MemberSpec("quantile" if i % 2 == 0 else "none", "none", int(shifts[i]), fingerprint, svd) for i in range(n)]


def class_permutation(e: int, n_classes: int, class_seed: int = 0) -> np.ndarray:
    return np.random.default_rng(class_seed + 7919 * (e + 1)).permutation(n_classes) if e else np.arange(n_classes)


# ---- fingerprint: 64-bit mix of the fp32 bit patterns of a row -> [0, 1) ---------------------------------------
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def fingerprint(X: np.ndarray) -> np.ndarray:
    X32 = np.ascontiguousarray(X, dtype=np.float32)
    X32 = np.where(X32 == 0, np.float32(0.0), X32)  # -0.0 -> +0.0
    bits = X32.view(np.uint32).astype(np.uint64)
    bits = np.where(np.isnan(X32), np.uint64(0x7FC00000), bits)
    h = np.full(X32.shape[0], 0x9E3779B97F4A7C15, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for f in range(X32.shape[1]):
            h = (h ^ bits[:, f]) * _M1
            h ^= h >> np.uint64(31)
        h ^= h >> np.uint64(29)
        h = h * _M2
        h ^= h >> np.uint64(32)
    return ((h >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(np.float32)


# ---- Yeo-Johnson ------------------------------------------------------------------------------------------------
def yeo_johnson(x: np.ndarray, lam: float) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    pos = x >= 0
    if abs(lam) < 1e-12:
        out[pos] = np.log1p(x[pos])
    else:
        out[pos] = (np.power(x[pos] + 1.0, lam) - 1.0) / lam
    if abs(lam - 2.0) < 1e-12:
        out[~pos] = -np.log1p(-x[~pos])
    else:
        out[~pos] = -(np.power(1.0 - x[~pos], 2.0 - lam) - 1.0) / (2.0 - lam)
    return out


def yeo_johnson_inverse(y: np.ndarray, lam: float) -> np.ndarray:
    """NaN where the inverse does not exist (lam < 0: y >= -1/lam; lam > 2: y <= -1/(lam - 2))."""
    y = np.asarray(y, dtype=np.float64)
    out = np.full_like(y, np.nan)
    pos = y >= 0
    with np.errstate(invalid="ignore", over="ignore"):
        if abs(lam) < 1e-12:
            out[pos] = np.expm1(y[pos])
        else:
            base = y[pos] * lam + 1.0
            out[pos] = np.where(base > 0, np.power(np.maximum(base, 1e-300), 1.0 / lam) - 1.0, np.nan)
        if abs(lam - 2.0) < 1e-12:
            out[~pos] = -np.expm1(-y[~pos])
        else:
            base = -(2.0 - lam) * y[~pos] + 1.0
            out[~pos] = np.where(base > 0, 1.0 - np.power(np.maximum(base, 1e-300), 1.0 / (2.0 - lam)), np.nan)
    return out


# ---- one member -------------------------------------------------------------------------------------------------
class OracleMember:
    def __init__(self, spec: MemberSpec):
        self.spec = spec

    def fit_x(self, X: np.ndarray):
        from sklearn.preprocessing import PowerTransformer, QuantileTransformer, StandardScaler
        X = np.asarray(X, dtype=np.float64)
        N = X.shape[0]
        self.keep = np.flatnonzero(X.std(axis=0) > 0) if N > 1 else np.arange(X.shape[1])
        if self.keep.size == 0:
            self.keep = np.arange(X.shape[1])
        Xk = X[:, self.keep]
        if self.spec.x_kind == "quantile":
            nq = max(min(N, max(N // 10, 2)), 2)
            self.qt = QuantileTransformer(n_quantiles=nq, output_distribution="uniform", subsample=10**9)
            base = np.concatenate([self.qt.fit_transform(Xk), Xk], axis=1)
            self.svd_k = 0
            if self.spec.svd and N > 1:
                k = max(1, min(N // 10 + 1, base.shape[1] // 2))
                self.svd_scale = base.std(axis=0)
                self.svd_scale[self.svd_scale == 0] = 1.0
                Z = base / self.svd_scale
                _u, _s, vt = np.linalg.svd(Z, full_matrices=False)
                vt = vt[:k]
                sign = np.sign(vt[np.arange(k), np.abs(vt).argmax(axis=1)])
                sign[sign == 0] = 1.0
                self.svd_vt = vt * sign[:, None]
                self.svd_k = k
        elif self.spec.x_kind == "none":
            pass
        else:
            self.sc_in = StandardScaler().fit(Xk)
            Z = self.sc_in.transform(Xk)
            self.pt = PowerTransformer(method="yeo-johnson", standardize=True).fit(Z)
        return self

    def transform_x(self, X: np.ndarray) -> np.ndarray:
        X = np.asarray(X, dtype=np.float64)
        Xk = X[:, self.keep]
        if self.spec.x_kind == "quantile":
            base = np.concatenate([self.qt.transform(Xk), Xk], axis=1)
            if self.svd_k:
                base = np.concatenate([base, (base / self.svd_scale) @ self.svd_vt.T], axis=1)
        elif self.spec.x_kind == "none":
            base = Xk
        else:
            base = self.pt.transform(self.sc_in.transform(Xk))
        if self.spec.fingerprint:
            base = np.concatenate([base, fingerprint(X)[:, None].astype(np.float64)], axis=1)
        perm = np.random.default_rng(self.spec.perm_seed).permutation(base.shape[1])
        return base[:, perm].astype(np.float32)

    def fit_y(self, yz: np.ndarray):
        """yz: the standardised target.  Returns what the model is fitted on (its own standardisation follows)."""
        if self.spec.y_kind == "none":
            self.lam_y = None
            return np.asarray(yz, dtype=np.float32)
        from sklearn.preprocessing import PowerTransformer
        pt = PowerTransformer(method="yeo-johnson", standardize=False).fit(np.asarray(yz, dtype=np.float64)[:, None])
        self.lam_y = float(pt.lambdas_[0])
        return yeo_johnson(yz, self.lam_y).astype(np.float32)

    def borders_z(self, borders_model: np.ndarray) -> np.ndarray:
        """Member borders in STANDARDISED target units; NaN where the inverse target transform does not exist."""
        b = np.asarray(borders_model, dtype=np.float64)
        return b if self.lam_y is None else yeo_johnson_inverse(b, self.lam_y)


def rebin_tables(member_borders: np.ndarray, common_borders: np.ndarray):
    """For every common border z_k: (idx, frac) with CDF_member(z_k) = C[idx] + p[idx] * frac, C the exclusive prefix
    sum of the member's (valid-bucket) probabilities; `valid[j]` marks buckets whose two borders exist and increase."""
    b = np.asarray(member_borders, dtype=np.float64)
    z = np.asarray(common_borders, dtype=np.float64)
    B = b.shape[0] - 1
    fin = np.isfinite(b)
    valid = fin[:-1] & fin[1:] & (np.diff(np.where(fin, b, 0.0)) > 0)
    vidx = np.flatnonzero(valid)
    assert vidx.size > 0 and np.all(np.diff(vidx) == 1), "valid buckets must form one contiguous range"
    lo, hi = vidx[0], vidx[-1]
    edges = b[lo:hi + 2]
    j = np.clip(np.searchsorted(edges, z, side="right") - 1, 0, hi - lo)
    frac = np.clip((z - edges[j]) / (edges[j + 1] - edges[j]), 0.0, 1.0)
    return (j + lo).astype(np.int32), frac.astype(np.float32), valid


def translate_probs(p: np.ndarray, idx: np.ndarray, frac: np.ndarray, valid: np.ndarray) -> np.ndarray:
    """p [M, B] member probabilities -> probabilities of the common buckets [M, B] (not renormalised)."""
    p = np.where(valid[None, :], np.asarray(p, dtype=np.float64), 0.0)
    p = p / p.sum(axis=1, keepdims=True)
    C = np.concatenate([np.zeros((p.shape[0], 1)), np.cumsum(p, axis=1)], axis=1)[:, :-1]
    cdf = C[:, idx] + p[:, idx] * frac[None, :].astype(np.float64)
    return np.maximum(np.diff(cdf, axis=1), 0.0)


def combine(logits: List[np.ndarray], tables: List[Optional[tuple]]) -> np.ndarray:
    """log(mean_e probs_e) on the common borders; tables[e] = None for a member that already lives on them."""
    acc = None
    for lg, tb in zip(logits, tables):
        lg = np.asarray(lg, dtype=np.float64)
        p = np.exp(lg - lg.max(axis=1, keepdims=True))
        p /= p.sum(axis=1, keepdims=True)
        if tb is not None:
            p = translate_probs(p, *tb)
        acc = p if acc is None else acc + p
    with np.errstate(divide="ignore"):
        return np.log(acc / len(logits)).astype(np.float32)


class OracleEnsembleRegressor:
    """`fit` / `predict(output_type="full")` -> {"criterion", "logits"} with `n_estimators` members."""

    def __init__(self, weights: Optional[PFNWeights] = None, softmax_temperature: float = 0.9, n_estimators: int = 8,
                 random_state: int = 0, fingerprint: bool = True, svd: bool = True, dtype=torch.float32,
                 chunk: int = 2048, **_ignored):
        self.w = weights or default_weights()
        self.temperature = float(softmax_temperature)
        self.specs = make_members(n_estimators, random_state, fingerprint, svd)
        self.dtype, self.chunk = dtype, chunk

    def fit(self, X, y):
        X = torch.as_tensor(X, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        self.y_mean, self.y_std, yz = y_standardise(y)
        self.common_z = self.w.borders.double().numpy()  # member-0 style borders in standardised units
        self.members, self.caches, self.tables = [], [], []
        for spec in self.specs:
            m = OracleMember(spec).fit_x(X.numpy())
            yt = torch.from_numpy(m.fit_y(yz.numpy()))
            mean_t, std_t, ytz = y_standardise(yt)
            Xt = torch.from_numpy(m.transform_x(X.numpy()))
            self.caches.append(model.prefill(self.w, Xt, ytz, dtype=self.dtype))
            if spec.y_kind == "none":
                # the model standardises its target itself; (mean_t, std_t) = (0, 1) up to rounding -> same borders
                self.tables.append(None)
            else:
                bz = m.borders_z(self.w.borders.double().numpy() * std_t + mean_t)
                self.tables.append(rebin_tables(bz, self.common_z))
            self.members.append(m)
        self.borders_orig = bar_head.renorm_borders(self.w.borders, self.y_mean, self.y_std)
        return self

    def member_logits(self, X) -> List[np.ndarray]:
        X = torch.as_tensor(X, dtype=torch.float32)
        out = []
        for m, cache in zip(self.members, self.caches):
            Xt = torch.from_numpy(m.transform_x(X.numpy()))
            lg = model.forward_test(self.w, cache, Xt, dtype=self.dtype, chunk=self.chunk).float()
            out.append((lg / np.float32(self.temperature)).numpy())
        return out

    def predict(self, X, output_type: str = "full", quantiles=None):
        assert output_type == "full"
        logits = torch.from_numpy(combine(self.member_logits(X), self.tables))
        return {"criterion": OracleCriterion(self.borders_orig), "logits": logits}


class OracleEnsembleClassifier:
    """`fit(X, y in {0..C-1})` / `predict_proba(X)` with `n_estimators` members (mean of member probabilities)."""

    def __init__(self, weights=None, softmax_temperature: float = 0.9, n_estimators: int = 4, random_state: int = 0,
                 fingerprint: bool = True, svd: bool = True, chunk: int = 2048, **_ignored):
        if weights is None:
            from npe_pfn_b200.estimator import default_classifier_weights
            weights = default_classifier_weights()
        self.w = weights
        self.temperature = float(softmax_temperature)
        self.specs = make_classifier_members(n_estimators, random_state, fingerprint, svd)
        self.random_state, self.chunk = random_state, chunk

    def fit(self, X, y):
        X = torch.as_tensor(X, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        self.n_classes = int(y.max().item()) + 1
        self.members, self.caches, self.perms = [], [], []
        for e, spec in enumerate(self.specs):
            m = OracleMember(spec).fit_x(X.numpy())
            cp = class_permutation(e, self.n_classes, self.random_state)
            yt = torch.from_numpy(cp[y.long().numpy()].astype(np.float32))
            self.caches.append(model.prefill(self.w, torch.from_numpy(m.transform_x(X.numpy())), yt))
            self.members.append(m)
            self.perms.append(cp)
        return self

    def predict_proba(self, X) -> np.ndarray:
        X = torch.as_tensor(X, dtype=torch.float32)
        acc = 0.0
        for m, cache, cp in zip(self.members, self.caches, self.perms):
            Xt = torch.from_numpy(m.transform_x(X.numpy()))
            lg = model.forward_test(self.w, cache, Xt, chunk=self.chunk).float() / np.float32(self.temperature)
            acc = acc + torch.softmax(lg[:, :self.n_classes], dim=-1).numpy()[:, cp]
        return acc / len(self.members)

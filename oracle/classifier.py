"""CPU ORACLE — test infrastructure only.

`TabPFNClassifier` restated over `oracle/tabpfn_oracle.py` for the density-ratio log-prob of the reference
(`/root/reference/npe_pfn/npe_pfn.py:603-704`: `fit(X[2n, d], y in {0,1})`, `predict_proba(X) -> ndarray[m, 2]`).
Same transformer as the regressor with the classifier's own (seeded random) weights, class indices fed unscaled
to the y-encoder, 10-way decoder, softmax over the first n_classes logits at temperature 0.9 (SURVEY.md Appendix
A.1 / A.4).  PARITY UNPINNED w.r.t. real `tabpfn` (absent offline).
"""
from __future__ import annotations

import numpy as np
import torch

from . import tabpfn_oracle as model


class OracleTabPFNClassifier:
    def __init__(self, weights=None, softmax_temperature: float = 0.9, n_estimators: int = 1, chunk: int = 2048,
                 **_ignored):
        assert n_estimators == 1
        if weights is None:
            from npe_pfn_b200.estimator import default_classifier_weights
            weights = default_classifier_weights()
        self.w = weights
        self.temperature = float(softmax_temperature)
        self.chunk = chunk
        self.cache = None

    def fit(self, X, y):
        X = torch.as_tensor(X, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        self.n_classes = int(y.max().item()) + 1
        self.cache = model.prefill(self.w, X, y)
        return self

    def predict_proba(self, X) -> np.ndarray:
        X = torch.as_tensor(X, dtype=torch.float32)
        logits = model.forward_test(self.w, self.cache, X, chunk=self.chunk).float() / np.float32(self.temperature)
        return torch.softmax(logits[:, :self.n_classes], dim=-1).numpy()

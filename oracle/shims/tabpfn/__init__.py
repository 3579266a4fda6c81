"""Shim so the UNMODIFIED reference package (`/root/reference/npe_pfn`, which does
`from tabpfn import TabPFNClassifier, TabPFNRegressor`, npe_pfn.py:8) can be driven
over the CPU oracle in this container.  Test infrastructure only."""
from oracle.estimator import OracleTabPFNRegressor as TabPFNRegressor  # noqa: F401


class TabPFNClassifier:  # ratio-based log-prob is a "next" row
    def __init__(self, **kw):
        pass

    def fit(self, X, y):
        raise NotImplementedError("classifier head is outside the round-1 oracle")

    def predict_proba(self, X):
        raise NotImplementedError("classifier head is outside the round-1 oracle")

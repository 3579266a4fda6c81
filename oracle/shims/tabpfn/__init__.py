"""Shim so the UNMODIFIED reference package (`/root/reference/npe_pfn`, which does
`from tabpfn import TabPFNClassifier, TabPFNRegressor`, npe_pfn.py:8) can be driven over the CPU oracle in this
container.  Test infrastructure only.  `n_estimators > 1` selects the ensemble oracle (oracle/ensemble.py)."""
from oracle.classifier import OracleTabPFNClassifier
from oracle.estimator import OracleTabPFNRegressor


def TabPFNRegressor(**kw):
    if int(kw.get("n_estimators", 1)) > 1:
        from oracle.ensemble import OracleEnsembleRegressor
        return OracleEnsembleRegressor(**kw)
    return OracleTabPFNRegressor(**kw)


def TabPFNClassifier(**kw):
    if int(kw.get("n_estimators", 1)) > 1:
        from oracle.ensemble import OracleEnsembleClassifier
        return OracleEnsembleClassifier(**kw)
    return OracleTabPFNClassifier(**kw)

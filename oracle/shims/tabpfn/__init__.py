"""Shim so the UNMODIFIED reference package (`/root/reference/npe_pfn`, which does
`from tabpfn import TabPFNClassifier, TabPFNRegressor`, npe_pfn.py:8) can be driven over the CPU oracle in this
container.  Test infrastructure only."""
from oracle.classifier import OracleTabPFNClassifier as TabPFNClassifier  # noqa: F401
from oracle.estimator import OracleTabPFNRegressor as TabPFNRegressor  # noqa: F401

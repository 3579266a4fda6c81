"""Generates tests/golden/ensemble_golden.pt from the sklearn-backed ensemble oracle (oracle/ensemble.py): frozen
per-member transformed test features, target-transform lambdas and combined logits for one small case.
`-m "not gpu"` tests check the oracle still reproduces them (guards sklearn / numpy drift), `-m gpu` tests compare the
CUDA ensemble path with them.

    python tests/golden/make_golden_ensemble.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from npe_pfn_b200.weights import PFNWeights  # noqa: E402
from oracle.ensemble import OracleEnsembleRegressor  # noqa: E402


def case():
    g = np.random.default_rng(2024)
    N, M, F = 64, 8, 3
    X = g.normal(size=(N + M, F))
    X[:, 1] = np.exp(X[:, 1])
    X[:, 2] = np.round(X[:, 2], 1)
    y = 0.5 * X[:, 0] + 0.3 * g.normal(size=N + M) + (X[:, 2] > 0) * 1.5
    X = X.astype(np.float32)
    return torch.from_numpy(X[:N]), torch.from_numpy(y[:N].astype(np.float32)), torch.from_numpy(X[N:])


def main():
    w = PFNWeights.random_init()
    Xc, yc, Xt = case()
    m = OracleEnsembleRegressor(weights=w, n_estimators=4, random_state=0).fit(Xc, yc)
    pd = m.predict(Xt)
    cols = torch.arange(0, w.cfg.num_buckets, 53)
    out = {
        "Xc": Xc, "yc": yc, "Xt": Xt, "cols": cols,
        "features": [torch.from_numpy(mm.transform_x(Xt.numpy())) for mm in m.members],
        "lam_y": [mm.lam_y for mm in m.members],
        "logits_cols": pd["logits"][:, cols].clone(), "lse": torch.logsumexp(pd["logits"], -1),
        "specs": [(s.x_kind, s.y_kind, s.perm_seed) for s in m.specs],
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ensemble_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""Generates tests/golden/oracle_golden.pt from the fp32 CPU oracle (seeded weights, seeded inputs).

The reference holds no golden vectors for this path and its arithmetic dependency (`tabpfn==2.2.1`) is not
importable offline, so these vectors freeze OUR restatement: `-m "not gpu"` tests check the
oracle still reproduces them, `-m gpu` tests compare the CUDA path against them.

    python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from npe_pfn_b200.weights import PFNWeights  # noqa: E402
from oracle.estimator import OracleTabPFNRegressor  # noqa: E402


def main():
    w = PFNWeights.random_init()
    cases = []
    for F in (1, 2, 3, 5, 10):
        g = torch.Generator().manual_seed(1000 + F)
        N, M = 16 if F < 10 else 40, 8
        Xc = torch.randn(N, F, generator=g)
        yc = 0.7 * Xc[:, 0] + 0.3 * torch.randn(N, generator=g) + 1.0
        Xt = torch.randn(M, F, generator=g)
        if F == 3:
            Xc[2, 1] = float("nan")
            Xt[1, 2] = float("inf")
        u = torch.rand(M, generator=g)
        y = torch.randn(M, generator=g) * 2 + 1.0
        m = OracleTabPFNRegressor(weights=w)
        m.fit(Xc, yc)
        pd = m.predict(Xt)
        logits = pd["logits"]
        cols = torch.arange(0, w.cfg.num_buckets, 97)
        theta, idx, _ = pd["criterion"].icdf_indices(logits, u)
        cases.append({
            "F": F, "Xc": Xc, "yc": yc, "Xt": Xt, "u": u, "y": y, "cols": cols,
            "logits_cols": logits[:, cols].clone(), "lse": torch.logsumexp(logits, -1),
            "logits_full0": logits[:1].clone(), "idx0": idx[0].clone(), "theta0": theta[0].clone(),
            "idx": idx.clone(), "theta": theta.clone(), "nll": pd["criterion"](logits, y),
            "y_mean": m.y_mean, "y_std": m.y_std,
        })
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.pt")
    torch.save({"cases": cases, "weights_seed": w.cfg.seed}, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Runs the BASELINE.json configuration shapes through the public API on one B200 and prints one JSON line each
(these are functional / timing checks of the other configs; the bench line is `bench.py`, configs[1]).

    python tools/run_configs.py [--scale 1.0]
"""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from npe_pfn_b200 import BoxUniform, NPE_PFN_Core, TabPFN_Based_NPE_PFN, run_tsnpe_pfn  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def two_moons_simulator(theta):
    """demo.ipynb:50-75 of the reference"""
    a = torch.rand(theta.shape[0]) * math.pi - math.pi / 2
    r = 0.1 + 0.01 * torch.randn(theta.shape[0])
    p = torch.stack([r * torch.cos(a) + 0.25, r * torch.sin(a)], 1)
    q = torch.stack([-(theta[:, 0] + theta[:, 1]).abs() / math.sqrt(2), (-theta[:, 0] + theta[:, 1]) / math.sqrt(2)], 1)
    return p + q


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    torch.manual_seed(42)
    out = []

    # 1. two_moons: 2-D theta, 1k simulations, 10k samples via .sample with the demo's plain Uniform prior
    prior = torch.distributions.Uniform(-torch.ones(2), torch.ones(2))
    theta = prior.sample((1000,))
    x = two_moons_simulator(theta)
    xo = two_moons_simulator(0.5 * torch.ones(1, 2))
    post = TabPFN_Based_NPE_PFN(prior=prior).append_simulations(theta, x)
    post.sample((100,), xo)
    s, dt = timed(lambda: post.sample((10_000,), xo))
    assert s.shape == (10_000, 2) and bool(((s >= -1) & (s <= 1)).all())
    out.append({"config": "two_moons", "samples": 10_000, "seconds": dt, "samples_per_s": 10_000 / dt,
                "acceptance": post.last_acceptance_rate})

    # 2b. gaussian_linear log_prob (sampling is bench.py): 10k context, log_prob of S rows, both modes
    S = int(20_000 * a.scale)
    g = torch.Generator().manual_seed(1)
    theta = math.sqrt(0.1) * torch.randn(10_000, 10, generator=g)
    x = theta + math.sqrt(0.1) * torch.randn(10_000, 10, generator=g)
    prior = torch.distributions.MultivariateNormal(torch.zeros(10), 0.1 * torch.eye(10))
    post = NPE_PFN_Core(prior=prior).append_simulations(theta, x)
    draws = post.sample((S,), x[:1], max_sampling_batch_size=S)
    lp, dt = timed(lambda: post.log_prob(draws, x[:1], max_sampling_batch_size=S))
    assert torch.isfinite(lp).all()
    out.append({"config": "gaussian_linear log_prob autoregressive", "rows": S, "seconds": dt, "rows_per_s": S / dt})
    lp2, dt = timed(lambda: post.log_prob(draws[:5000], x[:1], mode="ratio_based"))
    assert torch.isfinite(lp2).all()
    out.append({"config": "gaussian_linear log_prob ratio_based (fit on 5000+5000, 5000 rows)", "seconds": dt})

    # 3. slcp shape: 5-D theta / 8-D x, TSNPE rounds with the truncated prior (scaled: 2 rounds x 1000 simulations)
    prior = BoxUniform(-3 * torch.ones(5), 3 * torch.ones(5))

    def slcp(theta):
        m = theta[:, :2]
        s1, s2, rho = theta[:, 2] ** 2, theta[:, 3] ** 2, torch.tanh(theta[:, 4])
        eps = torch.randn(theta.shape[0], 4, 2)
        xs = torch.stack([m[:, None, 0] + s1[:, None] * eps[..., 0],
                          m[:, None, 1] + s2[:, None] * (rho[:, None] * eps[..., 0]
                                                        + torch.sqrt(1 - rho[:, None] ** 2) * eps[..., 1])], -1)
        return xs.reshape(theta.shape[0], 8)

    xo = slcp(prior.sample((1,)))
    post, dt = timed(lambda: run_tsnpe_pfn(slcp, prior, xo, num_simulations=int(2000 * a.scale), num_rounds=2,
                                           proposal_batch_size=1000, num_samples_to_estimate_support=2000,
                                           allowed_false_negatives=1e-3, log_prob_mode="autoregressive"))
    s, dt2 = timed(lambda: post.sample((int(50_000 * a.scale),), xo, max_sampling_batch_size=50_000))
    assert bool(prior.support.check(s).all())
    out.append({"config": "slcp TSNPE (2 rounds) + sample", "tsnpe_seconds": dt, "samples": s.shape[0], "seconds": dt2,
                "samples_per_s": s.shape[0] / dt2, "acceptance": post.last_acceptance_rate})

    # 4. bernoulli_glm shape: 10-D / 10-D, sample_batched over many observations
    g = torch.Generator().manual_seed(3)
    V = torch.randn(100, 10, generator=g)
    theta = math.sqrt(2.0) * torch.randn(10_000, 10, generator=g)
    z = torch.bernoulli(torch.sigmoid(theta @ V.T), generator=g)
    x = z @ V
    post = NPE_PFN_Core(prior=None).append_simulations(theta, x)
    n_obs, n = int(50 * a.scale), 1000
    s, dt = timed(lambda: post.sample_batched(x[:n_obs], (n,)))
    assert s.shape == (n_obs, n, 10) and torch.isfinite(s).all()
    out.append({"config": "bernoulli_glm sample_batched", "observations": n_obs, "samples_per_obs": n, "seconds": dt,
                "samples_per_s": n_obs * n / dt})

    # 5. sweep point: N = 50k context (beyond the 10k default), 2-D theta, log_prob of M rows (autoregressive)
    g = torch.Generator().manual_seed(5)
    N5, M5 = int(50_000 * min(a.scale, 1.0)), int(20_000 * a.scale)
    theta = torch.randn(N5, 2, generator=g)
    x = theta @ torch.randn(2, 10, generator=g) + 0.1 * torch.randn(N5, 10, generator=g) + 1.0
    post = NPE_PFN_Core(prior=None).append_simulations(theta, x)
    _, dt_pre = timed(lambda: post.prefill(x[:1]))
    th = torch.randn(M5, 2, generator=g)
    lp, dt = timed(lambda: post.log_prob(th, x[:1], max_sampling_batch_size=M5))
    assert lp.shape == (M5,) and torch.isfinite(lp).all()
    info = post.engine.slot_info(1)
    out.append({"config": "sweep N=50k context", "context_rows": N5, "prefill_seconds": dt_pre, "rows": M5, "seconds": dt,
                "rows_per_s": M5 / dt, "kv_cache_MB_dim1": info["kv_bytes"] / 1e6,
                "gpu_mem_GB": torch.cuda.max_memory_allocated() / 1e9})

    # 6. upstream's default 8-member preprocessing ensemble at the gaussian_linear shape (10k simulations):
    #     fit of all 8 x 10 member contexts, then S draws
    g = torch.Generator().manual_seed(1)
    theta = math.sqrt(0.1) * torch.randn(10_000, 10, generator=g)
    x = theta + math.sqrt(0.1) * torch.randn(10_000, 10, generator=g)
    prior = torch.distributions.MultivariateNormal(torch.zeros(10), 0.1 * torch.eye(10))
    post = NPE_PFN_Core(prior=prior, regressor_init_kwargs={"n_estimators": 8}).append_simulations(theta, x)
    _, dt_fit = timed(lambda: post.prefill(x[:1]))
    S6 = int(10_000 * a.scale)
    s, dt = timed(lambda: post.sample((S6,), x[:1], max_sampling_batch_size=S6))
    assert s.shape == (S6, 10) and torch.isfinite(s).all()
    out.append({"config": "gaussian_linear, n_estimators=8 ensemble", "fit_seconds_all_dims": dt_fit, "samples": S6,
                "seconds": dt, "samples_per_s": S6 / dt, "gpu_mem_GB": torch.cuda.max_memory_allocated() / 1e9})

    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE.json config 5 ("log_prob sweep: context N = 1k-50k x M test rows, roofline scaling") on one B200 through the
public API: autoregressive `log_prob` of M rows against N simulations for three problem shapes, one JSON line per
point with rows/s and the achieved algorithmic TFLOP/s (SURVEY.md 8d FLOP model, K/V caches prefilled once).

    python tools/sweep_logprob.py [--rows 100000] [--ratio]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from npe_pfn_b200 import NPE_PFN_Core  # noqa: E402

E, HID, L, BUCKETS = 192, 768, 12, 5000


def flops_row(T, N):
    return L * T * (28 * E * E + 4 * T * E + 4 * N * E) + 2 * E * HID + 2 * HID * BUCKETS


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--ratio", action="store_true", help="also time the ratio-based mode (classifier on 2 x 5000 rows)")
    a = ap.parse_args()
    torch.manual_seed(0)
    for dx, dth in ((2, 2), (5, 5), (10, 10)):
        for N in (1_000, 5_000, 10_000, 50_000):
            g = torch.Generator().manual_seed(N + dx)
            theta = torch.randn(N, dth, generator=g)
            x = theta @ torch.randn(dth, dx, generator=g) + 0.1 * torch.randn(N, dx, generator=g) + 1.0
            prior = torch.distributions.MultivariateNormal(torch.zeros(dth), torch.eye(dth))
            post = NPE_PFN_Core(prior=prior).append_simulations(theta, x)
            M = a.rows if N <= 10_000 else a.rows // 2
            th = torch.randn(M, dth, generator=g)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            post.prefill(x[:1])
            torch.cuda.synchronize()
            t_pre = time.perf_counter() - t0
            post.log_prob(th[:1000], x[:1], max_sampling_batch_size=M)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lp = post.log_prob(th, x[:1], max_sampling_batch_size=M)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert lp.shape == (M,) and torch.isfinite(lp).all()
            # dimension 0 is one forward row for all M targets (identical features)
            fl = sum(flops_row((dx + d + 1) // 2 + 1, N) for d in range(1, dth)) * M + flops_row((dx + 1) // 2 + 1, N)
            out = {"dx": dx, "dtheta": dth, "context_rows": N, "rows": M, "prefill_s": round(t_pre, 4), "seconds": round(dt, 4),
                   "rows_per_s": round(M / dt, 1), "tflops": round(fl / dt / 1e12, 1)}
            if a.ratio and N <= 10_000:
                post.log_prob(th[:1000], x[:1], mode="ratio_based")
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                lp2 = post.log_prob(th, x[:1], mode="ratio_based", max_sampling_batch_size=M)
                torch.cuda.synchronize()
                out["ratio_rows_per_s"] = round(M / (time.perf_counter() - t0), 1)
                assert torch.isfinite(lp2).all()
            print(json.dumps(out), flush=True)
            post.invalidate_cache()
            del post
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

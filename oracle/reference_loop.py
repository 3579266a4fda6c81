"""CPU ORACLE — test infrastructure only.

Restatement of the reference's autoregressive hot loops around the five-call
estimator protocol, used (a) as the parity checker for the CUDA path and (b) as
the `cpu_baseline` / `--impl reference` leg of bench.py on boxes where
`/root/reference` does not exist.

  sample_loop   <- NPE_PFN_Core._sample                /root/reference/npe_pfn/npe_pfn.py:111-169
  logprob_loop  <- NPE_PFN_Core._autoregressive_log_prob   /root/reference/npe_pfn/npe_pfn.py:462-524
  sample_batched_loop <- NPE_PFN_Core._sample_batched       /root/reference/npe_pfn/npe_pfn.py:171-251

Like the reference, every call re-fits the estimator once per dimension.
`uniforms[M, dim_theta]` may be injected so results are comparable draw by draw.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from .estimator import OracleTabPFNRegressor


def sample_loop(model: OracleTabPFNRegressor, x_ctx, theta_ctx, x_obs, M: int, *, with_log_prob=False,
                eps=1e-15, uniforms: Optional[torch.Tensor] = None, return_bins=False):
    samples_batch = x_obs.reshape(1, -1).repeat(M, 1) if x_obs.shape[0] == 1 or x_obs.ndim == 1 else x_obs
    joint = torch.cat([x_ctx, theta_ctx], dim=1)
    dx, dth = x_ctx.shape[1], theta_ctx.shape[1]
    lp = torch.zeros(samples_batch.shape[0]) if with_log_prob else None
    bins = []
    for d in range(dth):
        model.fit(joint[:, :dx + d], joint[:, dx + d])
        pd = model.predict(samples_batch, output_type="full", quantiles=[])
        crit, logits = pd["criterion"], pd["logits"]
        if uniforms is None:
            th = crit.sample(logits)
        else:
            th, idx, _ = crit.icdf_indices(logits, uniforms[:, d].contiguous())
            bins.append(idx)
        if with_log_prob:
            dlp = -crit(logits, th)
            dlp = torch.where(dlp == float("-inf"), torch.tensor(math.log(eps)), dlp)
            lp += dlp
        samples_batch = torch.cat([samples_batch, th[:, None]], dim=1)
    out = (samples_batch[:, dx:], lp)
    if return_bins:
        return out + (torch.stack(bins, 1) if bins else None,)
    return out


def logprob_loop(model: OracleTabPFNRegressor, x_ctx, theta_ctx, x_obs, theta, *, eps=1e-15):
    m = theta.shape[0]
    test_joint = torch.cat([x_obs.reshape(1, -1).repeat(m, 1), theta], dim=1)
    joint = torch.cat([x_ctx, theta_ctx], dim=1)
    dx, dth = x_ctx.shape[1], theta_ctx.shape[1]
    lp = torch.zeros(m)
    for d in range(dth):
        model.fit(joint[:, :dx + d], joint[:, dx + d])
        pd = model.predict(test_joint[:, :dx + d], output_type="full", quantiles=[])
        dlp = -pd["criterion"](pd["logits"], test_joint[:, dx + d].contiguous())
        dlp = torch.where(dlp == float("-inf"), torch.tensor(math.log(eps)), dlp)
        lp += dlp
    return lp


def sample_batched_loop(model: OracleTabPFNRegressor, x_ctx, theta_ctx, x_obs, n: int, *, with_log_prob=False,
                        eps=1e-15, uniforms: Optional[torch.Tensor] = None, return_bins=False):
    """Many observations against one shared, unfiltered context: test row o * n + i is draw i of observation o
    (`repeat_interleave`); -> theta [num_obs, n, dim_theta], log_probs [num_obs, n] | None (, bins)."""
    num_obs = x_obs.shape[0]
    out = sample_loop(model, x_ctx, theta_ctx, x_obs.repeat_interleave(n, dim=0), num_obs * n,
                      with_log_prob=with_log_prob, eps=eps, uniforms=uniforms, return_bins=return_bins)
    theta = out[0].reshape(num_obs, n, -1)
    lp = out[1].reshape(num_obs, n) if out[1] is not None else None
    if return_bins:
        return theta, lp, (out[2].reshape(num_obs, n, -1) if out[2] is not None else None)
    return theta, lp

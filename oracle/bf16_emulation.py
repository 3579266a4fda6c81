"""CPU ORACLE — test infrastructure only (never imported by the product path).

The same restatement as `oracle/tabpfn_oracle.py`, with bf16 roundings inserted EXACTLY where the sm_100a
kernels round (`npe_pfn_b200/csrc/`), so that a comparison against it isolates kernel defects from precision:

  * weights of every projection: bf16 (`f32_to_bf16_kernel`); the item-attention query rows are scaled by
    log2(e)/sqrt(32) BEFORE rounding (`scale_item_q_kernel`), scores therefore live in the log2 domain;
  * the residual stream is fp32 (`xf`), every GEMM reads its bf16 copy (`xb`); encoder weights stay fp32
    (`encode_kernel`);
  * GEMM outputs that feed another kernel are bf16: feature-attention qkv, item-attention q / k / v (the K/V cache),
    both attention outputs, the MLP hidden activation and the decoder hidden activation; accumulation is fp32;
  * attention probabilities are rounded to bf16 before the P V contraction, the row sum uses the unrounded fp32
    values (`feature_attn_mma_kernel`, `attn_tc_kernel`); the item-attention reference maximum is integer valued, so
    the rounded mantissa of P does not depend on it;
  * out-projection + residual + LayerNorm run in fp32 on the fp32 accumulator (`gemm_tc_epilogue<EPI_RESID_LN>`).

What is NOT reproduced (so agreement is close, not bitwise): the order of fp32 accumulation inside the tensor core,
`ex2.approx` / the degree-3 polynomial of the FMA-pipe exponentials (<= 7.5e-5 relative before the bf16 rounding of P)
and the 3-term erf of the fused GELU (|gelu error| <= 1.3e-5 |x|).  Any of these can flip a bf16 rounding, and a
flipped rounding propagates through the remaining layers; the measured residual is stated in
`tests/test_gpu_parity.py::test_bf16_emulated_oracle`.

PARITY UNPINNED w.r.t. real `tabpfn` (see `oracle/tabpfn_oracle.py`).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import tabpfn_oracle as base

ITEM_SCALE_LOG2 = float(torch.tensor(0.17677669529663687, dtype=torch.float32) *
                        torch.tensor(1.4426950408889634, dtype=torch.float32))


def r(x: torch.Tensor) -> torch.Tensor:
    """round to bf16 (nearest even), keep fp32 storage"""
    return x.to(torch.bfloat16).to(torch.float32)


def _feature_attn(w, l, x):
    cfg = w.cfg
    R, T, E = x.shape
    H, dh = cfg.nhead, cfg.head_dim
    qkv = r(r(x) @ r(w.feat_wqkv[l]).T)
    q, k, v = [a.reshape(R, T, H, dh).transpose(1, 2) for a in qkv.split(E, dim=-1)]  # [R, H, T, dh]
    sc = torch.tensor(0.17677669529663687, dtype=torch.float32) * torch.tensor(1.4426950408889634, dtype=torch.float32)
    s = q @ k.transpose(-1, -2)                         # raw scores, fp32
    mx = s.max(dim=-1, keepdim=True).values * sc
    p = torch.exp2(s * sc - mx)
    o = (r(p) @ v) / p.sum(-1, keepdim=True)
    o = r(o).transpose(1, 2).reshape(R, T, E)
    return base._ln(x + o @ r(w.feat_wo[l]).T, cfg.ln_eps)


def _mlp(w, l, x):
    h = r(F.gelu(r(x) @ r(w.mlp_w1[l]).T))
    return base._ln(x + h @ r(w.mlp_w2[l]).T, w.cfg.ln_eps)


def _item_attn(q, k, v, max_elems: int = 1 << 27):
    """q [B, Lq, dh] pre-scaled (log2 domain), k / v [B, Lk, dh] or [Lk, dh], all bf16-valued -> bf16-valued output.
    Processed in slices of the batch dimension so the score matrix stays below `max_elems` elements."""
    B, Lq = q.shape[0], q.shape[1]
    Lk = k.shape[-2]
    step = max(1, max_elems // max(Lq * Lk, 1))
    outs = []
    for b0 in range(0, B, step):
        qb = q[b0:b0 + step]
        kb = k[b0:b0 + step] if k.ndim == 3 and k.shape[0] == B else k
        vb = v[b0:b0 + step] if v.ndim == 3 and v.shape[0] == B else v
        s = qb @ kb.transpose(-1, -2)
        m = torch.round(s.max(dim=-1, keepdim=True).values)  # integer reference: P's mantissa is independent of it
        p = torch.exp2(s - m)
        outs.append(r((r(p) @ vb) / p.sum(-1, keepdim=True)))
    return torch.cat(outs, 0)


def _wq_scaled(w, l, E):
    return r(w.item_wqkv[l][:E] * ITEM_SCALE_LOG2)


def prefill(w, Xc: torch.Tensor, yc: torch.Tensor) -> base.ContextCache:
    cfg = w.cfg
    N, Fdim = Xc.shape
    G = (Fdim + 1) // 2
    T = G + 1
    E, H, dh = cfg.emsize, cfg.nhead, cfg.head_dim
    cache = base.ContextCache()
    cache.stats = st = base.EncoderStats(Xc, yc, G)
    cache.T, cache.N = T, N
    x = torch.cat([base.encode_x(w, st, Xc, torch.float32), base.encode_y_ctx(w, yc, torch.float32)[:, None, :]], dim=1)
    for l in range(cfg.nlayers):
        x = _feature_attn(w, l, x)
        xb = r(x)
        q = r(xb @ _wq_scaled(w, l, E).T)
        kv = r(xb @ r(w.item_wqkv[l][E:]).T)
        k, v = kv.split(E, dim=-1)
        q = q.reshape(N, T, H, dh).permute(1, 2, 0, 3)
        k = k.reshape(N, T, H, dh).permute(1, 2, 0, 3)
        v = v.reshape(N, T, H, dh).permute(1, 2, 0, 3)
        cache.k0.append(k[:, 0].contiguous())
        cache.v0.append(v[:, 0].contiguous())
        if l == cfg.nlayers - 1:
            break  # the context's final states are never read (the kernels stop here too)
        o = _item_attn(q.reshape(T * H, N, dh), k.reshape(T * H, N, dh), v.reshape(T * H, N, dh))
        o = o.reshape(T, H, N, dh).permute(2, 0, 1, 3).reshape(N, T, E)
        x = base._ln(x + o @ r(w.item_wo[l]).T, cfg.ln_eps)
        x = _mlp(w, l, x)
    return cache


def forward_test(w, cache: base.ContextCache, Xt: torch.Tensor, chunk: int = 2048) -> torch.Tensor:
    """raw decoder logits [M, num_buckets] (before the softmax temperature), bf16 roundings as in the kernels"""
    cfg = w.cfg
    E, H, dh = cfg.emsize, cfg.nhead, cfg.head_dim
    outs = []
    for s0 in range(0, Xt.shape[0], chunk):
        X = Xt[s0:s0 + chunk]
        M, T = X.shape[0], cache.T
        x = torch.cat([base.encode_x(w, cache.stats, X, torch.float32),
                       base.encode_y_test(w, cache.stats, M, torch.float32)[:, None, :]], dim=1)
        for l in range(cfg.nlayers):
            x = _feature_attn(w, l, x)
            q = r(r(x) @ _wq_scaled(w, l, E).T)
            q = q.reshape(M, T, H, dh).permute(1, 0, 2, 3).reshape(T, M * H, dh)
            o = _item_attn(q, cache.k0[l], cache.v0[l])  # k0 / v0 [T, N, dh]: one key set per column, shared by all heads
            o = o.reshape(T, M, H, dh).permute(1, 0, 2, 3).reshape(M, T, E)
            x = base._ln(x + o @ r(w.item_wo[l]).T, cfg.ln_eps)
            x = _mlp(w, l, x)
        h = r(F.gelu(r(x[:, -1]) @ r(w.dec_w1).T + w.dec_b1))
        outs.append(h @ r(w.dec_w2).T + w.dec_b2)
    return torch.cat(outs, 0)


class Bf16EmulatedRegressor:
    """`OracleTabPFNRegressor` with the kernels' roundings (same fit statistics, borders and head)."""

    def __init__(self, weights, softmax_temperature: float = 0.9, chunk: int = 2048):
        self.w = weights
        self.temperature = float(softmax_temperature)
        self.chunk = chunk

    def fit(self, X, y):
        from . import bar_head
        from .estimator import y_standardise
        X = torch.as_tensor(X, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        self.y_mean, self.y_std, yz = y_standardise(y)
        self.cache = prefill(self.w, X, yz)
        self.borders_orig = bar_head.renorm_borders(self.w.borders, self.y_mean, self.y_std)
        return self

    def fit_with_kv(self, X, y, kv: torch.Tensor):
        """Like `fit`, but the per-layer head-0 K/V of the context rows are GIVEN (`kv [L, T, N, 64]`, K 0..31 | V 32..63,
        e.g. a slot's cache exported from the device) instead of being recomputed: at 10 000 context rows the emulated
        context pass would take minutes on the CPU, while the test-row path against a given cache takes seconds."""
        from . import bar_head
        from .estimator import y_standardise
        X = torch.as_tensor(X, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        self.y_mean, self.y_std, yz = y_standardise(y)
        G = (X.shape[1] + 1) // 2
        cache = base.ContextCache()
        cache.stats = base.EncoderStats(X, yz, G)
        cache.T, cache.N = G + 1, X.shape[0]
        kv = kv.float()
        assert kv.shape[1:] == (cache.T, cache.N, 64)
        cache.k0 = [kv[l, :, :, :32].contiguous() for l in range(kv.shape[0])]
        cache.v0 = [kv[l, :, :, 32:].contiguous() for l in range(kv.shape[0])]
        self.cache = cache
        self.borders_orig = bar_head.renorm_borders(self.w.borders, self.y_mean, self.y_std)
        return self

    def predict(self, X, output_type: str = "full", quantiles=None):
        from .estimator import OracleCriterion
        assert output_type == "full"
        X = torch.as_tensor(X, dtype=torch.float32)
        logits = forward_test(self.w, self.cache, X, chunk=self.chunk) * torch.tensor(1.0 / self.temperature,
                                                                                      dtype=torch.float32)
        return {"criterion": OracleCriterion(self.borders_orig), "logits": logits}

"""CPU ORACLE — test infrastructure only (never imported by the product path).

fp32 (optionally fp64) pure-PyTorch restatement of the in-context model behind
the reference's five-call estimator protocol:

    TabPFNRegressor(**kw) / .fit(X, y) / .predict(X, output_type="full", quantiles=[])
        -> {"criterion", "logits"} / criterion.sample(logits) / criterion(logits, y)

called from `/root/reference/npe_pfn/npe_pfn.py:48, 140, 143-146, 149-151,
215-220, 226-228, 502-512`.

PARITY UNPINNED: the arithmetic lives in the third-party package
`tabpfn==2.2.1` (`/root/reference/poetry.lock:4455-4456`), which is neither in
`/root/reference` nor installable offline, and the reference's own tests hold
no golden values for this path (`/root/reference/tests/test_npe_pfn.py:69-71`
assert shapes / finiteness only).  This file restates the published
PerFeatureTransformer (TabPFNv2 regressor) algorithm as documented in
SURVEY.md Appendix A.2, with `n_estimators=1` and identity preprocessing:

  * per-group (2 features) encoder with train-row z-normalisation, NaN/inf
    indicators, scaling by used features, Linear(4->E), subspace feature
    positional embedding; y-encoder Linear(2->E)+bias with test rows' y = NaN
    -> (train mean, indicator -2);
  * 12 post-LayerNorm layers of {attention between features, attention
    between items, MLP(GELU)}; context rows use full multi-head item
    attention, test rows attend ONLY to context rows through head 0's K/V
    shared by all query heads (multiquery_item_attention_for_test_set);
  * decoder Linear(E->4E)+GELU+Linear(4E->5000) on the y-token of test rows.

Because context states do not depend on test rows, the forward is split into
`prefill` (context -> per-layer head-0 K/V cache) and `forward_test`;
`forward_joint` is the monolithic version used to check the split.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


def _ln(x: torch.Tensor, eps: float) -> torch.Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def _attn(q, k, v, explicit: bool):
    """q [..., Lq, dh], k/v [..., Lk, dh] -> softmax(q k^T / sqrt(dh)) v."""
    if explicit:
        s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
        return torch.softmax(s, dim=-1) @ v
    return F.scaled_dot_product_attention(q, k, v)


class EncoderStats:
    """Per-feature train-row statistics (Appendix A.2 step 2) + y fill value."""

    def __init__(self, Xc: torch.Tensor, yc: torch.Tensor, G: int):
        N, Fdim = Xc.shape
        Fp = 2 * G
        X = torch.zeros(N, Fp, dtype=torch.float64)
        X[:, :Fdim] = Xc.double()
        finite = torch.isfinite(X)
        cnt = finite.sum(0).clamp(min=1)
        mean = torch.where(finite, X, torch.zeros_like(X)).sum(0) / cnt
        filled = torch.where(finite, X, mean.expand_as(X))
        if N > 1:
            var = ((filled - mean) ** 2).sum(0) / (N - 1)
        else:
            var = torch.zeros(Fp, dtype=torch.float64)
        std = var.sqrt()
        self.mean = mean.float()  # [Fp]
        self.std = std.float()
        self.const = self.std == 0  # padded columns are constant by construction
        used = (~self.const).reshape(G, 2).sum(1).clamp(min=1).float()
        self.scale = torch.sqrt(2.0 / used)  # [G]
        self.y_fill = yc.double().mean().float()
        self.G = G
        self.F = Fdim


def encode_x(w, st: EncoderStats, X: torch.Tensor, dtype) -> torch.Tensor:
    """rows x F raw features -> [rows, G, E] x-tokens (incl. positional embedding)."""
    R, Fdim = X.shape
    G = st.G
    Xp = torch.zeros(R, 2 * G, dtype=torch.float32)
    Xp[:, :Fdim] = X.float()
    ind = torch.zeros_like(Xp)
    ind = torch.where(torch.isnan(Xp), torch.full_like(Xp, -2.0), ind)
    ind = torch.where(Xp == float("inf"), torch.full_like(Xp, 2.0), ind)
    ind = torch.where(Xp == float("-inf"), torch.full_like(Xp, 4.0), ind)
    filled = torch.where(torch.isfinite(Xp), Xp, st.mean.expand_as(Xp))
    xn = (filled - st.mean) / (st.std + 1e-16)
    xn = torch.where(st.const.expand_as(xn), torch.zeros_like(xn), xn).clamp(-100.0, 100.0)
    xn = xn.reshape(R, G, 2) * st.scale[None, :, None]
    ind = ind.reshape(R, G, 2)
    feat = torch.cat([xn, ind], dim=-1).to(dtype)  # [R, G, 4]
    tok = feat @ w.enc_x_w.to(dtype).T  # [R, G, E]
    return tok + w.pos_emb[:G].to(dtype)[None]


def encode_y_ctx(w, y: torch.Tensor, dtype) -> torch.Tensor:
    yy = torch.stack([y.float(), torch.zeros_like(y, dtype=torch.float32)], -1).to(dtype)
    return yy @ w.enc_y_w.to(dtype).T + w.enc_y_b.to(dtype)


def encode_y_test(w, st: EncoderStats, M: int, dtype) -> torch.Tensor:
    yy = torch.stack([st.y_fill.expand(M), torch.full((M,), -2.0)], -1).to(dtype)
    return yy @ w.enc_y_w.to(dtype).T + w.enc_y_b.to(dtype)


class ContextCache:
    """What `prefill` keeps for one (context, dimension): encoder statistics and,
    per layer, head 0's K and V of the context rows `[T, N, dh]`."""

    def __init__(self):
        self.stats: Optional[EncoderStats] = None
        self.k0: List[torch.Tensor] = []
        self.v0: List[torch.Tensor] = []
        self.T = 0
        self.N = 0


def _feature_attn(w, l, x, explicit, dtype):
    cfg = w.cfg
    R, T, E = x.shape
    H, dh = cfg.nhead, cfg.head_dim
    qkv = x @ w.feat_wqkv[l].to(dtype).T  # [R, T, 3E]
    q, k, v = qkv.split(E, dim=-1)
    q = q.reshape(R, T, H, dh).transpose(1, 2)
    k = k.reshape(R, T, H, dh).transpose(1, 2)
    v = v.reshape(R, T, H, dh).transpose(1, 2)
    o = _attn(q, k, v, explicit).transpose(1, 2).reshape(R, T, E)
    return _ln(x + o @ w.feat_wo[l].to(dtype).T, cfg.ln_eps)


def _mlp(w, l, x, dtype):
    h = F.gelu(x @ w.mlp_w1[l].to(dtype).T)
    return _ln(x + h @ w.mlp_w2[l].to(dtype).T, w.cfg.ln_eps)


def prefill(w, Xc: torch.Tensor, yc: torch.Tensor, *, dtype=torch.float32,
            explicit: bool = False, return_states: bool = False):
    """Context rows through all layers; returns the `ContextCache` (and optionally
    the final context states, for tests)."""
    cfg = w.cfg
    N, Fdim = Xc.shape
    G = (Fdim + 1) // 2
    T = G + 1
    E, H, dh = cfg.emsize, cfg.nhead, cfg.head_dim
    cache = ContextCache()
    cache.stats = st = EncoderStats(Xc, yc, G)
    cache.T, cache.N = T, N
    x = torch.cat([encode_x(w, st, Xc, dtype), encode_y_ctx(w, yc, dtype)[:, None, :]], dim=1)  # [N,T,E]
    for l in range(cfg.nlayers):
        x = _feature_attn(w, l, x, explicit, dtype)
        qkv = x @ w.item_wqkv[l].to(dtype).T  # [N, T, 3E]
        q, k, v = qkv.split(E, dim=-1)
        # [T, H, N, dh]: items are the sequence, columns the batch
        q = q.reshape(N, T, H, dh).permute(1, 2, 0, 3)
        k = k.reshape(N, T, H, dh).permute(1, 2, 0, 3)
        v = v.reshape(N, T, H, dh).permute(1, 2, 0, 3)
        cache.k0.append(k[:, 0].contiguous())  # [T, N, dh]
        cache.v0.append(v[:, 0].contiguous())
        o = _attn(q, k, v, explicit).permute(2, 0, 1, 3).reshape(N, T, E)
        x = _ln(x + o @ w.item_wo[l].to(dtype).T, cfg.ln_eps)
        x = _mlp(w, l, x, dtype)
    if return_states:
        return cache, x
    return cache


def forward_test(w, cache: ContextCache, Xt: torch.Tensor, *, dtype=torch.float32,
                 explicit: bool = False, chunk: int = 4096, return_states: bool = False):
    """Test rows against the cached context -> raw decoder logits [M, num_buckets]."""
    cfg = w.cfg
    E, H, dh = cfg.emsize, cfg.nhead, cfg.head_dim
    outs, states = [], []
    for s in range(0, Xt.shape[0], chunk):
        X = Xt[s:s + chunk]
        M = X.shape[0]
        T = cache.T
        x = torch.cat([encode_x(w, cache.stats, X, dtype),
                       encode_y_test(w, cache.stats, M, dtype)[:, None, :]], dim=1)
        for l in range(cfg.nlayers):
            x = _feature_attn(w, l, x, explicit, dtype)
            q = x @ w.item_wqkv[l][:E].to(dtype).T  # [M, T, E]
            q = q.reshape(M, T, H, dh).permute(1, 0, 2, 3).reshape(T, M * H, dh)
            o = _attn(q, cache.k0[l].to(dtype), cache.v0[l].to(dtype), explicit)  # [T, M*H, dh]
            o = o.reshape(T, M, H, dh).permute(1, 0, 2, 3).reshape(M, T, E)
            x = _ln(x + o @ w.item_wo[l].to(dtype).T, cfg.ln_eps)
            x = _mlp(w, l, x, dtype)
        h = F.gelu(x[:, -1] @ w.dec_w1.to(dtype).T + w.dec_b1.to(dtype))
        outs.append(h @ w.dec_w2.to(dtype).T + w.dec_b2.to(dtype))
        if return_states:
            states.append(x)
    logits = torch.cat(outs, 0)
    if return_states:
        return logits, torch.cat(states, 0)
    return logits


def forward_joint(w, Xc, yc, Xt, *, dtype=torch.float32) -> torch.Tensor:
    """Monolithic [context; test] forward the way upstream runs it
    (single_eval_pos = N): used only to check that prefill + forward_test is the
    same function."""
    cfg = w.cfg
    E, H, dh = cfg.emsize, cfg.nhead, cfg.head_dim
    N, M = Xc.shape[0], Xt.shape[0]
    G = (Xc.shape[1] + 1) // 2
    T = G + 1
    st = EncoderStats(Xc, yc, G)
    xs = encode_x(w, st, torch.cat([Xc, Xt], 0), dtype)
    ys = torch.cat([encode_y_ctx(w, yc, dtype), encode_y_test(w, st, M, dtype)], 0)
    x = torch.cat([xs, ys[:, None, :]], dim=1)  # [N+M, T, E]
    for l in range(cfg.nlayers):
        x = _feature_attn(w, l, x, True, dtype)
        qkv = x @ w.item_wqkv[l].to(dtype).T
        q, k, v = [a.reshape(N + M, T, H, dh) for a in qkv.split(E, dim=-1)]
        o = torch.empty_like(q)
        for t in range(T):
            for h in range(H):
                # train rows: full attention among train rows, own head
                o[:N, t, h] = _attn(q[:N, t, h], k[:N, t, h], v[:N, t, h], True)
                # test rows: keys/values = train rows, head 0 only
                o[N:, t, h] = _attn(q[N:, t, h], k[:N, t, 0], v[:N, t, 0], True)
        x = _ln(x + o.reshape(N + M, T, E) @ w.item_wo[l].to(dtype).T, cfg.ln_eps)
        x = _mlp(w, l, x, dtype)
    h = F.gelu(x[N:, -1] @ w.dec_w1.to(dtype).T + w.dec_b1.to(dtype))
    return h @ w.dec_w2.to(dtype).T + w.dec_b2.to(dtype)

"""TabPFNv2-regressor architecture constants and seeded random-init weights.

The upstream `tabpfn==2.2.1` package and its HF checkpoint are absent offline
(SURVEY.md §8c), so the transformer weights are generated from a seed with the
architecture of SURVEY.md Appendix A.1: E=192, 6 heads x 32, 12 layers, hidden
768, 5000 buckets, features_per_group=2, LayerNorm without affine, no biases in
attention / MLP linears, biases in the y-encoder and the decoder.

The same `PFNWeights` object feeds both the CPU oracle (`oracle/`) and the
CUDA engine (`npe_pfn_b200.csrc`), which is what makes parity meaningful.

`load_checkpoint` is the (blind, table-driven) loader for a real
`tabpfn-v2-regressor.ckpt` should one be present; key names are isolated in
`_CKPT_KEYMAP` so they can be fixed in one place.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, Optional

import torch


@dataclasses.dataclass(frozen=True)
class PFNConfig:
    emsize: int = 192
    nhead: int = 6
    nlayers: int = 12
    nhid: int = 768
    features_per_group: int = 2
    num_buckets: int = 5000
    pos_dim: int = 48  # "subspace" feature positional embedding (Appendix A.2 step 4)
    max_groups: int = 64  # pos-emb rows materialised (features <= 128)
    ln_eps: float = 1e-5
    seed: int = 0

    @property
    def head_dim(self) -> int:
        return self.emsize // self.nhead


#: order and shapes of the flat fp32 weight blob handed to the C-ABI
#: (`pfn_ctx_create`, include/npe_pfn_b200.h).  Per-layer tensors are stacked on
#: a leading [nlayers] axis.
def blob_layout(cfg: PFNConfig):
    E, H, B, L = cfg.emsize, cfg.nhid, cfg.num_buckets, cfg.nlayers
    return [
        ("enc_x_w", (E, 4)),
        ("enc_y_w", (E, 2)),
        ("enc_y_b", (E,)),
        ("pos_emb", (cfg.max_groups, E)),
        ("feat_wqkv", (L, 3 * E, E)),
        ("feat_wo", (L, E, E)),
        ("item_wqkv", (L, 3 * E, E)),
        ("item_wo", (L, E, E)),
        ("mlp_w1", (L, H, E)),
        ("mlp_w2", (L, E, H)),
        ("dec_w1", (H, E)),
        ("dec_b1", (H,)),
        ("dec_w2", (B, H)),
        ("dec_b2", (B,)),
        ("borders", (B + 1,)),
    ]


class PFNWeights:
    """fp32 CPU tensors of one TabPFNv2-architecture regressor."""

    def __init__(self, cfg: PFNConfig, tensors: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.t = tensors
        for name, shape in blob_layout(cfg):
            assert tuple(tensors[name].shape) == tuple(shape), (name, tensors[name].shape, shape)
            assert tensors[name].dtype == torch.float32

    def __getattr__(self, name):
        t = self.__dict__.get("t")
        if t is not None and name in t:
            return t[name]
        raise AttributeError(name)

    def to_blob(self) -> torch.Tensor:
        """Flat fp32 blob in `blob_layout` order (what `pfn_ctx_create` takes)."""
        return torch.cat([self.t[n].reshape(-1) for n, _ in blob_layout(self.cfg)]).contiguous()

    # ------------------------------------------------------------------
    @staticmethod
    def random_init(cfg: Optional[PFNConfig] = None, seed: Optional[int] = None) -> "PFNWeights":
        """Seeded random init (std = 1/sqrt(fan_in)); out-projections are NOT
        zero-initialised (SURVEY.md §7.2 "random-init realism")."""
        cfg = cfg or PFNConfig()
        g = torch.Generator(device="cpu")
        g.manual_seed(cfg.seed if seed is None else seed)
        E, H, B, L = cfg.emsize, cfg.nhid, cfg.num_buckets, cfg.nlayers

        def lin(*shape, fan_in, gain=1.0):
            return (torch.randn(*shape, generator=g) * (gain / math.sqrt(fan_in))).float()

        t: Dict[str, torch.Tensor] = {}
        t["enc_x_w"] = lin(E, 4, fan_in=2)
        t["enc_y_w"] = lin(E, 2, fan_in=2)
        t["enc_y_b"] = lin(E, fan_in=4)
        pos_raw = torch.randn(cfg.max_groups, cfg.pos_dim, generator=g)
        pos_w = lin(E, cfg.pos_dim, fan_in=cfg.pos_dim)
        pos_b = lin(E, fan_in=4)
        t["pos_emb"] = (pos_raw @ pos_w.T + pos_b).float().contiguous()
        t["feat_wqkv"] = lin(L, 3 * E, E, fan_in=E)
        t["feat_wo"] = lin(L, E, E, fan_in=E)
        t["item_wqkv"] = lin(L, 3 * E, E, fan_in=E, gain=1.5)
        t["item_wo"] = lin(L, E, E, fan_in=E)
        t["mlp_w1"] = lin(L, H, E, fan_in=E)
        t["mlp_w2"] = lin(L, E, H, fan_in=H)
        t["dec_w1"] = lin(H, E, fan_in=E)
        t["dec_b1"] = lin(H, fan_in=4)
        t["dec_w2"] = lin(B, H, fan_in=H, gain=3.0)
        t["dec_b2"] = lin(B, fan_in=4)
        t["borders"] = default_borders(B)
        return PFNWeights(cfg, t)

    # ------------------------------------------------------------------
    _CKPT_KEYMAP = {
        # ours -> upstream state_dict key pattern ({l} = layer index).  Blind:
        # recalled from tabpfn 2.x (`tabpfn/architectures/base/*`), unverified.
        "enc_x_w": "encoder.5.layer.weight",
        "enc_y_w": "y_encoder.2.layer.weight",
        "enc_y_b": "y_encoder.2.layer.bias",
        "pos_w": "feature_positional_embedding_embeddings.weight",
        "pos_b": "feature_positional_embedding_embeddings.bias",
        "feat_wqkv": "transformer_encoder.layers.{l}.self_attn_between_features._w_qkv",
        "feat_wo": "transformer_encoder.layers.{l}.self_attn_between_features._w_out",
        "item_wq": "transformer_encoder.layers.{l}.self_attn_between_items._w_q",
        "item_wkv": "transformer_encoder.layers.{l}.self_attn_between_items._w_kv",
        "item_wqkv": "transformer_encoder.layers.{l}.self_attn_between_items._w_qkv",
        "item_wo": "transformer_encoder.layers.{l}.self_attn_between_items._w_out",
        "mlp_w1": "transformer_encoder.layers.{l}.mlp.linear1.weight",
        "mlp_w2": "transformer_encoder.layers.{l}.mlp.linear2.weight",
        "dec_w1": "decoder_dict.standard.0.weight",
        "dec_b1": "decoder_dict.standard.0.bias",
        "dec_w2": "decoder_dict.standard.2.weight",
        "dec_b2": "decoder_dict.standard.2.bias",
        "borders": "criterion.borders",
    }

    @staticmethod
    def load_checkpoint(path: str, cfg: Optional[PFNConfig] = None, strict: bool = True) -> "PFNWeights":
        """Load a `tabpfn-v2-regressor.ckpt` / `tabpfn-v2-classifier.ckpt` (blind mapping, see `_CKPT_KEYMAP`).

        Upstream stores attention weights as `_w_qkv[3, nhead, d_k, E]` and
        `_w_out[nhead, d_v, E]`; they are flattened to our `[3E, E]` / `[E, E]`
        (out-projection transposed to `[E_out, E_in]`).

        The mapping could not be validated offline (no checkpoint, no `tabpfn`), so it refuses to guess: a missing
        key raises, a tensor whose element count does not match the target shape raises, and with `strict` (default)
        any state_dict entry the mapping did NOT consume raises too, listing the leftovers - a silently ignored
        tensor would mean a silently wrong posterior.  `tests/test_checkpoint_gated.py` compares the loaded model
        with upstream's own forward whenever a checkpoint and the `tabpfn` package are present."""
        cfg = cfg or PFNConfig()
        ck = torch.load(path, map_location="cpu", weights_only=False)
        sd = ck.get("state_dict", ck)
        km = PFNWeights._CKPT_KEYMAP
        E, L, H, B = cfg.emsize, cfg.nlayers, cfg.nhid, cfg.num_buckets
        used = set()

        def get(key, l=None, numel=None):
            k = km[key].format(l=l)
            for cand in (k, "model." + k):
                if cand in sd:
                    used.add(cand)
                    w = sd[cand].float()
                    if numel is not None and w.numel() != numel:
                        raise ValueError(f"checkpoint tensor {cand!r} has {w.numel()} elements, expected {numel} "
                                         f"(shape {tuple(w.shape)}); wrong architecture config or key map")
                    return w
            raise KeyError(f"checkpoint key {k!r} not found; fix PFNWeights._CKPT_KEYMAP")

        def qkv(prefix, l):
            try:
                w = get(prefix + "_wqkv", l, 3 * E * E)  # [3, H, dk, E]
                return w.reshape(3 * E, E)
            except KeyError:
                q = get(prefix + "_wq", l, E * E).reshape(E, E)
                kv = get(prefix + "_wkv", l, 2 * E * E).reshape(2 * E, E)
                return torch.cat([q, kv], 0)

        def wo(prefix, l):
            w = get(prefix + "_wo", l, E * E)  # [H, dv, E_out]
            return w.reshape(E, E).T.contiguous()

        t: Dict[str, torch.Tensor] = {}
        t["enc_x_w"] = get("enc_x_w", numel=4 * E).reshape(E, 4)
        t["enc_y_w"] = get("enc_y_w", numel=2 * E).reshape(E, 2)
        t["enc_y_b"] = get("enc_y_b", numel=E)
        # upstream draws the subspace embedding inputs from a CPU generator seeded with the model seed at every forward
        g = torch.Generator(device="cpu").manual_seed(cfg.seed)
        pos_raw = torch.randn(cfg.max_groups, cfg.pos_dim, generator=g)
        t["pos_emb"] = (pos_raw @ get("pos_w", numel=E * cfg.pos_dim).reshape(E, cfg.pos_dim).T + get("pos_b", numel=E)).contiguous()
        t["feat_wqkv"] = torch.stack([qkv("feat", l) for l in range(L)])
        t["feat_wo"] = torch.stack([wo("feat", l) for l in range(L)])
        t["item_wqkv"] = torch.stack([qkv("item", l) for l in range(L)])
        t["item_wo"] = torch.stack([wo("item", l) for l in range(L)])
        t["mlp_w1"] = torch.stack([get("mlp_w1", l, H * E).reshape(H, E) for l in range(L)])
        t["mlp_w2"] = torch.stack([get("mlp_w2", l, E * H).reshape(E, H) for l in range(L)])
        t["dec_w1"] = get("dec_w1", numel=H * E).reshape(H, E)
        t["dec_b1"] = get("dec_b1", numel=H)
        t["dec_w2"] = get("dec_w2", numel=B * H).reshape(B, H)
        t["dec_b2"] = get("dec_b2", numel=B)
        try:
            t["borders"] = get("borders", numel=B + 1)
        except KeyError:
            if B >= 100:
                raise  # a regressor checkpoint must carry its bucket borders
            t["borders"] = default_borders(B)  # classifier: unused
        leftover = sorted(k for k in sd if k not in used and torch.is_tensor(sd[k]) and sd[k].numel() > 1)
        if strict and leftover:
            raise ValueError(f"checkpoint {path!r}: {len(leftover)} tensors were not consumed by the key map "
                             f"(first: {leftover[:8]}); refusing to run a partially loaded model - extend "
                             f"PFNWeights._CKPT_KEYMAP or pass strict=False after checking them")
        return PFNWeights(cfg, {k: v.contiguous() for k, v in t.items()})

    @staticmethod
    def default(cfg: Optional[PFNConfig] = None, env: str = "NPE_PFN_B200_CKPT",
                allow_random_init: Optional[bool] = None) -> "PFNWeights":
        """Weights for an estimator constructed without explicit `weights`.

        `$NPE_PFN_B200_CKPT` (`$NPE_PFN_B200_CLASSIFIER_CKPT` for the classifier) must point at a TabPFNv2 checkpoint.
        Nothing else is guessed: a path that does not exist raises, and with no checkpoint configured this raises as
        well unless random initialisation was asked for explicitly (`allow_random_init=True`, or
        `NPE_PFN_B200_ALLOW_RANDOM_INIT=1` as the tests and bench.py set it) - an untrained network returns
        well-formed but meaningless posteriors, which must never happen silently."""
        path = os.environ.get(env, "")
        if path:
            if not os.path.exists(path):
                raise FileNotFoundError(f"${env} = {path!r} does not exist")
            return PFNWeights.load_checkpoint(path, cfg)
        if allow_random_init is None:
            allow_random_init = os.environ.get("NPE_PFN_B200_ALLOW_RANDOM_INIT", "") not in ("", "0")
        if not allow_random_init:
            raise RuntimeError(
                f"no TabPFNv2 checkpoint configured: set ${env} to the checkpoint file, pass `weights=` explicitly "
                "(regressor_init_kwargs / classifier_init_kwargs), or opt in to an UNTRAINED seeded random "
                "initialisation with NPE_PFN_B200_ALLOW_RANDOM_INIT=1 (benchmarks and tests only)")
        import warnings
        warnings.warn(f"${env} is not set: using a seeded RANDOM initialisation of the TabPFNv2 architecture "
                      "(NPE_PFN_B200_ALLOW_RANDOM_INIT); posteriors are meaningless", RuntimeWarning, stacklevel=2)
        return PFNWeights.random_init(cfg)


def classifier_config() -> PFNConfig:
    """TabPFNv2 classifier: the same transformer with its own weights and a 10-way decoder (max_num_classes = 10,
    SURVEY.md Appendix A.1); its `borders` are unused."""
    return PFNConfig(num_buckets=10, seed=1)


def default_borders(num_buckets: int) -> torch.Tensor:
    """Bucket borders for the random-init model: standard-normal quantiles on an
    even probability grid (upstream derives its borders from prior-data
    quantiles)."""
    if num_buckets < 100:  # classifier head: placeholder grid
        return torch.linspace(-1.0, 1.0, num_buckets + 1, dtype=torch.float32)
    p = torch.linspace(2e-4, 1.0 - 2e-4, num_buckets + 1, dtype=torch.float64)
    b = torch.distributions.Normal(0.0, 1.0).icdf(p)
    b = b.float()
    assert bool((b[1:] > b[:-1]).all())
    return b.contiguous()

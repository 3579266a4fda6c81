"""ctypes binding of the C-ABI in `include/npe_pfn_b200.h` (thin: pointers + sizes only).

`Engine` owns one `pfn_ctx` on one CUDA device.  Torch is used here only for device memory and the
current stream; every compute call goes through `libnpe_pfn_b200.so`.  There is no CPU fallback: if the
library is missing, cannot be built, or no sm_100 device is visible, construction raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import build as _build
from .weights import PFNConfig, PFNWeights

c = ctypes


class pfn_model_config(ctypes.Structure):
    _fields_ = [
        ("emsize", c.c_int32), ("nhead", c.c_int32), ("nlayers", c.c_int32), ("nhid", c.c_int32),
        ("num_buckets", c.c_int32), ("max_groups", c.c_int32), ("max_slots", c.c_int32), ("chunk_rows", c.c_int32),
        ("ln_eps", c.c_float), ("softmax_temperature", c.c_float),
    ]


#: every symbol `include/npe_pfn_b200.h` declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "pfn_abi_version", "pfn_last_error", "pfn_ctx_create", "pfn_ctx_destroy", "pfn_set_option", "pfn_prefill",
    "pfn_forward_logits", "pfn_head_sample", "pfn_head_nll", "pfn_sample", "pfn_logprob", "pfn_accept_compact",
    "pfn_accept_append", "pfn_uniform_box", "pfn_sample_rejection", "pfn_filter_context",
    "pfn_slot_info", "pfn_launch_count", "pfn_kernel_times", "pfn_slot_export", "pfn_slot_state", "pfn_slot_import",
    "pfn_debug_last_states", "pfn_member_transform", "pfn_ensemble_combine", "pfn_attn_debug_counts",
]

#: PFN_ABI_VERSION of include/npe_pfn_b200.h this binding was written against
ABI_VERSION = 2

_LIB = None


def load_library(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building if stale and nvcc is available) the in-tree shared library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("NPE_PFN_B200_LIB") or _build.LIB_PATH  # override: experimental builds (tuning sweeps)
    if path == _build.LIB_PATH and build_if_missing and _build.is_stale():
        try:
            _build.build_library()
        except Exception as e:  # a stale library may still be usable (the ABI check below decides); a missing one is fatal
            if not os.path.exists(path):
                raise RuntimeError(f"libnpe_pfn_b200.so is missing and could not be built: {e}") from e
            import warnings
            warnings.warn(f"libnpe_pfn_b200.so is older than its sources and the rebuild failed ({e}); using the stale "
                          f"library", RuntimeWarning)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = ctypes.CDLL(path)
    L.pfn_abi_version.restype = c.c_int
    if L.pfn_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path} exports ABI version {L.pfn_abi_version()}, this binding needs {ABI_VERSION}: rebuild "
                           f"with `python -c 'import __graft_entry__ as g; g.build()'`")
    vp, i64, u64, i32, f32 = c.c_void_p, c.c_int64, c.c_uint64, c.c_int32, c.c_float
    L.pfn_abi_version.restype = c.c_int
    L.pfn_last_error.restype = c.c_char_p
    L.pfn_ctx_create.restype = c.c_int
    L.pfn_ctx_create.argtypes = [c.POINTER(pfn_model_config), vp, c.c_size_t, c.c_int, vp, c.POINTER(vp)]
    L.pfn_ctx_destroy.restype = c.c_int
    L.pfn_ctx_destroy.argtypes = [vp]
    L.pfn_set_option.restype = c.c_int
    L.pfn_set_option.argtypes = [vp, c.c_char_p, i64]
    L.pfn_prefill.restype = c.c_int
    L.pfn_prefill.argtypes = [vp, c.c_int, vp, i64, vp, i64, c.c_int, vp]
    L.pfn_forward_logits.restype = c.c_int
    L.pfn_forward_logits.argtypes = [vp, c.c_int, vp, i64, i64, vp, i64, vp]
    L.pfn_head_sample.restype = c.c_int
    L.pfn_head_sample.argtypes = [vp, c.c_int, vp, i64, i64, i64, vp, u64, u64, u64, vp, i64, vp, vp, vp, f32, c.c_int, vp]
    L.pfn_head_nll.restype = c.c_int
    L.pfn_head_nll.argtypes = [vp, c.c_int, vp, i64, i64, i64, vp, i64, vp, vp, f32, c.c_int, vp]
    L.pfn_sample.restype = c.c_int
    L.pfn_sample.argtypes = [vp, c.c_int, vp, i64, i64, vp, u64, u64, u64, vp, i64, vp, vp, f32, c.c_int, vp]
    L.pfn_logprob.restype = c.c_int
    L.pfn_logprob.argtypes = [vp, c.c_int, vp, i64, i64, vp, i64, vp, f32, c.c_int, vp]
    L.pfn_accept_compact.restype = c.c_int
    L.pfn_accept_compact.argtypes = [vp, vp, i64, i64, c.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.pfn_accept_append.restype = c.c_int
    L.pfn_accept_append.argtypes = [vp, vp, i64, i64, c.c_int, vp, vp, vp, vp, vp, vp, vp, i64, vp, i64, vp, vp]
    L.pfn_uniform_box.restype = c.c_int
    L.pfn_uniform_box.argtypes = [vp, vp, vp, i64, c.c_int, u64, u64, vp, i64, vp]
    L.pfn_sample_rejection.restype = c.c_int
    L.pfn_sample_rejection.argtypes = [vp, c.POINTER(i32), vp, c.c_int, c.c_int, c.c_int, i64, vp, vp, u64, u64, f32, vp, i64,
                                       vp, i64, vp, vp]
    L.pfn_filter_context.restype = c.c_int
    L.pfn_filter_context.argtypes = [vp, vp, i64, i64, c.c_int, vp, i64, vp, vp, vp]
    L.pfn_slot_info.restype = c.c_int
    L.pfn_slot_info.argtypes = [vp, c.c_int, c.POINTER(i64), c.POINTER(i32), c.POINTER(i32), c.POINTER(i64)]
    L.pfn_launch_count.restype = i64
    L.pfn_launch_count.argtypes = [vp]
    L.pfn_kernel_times.restype = c.c_int
    L.pfn_kernel_times.argtypes = [vp, c.POINTER(c.c_double), c.POINTER(i64), c.POINTER(c.c_double), c.POINTER(c.c_double), c.c_int,
                                   c.c_int]
    L.pfn_slot_export.restype = c.c_int
    L.pfn_slot_export.argtypes = [vp, c.c_int, vp, vp, vp, vp, vp]
    L.pfn_slot_state.restype = c.c_int
    L.pfn_slot_state.argtypes = [vp, c.c_int, vp, vp]
    L.pfn_slot_import.restype = c.c_int
    L.pfn_slot_import.argtypes = [vp, c.c_int, i64, c.c_int, vp, vp, vp, vp]
    L.pfn_debug_last_states.restype = c.c_int
    L.pfn_debug_last_states.argtypes = [vp, vp, i64, vp]
    L.pfn_attn_debug_counts.restype = c.c_int
    L.pfn_attn_debug_counts.argtypes = [vp, c.POINTER(u64)]
    L.pfn_member_transform.restype = c.c_int
    L.pfn_member_transform.argtypes = [vp, vp, vp, i64, i64, vp, i64, vp]
    L.pfn_ensemble_combine.restype = c.c_int
    L.pfn_ensemble_combine.argtypes = [vp, vp, i64, i64, c.c_int, i64, vp, vp, vp, vp, i64, vp]
    _LIB = L
    return L


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else c.c_void_p(t.data_ptr())


def _f32_rows(t: torch.Tensor, device) -> torch.Tensor:
    """fp32 CUDA tensor with unit inner stride (rows may be strided)."""
    t = t.to(device=device, dtype=torch.float32)
    if t.ndim == 1:
        return t.contiguous()
    if t.stride(-1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


class Engine:
    """One `pfn_ctx`: bf16 weights, K/V-cache slots and workspace on one B200."""

    def __init__(self, weights: Optional[PFNWeights] = None, device: Optional[int] = None, max_slots: int = 16,
                 softmax_temperature: float = 0.9, chunk_rows: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("npe_pfn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.weights = weights or PFNWeights.default()
        cfg: PFNConfig = self.weights.cfg
        self.cfg = cfg
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.max_slots = max_slots
        self.temperature = float(softmax_temperature)
        mc = pfn_model_config(cfg.emsize, cfg.nhead, cfg.nlayers, cfg.nhid, cfg.num_buckets, cfg.max_groups, max_slots,
                              chunk_rows, cfg.ln_eps, self.temperature)
        blob = self.weights.to_blob().to(self.device)
        h = c.c_void_p()
        rc = self.lib.pfn_ctx_create(c.byref(mc), _ptr(blob), blob.numel(), self.device_index, self._stream(),
                                     c.byref(h))
        self._h = h if rc == 0 else None
        self._check(rc)
        del blob

    # ------------------------------------------------------------------
    def _stream(self):
        return c.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError("npe_pfn_b200: " + self.lib.pfn_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pfn_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        self._check(self.lib.pfn_set_option(self._h, key.encode(), int(value)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.pfn_launch_count(self._h))

    # ------------------------------------------------------------------
    def prefill(self, slot: int, X: torch.Tensor, y: torch.Tensor):
        X = _f32_rows(X, self.device)
        y = y.to(device=self.device, dtype=torch.float32).contiguous()
        assert X.ndim == 2 and y.ndim == 1 and X.shape[0] == y.shape[0]
        self._check(self.lib.pfn_prefill(self._h, slot, _ptr(X), X.stride(0) if X.shape[0] > 1 else X.shape[1],
                                         _ptr(y), X.shape[0], X.shape[1], self._stream()))

    def prefill_joint(self, slot: int, joint: torch.Tensor, n_features: int):
        """context = joint[:, :n_features], target = joint[:, n_features] (no copies of X)."""
        assert joint.is_cuda and joint.dtype == torch.float32 and joint.stride(1) == 1
        y = joint[:, n_features].contiguous()
        self._check(self.lib.pfn_prefill(self._h, slot, _ptr(joint), joint.stride(0), _ptr(y), joint.shape[0],
                                         n_features, self._stream()))

    def forward_logits(self, slot: int, X: torch.Tensor) -> torch.Tensor:
        X = _f32_rows(X, self.device)
        M = X.shape[0]
        B = self.cfg.num_buckets
        ld = (B + 3) // 4 * 4  # the library wants 16-byte aligned logits rows
        out = torch.empty(M, ld, dtype=torch.float32, device=self.device)
        self._check(self.lib.pfn_forward_logits(self._h, slot, _ptr(X), X.stride(0) if M > 1 else X.shape[1], M,
                                                _ptr(out), ld, self._stream()))
        return out if ld == B else out[:, :B]

    def head_sample(self, slot: int, logits: torch.Tensor, M: Optional[int] = None, uniforms=None, seed=0, row0=0,
                    offset=0, out_theta=None, ld_theta=1, with_log_prob=False, out_logp=None, eps=1e-15,
                    accumulate=False, return_bins=False, bins=None, group: int = 1):
        """logits [M, B] (or a single row broadcast to M rows when `M` is given and logits.shape[0] == 1; or
        logits [M / group, B] with `group` consecutive draws per logits row)."""
        logits = logits.to(self.device, torch.float32)
        assert logits.stride(-1) == 1
        if M is None:
            M = logits.shape[0]
            ld = logits.stride(0)
        else:
            assert logits.shape[0] == 1 or logits.shape[0] * group == M
            ld = 0 if logits.shape[0] == 1 and M != 1 else logits.stride(0)
        if out_theta is None:
            out_theta = torch.empty(M, dtype=torch.float32, device=self.device)
            ld_theta = 1
        if bins is None and return_bins:
            bins = torch.empty(M, dtype=torch.int32, device=self.device)
        if with_log_prob and out_logp is None:
            out_logp = torch.zeros(M, dtype=torch.float32, device=self.device)
        if uniforms is not None:
            uniforms = uniforms.to(self.device, torch.float32).contiguous()
        self._check(self.lib.pfn_head_sample(self._h, slot, _ptr(logits), ld, int(group), M, _ptr(uniforms), seed, row0, offset,
                                             _ptr(out_theta), ld_theta, _ptr(bins), None, _ptr(out_logp), eps,
                                             int(accumulate), self._stream()))
        return out_theta, bins, out_logp

    def head_nll(self, slot: int, logits: torch.Tensor, y: torch.Tensor, eps=1e-15, ld_y: int = 1, out_logp=None,
                 accumulate=False):
        """-log p(y_r) for logits [M, B], or one logits row evaluated against all M targets (logits.shape[0] == 1).
        With `out_logp` the clamped log-prob (-inf -> log eps) is written / accumulated there instead."""
        logits = logits.to(self.device, torch.float32)
        y = y.to(self.device, torch.float32)
        if ld_y == 1:
            y = y.reshape(-1).contiguous()
        M = y.shape[0] if y.ndim == 1 else y.numel()
        ld = 0 if (logits.shape[0] == 1 and M != 1) else logits.stride(0)
        assert logits.shape[0] in (1, M)
        out = None if out_logp is not None else torch.empty(M, dtype=torch.float32, device=self.device)
        self._check(self.lib.pfn_head_nll(self._h, slot, _ptr(logits), ld, 1, M, _ptr(y), ld_y, _ptr(out), _ptr(out_logp), eps,
                                          int(accumulate), self._stream()))
        return out if out_logp is None else out_logp

    def sample_step(self, slot: int, joint: torch.Tensor, n_features: int, out_col: int, uniforms=None, seed=0, row0=0,
                    offset=0, out_logp=None, eps=1e-15, accumulate=True, bins=None):
        """Fused step on the rows of `joint` [M, >= n_features+1] (fp32 CUDA, unit inner stride): reads features
        joint[:, :n_features], writes the draw into joint[:, out_col]."""
        assert joint.is_cuda and joint.dtype == torch.float32 and joint.stride(1) == 1
        M = joint.shape[0]
        ld = joint.stride(0)
        out_ptr = c.c_void_p(joint.data_ptr() + 4 * out_col)
        self._check(self.lib.pfn_sample(self._h, slot, _ptr(joint), ld, M, _ptr(uniforms), seed, row0, offset, out_ptr,
                                        ld, _ptr(bins), _ptr(out_logp), eps, int(accumulate), self._stream()))

    def logprob_step(self, slot: int, joint: torch.Tensor, n_features: int, y_col: int, out_logp: torch.Tensor,
                     eps=1e-15, accumulate=True):
        assert joint.is_cuda and joint.dtype == torch.float32 and joint.stride(1) == 1
        M = joint.shape[0]
        ld = joint.stride(0)
        y_ptr = c.c_void_p(joint.data_ptr() + 4 * y_col)
        self._check(self.lib.pfn_logprob(self._h, slot, _ptr(joint), ld, M, y_ptr, ld, _ptr(out_logp), eps,
                                         int(accumulate), self._stream()))

    def accept_compact(self, theta: torch.Tensor, lo=None, hi=None, mask=None, want_rows=True):
        """-> (idx[int64, M] (first `count` valid), rows[M, dim] or None, count (device int64 scalar))."""
        assert theta.is_cuda and theta.dtype == torch.float32 and theta.ndim == 2 and theta.stride(1) == 1
        M, dim = theta.shape
        idx = torch.empty(M, dtype=torch.int64, device=self.device)
        rows = torch.empty(M, dim, dtype=torch.float32, device=self.device) if want_rows else None
        count = torch.zeros((), dtype=torch.int64, device=self.device)
        lo, hi = self._bounds(lo, hi, dim)
        mask = None if mask is None else mask.to(self.device, torch.uint8).contiguous()
        self._check(self.lib.pfn_accept_compact(self._h, _ptr(theta), theta.stride(0) if M > 1 else dim, M, dim,
                                                _ptr(lo), _ptr(hi), _ptr(mask), _ptr(idx), _ptr(rows), _ptr(count),
                                                self._stream()))
        return idx, rows, count

    def _bounds(self, lo, hi, dim: int):
        """box bounds as [dim] fp32 device tensors (scalars / broadcastable bounds are expanded: the kernels index them
        per dimension)"""
        def fix(b):
            if b is None:
                return None
            b = torch.as_tensor(b, dtype=torch.float32).to(self.device).reshape(-1)
            if b.numel() == 1:
                b = b.expand(dim)
            if b.numel() != dim:
                raise ValueError(f"support bounds have {b.numel()} entries, theta has {dim} dimensions")
            return b.contiguous()
        return fix(lo), fix(hi)

    def accept_append(self, theta: torch.Tensor, out_rows: torch.Tensor, cursor: torch.Tensor, lo=None, hi=None, mask=None,
                      score=None, thr=None, logp=None, out_logp=None):
        """One link of an on-device rejection loop: append the accepted rows of `theta` to `out_rows` at `cursor[0]`
        (device int64[2] = accepted, proposed), in order.  No host synchronisation."""
        assert theta.is_cuda and theta.dtype == torch.float32 and theta.ndim == 2 and theta.stride(1) == 1
        assert out_rows.is_cuda and out_rows.dtype == torch.float32 and out_rows.stride(1) == 1
        assert cursor.is_cuda and cursor.dtype == torch.int64 and cursor.numel() >= 2
        M, dim = theta.shape
        lo, hi = self._bounds(lo, hi, dim)
        mask = None if mask is None else mask.to(self.device, torch.uint8).contiguous()
        if score is not None:
            score = score.to(self.device, torch.float32).contiguous()
            thr = torch.as_tensor(thr, dtype=torch.float32).to(self.device).reshape(1)
        self._check(self.lib.pfn_accept_append(self._h, _ptr(theta), theta.stride(0) if M > 1 else dim, M, dim, _ptr(lo),
                                               _ptr(hi), _ptr(mask), _ptr(score), _ptr(thr), _ptr(logp), _ptr(out_rows),
                                               out_rows.stride(0), _ptr(out_logp), out_rows.shape[0], _ptr(cursor),
                                               self._stream()))

    def uniform_box(self, lo, hi, M: int, seed: int, row0: int = 0) -> torch.Tensor:
        lo, hi = self._bounds(lo, hi, int(torch.as_tensor(lo).numel()))
        out = torch.empty(M, lo.numel(), dtype=torch.float32, device=self.device)
        self._check(self.lib.pfn_uniform_box(self._h, _ptr(lo), _ptr(hi), M, lo.numel(), seed, row0, _ptr(out), lo.numel(),
                                             self._stream()))
        return out

    def sample_rejection(self, slots, x_obs: torch.Tensor, dim_theta: int, n_rounds: int, round_rows: int, out_theta,
                         cursor, lo=None, hi=None, seed=0, row0=0, eps=1e-15, out_logp=None):
        """`n_rounds` proposal rounds of the autoregressive sampler + support check + ordered append, enqueued back to
        back with no host round trip (pfn_sample_rejection)."""
        x_obs = x_obs.to(self.device, torch.float32).reshape(-1).contiguous()
        lo, hi = self._bounds(lo, hi, dim_theta)
        arr = (c.c_int32 * dim_theta)(*[int(s) for s in slots])
        self._check(self.lib.pfn_sample_rejection(self._h, arr, _ptr(x_obs), x_obs.numel(), dim_theta, int(n_rounds),
                                                  int(round_rows), _ptr(lo), _ptr(hi), seed, row0, eps, _ptr(out_theta),
                                                  out_theta.stride(0), _ptr(out_logp), out_theta.shape[0], _ptr(cursor),
                                                  self._stream()))

    def attn_debug_counts(self):
        """(redone fast-path tiles, reference changes, general-path tiles) since set_option("attn_debug", 1)."""
        out = (c.c_uint64 * 3)()
        self._check(self.lib.pfn_attn_debug_counts(self._h, out))
        return tuple(int(v) for v in out)

    KERNEL_CLASSES = ["attn_test", "attn_ctx", "gemm", "other", "mlp", "head", "encode", "kv_cache", "compact"]

    def kernel_times(self, reset: bool = True, with_bytes: bool = False):
        """{class: (ms, launches, flops[, bytes])} of the launches recorded while option "time_kernels" was on."""
        n = len(self.KERNEL_CLASSES)
        ms, cnt, fl, by = (c.c_double * n)(), (c.c_int64 * n)(), (c.c_double * n)(), (c.c_double * n)()
        self._check(self.lib.pfn_kernel_times(self._h, ms, cnt, fl, by, n, int(reset)))
        if with_bytes:
            return {k: (ms[i], cnt[i], fl[i], by[i]) for i, k in enumerate(self.KERNEL_CLASSES)}
        return {k: (ms[i], cnt[i], fl[i]) for i, k in enumerate(self.KERNEL_CLASSES)}

    def filter_context(self, x_train: torch.Tensor, obs: torch.Tensor, k: int, want_dist: bool = False):
        """Indices (int64, CUDA) of the k simulations nearest to `obs` in z-scored x, ascending distance."""
        x_train = _f32_rows(x_train, self.device)
        obs = obs.to(self.device, torch.float32).reshape(-1).contiguous()
        N, dx = x_train.shape
        idx = torch.empty(k, dtype=torch.int64, device=self.device)
        dist = torch.empty(k, dtype=torch.float32, device=self.device) if want_dist else None
        self._check(self.lib.pfn_filter_context(self._h, _ptr(x_train), x_train.stride(0) if N > 1 else dx, N, dx,
                                                _ptr(obs), k, _ptr(idx), _ptr(dist), self._stream()))
        return (idx, dist) if want_dist else idx

    def slot_info(self, slot: int):
        N, F, T, kv = c.c_int64(), c.c_int32(), c.c_int32(), c.c_int64()
        self._check(self.lib.pfn_slot_info(self._h, slot, c.byref(N), c.byref(F), c.byref(T), c.byref(kv)))
        return {"N": N.value, "F": F.value, "T": T.value, "kv_bytes": kv.value}

    def slot_export(self, slot: int, want_kv: bool = False):
        info = self.slot_info(slot)
        G = info["T"] - 1
        stats = torch.empty(3 * 2 * G, dtype=torch.float32, device=self.device)
        ystats = torch.empty(3, dtype=torch.float32, device=self.device)
        borders = torch.empty(self.cfg.num_buckets + 1, dtype=torch.float32, device=self.device)
        kv = None
        if want_kv:
            kv = torch.empty(self.cfg.nlayers, info["T"], info["N"], 64, dtype=torch.bfloat16, device=self.device)
        self._check(self.lib.pfn_slot_export(self._h, slot, _ptr(stats), _ptr(ystats), _ptr(borders), _ptr(kv),
                                             self._stream()))
        Fp = 2 * G
        return {"mean": stats[:Fp], "std": stats[Fp:2 * Fp], "scale": stats[2 * Fp:2 * Fp + G], "y_mean": ystats[0],
                "y_std": ystats[1], "y_fill": ystats[2], "borders": borders, "kv": kv, **info}

    ENC_STATE_FLOATS = 2 * 128 + 64 + 4

    def slot_pack(self, slot: int):
        """Everything `prefill` produced for a slot, as device tensors (for a broadcast to other ranks)."""
        info = self.slot_info(slot)
        enc = torch.empty(self.ENC_STATE_FLOATS, dtype=torch.float32, device=self.device)
        self._check(self.lib.pfn_slot_state(self._h, slot, _ptr(enc), self._stream()))
        borders = torch.empty(self.cfg.num_buckets + 1, dtype=torch.float32, device=self.device)
        kv = torch.empty(self.cfg.nlayers, info["T"], info["N"], 64, dtype=torch.bfloat16, device=self.device)
        self._check(self.lib.pfn_slot_export(self._h, slot, None, None, _ptr(borders), _ptr(kv), self._stream()))
        return enc, borders, kv

    def slot_unpack(self, slot: int, N: int, F: int, enc: torch.Tensor, borders: torch.Tensor, kv: torch.Tensor):
        assert enc.is_cuda and borders.is_cuda and kv.is_cuda and kv.is_contiguous()
        self._check(self.lib.pfn_slot_import(self._h, slot, N, F, _ptr(enc), _ptr(borders), _ptr(kv), self._stream()))

    def last_states(self, rows: int, T: int) -> torch.Tensor:
        out = torch.empty(rows, T, self.cfg.emsize, dtype=torch.float32, device=self.device)
        self._check(self.lib.pfn_debug_last_states(self._h, _ptr(out), out.numel(), self._stream()))
        return out


_ENGINES = {}


def get_engine(device: Optional[int] = None, weights: Optional[PFNWeights] = None, **kw) -> Engine:
    """Process-wide engine per (device, weights identity, options)."""
    if not torch.cuda.is_available():
        raise RuntimeError("npe_pfn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)
    key = (dev, id(weights) if weights is not None else None, tuple(sorted(kw.items())))
    if key not in _ENGINES:
        _ENGINES[key] = Engine(weights=weights, device=dev, **kw)
    return _ENGINES[key]

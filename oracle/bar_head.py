"""CPU ORACLE — test infrastructure only.

ctypes binding of `oracle/bar_head.c` (the deterministic restatement of the
bar-distribution head) plus `textbook_*`: the plain fp32 torch formulation of
upstream `FullSupportBarDistribution` that the C
restatement is pinned against in tests/test_oracle.py.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpfn_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bar_head.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        c = ctypes
        L.pfn_oracle_exp_det.restype = c.c_float
        L.pfn_oracle_exp_det.argtypes = [c.c_float]
        L.pfn_oracle_quantize.restype = c.c_uint64
        L.pfn_oracle_quantize.argtypes = [c.c_float]
        L.pfn_oracle_philox_uniform.restype = c.c_float
        L.pfn_oracle_philox_uniform.argtypes = [c.c_uint64, c.c_uint64, c.c_uint64]
        L.pfn_oracle_philox4x32.restype = None
        L.pfn_oracle_philox4x32.argtypes = [c.c_uint64, c.c_uint64, c.c_uint64, c.c_void_p]
        L.pfn_oracle_renorm_borders.restype = None
        L.pfn_oracle_renorm_borders.argtypes = [c.c_void_p, c.c_int, c.c_float, c.c_float, c.c_void_p]
        L.pfn_oracle_sample.restype = None
        L.pfn_oracle_sample.argtypes = [c.c_void_p, c.c_int64, c.c_int, c.c_void_p, c.c_void_p,
                                        c.c_uint64, c.c_uint64, c.c_uint64, c.c_void_p, c.c_void_p, c.c_void_p]
        L.pfn_oracle_nll.restype = None
        L.pfn_oracle_nll.argtypes = [c.c_void_p, c.c_int64, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p]
        _lib = L
    return _lib


def _f32(t) -> torch.Tensor:
    return torch.as_tensor(t, dtype=torch.float32).contiguous().cpu()


def renorm_borders(borders: torch.Tensor, y_mean: float, y_std: float) -> torch.Tensor:
    b = _f32(borders)
    out = torch.empty_like(b)
    lib().pfn_oracle_renorm_borders(b.data_ptr(), b.numel(), float(np.float32(y_mean)), float(np.float32(y_std)),
                                    out.data_ptr())
    return out


def philox_uniforms(seed: int, row0: int, offset: int, M: int) -> torch.Tensor:
    L = lib()
    return torch.tensor([L.pfn_oracle_philox_uniform(seed, row0 + r, offset) for r in range(M)], dtype=torch.float32)


def sample(logits: torch.Tensor, borders: torch.Tensor, uniforms=None, seed: int = 0, row0: int = 0,
           offset: int = 0):
    """-> (theta[M] fp32, idx[M] int32, u[M] fp32)"""
    lg = _f32(logits)
    M, B = lg.shape
    b = _f32(borders)
    assert b.numel() == B + 1
    theta = torch.empty(M, dtype=torch.float32)
    idx = torch.empty(M, dtype=torch.int32)
    u_out = torch.empty(M, dtype=torch.float32)
    u = _f32(uniforms) if uniforms is not None else None
    lib().pfn_oracle_sample(lg.data_ptr(), M, B, b.data_ptr(), u.data_ptr() if u is not None else None,
                            seed, row0, offset, theta.data_ptr(), idx.data_ptr(), u_out.data_ptr())
    return theta, idx, u_out


def nll(logits: torch.Tensor, borders: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    lg = _f32(logits)
    M, B = lg.shape
    b = _f32(borders)
    yy = _f32(y)
    out = torch.empty(M, dtype=torch.float32)
    lib().pfn_oracle_nll(lg.data_ptr(), M, B, b.data_ptr(), yy.data_ptr(), out.data_ptr())
    return out


# ---------------------------------------------------------------------------
# textbook fp32 formulation (Appendix A.3) — used to pin the C restatement
# ---------------------------------------------------------------------------
def textbook_icdf(logits: torch.Tensor, borders: torch.Tensor, u: torch.Tensor):
    p = torch.softmax(logits.double(), -1)
    c = torch.cumsum(p, -1)
    idx = torch.searchsorted(c, u.double()[:, None]).squeeze(-1).clamp(0, p.shape[-1] - 1)
    c0 = torch.cat([torch.zeros(c.shape[0], 1, dtype=c.dtype), c], -1)
    rest = u.double() - c0.gather(-1, idx[:, None]).squeeze(-1)
    lo = borders.double()[idx]
    hi = borders.double()[idx + 1]
    return (lo + (hi - lo) * rest / p.gather(-1, idx[:, None]).squeeze(-1)).float(), idx.int()


def textbook_nll(logits: torch.Tensor, borders: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    B = logits.shape[-1]
    borders = borders.double()
    y = y.double()
    widths = borders[1:] - borders[:-1]
    idx = (torch.searchsorted(borders, y) - 1).clamp(0, B - 1)
    logp = torch.log_softmax(logits.double(), -1).gather(-1, idx[:, None]).squeeze(-1) - torch.log(widths[idx])
    icdf_half = torch.distributions.HalfNormal(torch.tensor(1.0, dtype=torch.float64)).icdf(
        torch.tensor(0.5, dtype=torch.float64))
    for edge, sel, v in ((0, idx == 0, borders[1] - y), (B - 1, idx == B - 1, y - borders[-2])):
        if sel.any():
            hn = torch.distributions.HalfNormal(widths[edge] / icdf_half)
            logp[sel] = logp[sel] + hn.log_prob(v[sel].clamp(min=1e-8)) + torch.log(widths[edge])
    return (-logp).float()

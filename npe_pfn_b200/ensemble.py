"""`n_estimators > 1`: the member ensemble upstream `TabPFNRegressor()` runs by default for the reference's
constructor call (`/root/reference/npe_pfn/npe_pfn.py:48`), on the B200 engine.

Per member: constant-feature removal, one of two feature pipelines (quantile-uniform with the original columns and
truncated-SVD components appended; standardise -> Yeo-Johnson -> standardise), a fingerprint feature, a seeded
feature shuffle, and one of two target transforms (none; Yeo-Johnson on the standardised target).  Everything that is
*fitted* (quantile tables, lambdas, SVD basis) is computed once per (context, dimension) from the context rows with
torch on the device; everything that touches the M test rows runs in the library's own kernels:
`pfn_member_transform` (feature pipeline), the ordinary per-slot forward (`pfn_forward_logits`, each member is a slot
with its own K/V cache) and `pfn_ensemble_combine` (softmax, re-binning across the members' borders, mean, log).
The arithmetic specification and its sklearn-backed oracle are `oracle/ensemble.py`; composition, seeds, fingerprint
hash and SVD solver follow that specification (real `tabpfn` is absent offline: parity unpinned).
"""
from __future__ import annotations

import ctypes as c
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from .engine import Engine, _ptr


@dataclass
class MemberSpec:
    x_kind: str  # "quantile" | "safepower" | "none"
    y_kind: str  # "none" | "safepower"
    perm_seed: int
    fingerprint: bool = True
    svd: bool = True


def make_members(n: int, random_state: int = 0, fingerprint: bool = True, svd: bool = True) -> List[MemberSpec]:
    """First ceil(n/2) members: quantile pipeline, the rest safepower; target transforms alternate none / safepower
    inside each half (member 0 always keeps the target as it is); shuffle seeds = seeded permutation of a range."""
    rng = np.random.default_rng(random_state)
    start = int(rng.integers(0, 1000))
    shifts = rng.permutation(np.arange(start, start + n))
    half = (n + 1) // 2
    out = []
    for i in range(n):
        j = i if i < half else i - half
        out.append(MemberSpec("quantile" if i < half else "safepower", "none" if j % 2 == 0 else "safepower",
                              int(shifts[i]), fingerprint, svd))
    return out


def make_classifier_members(n: int, random_state: int = 0, fingerprint: bool = True, svd: bool = True) -> List[MemberSpec]:
    """Classifier ensemble (upstream default 4 members): quantile pipeline and untouched features alternate, the
    target is never transformed (class indices are permuted per member instead, see `EnsembleDim.fit`)."""
    rng = np.random.default_rng(random_state)
    start = int(rng.integers(0, 1000))
    shifts = rng.permutation(np.arange(start, start + n))
    return [MemberSpec("quantile" if i % 2 == 0 else "none", "none", int(shifts[i]), fingerprint, svd) for i in range(n)]


class pfn_member_desc(c.Structure):
    _fields_ = [("n_features_in", c.c_int32), ("n_keep", c.c_int32), ("keep", c.c_void_p), ("kind", c.c_int32),
                ("n_quantiles", c.c_int32), ("quantiles", c.c_void_p), ("safepower", c.c_void_p),
                ("svd_k", c.c_int32), ("svd_inv_scale", c.c_void_p), ("svd_vt", c.c_void_p),
                ("fingerprint", c.c_int32), ("n_out", c.c_int32), ("perm", c.c_void_p)]


# ---- Yeo-Johnson (torch, fp64, broadcast over columns) ------------------------------------------------------------
def yeo_johnson(x: torch.Tensor, lam: torch.Tensor) -> torch.Tensor:
    lam = lam.to(x.dtype)
    pos = x >= 0
    xp = torch.where(pos, x, torch.zeros_like(x))
    xn = torch.where(pos, torch.zeros_like(x), x)
    l0 = lam.abs() < 1e-12
    l2 = (lam - 2).abs() < 1e-12
    safe_l = torch.where(l0, torch.ones_like(lam), lam)
    safe_2 = torch.where(l2, torch.ones_like(lam), 2 - lam)
    out_p = torch.where(l0, torch.log1p(xp), (torch.pow(xp + 1, safe_l) - 1) / safe_l)
    out_n = torch.where(l2, -torch.log1p(-xn), -(torch.pow(1 - xn, safe_2) - 1) / safe_2)
    return torch.where(pos, out_p, out_n)


def yeo_johnson_inverse(y: torch.Tensor, lam: float) -> torch.Tensor:
    """NaN where the inverse does not exist."""
    y = y.double()
    pos = y >= 0
    nan = torch.full_like(y, float("nan"))
    if abs(lam) < 1e-12:
        out_p = torch.expm1(y)
    else:
        base = y * lam + 1
        out_p = torch.where(base > 0, torch.pow(base.clamp_min(1e-300), 1.0 / lam) - 1, nan)
    if abs(lam - 2.0) < 1e-12:
        out_n = -torch.expm1(-y)
    else:
        base = -(2 - lam) * y + 1
        out_n = torch.where(base > 0, 1 - torch.pow(base.clamp_min(1e-300), 1.0 / (2 - lam)), nan)
    return torch.where(pos, out_p, out_n)


def fit_yeo_johnson_lambda(X: torch.Tensor) -> torch.Tensor:
    """Maximum-likelihood lambda per column of X [N, F] (fp64), the objective sklearn's PowerTransformer minimises
    with Brent (`-n/2 log var(psi(x, l)) + (l - 1) sum sign(x) log1p|x|`): coarse grid on [-5, 7], then a
    golden-section refinement of the best cell.  Constant columns get lambda = 1."""
    X = X.double()
    N, F = X.shape
    const = torch.sum(torch.sign(X) * torch.log1p(X.abs()), dim=0)

    def nll(lam):  # lam [..., F] -> [..., F]
        t = yeo_johnson(X.unsqueeze(0), lam.reshape(-1, 1, F))  # [G, N, F]
        var = t.var(dim=1, unbiased=False)
        out = 0.5 * N * torch.log(var.clamp_min(1e-300)) - (lam.reshape(-1, F) - 1) * const
        return torch.where(var < 1e-300, torch.full_like(out, float("inf")), out)

    grid = torch.linspace(-5.0, 7.0, 49, dtype=torch.float64, device=X.device)
    vals = torch.cat([nll(grid[i:i + 7].unsqueeze(1).expand(-1, F)) for i in range(0, 49, 7)], dim=0)  # [49, F]
    best = vals.argmin(dim=0)
    a = grid[(best - 1).clamp_min(0)]
    b = grid[(best + 1).clamp_max(48)]
    gr = (5 ** 0.5 - 1) / 2
    x1 = b - gr * (b - a)
    x2 = a + gr * (b - a)
    f1, f2 = nll(x1)[0], nll(x2)[0]
    for _ in range(48):
        left = f1 < f2
        b = torch.where(left, x2, b)
        a = torch.where(left, a, x1)
        nx1 = torch.where(left, b - gr * (b - a), x2)
        nx2 = torch.where(left, x1, a + gr * (b - a))
        nf1 = torch.where(left, nll(nx1)[0], f2)
        nf2 = torch.where(left, f1, nll(nx2)[0])
        x1, x2, f1, f2 = nx1, nx2, nf1, nf2
    lam = 0.5 * (a + b)
    return torch.where(X.std(dim=0, unbiased=False) > 0, lam, torch.ones_like(lam))


def rebin_tables(member_borders: torch.Tensor, common_borders: torch.Tensor):
    """(idx int32 [B+1], frac fp32 [B+1], valid uint8 [B]) for `pfn_ensemble_combine` (include/npe_pfn_b200.h)."""
    b = member_borders.double()
    z = common_borders.double().to(b.device)
    fin = torch.isfinite(b)
    bf = torch.where(fin, b, torch.zeros_like(b))
    valid = fin[:-1] & fin[1:] & ((bf[1:] - bf[:-1]) > 0)
    vidx = torch.nonzero(valid).flatten()
    if vidx.numel() == 0 or not bool(torch.all(vidx[1:] - vidx[:-1] == 1)):
        raise RuntimeError("target transform leaves no contiguous range of valid buckets")
    lo, hi = int(vidx[0]), int(vidx[-1])
    edges = b[lo:hi + 2].contiguous()
    j = (torch.searchsorted(edges, z, right=True) - 1).clamp(0, hi - lo)
    frac = ((z - edges[j]) / (edges[j + 1] - edges[j])).clamp(0.0, 1.0)
    return (j + lo).to(torch.int32), frac.to(torch.float32), valid.to(torch.uint8)


class _Member:
    """Fitted state of one member for one (context, dimension): device tables + the C descriptor pointing at them."""

    def __init__(self, spec: MemberSpec):
        self.spec = spec
        self.desc = pfn_member_desc()
        self._keepalive = []
        self.lam_y: Optional[float] = None
        self.n_out = 0

    def _set(self, field: str, t: Optional[torch.Tensor]):
        if t is not None:
            self._keepalive.append(t)
        setattr(self.desc, field, None if t is None else t.data_ptr())


class EnsembleDim:
    """All members fitted for one (context, dimension): slots `slot0 .. slot0 + E - 1` of the engine."""

    def __init__(self, engine: Engine, specs: List[MemberSpec], slot0: int):
        self.engine, self.specs, self.slot0 = engine, specs, slot0
        self.members: List[_Member] = []
        self.E = len(specs)
        self.chunk_rows = 4096

    @property
    def head_slot(self) -> int:
        """Slot whose borders are the common borders in original target units (member 0: no target transform)."""
        return self.slot0

    # -- descriptor-driven kernel call -------------------------------------------------------------------
    def _transform(self, m: _Member, X: torch.Tensor) -> torch.Tensor:
        eng = self.engine
        M = X.shape[0]
        out = torch.empty(M, m.desc.n_out, dtype=torch.float32, device=eng.device)
        eng._check(eng.lib.pfn_member_transform(eng._h, c.byref(m.desc), _ptr(X), X.stride(0) if M > 1 else X.shape[1],
                                                M, _ptr(out), m.desc.n_out, eng._stream()))
        return out

    def fit(self, X: torch.Tensor, y: torch.Tensor, n_classes: int = 0, class_seed: int = 0):
        """X [N, F] fp32 on the engine's device (rows may be strided), y [N] raw targets.  `n_classes > 0`: y holds
        class indices; member e is fitted on a seeded permutation of them (`self.class_perms[e]`, upstream's class
        shuffle) and no target transform applies."""
        eng = self.engine
        self.n_classes = int(n_classes)
        self.class_perms = []
        dev = eng.device
        N, F = X.shape
        Xd = X.double()
        yd = y.double()
        y_mean = yd.mean()
        y_std = yd.std(unbiased=False) if N > 1 else torch.zeros((), dtype=torch.float64, device=dev)
        y_mean32 = y_mean.float()
        y_std32 = (y_std + 1e-20).float() if float(y_std) > 0 else torch.zeros((), dtype=torch.float32, device=dev)
        if not bool(torch.isfinite(y_std32)) or float(y_std32) == 0.0:
            y_std32 = torch.ones((), dtype=torch.float32, device=dev)
        yz = (y.float() - y_mean32) / y_std32
        keep = torch.nonzero(Xd.std(dim=0, unbiased=False) > 0).flatten() if N > 1 else torch.arange(F, device=dev)
        if keep.numel() == 0:
            keep = torch.arange(F, device=dev)
        nk = int(keep.numel())
        Xk = Xd[:, keep]
        lam_y_cache = None
        B = eng.cfg.num_buckets
        idx = torch.full((self.E, B + 1), -1, dtype=torch.int32, device=dev)
        frac = torch.zeros(self.E, B + 1, dtype=torch.float32, device=dev)
        valid = torch.ones(self.E, B, dtype=torch.uint8, device=dev)
        common_z = eng.weights.borders.double().to(dev)
        self.members = []
        for e, spec in enumerate(self.specs):
            m = _Member(spec)
            d = m.desc
            d.n_features_in, d.n_keep, d.fingerprint = F, nk, int(spec.fingerprint)
            m._set("keep", keep.to(torch.int32).contiguous())
            if spec.x_kind == "quantile":
                d.kind = 0
                nq = max(min(N, max(N // 10, 2)), 2)
                refs = torch.linspace(0.0, 1.0, nq, dtype=torch.float64, device=dev)
                q = torch.quantile(Xk, refs, dim=0, interpolation="linear") if N > 1 else Xk.expand(nq, nk)
                d.n_quantiles = nq
                m._set("quantiles", torch.cummax(q, dim=0).values.t().contiguous().float())
                d.svd_k = 0
                n_el = 2 * nk
                if spec.svd and N > 1:
                    # base columns of the context through the same kernel (no SVD / fingerprint / shuffle yet)
                    d.n_out = n_el
                    d.fingerprint = 0
                    m._set("perm", torch.arange(n_el, dtype=torch.int32, device=dev))
                    base = self._transform(m, X).double()
                    d.fingerprint = int(spec.fingerprint)
                    k = max(1, min(N // 10 + 1, n_el // 2))
                    scale = base.std(dim=0, unbiased=False)
                    scale = torch.where(scale == 0, torch.ones_like(scale), scale)
                    # right singular vectors of the scaled base matrix = eigenvectors of its small Gram matrix
                    # (n_el x n_el, fp64): much cheaper on the device than an SVD of the tall [N, n_el] matrix
                    Z = base / scale
                    _w, vecs = torch.linalg.eigh(Z.t() @ Z)
                    vt = vecs.flip(1).t()[:k].contiguous()
                    big = vt.abs().argmax(dim=1)
                    sign = torch.sign(vt[torch.arange(k, device=dev), big])
                    sign = torch.where(sign == 0, torch.ones_like(sign), sign)
                    d.svd_k = k
                    m._set("svd_inv_scale", (1.0 / scale).float().contiguous())
                    m._set("svd_vt", (vt * sign[:, None]).float().contiguous())
                n_base = n_el + d.svd_k + int(spec.fingerprint)
            elif spec.x_kind == "none":
                d.kind = 2
                n_base = nk + int(spec.fingerprint)
            else:
                d.kind = 1
                mean_in = Xk.mean(dim=0)
                std_in = Xk.std(dim=0, unbiased=False)
                std_in = torch.where(std_in == 0, torch.ones_like(std_in), std_in)
                Z = (Xk - mean_in) / std_in
                lam = fit_yeo_johnson_lambda(Z)
                T = yeo_johnson(Z, lam)
                mean_out = T.mean(dim=0)
                std_out = T.std(dim=0, unbiased=False)
                std_out = torch.where(std_out == 0, torch.ones_like(std_out), std_out)
                m.lam_x = lam
                m._set("safepower", torch.stack([mean_in, 1.0 / std_in, lam, mean_out, 1.0 / std_out]).float().contiguous())
                n_base = nk + int(spec.fingerprint)
            perm = np.random.default_rng(spec.perm_seed).permutation(n_base)
            d.n_out = n_base
            m.n_out = n_base
            m._set("perm", torch.from_numpy(perm.astype(np.int32)).to(dev))
            # target
            if self.n_classes:
                cp = np.random.default_rng(class_seed + 7919 * (e + 1)).permutation(self.n_classes) if e else np.arange(self.n_classes)
                cp = torch.from_numpy(cp).to(dev)
                self.class_perms.append(cp)
                target = cp[y.long()].float()
            elif spec.y_kind == "none":
                target = y.float()
            else:
                if lam_y_cache is None:
                    lam_y_cache = float(fit_yeo_johnson_lambda(yz.double().unsqueeze(1))[0])
                m.lam_y = lam_y_cache
                target = yeo_johnson(yz.double(), torch.tensor([m.lam_y], dtype=torch.float64, device=dev)).float()
            slot = self.slot0 + e
            eng.prefill(slot, self._transform(m, X), target)
            if m.lam_y is not None:
                bz = yeo_johnson_inverse(eng.slot_export(slot)["borders"], m.lam_y)
                idx[e], frac[e], valid[e] = rebin_tables(bz, common_z)
            self.members.append(m)
        self.idx, self.frac, self.valid = idx, frac, valid
        return self

    def class_probabilities(self, X: torch.Tensor) -> torch.Tensor:
        """Classifier ensemble: mean over members of softmax(first n_classes logits), class permutation undone."""
        eng = self.engine
        acc = None
        for e, m in enumerate(self.members):
            lg = eng.forward_logits(self.slot0 + e, self._transform(m, X))[:, :self.n_classes]
            p = torch.softmax(lg, dim=-1)[:, self.class_perms[e]]  # column c <- member column perm[c]
            acc = p if acc is None else acc + p
        return acc / self.E

    def logits(self, X: torch.Tensor) -> torch.Tensor:
        """Combined logits [M, B] (= log of the member-averaged bucket probabilities) for raw test rows X [M, F]."""
        eng = self.engine
        M = X.shape[0]
        B = eng.cfg.num_buckets
        ld = (B + 3) // 4 * 4
        out = torch.empty(M, ld, dtype=torch.float32, device=eng.device)
        rows = min(self.chunk_rows, max(M, 1))
        buf = torch.empty(self.E, rows, ld, dtype=torch.float32, device=eng.device)
        for r0 in range(0, M, rows):
            r1 = min(r0 + rows, M)
            n = r1 - r0
            Xc = X[r0:r1]
            for e, m in enumerate(self.members):
                Xt = self._transform(m, Xc)
                eng._check(eng.lib.pfn_forward_logits(eng._h, self.slot0 + e, _ptr(Xt), Xt.stride(0) if n > 1 else Xt.shape[1],
                                                      n, _ptr(buf[e]), ld, eng._stream()))
            eng._check(eng.lib.pfn_ensemble_combine(eng._h, _ptr(buf), ld, buf.stride(0), self.E, n, _ptr(self.idx),
                                                    _ptr(self.frac), _ptr(self.valid), _ptr(out[r0:r1]), ld,
                                                    eng._stream()))
        return out if ld == B else out[:, :B]

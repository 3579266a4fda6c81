"""Sample-set comparison helpers for the parity tests (test infrastructure).

`c2st` restates the reference's classifier two-sample test recipe
(`/root/reference/scripts/evaluate_ropefm_batched.py:343-475`, `classifier_two_samples_test_torch` with its default
model `DefaultMLP` :255-293): z-score both sets with the statistics of the first, stratified k-fold with shuffling,
an MLP `dim -> 4 dim -> 8 dim -> 8 dim -> 4 dim -> 2` with ReLU, Adam (lr 1e-3) on the cross-entropy, accuracy on the
held-out fold averaged over folds.  0.5 = indistinguishable, 1.0 = perfectly separable.  `epochs` and `batch_size` are
the recipe's `training_kwargs` (defaults there: 100 / 128); the tests pass smaller values to bound their run time.

`ks_pvalues` is the per-dimension two-sample Kolmogorov-Smirnov check of
`/root/reference/notebooks/benchmark_sample_batched.ipynb` (cell "Statistical test").
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


def _mlp(dim: int, hidden_mult: int = 8) -> nn.Module:
    h = hidden_mult * dim
    return nn.Sequential(nn.Linear(dim, h // 2), nn.ReLU(), nn.Linear(h // 2, h), nn.ReLU(), nn.Linear(h, h), nn.ReLU(),
                         nn.Linear(h, h // 2), nn.ReLU(), nn.Linear(h // 2, 2))


def c2st(X: torch.Tensor, Y: torch.Tensor, seed: int = 1, n_folds: int = 5, epochs: int = 100, batch_size: int = 128,
         lr: float = 1e-3, device=None) -> float:
    from sklearn.model_selection import StratifiedKFold
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    torch.manual_seed(seed)
    X, Y = X.detach().float().cpu(), Y.detach().float().cpu()
    mean, std = X.mean(0), X.std(0)
    std = torch.where(std == 0, torch.ones_like(std), std)
    X, Y = (X - mean) / std, (Y - mean) / std
    data = torch.cat([X, Y]).to(device)
    labels = torch.cat([torch.zeros(len(X)), torch.ones(len(Y))]).long().to(device)
    labels_np = labels.cpu().numpy()
    scores = []
    gen = torch.Generator().manual_seed(seed)
    for tr, te in StratifiedKFold(n_splits=n_folds, shuffle=True, random_state=seed).split(np.zeros(len(labels_np)), labels_np):
        tr_t, te_t = torch.as_tensor(tr, device=device), torch.as_tensor(te, device=device)
        net = _mlp(data.shape[1]).to(device)
        opt = torch.optim.Adam(net.parameters(), lr=lr)
        loss_fn = nn.CrossEntropyLoss()
        net.train()
        for _ in range(epochs):
            perm = tr_t[torch.randperm(len(tr_t), generator=gen).to(device)]
            for i in range(0, len(perm), batch_size):
                idx = perm[i:i + batch_size]
                opt.zero_grad()
                loss_fn(net(data[idx]), labels[idx]).backward()
                opt.step()
        net.eval()
        with torch.no_grad():
            scores.append(float((net(data[te_t]).argmax(1) == labels[te_t]).float().mean()))
    return float(np.mean(scores))


def ks_pvalues(a: torch.Tensor, b: torch.Tensor):
    from scipy import stats
    a, b = a.detach().cpu().numpy(), b.detach().cpu().numpy()
    return [float(stats.ks_2samp(a[:, d], b[:, d]).pvalue) for d in range(a.shape[1])]


def two_moons_simulator(theta: torch.Tensor, generator=None) -> torch.Tensor:
    """The simulator of `/root/reference/demo.ipynb` (cell 2): crescent of radius N(0.1, 0.01) plus a folded rotation."""
    n = theta.shape[0]
    a = (torch.rand(n, generator=generator) - 0.5) * np.pi
    r = 0.1 + 0.01 * torch.randn(n, generator=generator)
    p = torch.stack([r * torch.cos(a) + 0.25, r * torch.sin(a)], 1)
    q = torch.stack([-(theta[:, 0] + theta[:, 1]).abs() / np.sqrt(2), (-theta[:, 0] + theta[:, 1]) / np.sqrt(2)], 1)
    return p + q

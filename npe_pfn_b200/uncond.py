"""Unconditional density estimator on top of the same hot loops (experimental in the reference too).

Behaviour of `/root/reference/npe_pfn/npe_pfn.py:747-900` (`TabPFN_Based_Uncond_Estimator`): the "observation" is
a dummy 1-D standard-normal column, theta is (optionally) split into k-means clusters, and a mixture over clusters
is sampled / evaluated: cluster weights = cluster sizes, per-cluster density = the autoregressive estimator with
that cluster's (at most 10 000, shuffled) points as context.
"""
from __future__ import annotations

from typing import Mapping

import numpy as np
import torch
from torch import Tensor

from .npe_pfn import NPE_PFN_Core


class TabPFN_Based_Uncond_Estimator(NPE_PFN_Core):
    def __init__(self, num_clusters: int = 1, show_progress_bars: bool = False, regressor_init_kwargs: Mapping = {},
                 classifier_init_kwargs: Mapping = {}):
        super().__init__(prior=None, show_progress_bars=show_progress_bars, regressor_init_kwargs=regressor_init_kwargs,
                         classifier_init_kwargs=classifier_init_kwargs)
        self.context_size = 10_000  # anything beyond is sliced off (npe_pfn.py:764-765)
        self.num_clusters = num_clusters
        self.cluster_state = 0
        self.kmeans = None
        self.counts = None

    def set_cluster_state(self, cluster_idx: int = 0):
        self.cluster_state = cluster_idx

    # the context is the current cluster's points; the K/V caches are keyed by the cluster as well
    def get_context(self, x: Tensor):
        members = torch.from_numpy(self.kmeans.labels_ == self.cluster_state)
        return self._theta_train[members][: self.context_size], self._x_train[members][: self.context_size]

    def _context_key(self, x: Tensor):
        return (self._ctx_version, "cluster", self.cluster_state)

    def append_simulations(self, theta: Tensor, x: Tensor = None):
        from sklearn.cluster import KMeans
        self._theta_train = None
        self._x_train = None
        theta = self._validate_theta(theta)
        self._theta_train = theta[torch.randperm(theta.shape[0])]  # shuffled: only the first 10k per cluster are used
        self._x_train = torch.randn(theta.shape[0], 1)
        self.kmeans = KMeans(n_clusters=self.num_clusters).fit(self._theta_train.cpu().numpy())
        _labels, counts = np.unique(self.kmeans.labels_, return_counts=True)
        assert np.min(counts) > 1, "Too few samples in some clusters, need at least 2."
        self.counts = counts
        self._ctx_version += 1
        self._ctx = None
        return self

    def _weights(self) -> np.ndarray:
        return self.counts / self.counts.sum()

    def sample(self, sample_shape=torch.Size(), x=None, max_sampling_batch_size=10000, with_log_prob=False, eps=1e-15):
        per_cluster = np.random.multinomial(torch.Size(sample_shape)[0], self._weights())
        draws, lps = [], []
        try:
            for k, n in enumerate(per_cluster):
                if n == 0:
                    continue
                self.set_cluster_state(k)
                s, lp = self._sample(max_sampling_batch_size, torch.randn(int(n), 1), repeat_x=False,
                                     with_log_prob=with_log_prob, eps=eps)
                draws.append(s)
                lps.append(lp)
        finally:
            self.set_cluster_state()
        samples = torch.cat(draws, dim=0)
        order = torch.randperm(samples.shape[0])
        if with_log_prob:
            return samples[order], torch.cat(lps, dim=0)[order]
        return samples[order]

    def log_prob(self, theta: Tensor, x=None, max_sampling_batch_size=10000, mode="autoregressive", eps=1e-15,
                 **ratio_kwargs):
        if mode not in ("autoregressive", "ratio_based"):
            raise ValueError(f"Invalid mode: {mode}")
        theta = self._validate_theta(theta)
        labels = torch.from_numpy(self.kmeans.predict(theta.cpu().numpy()))
        log_w = np.log(self._weights())
        out = torch.zeros(theta.shape[0])
        try:
            for k in range(self.num_clusters):
                mine = labels == k
                th_k = theta[mine]
                if th_k.shape[0] == 0:
                    continue
                self.set_cluster_state(k)
                lp_k = torch.zeros(th_k.shape[0])
                for i in range(0, th_k.shape[0], max_sampling_batch_size):
                    part = th_k[i:i + max_sampling_batch_size]
                    if mode == "autoregressive":
                        lp_k[i:i + max_sampling_batch_size] = self._autoregressive_log_prob(
                            part, torch.randn(part.shape[0], 1), repeat_x=False, eps=eps)
                    else:  # the classifier is fitted on unconditional draws of this estimator
                        lp_k[i:i + max_sampling_batch_size] = self._ratio_based_log_prob(
                            part, torch.zeros(1, 1), eps=eps, **ratio_kwargs)
                out[mine] = lp_k + float(log_w[k])
        finally:
            self.set_cluster_state()
        return out

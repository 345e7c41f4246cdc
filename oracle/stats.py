"""Oracle: streaming mean / covariance (reference data.py:37-58, 77-106)."""
from __future__ import annotations

import numpy as np


def batch_stats(x: np.ndarray, compute_dtype=None):
    """mean and unbiased covariance of one batch, as AudioMetricsData.add does.

    data.py:39  mean = torch.mean(embeddings, 0).to(float64)   (computed in the input dtype)
    data.py:40-44  n == 1 -> zeros((d, d)); else torch.cov(embeddings.T) (correction=1,
    centred product computed in the input dtype) .to(float64).
    ``compute_dtype=np.float64`` gives the noise-free variant used to bound the
    reference's own fp32 rounding.
    """
    x = np.asarray(x)
    cd = x.dtype if compute_dtype is None else compute_dtype
    xc = x.astype(cd, copy=False)
    n, d = xc.shape
    mean = xc.mean(axis=0, dtype=cd)
    if n == 1:
        cov = np.zeros((d, d), dtype=np.float64)
    else:
        c = xc - mean
        cov = (c.T @ c) / cd.type(n - 1) if hasattr(cd, "type") else (c.T @ c) / (n - 1)
        cov = cov.astype(np.float64)
    return mean.astype(np.float64), cov, n


def chan_merge(n1, mean1, cov1, n2, mean2, cov2):
    """Pairwise merge of two (n, mean, cov) triples (data.py:77-94 _update_stats)."""
    if n1 is None:
        return n2, mean2, cov2
    n_prod = n1 * n2
    n_total = n1 + n2
    new_mean = (n1 * mean1 + n2 * mean2) / n_total
    diff = mean1 - mean2
    diff_mat = np.einsum("i,j->ij", diff, diff)
    w_self = (n1 - 1) / (n_total - 1)
    w_other = (n2 - 1) / (n_total - 1)
    w_diff = (n_prod / n_total) / (n_total - 1)
    new_cov = w_self * cov1 + w_other * cov2 + w_diff * diff_mat
    return n_total, new_mean, new_cov


class StreamingStats:
    """AudioMetricsData's statistics state (data.py:18-58, 96-106), embeddings optional."""

    def __init__(self, store_embeddings=True):
        self.n = None
        self.mean = None
        self.cov = None
        self.store_embeddings = store_embeddings
        self.embeddings = None

    def add(self, x):
        mean, cov, n = batch_stats(x)
        self.n, self.mean, self.cov = chan_merge(self.n, self.mean, self.cov, n, mean, cov)
        if self.store_embeddings:  # data.py:68-72
            x = np.asarray(x)
            self.embeddings = x.copy() if self.embeddings is None else np.concatenate((self.embeddings, x))

    def recompute_stats(self):
        """data.py:49-58 (n == 1 gives a (1, 1) zero matrix there — shape quirk kept)."""
        if self.embeddings is not None:
            self.n = len(self.embeddings)
            if self.n == 1:
                self.mean = self.embeddings.mean(axis=0).astype(np.float64)
                self.cov = np.zeros((1, 1))
            else:
                self.mean, self.cov, _ = batch_stats(self.embeddings)

    def merge(self, other: "StreamingStats"):
        """data.py:96-106 __iadd__."""
        if other.n is None:
            return self
        if self.n is None:
            self.store_embeddings = other.store_embeddings
        assert self.store_embeddings == other.store_embeddings
        self.n, self.mean, self.cov = chan_merge(self.n, self.mean, self.cov, other.n, other.mean, other.cov)
        if self.store_embeddings:
            self.embeddings = (other.embeddings.copy() if self.embeddings is None
                               else np.concatenate((self.embeddings, other.embeddings)))
        return self

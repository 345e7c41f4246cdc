"""PCA projection on the device (reference projection.py:6-46: sklearn's IncrementalPCA with
torch-serialisable state; ``transform`` returns a float64 tensor).

Same attributes, same state-dict keys (state files written by the reference load here and
vice versa), same numbers — but no host round trip and no sklearn on the path:

  partial_fit   first call: sklearn centres the batch and takes its SVD (X - mean = U S Vt;
                components_ = Vt[:k], singular_values_ = S[:k]).  Vt and S^2 are the eigenvectors
                and eigenvalues of the d x d scatter matrix (n - 1) cov, which the covariance
                kernels already produce from one pass over the embeddings, so the fit is
                ``amb_sym_eig`` (one-sided Jacobi, fp64) of that matrix.  Later calls: sklearn
                stacks [S_old * components_old ; X - batch mean ; sqrt(n_old n_b / n) (mean_old -
                mean_b)] and takes the SVD of that — equally the eigen-decomposition of
                  components_old^T diag(S_old^2) components_old + scatter_b + (n_old n_b / n) dd^T
  transform     (X - mean_) @ components_.T as one HBM-bound pass (``amb_pca_transform``), fp64 out.

Sign convention: sklearn's ``svd_flip(u_based_decision=False)`` — the entry of largest magnitude
of every component is positive.
"""
from __future__ import annotations

import torch

from . import _lib

_ARRAYS = ("components_", "mean_", "var_", "singular_values_", "explained_variance_", "explained_variance_ratio_")


class IncrementalPCA:
    def __init__(self, n_components=None, *, whiten=False, copy=True, batch_size=None, device=None):
        self.n_components = n_components
        self.whiten = whiten
        self.copy = copy
        self.batch_size = batch_size
        self._device = device

    # ------------------------------------------------------------------ helpers
    @property
    def device(self):
        if self._device is None or not isinstance(self._device, torch.device):
            self._device = _lib.require_cuda(self._device)
        return self._device

    def _stats(self, x):
        """(n, mean [d], cov [d, d]) of a batch, fp64 on the device."""
        from .data import AudioMetricsData

        if isinstance(x, AudioMetricsData):
            c = x
        else:
            c = AudioMetricsData(store_embeddings=False, device=self.device)
            c.add(x)
        # under torch.distributed the containers hold each rank's rows: fit on the statistics of the whole set
        from .dist import global_stats

        n, mean, cov = global_stats(c)
        d = mean.shape[0]
        if tuple(cov.shape) != (d, d):
            cov = torch.zeros((d, d), dtype=torch.float64, device=self.device)
        return n, mean, cov

    # ---------------------------------------------------------------------- fit
    def partial_fit(self, x, y=None, check_input=True):
        """sklearn IncrementalPCA.partial_fit.  ``x``: [n, d] embeddings (tensor / array) or an
        AudioMetricsData whose statistics are then reused instead of recomputed."""
        dev = self.device
        n_b, mean_b, cov_b = self._stats(x)
        d = mean_b.shape[0]
        first = not hasattr(self, "components_")
        if first:
            self.n_samples_seen_ = 0
        if self.n_components is None:
            k = min(n_b, d) if first else self.components_.shape[0]
        elif not self.n_components <= d:
            raise ValueError(f"n_components={self.n_components} invalid for n_features={d}, need more rows than columns "
                             "for IncrementalPCA processing")
        elif self.n_components > n_b and first:
            raise ValueError(f"n_components={self.n_components} must be less or equal to the batch number of "
                             f"samples {n_b} for the first partial_fit call.")
        else:
            k = self.n_components
        if not first and self.components_.shape[0] != k:
            raise ValueError("Number of input features has changed")
        self.n_components_ = k
        self.n_features_in_ = d

        scatter = cov_b * max(n_b - 1, 0)                              # (X - mean_b)^T (X - mean_b)
        n_old = self.n_samples_seen_
        n_total = n_old + n_b
        if n_old == 0:
            mean, var_sum = mean_b, scatter.diagonal().clone()         # var_ is the biased variance (np.var)
        else:
            delta = self.mean_ - mean_b
            mean = (self.mean_ * n_old + mean_b * n_b) / n_total
            var_sum = self.var_ * n_old + scatter.diagonal() + delta * delta * (n_old * n_b / n_total)
            s2 = self.singular_values_ ** 2
            scatter = (self.components_.T * s2) @ self.components_ + scatter \
                + torch.outer(delta, delta) * (n_old * n_b / n_total)
        L = _lib.lib()
        evals = torch.empty(d, dtype=torch.float64, device=dev)
        evecs = torch.empty((d, d), dtype=torch.float64, device=dev)
        scatter = scatter.contiguous()
        ws = _lib.workspace(L.amb_sym_eig_ws_bytes(d), dev)
        _lib.check(L.amb_sym_eig(dev.index, _lib.stream_ptr(dev), d, scatter.data_ptr(), evals.data_ptr(),
                                 evecs.data_ptr(), ws.data_ptr(), ws.numel()))
        # sklearn keeps min(rows of the stacked matrix, d) singular values
        n_sv = min(d, n_b if n_old == 0 else k + n_b + 1)
        s2 = evals[:n_sv].clamp_min(0)
        explained_variance = s2 / (n_total - 1)
        explained_variance_ratio = s2 / var_sum.sum()
        self.n_samples_seen_ = n_total
        self.components_ = evecs[:k].contiguous()
        self.singular_values_ = s2[:k].sqrt()
        self.mean_ = mean
        self.var_ = var_sum / n_total
        self.explained_variance_ = explained_variance[:k].clone()
        self.explained_variance_ratio_ = explained_variance_ratio[:k].clone()
        if k not in (n_b, d):                                          # sklearn: noise variance of the discarded axes
            self.noise_variance_ = float(explained_variance[k:].mean())
        else:
            self.noise_variance_ = 0.0
        return self

    def fit(self, x, y=None):
        for a in _ARRAYS + ("n_samples_seen_", "noise_variance_", "n_components_", "n_features_in_"):
            if hasattr(self, a):
                delattr(self, a)
        return self.partial_fit(x)

    # ---------------------------------------------------------------- transform
    def transform(self, x):
        """projection.py:20-21 — [n, n_components] float64, on the device."""
        if not hasattr(self, "components_"):
            raise RuntimeError("This IncrementalPCA instance is not fitted yet")
        dev = self.device
        x = _lib.as_device_matrix(x, dev)
        n, d = x.shape
        k = self.components_.shape[0]
        if d != self.components_.shape[1]:
            raise ValueError(f"X has {d} features, but IncrementalPCA is expecting {self.components_.shape[1]}")
        out = torch.empty((n, k), dtype=torch.float64, device=dev)
        comp = self.components_.to(dev, torch.float64).contiguous()
        mean = self.mean_.to(dev, torch.float64).contiguous()
        _lib.check(_lib.lib().amb_pca_transform(dev.index, _lib.stream_ptr(dev), x.data_ptr(), _lib.dtype_code(x), n, d,
                                                x.stride(0), mean.data_ptr(), comp.data_ptr(), k, out.data_ptr()))
        if self.whiten:
            out = out / self.explained_variance_.sqrt()
        return out

    def fit_transform(self, x, y=None):
        return self.fit(x).transform(x)

    # -------------------------------------------------------------------- state
    def __getstate__(self):
        """projection.py:23-33: sklearn's parameters and fitted attributes, arrays as CPU tensors."""
        state = {"n_components": self.n_components, "whiten": self.whiten, "copy": self.copy,
                 "batch_size": self.batch_size}
        for k in ("n_features_in_", "n_components_", "n_samples_seen_"):
            if hasattr(self, k):
                state[k] = int(getattr(self, k))
        for k in _ARRAYS:
            if hasattr(self, k):
                state[k] = getattr(self, k).detach().to("cpu", torch.float64)
        if hasattr(self, "noise_variance_"):
            state["noise_variance_"] = float(self.noise_variance_)
        return state

    def __setstate__(self, state):
        """projection.py:35-46; accepts what the reference wrote."""
        state = dict(state)
        device = state.pop("_device", getattr(self, "_device", None))
        state.pop("_sklearn_version", None)
        self._device = device
        for k, v in state.items():
            if k in _ARRAYS:
                v = torch.as_tensor(v).to(self.device, torch.float64)
            elif k in ("n_samples_seen_", "n_features_in_", "n_components_"):
                v = int(v)
            elif k == "noise_variance_":
                v = float(v)
            setattr(self, k, v)

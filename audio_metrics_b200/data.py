"""``AudioMetricsData`` — the reference's set-statistics container (data.py:18-112)
with the same public attributes and methods, backed by the CUDA library.

State mirrors the reference field for field so that saved state files
interoperate (``serialize`` / ``deserialize``, data.py:28-35): ``mean`` [d] fp64,
``cov`` [d, d] fp64, ``n`` int, ``store_embeddings``, ``embeddings`` [n, d] in the
input dtype, ``radii`` {"radii_<k>": [n]}, ``dtype``.  Tensors live on the CUDA
device; ``serialize()`` hands out CPU copies.

Differences, all supersets of the reference behaviour:
  * embeddings are appended into a geometrically grown device buffer instead of
    ``torch.cat`` per batch (data.py:68-72 re-copies the whole store every call);
  * batch statistics are accumulated in fp64 (the reference computes them in the
    input dtype, then casts — data.py:39,44);
  * the radii cache is dropped when embeddings are added (the reference never
    invalidates it, so a second ``add_reference`` followed by ``evaluate`` fails).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def ensure_tensor(x, device=None):
    """data.py:6-9."""
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    return x.to(device, non_blocking=True) if device else x


def ensure_ndarray(x):
    """data.py:12-15."""
    if isinstance(x, torch.Tensor):
        x = x.cpu().numpy()
    return x


class AudioMetricsData:
    def __init__(self, store_embeddings=True, device=None):
        self.mean = None
        self.n = None
        self.cov = None
        self.store_embeddings = store_embeddings
        self.radii = {}
        self.dtype = torch.float64
        self._device = device
        self._buf = None       # capacity-grown embedding store [cap, d]
        self._n_stored = 0
        self._packed = None    # cached tensor-core operand blob for the current embeddings

    # ------------------------------------------------------------------ state
    @property
    def device(self) -> torch.device:
        if self._device is None or not isinstance(self._device, torch.device):
            self._device = _lib.require_cuda(self._device)
        return self._device

    @property
    def embeddings(self):
        if self._buf is None:
            return None
        return self._buf[: self._n_stored]

    @embeddings.setter
    def embeddings(self, value):
        if value is None:
            self._buf = None
            self._n_stored = 0
        else:
            value = _lib.as_device_matrix(value, self.device)
            self._buf = value
            self._n_stored = value.shape[0]
        self._invalidate()

    def _invalidate(self):
        self.radii = {}
        self._packed = None

    def serialize(self):
        """data.py:28-29 — the reference's field names, CPU tensors."""
        cpu = lambda t: None if t is None else t.detach().cpu()
        return {
            "mean": cpu(self.mean),
            "n": self.n,
            "cov": cpu(self.cov),
            "store_embeddings": self.store_embeddings,
            "embeddings": cpu(self.embeddings),
            "radii": {k: cpu(v) for k, v in self.radii.items()},
            "dtype": self.dtype,
        }

    @classmethod
    def deserialize(cls, state, device=None):
        """data.py:31-35."""
        self = cls(device=device)
        self.store_embeddings = state.get("store_embeddings", True)
        self.n = state.get("n")
        dev = self.device
        self.mean = None if state.get("mean") is None else ensure_tensor(state["mean"]).to(dev, torch.float64)
        self.cov = None if state.get("cov") is None else ensure_tensor(state["cov"]).to(dev, torch.float64)
        emb = state.get("embeddings")
        if emb is not None:
            self.embeddings = emb
        self.radii = {k: ensure_tensor(v).to(dev) for k, v in (state.get("radii") or {}).items()}
        self.dtype = state.get("dtype", torch.float64)
        return self

    def __len__(self):
        return self.n or 0

    # ------------------------------------------------------------- statistics
    def _batch_stats(self, x: torch.Tensor):
        """fp64 (mean, cov) of one [n, d] device batch — data.py:38-44."""
        dev = self.device
        n, d = x.shape
        L = _lib.lib()
        sums = torch.zeros(d, dtype=torch.float64, device=dev)
        gram = torch.zeros((d, d), dtype=torch.float64, device=dev)
        mean = torch.empty(d, dtype=torch.float64, device=dev)
        cov = torch.empty((d, d), dtype=torch.float64, device=dev)
        ws = _lib.workspace(L.amb_cov_ws_bytes(n, d), dev)
        st = _lib.stream_ptr(dev)
        _lib.check(L.amb_cov_accumulate(dev.index, st, x.data_ptr(), _lib.dtype_code(x), n, d, x.stride(0),
                                        sums.data_ptr(), gram.data_ptr(), ws.data_ptr(), ws.numel()))
        _lib.check(L.amb_cov_finalize(dev.index, st, n, d, sums.data_ptr(), gram.data_ptr(), mean.data_ptr(),
                                      cov.data_ptr()))
        return mean, cov

    def add(self, embeddings):
        """data.py:37-47."""
        x = _lib.as_device_matrix(embeddings, self.device)
        n = len(x)
        if n == 0:
            return
        mean, cov = self._batch_stats(x)
        self._update_stats(mean, cov, n)
        if self.store_embeddings:
            self._update_embeddings(x)

    def recompute_stats(self):
        """data.py:49-58 (n == 1 yields a (1, 1) zero covariance there; kept)."""
        if self.embeddings is not None:
            self.n = len(self.embeddings)
            mean, cov = self._batch_stats(self.embeddings)
            self.mean = mean
            self.cov = torch.zeros((1, 1), dtype=self.dtype, device=self.device) if self.n == 1 else cov

    def get_radii(self, k_neighbor):
        """data.py:60-66."""
        key = f"radii_{k_neighbor}"
        radii = self.radii.get(key)
        if radii is None and self.embeddings is not None:
            from .metrics.prdc import nearest_neighbour_distances

            radii = nearest_neighbour_distances(self, k_neighbor)
            self.radii[key] = radii
        return radii

    def packed(self) -> torch.Tensor:
        """Tensor-core operand blob (fp16 hi/lo planes + norms) of the stored embeddings."""
        if self._packed is None:
            x = self.embeddings
            if x is None:
                raise ValueError("no embeddings stored")
            dev = self.device
            L = _lib.lib()
            n, d = x.shape
            blob = _lib.workspace(L.amb_packed_bytes(n, d), dev)
            _lib.check(L.amb_pack(dev.index, _lib.stream_ptr(dev), x.data_ptr(), _lib.dtype_code(x), n, d,
                                  x.stride(0), blob.data_ptr()))
            self._packed = blob
        return self._packed

    def _update_embeddings(self, x: torch.Tensor):
        """data.py:68-72, amortised O(1) append."""
        if self._buf is None:
            self._buf = x.clone()
            self._n_stored = x.shape[0]
        else:
            if x.dtype != self._buf.dtype:
                x = x.to(self._buf.dtype)
            need = self._n_stored + x.shape[0]
            if need > self._buf.shape[0]:
                cap = max(need, int(self._buf.shape[0] * 2))
                grown = torch.empty((cap, self._buf.shape[1]), dtype=self._buf.dtype, device=self._buf.device)
                grown[: self._n_stored] = self._buf[: self._n_stored]
                self._buf = grown
            self._buf[self._n_stored:need] = x
            self._n_stored = need
        self._invalidate()

    def _update_stats(self, mean, cov, n):
        """data.py:77-94 (Chan merge, on device)."""
        if self.n is None:
            self.mean = mean
            self.cov = cov
            self.n = n
            return
        dev = self.device
        d = mean.shape[0]
        if self.cov.shape != (d, d):   # (1, 1) quirk after recompute_stats with one row
            self.cov = torch.zeros((d, d), dtype=torch.float64, device=dev)
        scratch = torch.empty(d, dtype=torch.float64, device=dev)
        mean = mean.to(dev, torch.float64).contiguous()
        cov = cov.to(dev, torch.float64).contiguous()
        self.mean = self.mean.to(dev, torch.float64).contiguous()
        self.cov = self.cov.to(dev, torch.float64).contiguous()
        _lib.check(_lib.lib().amb_stats_merge(dev.index, _lib.stream_ptr(dev), d, self.n, self.mean.data_ptr(),
                                              self.cov.data_ptr(), n, mean.data_ptr(), cov.data_ptr(),
                                              scratch.data_ptr()))
        self.n = self.n + n

    def __iadd__(self, other):
        """data.py:96-106."""
        assert isinstance(other, AudioMetricsData)
        if other.n is None:
            return self
        if self.n is None:
            self.store_embeddings = other.store_embeddings
        assert self.store_embeddings == other.store_embeddings
        self._update_stats(other.mean.clone(), other.cov.clone(), other.n)
        if self.store_embeddings:
            self._update_embeddings(other.embeddings.to(self.device))
        return self

    def __add__(self, other):
        """data.py:108-112."""
        new = AudioMetricsData(device=self._device)
        new += self
        new += other
        return new

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
  * statistics are DEFERRED: ``add`` only appends (or, without an embedding store,
    adds the batch's raw fp64 moments to an accumulator — one kernel launch), and
    ``mean`` / ``cov`` fold everything outstanding in when they are read.  The
    reference finalises and Chan-merges a d x d covariance on every 32-row batch
    (embed.py:226-236 -> data.py:37-47; 10 s for 200k rows).  Values agree with the
    per-batch merge to fp64 round-off (the reference computes batch statistics in
    the input dtype, then casts — data.py:39,44);
  * host inputs are copied on a separate copy stream; kernels that read the store
    wait for that copy on the device, so an ``add`` from pinned memory returns
    immediately and the copy overlaps whatever the GPU is doing;
  * the radii cache is dropped when embeddings are added (the reference never
    invalidates it, so a second ``add_reference`` followed by ``evaluate`` fails).
"""
from __future__ import annotations

import torch

from . import _lib

_COPY_STREAMS = {}


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    """One host-to-device copy stream per device for the life of the process (a fresh pooled
    stream per call eventually aliases the main stream's hardware queue)."""
    s = _COPY_STREAMS.get(dev.index)
    if s is None:
        s = _COPY_STREAMS[dev.index] = torch.cuda.Stream(dev)
    return s


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
    # Batches of at most this many rows go through the one-launch fp64 moment kernel when there
    # is no embedding store to defer to; larger ones through amb_cov_accumulate.
    def __init__(self, store_embeddings=True, device=None):
        self.store_embeddings = store_embeddings
        self.radii = {}
        self.dtype = torch.float64
        self._device = device
        self._mean = None      # statistics of the `_n - outstanding` rows folded in so far
        self._cov = None
        self._n = None         # rows added (python int; never needs the device)
        self._n_folded = 0
        self._buf = None       # capacity-grown embedding store [cap, d]
        self._n_stored = 0
        self._lazy_from = None # first stored row whose statistics are still outstanding
        self._acc = None       # raw fp64 moments [d + d*d] of outstanding rows that are not stored
        self._acc_n = 0
        self._ready = None     # (event, stream): host-to-device copy of the store still in flight
        self._packed = None    # cached tensor-core operand blob for the current embeddings
        self._cache = {}       # derived data keyed by the evaluation code (gathered sets, ...)

    # ------------------------------------------------------------------ state
    @property
    def device(self) -> torch.device:
        if self._device is None or not isinstance(self._device, torch.device):
            self._device = _lib.require_cuda(self._device)
        return self._device

    def _wait_ready(self):
        """Order the current stream after an in-flight host-to-device copy of the store."""
        if self._ready is not None:
            ev, main = self._ready
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            if cur == main:
                self._ready = None

    @property
    def embeddings(self):
        if self._buf is None:
            return None
        self._wait_ready()
        return self._buf[: self._n_stored]

    @embeddings.setter
    def embeddings(self, value):
        """Replaces the store (statistics are left alone, as assigning the attribute does in the
        reference)."""
        self._ready = None
        if value is None:
            self._buf = None
            self._n_stored = 0
        else:
            value, ready = self._to_device(value)
            self._buf = value
            self._n_stored = value.shape[0]
            self._ready = ready
        self._lazy_from = None
        self._invalidate()

    @property
    def n(self):
        return self._n

    @n.setter
    def n(self, value):
        self._n = value
        self._drop_outstanding()

    @property
    def mean(self):
        self._fold()
        return self._mean

    @mean.setter
    def mean(self, value):
        self._drop_outstanding()
        self._mean = value

    @property
    def cov(self):
        self._fold()
        return self._cov

    @cov.setter
    def cov(self, value):
        self._drop_outstanding()
        self._cov = value

    def _drop_outstanding(self):
        self._lazy_from = None
        self._acc = None
        self._acc_n = 0
        self._n_folded = self._n or 0

    def _invalidate(self):
        self.radii = {}
        self._packed = None
        self._cache = {}

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
        """data.py:31-35; accepts what the reference package wrote (CPU tensors)."""
        self = cls(device=device)
        self.store_embeddings = state.get("store_embeddings", True)
        dev = self.device
        emb = state.get("embeddings")
        if emb is not None:
            self.embeddings = emb
        self._n = state.get("n")
        self._mean = None if state.get("mean") is None else ensure_tensor(state["mean"]).to(dev, torch.float64)
        self._cov = None if state.get("cov") is None else ensure_tensor(state["cov"]).to(dev, torch.float64)
        self._drop_outstanding()
        self.radii = {k: ensure_tensor(v).to(dev) for k, v in (state.get("radii") or {}).items()}
        self.dtype = state.get("dtype", torch.float64)
        return self

    def __len__(self):
        return self.n or 0

    # ---------------------------------------------------------------- transfers
    def _to_device(self, x):
        """([n, d] device matrix, ready) — ``ready`` = (event, main stream) when the rows are still
        being copied from host memory on the copy stream, else None."""
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(x)
        if x.ndim != 2:
            raise ValueError(f"expected a 2-D [n, d] array of embeddings, got shape {tuple(x.shape)}")
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float32)
        dev = self.device
        if x.is_cuda or x.numel() == 0:
            return _lib.as_device_matrix(x, dev), None
        x = x.contiguous()
        main = torch.cuda.current_stream(dev)
        cs = _copy_stream(dev)
        with torch.cuda.stream(cs):
            xd = x.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        xd.record_stream(main)   # allocated on the copy stream, consumed on the main stream
        return xd, (ev, main)

    # ------------------------------------------------------------- statistics
    SMALL_BATCH = 2048     # rows up to which the one-launch moment kernel is used

    def _moments(self, x: torch.Tensor, out: torch.Tensor = None, mask=None, mask_value=0) -> torch.Tensor:
        """out[d + d*d] (+)= raw fp64 moments (column sums | Gram) of a device batch (of its rows
        with mask == mask_value when an int32 device mask is given)."""
        dev = self.device
        n, d = x.shape
        L = _lib.lib()
        if out is None:
            out = torch.zeros(d + d * d, dtype=torch.float64, device=dev)
        if mask is not None or n <= self.SMALL_BATCH:
            _lib.check(L.amb_cov_accumulate_masked(dev.index, _lib.stream_ptr(dev), x.data_ptr(), _lib.dtype_code(x),
                                                   n, d, x.stride(0), None if mask is None else mask.data_ptr(),
                                                   int(mask_value), out.data_ptr(), out[d:].data_ptr()))
            return out
        ws = _lib.workspace(L.amb_cov_ws_bytes(n, d), dev)
        _lib.check(L.amb_cov_accumulate(dev.index, _lib.stream_ptr(dev), x.data_ptr(), _lib.dtype_code(x), n, d,
                                        x.stride(0), out.data_ptr(), out[d:].data_ptr(), ws.data_ptr(), ws.numel()))
        return out

    def _finalize(self, mom: torch.Tensor, n: int, d: int):
        """(mean, cov) from raw moments — data.py:38-44 (zeros for n == 1)."""
        dev = self.device
        mean = torch.empty(d, dtype=torch.float64, device=dev)
        cov = torch.empty((d, d), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().amb_cov_finalize(dev.index, _lib.stream_ptr(dev), n, d, mom.data_ptr(),
                                               mom[d:].data_ptr(), mean.data_ptr(), cov.data_ptr()))
        return mean, cov

    def _batch_stats(self, x: torch.Tensor):
        """fp64 (mean, cov) of one [n, d] device batch — data.py:38-44."""
        n, d = x.shape
        return self._finalize(self._moments(x), n, d)

    def _fold(self):
        """Fold the outstanding rows into (mean, cov): one moment pass over the stored rows that
        were added since the last read, or the finalisation of the running accumulator."""
        if self._lazy_from is not None:
            lo, self._lazy_from = self._lazy_from, None
            self._wait_ready()
            x = self._buf[lo: self._n_stored]
            if len(x):
                mean, cov = self._batch_stats(x)
                self._merge(mean, cov, len(x))
        if self._acc_n:
            n, self._acc_n = self._acc_n, 0
            acc, self._acc = self._acc, None
            d = int(round((-1 + (1 + 4 * acc.numel()) ** 0.5) / 2))
            mean, cov = self._finalize(acc, n, d)
            self._merge(mean, cov, n)

    def add(self, embeddings):
        """data.py:37-47."""
        x, ready = self._to_device(embeddings)
        n = len(x)
        if n == 0:
            return
        if self.store_embeddings:
            first = self._n_stored
            self._update_embeddings(x, ready)
            if self._lazy_from is None:
                self._lazy_from = first
        else:
            if ready is not None:
                torch.cuda.current_stream(self.device).wait_event(ready[0])
            d = x.shape[1]
            if self._acc is not None and self._acc.numel() != d + d * d:
                raise ValueError("embedding width changed between batches")
            self._acc = self._moments(x, self._acc)
            self._acc_n += n
            self._cache = {}
        self._n = (self._n or 0) + n

    def add_masked(self, embeddings, mask, mask_value, count):
        """``add(embeddings[mask == mask_value])`` for a device batch and an int32 device mask of which
        ``count`` rows match (the caller built the mask on the host and knows the count): what the
        embedding pipeline does per item category (embed.py:231-236), without materialising the
        selection when only statistics are kept."""
        if count == 0:
            return
        if self.store_embeddings:
            sel = torch.nonzero(mask == int(mask_value)).squeeze(1)
            return self.add(embeddings.index_select(0, sel))
        x = _lib.as_device_matrix(embeddings, self.device)
        d = x.shape[1]
        if self._acc is not None and self._acc.numel() != d + d * d:
            raise ValueError("embedding width changed between batches")
        self._acc = self._moments(x, self._acc, mask=mask, mask_value=mask_value)
        self._acc_n += int(count)
        self._cache = {}
        self._n = (self._n or 0) + int(count)

    def recompute_stats(self):
        """data.py:49-58 (n == 1 yields a (1, 1) zero covariance there; kept)."""
        if self._buf is not None:
            x = self.embeddings
            self._drop_outstanding()
            self._n = self._n_folded = len(x)
            if self._n:
                self._mean, cov = self._batch_stats(x)
                self._cov = torch.zeros((1, 1), dtype=self.dtype, device=self.device) if self._n == 1 else cov

    def get_radii(self, k_neighbor):
        """data.py:60-66."""
        key = f"radii_{k_neighbor}"
        radii = self.radii.get(key)
        if radii is None and self._buf is not None:
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

    def _update_embeddings(self, x: torch.Tensor, ready=None):
        """data.py:68-72, amortised O(1) append."""
        if self._buf is None:
            if ready is not None:
                self._buf = x            # our own fresh copy of host data: no clone needed
                self._ready = ready
            else:
                self._buf = x.clone()    # the caller keeps its tensor (data.py:70)
            self._n_stored = x.shape[0]
        else:
            cur = torch.cuda.current_stream(self.device)
            if ready is not None:
                cur.wait_event(ready[0])
            self._wait_ready()
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

    def _merge(self, mean, cov, n):
        """data.py:77-94 (Chan merge, on device) of a block of ``n`` rows into the folded statistics."""
        if self._mean is None or not self._n_folded:
            self._mean = mean
            self._cov = cov
            self._n_folded = n
            return
        dev = self.device
        d = mean.shape[0]
        # (1, 1) quirk after recompute_stats with one row (data.py:56), on either side of the merge:
        # the reference broadcasts that zero (its weight (n - 1) / (n_total - 1) is 0 anyway)
        if self._cov.shape != (d, d):
            self._cov = torch.zeros((d, d), dtype=torch.float64, device=dev)
        if tuple(cov.shape) != (d, d):
            if cov.numel() != 1:
                raise ValueError(f"covariance of shape {tuple(cov.shape)} cannot be merged into {(d, d)}")
            cov = torch.zeros((d, d), dtype=torch.float64, device=dev) + cov.to(dev, torch.float64).reshape(())
        scratch = torch.empty(d, dtype=torch.float64, device=dev)
        mean = mean.to(dev, torch.float64).contiguous()
        cov = cov.to(dev, torch.float64).contiguous()
        self._mean = self._mean.to(dev, torch.float64).contiguous()
        self._cov = self._cov.to(dev, torch.float64).contiguous()
        _lib.check(_lib.lib().amb_stats_merge(dev.index, _lib.stream_ptr(dev), d, self._n_folded, self._mean.data_ptr(),
                                              self._cov.data_ptr(), n, mean.data_ptr(), cov.data_ptr(),
                                              scratch.data_ptr()))
        self._n_folded += n

    def _update_stats(self, mean, cov, n):
        """data.py:77-94: merge externally computed statistics of ``n`` further rows."""
        self._fold()
        self._merge(mean, cov, n)
        self._n = (self._n or 0) + n

    def local_moments(self) -> torch.Tensor:
        """Raw fp64 moments [d + d*d] (column sums | Gram) of all rows of this container, rebuilt
        from (n, mean, cov): what row-sharded evaluation all-reduces (dist.py)."""
        mean, cov, n = self.mean, self.cov, self.n
        d = mean.shape[0]
        if tuple(cov.shape) != (d, d):
            cov = torch.zeros((d, d), dtype=torch.float64, device=mean.device)
        out = torch.empty(d + d * d, dtype=torch.float64, device=mean.device)
        out[:d] = mean * n
        out[d:] = (cov * (n - 1) + torch.outer(mean, mean) * n).reshape(-1)
        return out

    def __iadd__(self, other):
        """data.py:96-106."""
        assert isinstance(other, AudioMetricsData)
        if other.n is None:
            return self
        if self.n is None:
            self.store_embeddings = other.store_embeddings
        assert self.store_embeddings == other.store_embeddings
        self._fold()
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

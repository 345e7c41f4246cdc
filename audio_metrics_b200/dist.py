"""Row-sharded evaluation of FAD / KD / PRDC across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink).  Every sub-path is
row-independent, so the only exchanges are:

  covariance   allreduce(sum) of the fp64 raw moments [sum | gram] of both sets
  embeddings   allgather of the row shards (every rank needs all columns)
  radii        allgather of each rank's slice of k-NN radii
  counts       allreduce(sum) of the per-candidate counts and of two row totals
  KD           subsets dealt round-robin to ranks, allgather of the 100 MMD values

Row shards are contiguous and aligned to 256 rows (two tensor-core row tiles).
The arithmetic is delegated to an ``ops`` object so that the sharding and
reduction logic can be exercised on CPU (gloo, world_size 2) in the test-suite
with stand-in kernels; the product ``CudaOps`` calls the C ABI and nothing else.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

_SIDE_STREAMS = {}
ROW_ALIGN = 256     # two row tiles: shards start on an even tile, so the CTA-pair engine applies


def shard_rows(n: int, world: int, rank: int):
    """(row0, nrows, chunk): contiguous 128-aligned row range of ``rank``."""
    chunk = -(-n // world)
    chunk = -(-chunk // ROW_ALIGN) * ROW_ALIGN
    row0 = min(rank * chunk, n)
    return row0, max(0, min(n, row0 + chunk) - row0), chunk


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _allgather_rows(x: torch.Tensor, n: int, chunk: int, group):
    """Gather equally padded row shards and cut the result back to n rows."""
    world, _ = _world(group)
    if world == 1:
        return x[:n]
    pad = torch.zeros((chunk,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    out = torch.empty((world * chunk,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n]


def _null():
    import contextlib

    return contextlib.nullcontext()


def _allreduce(t: torch.Tensor, group):
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class CudaOps:
    """The product kernels behind the C ABI (device = this rank's GPU)."""

    def __init__(self, device=None):
        from . import _lib

        self._lib = _lib
        self.device = _lib.require_cuda(device)
        self._side = None

    # -- the N-independent FAD kernels (Cholesky, Jacobi: ~30 CTAs, latency-bound) run on a
    #    high-priority side stream beside the PRDC sweeps, whose CTA pairs take their work
    #    items dynamically and simply give up the SMs the side stream is holding
    def side(self, *tensors):
        import contextlib

        @contextlib.contextmanager
        def ctx():
            if self._side is None:
                # ONE side stream per device for the life of the process: torch hands out pooled
                # streams round-robin, and a fresh one per call eventually lands on the hardware queue
                # of the main stream, which serialises the two (seen as 100-200 ms outlier steps)
                key = self.device.index
                if key not in _SIDE_STREAMS:
                    _SIDE_STREAMS[key] = torch.cuda.Stream(self.device, priority=-1)
                self._side = _SIDE_STREAMS[key]
            main = torch.cuda.current_stream(self.device)
            self._side.wait_stream(main)
            # (no record_stream: the caller keeps `tensors` alive until join(), and marking them would
            #  make the caching allocator hold their blocks back across streams)
            L = self._lib.lib()
            L.amb_set_option(b"fad_ctas", self.SHARED_SMS)     # the sweep beside it leaves these SMs alone
            try:
                with torch.cuda.stream(self._side):
                    if self.trace is not None:
                        e0 = torch.cuda.Event(enable_timing=True); e0.record()
                    yield
                    if self.trace is not None:
                        e1 = torch.cuda.Event(enable_timing=True); e1.record()
                        self.trace.append((e0, e1))
            finally:
                L.amb_set_option(b"fad_ctas", 0)
        return ctx()

    import os as _os
    SHARED_SMS = int(_os.environ.get("AMB_SHARED_SMS", "16"))
    trace = None     # set to a list to collect (start, end) CUDA events of the side-stream section

    def reserve_sms(self, on):
        """The next all-pairs sweeps leave SHARED_SMS SMs to the side stream (on) / use them all (off)."""
        self._lib.lib().amb_set_option(b"engine_reserve_sms", self.SHARED_SMS if on else 0)

    def use_side(self):
        import os

        return os.environ.get("AMB_FAD_SIDE", "1") != "0"

    def join(self, *tensors):
        if self._side is not None:
            main = torch.cuda.current_stream(self.device)
            main.wait_stream(self._side)

    def moments(self, x):
        """fp64 [d + d*d] raw moments (column sums | Gram) of a row shard."""
        L, dev = self._lib.lib(), self.device
        d = x.shape[1]
        buf = torch.zeros(d + d * d, dtype=torch.float64, device=dev)
        if x.shape[0]:
            ws = self._lib.workspace(L.amb_cov_ws_bytes(x.shape[0], d), dev)
            self._lib.check(L.amb_cov_accumulate(dev.index, self._lib.stream_ptr(dev), x.data_ptr(),
                                                 self._lib.dtype_code(x), x.shape[0], d, x.stride(0), buf.data_ptr(),
                                                 buf[d:].data_ptr(), ws.data_ptr(), ws.numel()))
        return buf

    def stats_from_moments(self, buf, n, d):
        L, dev = self._lib.lib(), self.device
        mean = torch.empty(d, dtype=torch.float64, device=dev)
        cov = torch.empty((d, d), dtype=torch.float64, device=dev)
        self._lib.check(L.amb_cov_finalize(dev.index, self._lib.stream_ptr(dev), n, d, buf.data_ptr(),
                                           buf[d:].data_ptr(), mean.data_ptr(), cov.data_ptr()))
        return mean, cov

    def frechet(self, sx, sy):
        from .metrics.fad import frechet_distances

        class _S:
            pass

        a, b = _S(), _S()
        a.mean, a.cov = sx
        b.mean, b.cov = sy
        return frechet_distances([(a, b)], device=self.device, as_tensor=True)[0]   # stays on the device

    def container(self, x):
        from .data import AudioMetricsData

        c = AudioMetricsData(store_embeddings=True, device=self.device)
        c.embeddings = x
        return c

    def radii_rows(self, c, row0, nrows, k):
        from .metrics.prdc import nearest_neighbour_distances

        return nearest_neighbour_distances(c, k, row_range=(row0, nrows))

    def count_rows(self, cref, ccand, r_ref, r_cand, row0, nrows, k):
        """(col_count [m] int32, totals int64 [n_recalled, n_covered, n_uncertain])."""
        from .metrics.prdc import prdc_totals

        col, rec, cov, totals = prdc_totals(cref, ccand, k, row_range=(row0, nrows), ref_radii=r_ref,
                                            cand_radii=r_cand)
        t = torch.stack([rec.sum(dtype=torch.int64), cov.sum(dtype=torch.int64), totals[4]])
        return col, t

    def check_uncertain(self, uncertain, n_ref, n_cand):
        """The refine list has a fixed capacity; pairs beyond it were not re-decided."""
        cap = self._lib.lib().amb_prdc_list_cap(n_ref, n_cand)
        if uncertain > cap * max(1, _world(None)[0]):
            raise self._lib.AmbError(f"{uncertain} near-tie pairs exceed the refine list capacity {cap}")

    def kd_mmds(self, f1, f2, idx, gamma, coef0, degree):
        L, dev = self._lib.lib(), self.device
        S, _, m = idx.shape
        out = torch.empty(S, dtype=torch.float64, device=dev)
        if S == 0:
            return out
        idx_dev = torch.from_numpy(np.ascontiguousarray(idx)).to(dev)
        ws = self._lib.workspace(L.amb_kd_ws_bytes(S, m, f1.shape[1]), dev)
        self._lib.check(L.amb_kd_subsets(dev.index, self._lib.stream_ptr(dev), f1.data_ptr(), f1.shape[0],
                                         f1.stride(0), f2.data_ptr(), f2.shape[0], f2.stride(0), f1.shape[1],
                                         self._lib.dtype_code(f1), idx_dev.data_ptr(), S, m, self._lib.AMB_KERNEL_POLY,
                                         float(gamma), float(coef0), int(degree), 1.0, out.data_ptr(), None,
                                         ws.data_ptr(), ws.numel()))
        return out


def evaluate_sharded(ref_shard, cand_shard, n_ref, n_cand, metrics=("fad", "kd", "prdc"), nearest_k=5,
                     group=None, ops=None, kd_subsets=100, kd_subset_size=1000, kd_seed=1234, ready=None):
    """FAD / KD / PRDC of (reference, candidate) given this rank's row shards.

    ``ref_shard`` / ``cand_shard`` are this rank's rows per ``shard_rows``; every
    rank returns the same result dict (keys as AudioMetrics.evaluate,
    audio_metrics.py:254-274).  With an uninitialised process group this is the
    single-GPU path.

    Everything is enqueued without a host synchronisation and read back once at the
    end (the reference makes one ``.item()`` per metric): the KD subset indices are
    drawn on the host (numpy, kd.py:176-186) while the GPU is busy with PRDC.
    ``ready = (event_ref, event_cand)`` (CUDA events, either may be None) lets the caller
    hand over shards whose host-to-device copies are still in flight on another stream:
    all reference-only work (moments, packing, radii) is queued before the candidate
    shard is first touched.
    """
    from .metrics.kd import draw_subset_indices, KID_DEGREE, KID_COEF0

    ops = ops or CudaOps()
    world, rank = _world(group)
    d = ref_shard.shape[1]
    r_row0, r_nrows, r_chunk = shard_rows(n_ref, world, rank)
    c_row0, c_nrows, c_chunk = shard_rows(n_cand, world, rank)
    assert ref_shard.shape[0] == r_nrows and cand_shard.shape[0] == c_nrows, "shards must follow shard_rows()"
    want_fad, want_kd, want_prdc = "fad" in metrics, "kd" in metrics, "prdc" in metrics
    k = nearest_k

    def wait(ev):
        if ev is not None:
            torch.cuda.current_stream(ref_shard.device).wait_event(ev)

    ev_ref, ev_cand = ready if ready is not None else (None, None)
    pending = {}                                     # device-side results, read back at the end

    overlap = want_fad and want_prdc and hasattr(ops, "side") and ops.use_side()   # alone, FAD runs at full width
    # One GPU: reference-only work first, so that a candidate shard still in flight (ready=) is touched
    # late; the FAD kernels then run beside the candidate radii sweep.  Several GPUs: the sweeps are
    # 1/world as long while FAD is N-independent, so FAD is started first and all sweeps leave it room.
    fad_first = overlap and world > 1

    def fad(mom_ref, mom_cand):
        mom = torch.cat([mom_ref, mom_cand])
        _allreduce(mom, group)                       # one message: 2 (d + d^2) doubles
        half = d + d * d
        with ops.side(mom) if overlap else _null():
            s_ref = ops.stats_from_moments(mom[:half], n_ref, d)
            s_cand = ops.stats_from_moments(mom[half:], n_cand, d)
            pending["fad"] = ops.frechet(s_cand, s_ref)  # (cand, ref) as audio_metrics.py:257
        pending["_mom"] = mom                        # alive until the read-back (used on the side stream)

    def narrowed(on):
        if overlap:
            ops.reserve_sms(on)

    # ---- reference-only work
    wait(ev_ref)
    if want_fad:
        mom_ref = ops.moments(ref_shard)
    if fad_first:
        wait(ev_cand)
        fad(mom_ref, ops.moments(cand_shard))
    if want_kd or want_prdc:
        ref = _allgather_rows(ref_shard, n_ref, r_chunk, group)
    try:
        if want_prdc:
            cref = ops.container(ref)
            if fad_first:
                cref.packed()
                narrowed(True)
            r_ref = _allgather_rows(ops.radii_rows(cref, r_row0, r_nrows, k), n_ref, r_chunk, group).contiguous()

        # ---- candidate
        wait(ev_cand)
        if want_fad and not fad_first:
            fad(mom_ref, ops.moments(cand_shard))
        if want_kd or want_prdc:
            cand = _allgather_rows(cand_shard, n_cand, c_chunk, group)
        if want_prdc:
            ccand = ops.container(cand)
            if overlap and not fad_first:
                ccand.packed()            # (the pack kernels are not part of the sweep that shares the GPU)
                narrowed(True)            # FAD (about 20 ms) runs beside the candidate radii sweep (about 35 ms)
            r_cand = _allgather_rows(ops.radii_rows(ccand, c_row0, c_nrows, k), n_cand, c_chunk, group).contiguous()
            if not fad_first:
                narrowed(False)           # ... and the count sweep is full width again
            col, t = ops.count_rows(cref, ccand, r_ref, r_cand, r_row0, r_nrows, k)
            _allreduce(col, group)                       # [m] int32
            _allreduce(t, group)                         # 3 int64
            pending["prdc"] = torch.cat([torch.stack([(col > 0).sum(dtype=torch.int64), col.sum(dtype=torch.int64)]),
                                         t.to(torch.int64)])
    finally:
        narrowed(False)

    if want_kd:
        n_s = min(n_ref, n_cand)
        m = kd_subset_size if kd_subset_size < n_s else max(1, n_s // 2)        # kd.py:160-168
        idx = draw_subset_indices(n_cand, n_ref, m, kd_subsets, kd_seed)         # features_1 = candidate
        mine = idx[rank::world]
        per = -(-kd_subsets // world)
        local = ops.kd_mmds(cand, ref, mine, 1.0 / d, KID_COEF0, KID_DEGREE)
        if world > 1:
            pad = torch.zeros(per, dtype=torch.float64, device=local.device)
            pad[: local.shape[0]] = local
            allv = torch.empty(world * per, dtype=torch.float64, device=local.device)
            dist.all_gather_into_tensor(allv, pad, group=group)
            order = torch.tensor([(s % world) * per + s // world for s in range(kd_subsets)], device=local.device)
            pending["kd"] = allv[order]
        else:
            pending["kd"] = local

    # ---- the one read-back
    if want_fad and hasattr(ops, "join"):
        ops.join(pending["fad"])
    result = {}
    if want_fad:
        result["fad"] = float(pending["fad"])
    if want_kd:
        mm = pending["kd"].cpu().numpy()
        result["kernel_distance_mean"] = float(np.mean(mm))                       # kd.py:190
        result["kernel_distance_std"] = float(np.std(mm))                         # kd.py:191
    if want_prdc:
        hits, total, recalled, covered, uncertain = pending["prdc"].tolist()
        ops.check_uncertain(uncertain, n_ref, n_cand)
        result.update(precision=hits / n_cand, recall=recalled / n_ref,
                      density=(1.0 / float(k)) * (total / n_cand), coverage=covered / n_ref)   # prdc.py:36-48
    return result

"""The fused evaluation step: FAD / KD / PRDC / APA of (reference, candidate) in one
asynchronous schedule with a single read-back, on one GPU or row-sharded across the
GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink).  Every sub-path is
row-independent, so the only exchanges are:

  covariance   allreduce(sum) of the fp64 raw moments [sum | gram] of every set (one message)
  embeddings   allgather of the row shards (every rank needs all columns)
  radii        allgather of each rank's slice of k-NN radii
  counts       allreduce(sum) of the per-candidate counts and of two row totals,
               allreduce(max) of the per-rank near-tie counts (refine-list overflow check)
  KD           subsets dealt round-robin to ranks, allgather of the 100 MMD values

Row shards are contiguous and aligned to 256 rows (two tensor-core row tiles).
Two front ends share the schedule:

  evaluate_sharded      raw row shards (tensors) in, result dict out
  evaluate_containers   AudioMetricsData containers in (what ``AudioMetrics.evaluate``
                        calls): statistics, gathered sets, packed operands and k-NN radii
                        are cached on the containers, so a reference set is swept once

The arithmetic is delegated to an ``ops`` object so that the sharding and
reduction logic can be exercised on CPU (gloo, world_size 2) in the test-suite
with stand-in kernels; the product ``CudaOps`` calls the C ABI and nothing else.
"""
from __future__ import annotations

import functools
import os

import numpy as np
import torch
import torch.distributed as dist

ROW_ALIGN = 256     # two row tiles: shards start on an even tile, so the CTA-pair engine applies
EXACT = "exact"     # refine-list capacity value that selects the exhaustive count kernel


def shard_rows(n: int, world: int, rank: int):
    """(row0, nrows, chunk): contiguous 256-aligned row range of ``rank``."""
    chunk = -(-n // world)
    chunk = -(-chunk // ROW_ALIGN) * ROW_ALIGN
    row0 = min(rank * chunk, n)
    return row0, max(0, min(n, row0 + chunk) - row0), chunk


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _allgather_rows_start(x: torch.Tensor, counts, group):
    """Asynchronous form of _allgather_rows: issues the collective now (it runs on NCCL's own stream,
    beside whatever is queued on the current stream afterwards) and returns a function that makes the
    current stream wait for it and hands out the gathered rows."""
    world, _ = _world(group)
    if world == 1:
        return lambda: x
    maxc = max(max(counts), 1)
    if x.shape[0] == maxc and x.is_contiguous():
        pad = x
    else:
        pad = torch.zeros((maxc,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[: x.shape[0]] = x
    out = torch.empty((world * maxc,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    work = dist.all_gather_into_tensor(out, pad, group=group, async_op=True)

    def finish(pad=pad):
        work.wait()
        if all(c == maxc for c in counts[:-1]):
            return out[: sum(counts)]
        return torch.cat([out[r * maxc: r * maxc + counts[r]] for r in range(world)])

    return finish


def _allgather_rows(x: torch.Tensor, counts, group):
    """All ranks' row blocks, concatenated in rank order.  ``counts[r]`` = rows rank r contributes
    (any sizes: who holds which rows is the caller's business, e.g. wherever the embedder left them).
    One NCCL allgather of blocks padded to the largest; when only the last block is short (the
    regular case) the result is a view of the receive buffer, otherwise one compaction copy."""
    world, _ = _world(group)
    if world == 1:
        return x
    maxc = max(max(counts), 1)
    if x.shape[0] == maxc and x.is_contiguous():
        pad = x
    else:
        pad = torch.zeros((maxc,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[: x.shape[0]] = x
    out = torch.empty((world * maxc,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(c == maxc for c in counts[:-1]):
        return out[: sum(counts)]
    return torch.cat([out[r * maxc: r * maxc + counts[r]] for r in range(world)])


def _gather_meta(values, device, group):
    """Every rank's small list of python ints (row counts, widths): ONE collective and the only
    host synchronisation of a step besides the final read-back — issued first, before any sweep is
    queued, so it costs one allgather latency and not a pipeline drain."""
    world, _ = _world(group)
    if world == 1:
        return [list(values)]
    mine = torch.tensor(list(values), dtype=torch.int64, device=device)
    out = torch.zeros(world * len(values), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    flat = out.tolist()
    return [flat[r * len(values):(r + 1) * len(values)] for r in range(world)]


# ---- work partition.  After the allgather every rank holds ALL rows, so which rows a rank SWEEPS is
# independent of which rows it contributed.  The Frechet distance is N-independent work (pivoted
# Cholesky + 69 fp64 GEMMs of d^3: 3.8 ms at d = 512 on a B200) that only one rank needs to do; that
# rank gets a proportionally smaller share of the all-pairs sweeps so that all ranks finish together.
FAD_RANK = 0
FAD_SIDE_STREAM = os.environ.get("AMB_FAD_SIDE", "1") != "0"   # A/B switch of the side-stream placement


def fad_seconds(d: int) -> float:
    """Measured cost model of amb_frechet on a B200 (profiles/r02_*): latency-bound pivoted Cholesky
    (~d steps), 69 GEMMs (launch-bound below d ~ 256, d^3 above), fixed glue."""
    dp = -(-d // 64) * 64
    return 1e-3 * (0.7 + 1.4 * (d / 512.0) + 1.7 * max((dp / 512.0) ** 3, 0.05))


def sweep_seconds(n_ref: int, n_cand: int, d: int) -> float:
    """The three all-pairs sweeps on one B200: 1.25e12 pairs/s at d = 512 (power-capped tensor pipe),
    proportional to the padded width."""
    kb = -(-d // 32)
    return (float(n_ref) * n_ref + float(n_cand) * n_cand + float(n_ref) * n_cand) * (kb / 16.0) / 1.25e12


def work_weights(world: int, n_ref: int, n_cand: int, d: int, with_fad: bool):
    """Relative share of the sweeps per rank: equal, except that the rank that also computes the
    Frechet distance takes less.  With W the sweep time of the whole problem on one GPU and F the
    FAD time, all ranks finish together at (W + F) / world when rank FAD_RANK sweeps the fraction
    1 / world - F / W ... i.e. a weight of 1 - world F / (W + F) relative to the others."""
    w = [1.0] * world
    if world > 1 and with_fad:
        W, F = sweep_seconds(n_ref, n_cand, d), fad_seconds(d)
        w[FAD_RANK] = min(1.0, max(0.25, 1.0 - world * F / (W + F)))
    return w


def work_rows(n: int, weights, rank: int):
    """(row0, nrows) of the rows ``rank`` sweeps: boundaries at the weighted quantiles of [0, n),
    rounded to ROW_ALIGN (the CTA-pair engine takes row tiles two at a time)."""
    total = sum(weights)
    def bound(r):
        if r >= len(weights):
            return n
        b = int(round(n * sum(weights[:r]) / total / ROW_ALIGN)) * ROW_ALIGN
        return min(b, n)
    b0, b1 = bound(rank), bound(rank + 1)
    return b0, max(0, b1 - b0)


def _allreduce(t: torch.Tensor, group, op=None):
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=group)
    return t


@functools.lru_cache(maxsize=8)
def kd_subset_indices(n1: int, n2: int, m: int, subsets: int, seed: int) -> np.ndarray:
    """The index stream of kd.py:176,185-186 ([S, 2, m] int32).  It depends only on the set sizes
    and the seed, so repeated evaluations against the same sizes draw it once (200 numpy
    ``choice`` calls: about 4 ms at 200k rows)."""
    from .metrics.kd import draw_subset_indices

    idx = draw_subset_indices(n1, n2, m, subsets, seed)
    idx.setflags(write=False)
    return idx


_SIDE_STREAMS = {}


def _side_stream(device):
    """One side stream per device for the life of the process (the Frechet distance of the rank that
    owns it runs there, in the gaps its smaller sweep share leaves before each collective)."""
    st = _SIDE_STREAMS.get(device.index)
    if st is None:
        st = _SIDE_STREAMS[device.index] = torch.cuda.Stream(device)
    return st


TRACE = None     # set to a list to collect (label, CUDA event) marks of the next fused step (diagnostics)


def _mark(label, device):
    if TRACE is not None and device.type == "cuda":
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(device))
        TRACE.append((label, ev))


def trace_report():
    """[(label, ms since the previous mark)] of the marks collected in TRACE (synchronises)."""
    if not TRACE:
        return []
    torch.cuda.synchronize()
    return [(b[0], a[1].elapsed_time(b[1])) for a, b in zip(TRACE[:-1], TRACE[1:])]


_KD_ORDER = {}


def _kd_order(subsets, world, per, device):
    """Index tensor that puts the round-robin dealt subset values back in subset order (cached: it
    depends on nothing but the counts)."""
    key = (subsets, world, per, str(device))
    t = _KD_ORDER.get(key)
    if t is None:
        t = _KD_ORDER[key] = torch.tensor([(s % world) * per + s // world for s in range(subsets)], device=device)
    return t


class CudaOps:
    """The product kernels behind the C ABI (device = this rank's GPU)."""

    def __init__(self, device=None):
        from . import _lib

        self._lib = _lib
        self.device = _lib.require_cuda(device)
        self._idx_dev = {}

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

    def frechet_batch(self, pairs):
        """[(stats_x, stats_y)] with stats = (mean, cov) -> fp64 device tensor [len(pairs)], no sync."""
        from .metrics.fad import frechet_distances

        class _S:
            def __init__(self, s):
                self.mean, self.cov = s

        return frechet_distances([(_S(a), _S(b)) for a, b in pairs], device=self.device, as_tensor=True)

    def container(self, x):
        from .data import AudioMetricsData

        c = AudioMetricsData(store_embeddings=True, device=self.device)
        c.embeddings = x
        return c

    def radii_rows(self, c, row0, nrows, k):
        from .metrics.prdc import nearest_neighbour_distances

        return nearest_neighbour_distances(c, k, row_range=(row0, nrows))

    def count_rows(self, cref, ccand, r_ref, r_cand, row0, nrows, k, list_cap=None):
        """(col_count [m] int32, int64 [n_recalled, n_covered], int64 [n_uncertain, list capacity])
        for one shard of reference rows; ``list_cap`` as metrics.prdc.prdc_totals."""
        from .metrics.prdc import prdc_totals

        col, rec, cov, totals = prdc_totals(cref, ccand, k, row_range=(row0, nrows), ref_radii=r_ref,
                                            cand_radii=r_cand, list_cap=list_cap)
        return col, torch.stack([rec.sum(dtype=torch.int64), cov.sum(dtype=torch.int64)]), totals[4:6]

    def next_list_cap(self, uncertain, n_ref, n_cand):
        from .metrics.prdc import next_list_cap

        return next_list_cap(uncertain, n_ref, n_cand, self.device)

    def kd_mmds(self, f1, f2, idx, gamma, coef0, degree, key=None):
        L, dev = self._lib.lib(), self.device
        S, _, m = idx.shape
        out = torch.empty(S, dtype=torch.float64, device=dev)
        if S == 0:
            return out
        idx_dev = self._idx_dev.get(key) if key is not None else None
        if idx_dev is None:
            idx_dev = torch.from_numpy(np.array(idx, copy=True)).to(dev)
            if key is not None:
                self._idx_dev = {key: idx_dev}     # keep the last one: same sizes evaluate after evaluate
        ws = self._lib.workspace(L.amb_kd_ws_bytes(S, m, f1.shape[1]), dev)
        self._lib.check(L.amb_kd_subsets(dev.index, self._lib.stream_ptr(dev), f1.data_ptr(), f1.shape[0],
                                         f1.stride(0), f2.data_ptr(), f2.shape[0], f2.stride(0), f1.shape[1],
                                         self._lib.dtype_code(f1), idx_dev.data_ptr(), S, m, self._lib.AMB_KERNEL_POLY,
                                         float(gamma), float(coef0), int(degree), 1.0, self._lib.AMB_MMD_UNBIASED,
                                         out.data_ptr(), None, ws.data_ptr(), ws.numel()))
        return out


_OPS = {}


def default_ops(device=None):
    """One CudaOps per device (it only caches the uploaded KD index tensor)."""
    from . import _lib

    dev = _lib.require_cuda(device)
    ops = _OPS.get(dev.index)
    if ops is None:
        ops = _OPS[dev.index] = CudaOps(dev)
    return ops


class _Shard:
    """A row shard handed in as a plain tensor (evaluate_sharded)."""

    def __init__(self, rows, n_total, ops, ready=None, counts=None):
        self.rows_, self.n_total, self.ops, self.ready = rows, n_total, ops, ready
        self.counts = counts            # rows held by every rank
        self.cache = {}
        self.d = rows.shape[1]
        self.device = rows.device

    def rows(self):
        if self.ready is not None:
            torch.cuda.current_stream(self.device).wait_event(self.ready)
            self.ready = None
        return self.rows_

    def moments(self):
        return self.ops.moments(self.rows())

    def full_container(self, gathered):
        return self.ops.container(gathered)


def _width(c):
    """Embedding width of a container, None if it has never seen a row."""
    if c._buf is not None:
        return c._buf.shape[1]
    if c.n and c.mean is not None:
        return c.mean.shape[0]
    return None


class _Held:
    """A row shard held by an AudioMetricsData (evaluate_containers): statistics come from the
    container, derived data is cached on it and dropped when rows are added."""

    def __init__(self, c, n_total, ops, d=None, counts=None):
        self.c, self.n_total, self.ops = c, n_total, ops
        self.counts = counts            # rows held by every rank
        self.cache = c._cache
        self.empty = not c.n                       # this rank holds no rows of the set (tiny sets, many ranks)
        self.d = _width(c) or d
        self.device = c.device

    def rows(self):
        if self.empty:
            return torch.empty((0, self.d), dtype=torch.float32, device=self.device)
        x = self.c.embeddings
        if x is None:
            raise ValueError("this metric needs stored embeddings (store_embeddings=True)")
        return x

    def moments(self):
        if self.empty:
            return torch.zeros(self.d + self.d * self.d, dtype=torch.float64, device=self.device)
        return self.c.local_moments()

    def full_container(self, gathered):
        return self.ops.container(gathered)

    def stats(self):
        """(mean, cov) of the rows held here — final on one GPU, no moment exchange needed."""
        mean, cov = self.c.mean, self.c.cov
        if tuple(cov.shape) != (self.d, self.d):
            cov = torch.zeros((self.d, self.d), dtype=torch.float64, device=mean.device)
        return mean, cov


def _fused(ops, ref, cand, metrics, nearest_k, group, extra_fad=(), kd_subsets=100, kd_subset_size=1000,
           kd_seed=1234):
    """The schedule.  ``ref`` / ``cand``: _Shard or _Held.  ``extra_fad``: [(name, x, y)] further
    Frechet distances (x, y: _Held) evaluated in the same batched launch (APA)."""
    from .metrics.kd import KID_COEF0, KID_DEGREE

    world, rank = _world(group)
    want_fad, want_kd, want_prdc = "fad" in metrics, "kd" in metrics, "prdc" in metrics
    pending = {}
    n_ref, n_cand = (ref.n_total, cand.n_total) if ref is not None else (0, 0)
    k = nearest_k if nearest_k is not None else max(1, min(10, n_ref, n_cand))   # audio_metrics.py:263
    d = ref.d if ref is not None else None

    with_fad = bool(want_fad or extra_fad)
    weights = work_weights(world, n_ref, n_cand, d, with_fad) if ref is not None else [1.0] * world

    def full(s, prefetch=False):
        """Container over ALL rows of the set — gathered once per set (cached with the set).
        ``prefetch``: only start the allgather (it then overlaps the work queued next: the candidate's
        rows travel while the reference sweep runs); the next call completes it."""
        if world == 1 and isinstance(s, _Held):
            return s.c                       # the container itself (never cached inside itself: no cycles)
        key = ("full", world, id(group))
        hit = s.cache.get(key)
        if hit is None:
            assert sum(s.counts) == s.n_total, "row counts of the ranks do not add up to the set size"
            hit = s.cache[key] = ("pending", _allgather_rows_start(s.rows(), s.counts, group))
        if isinstance(hit, tuple) and not prefetch:
            hit = s.cache[key] = s.full_container(hit[1]())
        return None if isinstance(hit, tuple) else hit

    def radii(s, n):
        c = full(s)
        key = f"radii_{k}"
        r = c.radii.get(key)
        if r is None:
            row0, nrows = work_rows(n, weights, rank)
            counts = [work_rows(n, weights, q)[1] for q in range(world)]
            r = _allgather_rows(ops.radii_rows(c, row0, nrows, k), counts, group).contiguous()
            c.radii[key] = r
        return r

    # ---- reference-only work first: a candidate whose host-to-device copy is still in flight is
    #      touched as late as possible
    stat_sets, fad_pairs = [], []
    if want_fad:
        fad_pairs.append(("fad", cand, ref))                    # (cand, ref) as audio_metrics.py:257
    fad_pairs.extend(extra_fad)
    for _, x, y in fad_pairs:
        for s in (y, x):                                        # references before candidates
            if all(s is not t for t in stat_sets):
                stat_sets.append(s)
    # One GPU: the candidate's moments last, after the reference sweep (its host-to-device copy may
    # still be in flight).  Several ranks: all moments first, so that the Frechet distance can start
    # on its side stream before the first sweep.
    stat_sets.sort(key=lambda s: s is cand)
    early_stats = world > 1
    local_stats = world == 1 and all(isinstance(s, _Held) for s in stat_sets)   # statistics are already final
    tdev = (ref if ref is not None else stat_sets[0]).device
    _mark("start", tdev)
    moms = []
    done_ref_sweep = False
    for s in stat_sets:
        if s is cand and want_prdc and not done_ref_sweep and not early_stats:
            _mark("moments (reference)", tdev)
            full(ref)
            _mark("allgather reference", tdev)
            full(ref).packed() if hasattr(full(ref), "packed") else None
            _mark("pack reference", tdev)
            r_ref = radii(ref, n_ref)
            _mark("radii reference (sweep, refine, allgather)", tdev)
            done_ref_sweep = True
        moms.append(s.stats() if local_stats else s.moments())
    _mark("moments", tdev)
    if fad_pairs:
        if local_stats:
            stats = moms
        else:
            sizes = [s.d + s.d * s.d for s in stat_sets]
            mom = torch.cat(moms) if len(moms) > 1 else moms[0]
            _allreduce(mom, group)                              # one message: sum of (d + d^2) doubles per set
            stats, off = [], 0
            for s, sz in zip(stat_sets, sizes):
                stats.append(ops.stats_from_moments(mom[off:off + sz], s.n_total, s.d))
                off += sz
            pending["_mom"] = mom
        lookup = lambda s: stats[[t is s for t in stat_sets].index(True)]
        pairs = [(lookup(x), lookup(y)) for _, x, y in fad_pairs]
        if rank == FAD_RANK:
            # N-independent work: ONE rank computes it.  That rank sweeps fewer rows (work_weights), so it
            # reaches every collective of the sweeps early; the Frechet kernels run on a side stream and
            # fill exactly those waits instead of holding all ranks up at the next collective.  On one GPU
            # the same side stream fills the tails of the persistent sweeps (the last partial round of
            # work items leaves SMs idle) with the small, latency-bound Frechet kernels.
            sdev = torch.device(stats[0][0].device if isinstance(stats[0], (tuple, list)) else tdev)
            if sdev.type == "cuda" and (want_prdc or want_kd) and FAD_SIDE_STREAM:
                side = _side_stream(sdev)
                side.wait_stream(torch.cuda.current_stream(sdev))
                with torch.cuda.stream(side):
                    pending["fad"] = ops.frechet_batch(pairs)
                pending["_fad_side"] = side
            else:
                pending["fad"] = ops.frechet_batch(pairs)
            pending["_fad_inputs"] = stats       # alive until the side stream is joined
        else:
            pending["fad"] = torch.zeros(len(fad_pairs), dtype=torch.float64, device=mom.device)
        _mark("allreduce moments" + (", Frechet distances" if world == 1 else ""), tdev)

    if want_prdc and not done_ref_sweep:
        full(ref)
        _mark("allgather reference", tdev)
        if world > 1:
            full(cand, prefetch=True)        # the candidate's rows travel while the reference sweep runs
        full(ref).packed() if hasattr(full(ref), "packed") else None
        _mark("pack reference", tdev)
        r_ref = radii(ref, n_ref)
        _mark("radii reference (sweep, refine, allgather)", tdev)

    # ---- PRDC
    def counts(list_cap):
        cref, ccand = full(ref), full(cand)
        r_row0, r_nrows = work_rows(n_ref, weights, rank)
        col, t, unc = ops.count_rows(cref, ccand, r_ref, r_cand, r_row0, r_nrows, k, list_cap=list_cap)
        _allreduce(col, group)                                  # [m] int32
        _allreduce(t, group)                                    # 2 int64
        unc = unc.clone()
        _allreduce(unc, group, dist.ReduceOp.MAX if world > 1 else None)   # worst rank decides: lists are per rank
        return torch.cat([torch.stack([(col > 0).sum(dtype=torch.int64), col.sum(dtype=torch.int64)]),
                          t.to(torch.int64), unc.to(torch.int64)])

    if want_prdc:
        full(cand)
        _mark("allgather candidate", tdev)
        full(cand).packed() if hasattr(full(cand), "packed") else None
        _mark("pack candidate", tdev)
        r_cand = radii(cand, n_cand)
        _mark("radii candidate (sweep, refine, allgather)", tdev)
        pending["prdc"] = counts(None)
        _mark("counts (sweep, refine, allreduce)", tdev)

    # ---- KD: subsets dealt round-robin to the ranks
    if want_kd:
        n_s = min(n_ref, n_cand)
        m = kd_subset_size if kd_subset_size < n_s else max(1, n_s // 2)        # kd.py:160-168
        idx = kd_subset_indices(n_cand, n_ref, m, kd_subsets, kd_seed)           # features_1 = candidate
        mine = idx[rank::world]
        per = -(-kd_subsets // world)
        f1, f2 = full(cand).embeddings, full(ref).embeddings
        local = ops.kd_mmds(f1, f2, mine, 1.0 / d, KID_COEF0, KID_DEGREE,
                            key=(n_cand, n_ref, m, kd_subsets, kd_seed, rank, world))
        if world > 1:
            pad = torch.zeros(per, dtype=torch.float64, device=local.device)
            pad[: local.shape[0]] = local
            allv = torch.empty(world * per, dtype=torch.float64, device=local.device)
            dist.all_gather_into_tensor(allv, pad, group=group)
            pending["kd"] = allv[_kd_order(kd_subsets, world, per, local.device)]
        else:
            pending["kd"] = local
        _mark("kernel distance", tdev)

    # ---- the one read-back
    result = {}
    if "_fad_side" in pending:
        torch.cuda.current_stream(pending["fad"].device).wait_stream(pending["_fad_side"])
    if fad_pairs and world > 1:
        dist.broadcast(pending["fad"], src=dist.get_global_rank(group, FAD_RANK) if group is not None else FAD_RANK,
                       group=group)
        _mark("Frechet distances joined, broadcast", tdev)
    if fad_pairs:
        vals = pending["fad"].tolist()
        for (name, _, _), v in zip(fad_pairs, vals):
            result[name] = float(v)
    if want_kd:
        mm = pending["kd"].cpu().numpy()
        result["kernel_distance_mean"] = float(np.mean(mm))                       # kd.py:190
        result["kernel_distance_std"] = float(np.std(mm))                         # kd.py:191
    if want_prdc:
        hits, total, recalled, covered, uncertain, cap = pending["prdc"].tolist()
        list_cap = None
        while list_cap != EXACT and uncertain > cap:
            # More near-tie pairs than the refine list of some rank holds (prdc.py:18-50 has no such
            # limit): repeat the count sweep with a list of the reported size, or exhaustively.
            # Every rank sees the same reduced numbers, so all of them take this branch together.
            list_cap = ops.next_list_cap(uncertain, n_ref, n_cand)
            hits, total, recalled, covered, uncertain, cap = counts(list_cap).tolist()
        result.update(precision=hits / n_cand, recall=recalled / n_ref,
                      density=(1.0 / float(k)) * (total / n_cand), coverage=covered / n_ref)   # prdc.py:36-48
    return result


def evaluate_sharded(ref_shard, cand_shard, n_ref, n_cand, metrics=("fad", "kd", "prdc"), nearest_k=5,
                     group=None, ops=None, kd_subsets=100, kd_subset_size=1000, kd_seed=1234, ready=None):
    """FAD / KD / PRDC of (reference, candidate) given this rank's row shards.

    ``ref_shard`` / ``cand_shard`` are the rows this rank holds — in rank order they make up the
    sets; ``shard_rows`` gives an even split, but any sizes are accepted.  Every
    rank returns the same result dict (keys as AudioMetrics.evaluate,
    audio_metrics.py:254-274).  With an uninitialised process group this is the
    single-GPU path.

    Everything is enqueued without a host synchronisation and read back once at the
    end (the reference makes one ``.item()`` per metric).
    ``ready = (event_ref, event_cand)`` (CUDA events, either may be None) lets the caller
    hand over shards whose host-to-device copies are still in flight on another stream:
    all reference-only work (moments, packing, radii) is queued before the candidate
    shard is first touched.
    """
    ops = ops or default_ops(ref_shard.device if ref_shard.is_cuda else None)
    ev_ref, ev_cand = ready if ready is not None else (None, None)
    meta = _gather_meta([ref_shard.shape[0], cand_shard.shape[0]], ref_shard.device, group)
    return _fused(ops, _Shard(ref_shard, n_ref, ops, ev_ref, [m[0] for m in meta]),
                  _Shard(cand_shard, n_cand, ops, ev_cand, [m[1] for m in meta]), metrics,
                  nearest_k, group, kd_subsets=kd_subsets, kd_subset_size=kd_subset_size, kd_seed=kd_seed)


def evaluate_containers(ref, cand, metrics=("fad", "kd", "prdc"), nearest_k=5, group=None, apa=None,
                        kd_subsets=100, kd_subset_size=1000, kd_seed=1234):
    """The same step on AudioMetricsData containers — what ``AudioMetrics.evaluate`` runs.

    ``ref`` / ``cand`` hold this rank's rows (all rows when no process group is initialised).
    ``apa = (apa_cand, apa_ref, apa_anti, d_x_xp)``: the mix containers of the APA score; their
    Frechet distances join the same batched launch, the result dict then carries "_d_y_x",
    "_d_y_xp" and (when ``d_x_xp`` is None) "_d_x_xp" for apa.py:22-32 to combine.
    ``ref`` / ``cand`` may be None when only APA is wanted.
    """
    some = ref if ref is not None else apa[0]
    ops = default_ops(some.device)
    world, _ = _world(group)

    sets = [c for c in (ref, cand) + (tuple(apa[:3]) if apa is not None else ()) if c is not None]
    key = ("meta", world, id(group))
    if any(key not in c._cache for c in sets):
        # rows and width of every set on every rank, in one exchange (cached with each container's state)
        meta = _gather_meta([v for c in sets for v in (c.n or 0, _width(c) or 0)], ops.device, group)
        for i, c in enumerate(sets):
            counts = [m[2 * i] for m in meta]
            c._cache[key] = (sum(counts), max(m[2 * i + 1] for m in meta), counts)
    views = {}

    def held(c):
        # (a view per call: caching it on the container would tie the container into a reference
        #  cycle, and its device buffers would then wait for the cyclic garbage collector)
        v = views.get(id(c))
        if v is None:
            n_total, d, counts = c._cache[key]
            v = views[id(c)] = _Held(c, n_total, ops, d, counts)
        return v

    extra = []
    if apa is not None:
        a_cand, a_ref, a_anti, d_x_xp = apa
        extra = [("_d_y_x", held(a_cand), held(a_ref)), ("_d_y_xp", held(a_cand), held(a_anti))]
        if d_x_xp is None:
            extra.append(("_d_x_xp", held(a_ref), held(a_anti)))
    if ref is None:
        metrics = ()
    return _fused(ops, held(ref) if ref is not None else None, held(cand) if cand is not None else None, metrics,
                  nearest_k, group, extra_fad=extra, kd_subsets=kd_subsets, kd_subset_size=kd_subset_size,
                  kd_seed=kd_seed)


def global_stats(c, group=None):
    """(n, mean, cov) of a set whose rows are spread over the ranks, from each rank's container: one
    small exchange of row counts and one allreduce of the raw moments.  With no process group this
    is the container's own statistics."""
    world, _ = _world(group)
    if world == 1:
        return c.n, c.mean, c.cov
    ops = default_ops(c.device)
    meta = _gather_meta([c.n or 0, _width(c) or 0], ops.device, group)
    n, d = sum(m[0] for m in meta), max(m[1] for m in meta)
    mom = c.local_moments() if c.n else torch.zeros(d + d * d, dtype=torch.float64, device=ops.device)
    _allreduce(mom, group)
    mean, cov = ops.stats_from_moments(mom, n, d)
    return n, mean, cov


# ------------------------------------------------------------------ one process, several GPUs
def _replica(c, dev):
    """A container with the same rows on another device of this process (peer copy over NVLink),
    cached on the original so that a reference set is replicated — and swept — once."""
    from .data import AudioMetricsData

    if dev == c.device:
        return c
    key = ("replica", dev.index)
    r = c._cache.get(key)
    if r is None:
        r = AudioMetricsData(store_embeddings=True, device=dev)
        x = c.embeddings
        with torch.cuda.device(dev):
            r.embeddings = x.to(dev, non_blocking=True)
        c._cache[key] = r
    return r


def evaluate_devices(ref, cand, devices, metrics=("fad", "kd", "prdc"), nearest_k=None, apa=None,
                     kd_subsets=100, kd_subset_size=1000, kd_seed=1234):
    """The fused step of ONE process over several GPUs — ``AudioMetrics(device_indices=[0, 1, ...])``.

    The reference's only multi-GPU mechanism is thread-per-GPU data parallelism inside one process
    (util/gpu_parallel.py:20-76), so the drop-in keeps that shape: no process group, one host thread
    that enqueues asynchronous work on every device.  The containers live on ``devices[0]`` (where the
    embedder produced them); their rows are replicated to the other devices by peer copies, every
    device takes a 256-aligned row shard of both all-pairs sweeps against all columns, radii slices
    and per-candidate counts travel back as small peer copies, and the N-independent parts (Frechet
    distances, kernel distance) run on ``devices[0]`` beside them.  Same results as one device
    (the counts are exact integers, the radii exact roundings)."""
    from .metrics.kd import KID_COEF0, KID_DEGREE

    devices = [torch.device(d) if not isinstance(d, torch.device) else d for d in devices]
    world = len(devices)
    dev0 = devices[0]
    ops0 = default_ops(dev0)
    want_prdc = "prdc" in metrics and ref is not None
    rest = tuple(m for m in metrics if m != "prdc")
    if world == 1 or not want_prdc:
        return evaluate_containers(ref, cand, metrics, nearest_k=nearest_k, apa=apa, kd_subsets=kd_subsets,
                                   kd_subset_size=kd_subset_size, kd_seed=kd_seed)
    n_ref, n_cand = ref.n, cand.n
    k = nearest_k if nearest_k is not None else max(1, min(10, n_ref, n_cand))
    # devices[0] also computes the N-independent metrics: it sweeps fewer rows (work_weights)
    weights = work_weights(world, n_ref, n_cand, _width(ref), bool(rest) or apa is not None)

    # ---- replicate (reference first), then every device sweeps its row shard
    def sweep_radii(c, n):
        key = f"radii_{k}"
        if c.radii.get(key) is not None:
            return [_replica(c, d).radii.setdefault(key, c.radii[key].to(d, non_blocking=True)) for d in devices]
        parts = []
        for i, d in enumerate(devices):
            row0, nrows = work_rows(n, weights, i)
            with torch.cuda.device(d):
                parts.append(default_ops(d).radii_rows(_replica(c, d), row0, nrows, k))
        full0 = torch.cat([p.to(dev0, non_blocking=True) for p in parts])      # peer copies of [n / world] floats
        out = []
        for d in devices:
            with torch.cuda.device(d):
                r = full0 if d == dev0 else full0.to(d, non_blocking=True)
            _replica(c, d).radii[key] = r
            out.append(r)
        return out

    r_ref = sweep_radii(ref, n_ref)
    r_cand = sweep_radii(cand, n_cand)

    def counts(list_cap):
        cols, ts, uncs = [], [], []
        for i, d in enumerate(devices):
            row0, nrows = work_rows(n_ref, weights, i)
            with torch.cuda.device(d):
                col, t, unc = default_ops(d).count_rows(_replica(ref, d), _replica(cand, d), r_ref[i], r_cand[i], row0,
                                                        nrows, k, list_cap=list_cap)
            cols.append(col); ts.append(t); uncs.append(unc)
        with torch.cuda.device(dev0):
            col = torch.stack([c.to(dev0, non_blocking=True) for c in cols]).sum(dim=0, dtype=torch.int64)
            t = torch.stack([x.to(dev0, non_blocking=True) for x in ts]).sum(dim=0)
            unc = torch.stack([x.to(dev0, non_blocking=True) for x in uncs]).max(dim=0).values   # lists are per device
            return torch.cat([torch.stack([(col > 0).sum(), col.sum()]), t.to(torch.int64), unc.to(torch.int64)])

    pending = counts(None)
    # ---- the N-independent metrics on devices[0], queued behind its share of the sweeps
    with torch.cuda.device(dev0):
        result = evaluate_containers(ref if rest else None, cand if rest else None, rest, nearest_k=k, apa=apa,
                                     kd_subsets=kd_subsets, kd_subset_size=kd_subset_size, kd_seed=kd_seed) \
            if (rest or apa is not None) else {}
        hits, total, recalled, covered, uncertain, cap = pending.tolist()
        list_cap = None
        while list_cap != EXACT and uncertain > cap:
            list_cap = ops0.next_list_cap(uncertain, n_ref, n_cand)
            hits, total, recalled, covered, uncertain, cap = counts(list_cap).tolist()
    result.update(precision=hits / n_cand, recall=recalled / n_ref, density=(1.0 / float(k)) * (total / n_cand),
                  coverage=covered / n_ref)                                        # prdc.py:36-48
    return result

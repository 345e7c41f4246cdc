"""k-NN radii and precision / recall / density / coverage (reference metrics/prdc.py)."""
from __future__ import annotations

import torch

from .. import _lib


def _container(x):
    from ..data import AudioMetricsData

    if isinstance(x, AudioMetricsData):
        return x
    c = AudioMetricsData(store_embeddings=True)
    c.embeddings = x
    return c


def nearest_neighbour_distances(input_features, nearest_k, row_range=None, info=None):
    """prdc.py:4-14: distance to the (k+1)-th nearest row (self included) for every
    row, as an fp32 tensor on the device.  Accepts a tensor/array or an
    AudioMetricsData.  ``row_range=(row0, nrows)`` restricts to a shard of rows
    (row0 a multiple of 128) against all columns.  ``info``: a dict that receives
    ``n_exhaustive`` (device int64: rows the exhaustive exact scan had to resolve)."""
    c = _container(input_features)
    x = c.embeddings
    if x is None:
        raise ValueError("nearest_neighbour_distances needs stored embeddings")
    dev = c.device
    n, d = x.shape
    row0, nrows = (0, n) if row_range is None else row_range
    L = _lib.lib()
    radii = torch.empty(nrows, dtype=torch.float32, device=dev)
    n_ex = torch.zeros(1, dtype=torch.int64, device=dev) if info is not None else None
    if info is not None:
        info["n_exhaustive"] = n_ex
    if nrows == 0:
        return radii
    ws = _lib.workspace(L.amb_knn_ws_bytes(nrows, n, d, int(nearest_k)), dev)
    _lib.check(L.amb_knn_radii(dev.index, _lib.stream_ptr(dev), x.data_ptr(), _lib.dtype_code(x), x.stride(0),
                               c.packed().data_ptr(), n, d, row0, nrows, int(nearest_k), radii.data_ptr(),
                               n_ex.data_ptr() if n_ex is not None else None, ws.data_ptr(), ws.numel()))
    return radii


# Modes of prdc_totals: the tensor-core filter with a refine list of `list_cap` entries (None: the
# library default), or the exhaustive exact kernel.
EXACT = "exact"


def prdc_totals(reference, candidate, nearest_k, row_range=None, ref_radii=None, cand_radii=None, list_cap=None):
    """Integer numerators of prdc.py:36-48 for a shard of reference rows, without synchronising:
    (col_count [m] int32, row_recall [nrows] uint8, row_cover [nrows] uint8, totals) — all on the
    device; ``totals`` is int64 [6]: slots 0-3 are filled by amb_prdc_reduce (#cols with count > 0,
    sum of counts, #rows recalled, #rows covered), slot 4 = pairs that fell inside the filter's
    band, slot 5 = capacity of the refine list this call used.  totals[4] > totals[5] means the
    counts are invalid and the call must be repeated with ``list_cap=int(totals[4])`` or
    ``list_cap=EXACT`` (see ``prdc`` below and amb200.h)."""
    ref, cand = _container(reference), _container(candidate)
    dev = ref.device
    xr, xc = ref.embeddings, cand.embeddings
    if xr.dtype != xc.dtype:
        raise ValueError("reference and candidate embeddings must share a dtype")
    n, d = xr.shape
    m = xc.shape[0]
    if ref_radii is None:
        ref_radii = ref.get_radii(nearest_k)       # prdc.py:31
    if cand_radii is None:
        cand_radii = cand.get_radii(nearest_k)     # prdc.py:32
    row0, nrows = (0, n) if row_range is None else row_range
    L = _lib.lib()
    col_count = torch.zeros(m, dtype=torch.int32, device=dev)
    rec = torch.zeros(max(nrows, 1), dtype=torch.uint8, device=dev)
    cov = torch.zeros(max(nrows, 1), dtype=torch.uint8, device=dev)
    totals = torch.zeros(6, dtype=torch.int64, device=dev)
    st = _lib.stream_ptr(dev)
    if list_cap == EXACT:
        _lib.check(L.amb_prdc_counts_exact(dev.index, st, xr.data_ptr(), xr.stride(0), n, ref_radii.data_ptr(),
                                           xc.data_ptr(), xc.stride(0), m, cand_radii.data_ptr(), d,
                                           _lib.dtype_code(xr), row0, nrows, col_count.data_ptr(), rec.data_ptr(),
                                           cov.data_ptr()))
        return col_count, rec[:nrows], cov[:nrows], totals
    nbytes = L.amb_prdc_ws_bytes(n, m) if list_cap is None else L.amb_prdc_ws_bytes_cap(n, m, max(1, int(list_cap)))
    ws = _lib.workspace(nbytes, dev)
    # (fill_ passes the scalar as a kernel argument; `totals[5] = value` would stage it through a
    #  synchronous host-to-device copy and drain the stream in the middle of the step)
    totals[5:].fill_(int(L.amb_prdc_ws_list_cap(n, m, ws.numel())))
    _lib.check(L.amb_prdc_counts(dev.index, st, xr.data_ptr(), xr.stride(0), ref.packed().data_ptr(), n,
                                 ref_radii.data_ptr(), xc.data_ptr(), xc.stride(0), cand.packed().data_ptr(), m,
                                 cand_radii.data_ptr(), d, _lib.dtype_code(xr), row0, nrows, col_count.data_ptr(),
                                 rec.data_ptr(), cov.data_ptr(), totals[4:].data_ptr(), ws.data_ptr(), ws.numel()))
    return col_count, rec[:nrows], cov[:nrows], totals


def next_list_cap(uncertain: int, n: int, m: int, device) -> object:
    """The next rung of the overflow ladder after a call reported ``uncertain`` pairs in the band:
    a refine list of exactly that many entries when the device can hold it, else the exhaustive
    kernel."""
    need = _lib.lib().amb_prdc_ws_bytes_cap(n, m, int(uncertain))
    free, _ = torch.cuda.mem_get_info(device)
    return int(uncertain) if need < free // 2 else EXACT


def prdc(reference, candidate, nearest_k):
    """prdc.py:18-50: dict(precision, recall, density, coverage) of python floats.

    Never fails where the reference returns: when more pairs fall inside the tensor-core
    filter's error band than the default refine list holds (collinear / rank-1 sets, rows
    spanning many orders of magnitude) the count is repeated with a list sized to the reported
    number, or with the exhaustive exact kernel."""
    ref, cand = _container(reference), _container(candidate)
    dev = ref.device
    n, m = len(ref.embeddings), len(cand.embeddings)
    L = _lib.lib()
    cap = None
    while True:
        col_count, rec, cov, totals = prdc_totals(ref, cand, nearest_k, list_cap=cap)
        _lib.check(L.amb_prdc_reduce(dev.index, _lib.stream_ptr(dev), col_count.data_ptr(), m, rec.data_ptr(),
                                     cov.data_ptr(), n, totals.data_ptr()))
        hits, total, recalled, covered, uncertain, used_cap = totals.tolist()   # the single device->host read
        if cap == EXACT or uncertain <= used_cap:
            break
        cap = next_list_cap(uncertain, n, m, dev)
    return dict(
        precision=hits / m,                                        # prdc.py:36-38
        recall=recalled / n,                                       # prdc.py:40-42
        density=(1.0 / float(nearest_k)) * (total / m),            # prdc.py:44-46
        coverage=covered / n,                                      # prdc.py:48
    )

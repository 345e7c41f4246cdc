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


def nearest_neighbour_distances(input_features, nearest_k, row_range=None):
    """prdc.py:4-14: distance to the (k+1)-th nearest row (self included) for every
    row, as an fp32 tensor on the device.  Accepts a tensor/array or an
    AudioMetricsData.  ``row_range=(row0, nrows)`` restricts to a shard of rows
    (row0 a multiple of 128) against all columns."""
    c = _container(input_features)
    x = c.embeddings
    if x is None:
        raise ValueError("nearest_neighbour_distances needs stored embeddings")
    dev = c.device
    n, d = x.shape
    row0, nrows = (0, n) if row_range is None else row_range
    L = _lib.lib()
    radii = torch.empty(nrows, dtype=torch.float32, device=dev)
    if nrows == 0:
        return radii
    ws = _lib.workspace(L.amb_knn_ws_bytes(nrows, n, d, int(nearest_k)), dev)
    _lib.check(L.amb_knn_radii(dev.index, _lib.stream_ptr(dev), x.data_ptr(), _lib.dtype_code(x), x.stride(0),
                               c.packed().data_ptr(), n, d, row0, nrows, int(nearest_k), radii.data_ptr(),
                               ws.data_ptr(), ws.numel()))
    return radii


def prdc_totals(reference, candidate, nearest_k, row_range=None, ref_radii=None, cand_radii=None):
    """Integer numerators of prdc.py:36-48 for a shard of reference rows: a CPU
    int64 tensor [#cols with count > 0, sum of counts, #rows recalled, #rows
    covered] plus the per-candidate count vector (device, int32)."""
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
    rec = torch.empty(max(nrows, 1), dtype=torch.uint8, device=dev)
    cov = torch.empty(max(nrows, 1), dtype=torch.uint8, device=dev)
    totals = torch.zeros(5, dtype=torch.int64, device=dev)
    ws = _lib.workspace(L.amb_prdc_ws_bytes(n, m), dev)
    st = _lib.stream_ptr(dev)
    _lib.check(L.amb_prdc_counts(dev.index, st, xr.data_ptr(), xr.stride(0), ref.packed().data_ptr(), n,
                                 ref_radii.data_ptr(), xc.data_ptr(), xc.stride(0), cand.packed().data_ptr(), m,
                                 cand_radii.data_ptr(), d, _lib.dtype_code(xr), row0, nrows, col_count.data_ptr(),
                                 rec.data_ptr(), cov.data_ptr(), totals[4:].data_ptr(), ws.data_ptr(), ws.numel()))
    return col_count, rec[:nrows], cov[:nrows], totals


def prdc(reference, candidate, nearest_k):
    """prdc.py:18-50: dict(precision, recall, density, coverage) of python floats."""
    ref, cand = _container(reference), _container(candidate)
    dev = ref.device
    n, m = len(ref.embeddings), len(cand.embeddings)
    col_count, rec, cov, totals = prdc_totals(ref, cand, nearest_k)
    L = _lib.lib()
    _lib.check(L.amb_prdc_reduce(dev.index, _lib.stream_ptr(dev), col_count.data_ptr(), m, rec.data_ptr(),
                                 cov.data_ptr(), n, totals.data_ptr()))
    hits, total, recalled, covered, uncertain = totals.tolist()   # the single device->host read
    if uncertain > L.amb_prdc_list_cap(n, m):
        raise _lib.AmbError(f"{uncertain} near-tie pairs exceed the refine list capacity")
    return dict(
        precision=hits / m,                                        # prdc.py:36-38
        recall=recalled / n,                                       # prdc.py:40-42
        density=(1.0 / float(nearest_k)) * (total / m),            # prdc.py:44-46
        coverage=covered / n,                                      # prdc.py:48
    )

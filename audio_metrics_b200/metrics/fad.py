"""Frechet distance (reference metrics/fad.py:8-31) on the CUDA library."""
from __future__ import annotations

import torch

from .. import _lib


def frechet_distances(pairs, device=None, as_tensor=False):
    """FAD for several (x, y) container pairs in one batched launch sequence.

    Each element of ``pairs`` is ``(x, y)`` with ``.mean`` / ``.cov`` attributes
    (fad.py:8-13).  Returns a list of python floats, or with ``as_tensor`` the fp64
    device tensor without synchronising.
    """
    dev = _lib.require_cuda(device if device is not None else pairs[0][0].mean.device
                            if isinstance(pairs[0][0].mean, torch.Tensor) and pairs[0][0].mean.is_cuda else None)
    tod = lambda t: torch.as_tensor(t).to(dev, torch.float64)
    mu_x = torch.stack([tod(x.mean) for x, _ in pairs]).contiguous()
    mu_y = torch.stack([tod(y.mean) for _, y in pairs]).contiguous()
    cov_x = torch.stack([tod(x.cov) for x, _ in pairs]).contiguous()
    cov_y = torch.stack([tod(y.cov) for _, y in pairs]).contiguous()
    batch, d = mu_x.shape
    if cov_x.shape != (batch, d, d) or cov_y.shape != (batch, d, d) or mu_y.shape != (batch, d):
        raise ValueError(f"inconsistent statistics shapes: mean {tuple(mu_x.shape)}, cov {tuple(cov_x.shape)}")
    L = _lib.lib()
    out = torch.empty(batch, dtype=torch.float64, device=dev)
    ws = _lib.workspace(L.amb_frechet_ws_bytes(batch, d), dev)
    _lib.check(L.amb_frechet(dev.index, _lib.stream_ptr(dev), batch, d, mu_x.data_ptr(), cov_x.data_ptr(),
                             mu_y.data_ptr(), cov_y.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel()))
    if as_tensor:
        return out
    return out.tolist()   # the one device->host read fad.py:13 (.item()) makes


def frechet_distance(x, y, device=None):
    """fad.py:8-13: FAD between two AudioMetricsData, as a python float."""
    return frechet_distances([(x, y)], device=device)[0]

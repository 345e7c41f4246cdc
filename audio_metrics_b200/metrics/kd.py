"""Kernel distance / KID-style MMD^2 (reference metrics/kd.py) on the CUDA library.

The subset indices are drawn on the host with numpy's PCG64 stream exactly as
kd.py:176,185-186 does (same seed, same call order), uploaded once, and the
100 x 3 kernel blocks run as one tensor-core launch whose epilogue applies the
kernel and reduces each block to a scalar — no m x m matrix reaches memory.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .. import _lib

KEY_METRIC_KID_MEAN = "kernel_distance_mean"   # kd.py:17
KEY_METRIC_KID_STD = "kernel_distance_std"     # kd.py:18
KID_SUBSETS = 100
KID_SUBSET_SIZE = 1000
KID_DEGREE = 3
KID_GAMMA = None
KID_COEF0 = 1
KID_SIGMA = 10.0


def kernel_distance(x, y):
    """kd.py:29-35 — x = candidate, y = reference containers."""
    return kid_features_to_metric(x.embeddings, y.embeddings)


def draw_subset_indices(n1, n2, m, subsets, seed):
    """kd.py:176,185-186: one generator, alternating choice(n1), choice(n2)."""
    rng = np.random.default_rng(seed)
    idx = np.empty((subsets, 2, m), dtype=np.int32)
    for i in range(subsets):
        idx[i, 0] = rng.choice(n1, m, replace=False)
        idx[i, 1] = rng.choice(n2, m, replace=False)
    return idx


def kid_features_to_metric(features_1, features_2, **kwargs):
    """kd.py:127-194 with the same keyword arguments; ``return_mmds=True`` adds the
    per-subset values under "mmds"."""
    kernel_type = kwargs.get("kernel_type", "polynomial")
    if kernel_type == "polynomial":
        ktype = _lib.AMB_KERNEL_POLY
    elif kernel_type == "rbf":
        ktype = _lib.AMB_KERNEL_RBF
    else:
        raise NotImplementedError(f'Unknown kernel_type "{kernel_type}"')   # kd.py:142
    if features_1 is None or features_2 is None:
        raise ValueError("kernel distance needs stored embeddings")
    dev = _lib.require_cuda(features_1.device if isinstance(features_1, torch.Tensor) and features_1.is_cuda else None)
    f1 = _lib.as_device_matrix(features_1, dev)
    f2 = _lib.as_device_matrix(features_2, dev)
    assert f1.shape[1] == f2.shape[1]                                       # kd.py:151
    if f1.dtype != f2.dtype:
        f1, f2 = f1.to(torch.float64), f2.to(torch.float64)

    kid_subsets = kwargs.get("kid_subsets", KID_SUBSETS)
    kid_subset_size = kwargs.get("kid_subset_size", KID_SUBSET_SIZE)
    verbose = kwargs.get("verbose", False)
    n1, n2 = len(f1), len(f2)
    assert n1 and n2, "Cannot compute KID on empty features tensor"          # kd.py:158
    n_samples = min(n1, n2)
    if kid_subset_size >= n_samples:                                        # kd.py:160-168
        new_ss = max(1, n_samples // 2)
        if verbose:
            logging.warning(f"Reducing KID subset size from {kid_subset_size} to {new_ss} "
                            "to accommodate small sample size")
        kid_subset_size = new_ss
    d = f1.shape[1]
    gamma = kwargs.get("kid_gamma", KID_GAMMA)
    if gamma is None:
        gamma = 1.0 / d                                                     # kd.py:113-114
    degree = int(kwargs.get("kid_degree", KID_DEGREE))
    coef0 = float(kwargs.get("kid_coef0", KID_COEF0))
    sigma = float(kwargs.get("kid_sigma", KID_SIGMA))

    idx = draw_subset_indices(n1, n2, kid_subset_size, kid_subsets, kwargs.get("rng_seed", 1234))
    idx_dev = torch.from_numpy(idx).to(dev, non_blocking=True)
    L = _lib.lib()
    mmds = torch.empty(kid_subsets, dtype=torch.float64, device=dev)
    stats = torch.empty(2, dtype=torch.float64, device=dev)
    ws = _lib.workspace(L.amb_kd_ws_bytes(kid_subsets, kid_subset_size, d), dev)
    _lib.check(L.amb_kd_subsets(dev.index, _lib.stream_ptr(dev), f1.data_ptr(), n1, f1.stride(0), f2.data_ptr(), n2,
                                f2.stride(0), d, _lib.dtype_code(f1), idx_dev.data_ptr(), kid_subsets,
                                kid_subset_size, ktype, float(gamma), coef0, degree, sigma, mmds.data_ptr(),
                                stats.data_ptr(), ws.data_ptr(), ws.numel()))
    mean, std = stats.tolist()
    out = {KEY_METRIC_KID_MEAN: float(mean), KEY_METRIC_KID_STD: float(std)}   # kd.py:189-192
    if kwargs.get("return_mmds", False):
        out["mmds"] = mmds.cpu().numpy()
    return out

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


_MMD_EST = {"unbiased": _lib.AMB_MMD_UNBIASED, "biased": _lib.AMB_MMD_BIASED, "u-statistic": _lib.AMB_MMD_USTAT}


def mmd2(K_XX, K_XY, K_YY, unit_diagonal=False, mmd_est="unbiased"):
    """kd.py:38-83 on precomputed m x m kernel matrices (arrays or tensors), fp64 on the device.
    ``kid_features_to_metric(..., mmd_est=...)`` evaluates the same three estimators without ever
    forming the matrices; this function exists for callers that already hold them."""
    assert mmd_est in _MMD_EST, "Invalid value of mmd_est"                   # kd.py:39-43
    dev = _lib.require_cuda(K_XX.device if isinstance(K_XX, torch.Tensor) and K_XX.is_cuda else None)
    K_XX, K_XY, K_YY = (torch.as_tensor(k).to(dev, torch.float64) for k in (K_XX, K_XY, K_YY))
    m = K_XX.shape[0]
    assert K_XX.shape == (m, m) and K_XY.shape == (m, m) and K_YY.shape == (m, m)   # kd.py:45-48
    if unit_diagonal:
        diag_X = diag_Y = torch.ones((), dtype=torch.float64, device=dev)
        sum_diag_X = sum_diag_Y = float(m)
    else:
        diag_X, diag_Y = K_XX.diagonal(), K_YY.diagonal()
        sum_diag_X, sum_diag_Y = diag_X.sum(), diag_Y.sum()
    Kt_XX_sum = (K_XX.sum(dim=1) - diag_X).sum()
    Kt_YY_sum = (K_YY.sum(dim=1) - diag_Y).sum()
    K_XY_sum = K_XY.sum(dim=0).sum()
    if mmd_est == "biased":
        out = (Kt_XX_sum + sum_diag_X) / (m * m) + (Kt_YY_sum + sum_diag_Y) / (m * m) - 2 * K_XY_sum / (m * m)
    else:
        out = (Kt_XX_sum + Kt_YY_sum) / (m * (m - 1))
        if mmd_est == "unbiased":
            out = out - 2 * K_XY_sum / (m * m)
        else:
            out = out - 2 * (K_XY_sum - K_XY.trace()) / (m * (m - 1))
    return float(out)


def kid_features_to_metric(features_1, features_2, **kwargs):
    """kd.py:127-194 with the same keyword arguments; ``return_mmds=True`` adds the
    per-subset values under "mmds".  Two further keywords select what the reference only
    reaches through ``mmd2`` (kd.py:38-83): ``mmd_est`` in ("unbiased", "biased",
    "u-statistic") and ``unit_diagonal``."""
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
    mmd_est = kwargs.get("mmd_est", "unbiased")
    assert mmd_est in _MMD_EST, "Invalid value of mmd_est"                   # kd.py:39-43
    est = _MMD_EST[mmd_est] | (_lib.AMB_MMD_UNIT_DIAGONAL if kwargs.get("unit_diagonal", False) else 0)

    from ..dist import kd_subset_indices

    seed = kwargs.get("rng_seed", 1234)
    if seed is None:       # numpy then seeds from the OS: a fresh draw every call, nothing to cache
        idx = draw_subset_indices(n1, n2, int(kid_subset_size), int(kid_subsets), None)
    else:
        idx = kd_subset_indices(n1, n2, int(kid_subset_size), int(kid_subsets), int(seed))
    idx_dev = torch.from_numpy(np.array(idx, copy=True)).to(dev, non_blocking=True)
    L = _lib.lib()
    mmds = torch.empty(kid_subsets, dtype=torch.float64, device=dev)
    stats = torch.empty(2, dtype=torch.float64, device=dev)
    ws = _lib.workspace(L.amb_kd_ws_bytes(kid_subsets, kid_subset_size, d), dev)
    _lib.check(L.amb_kd_subsets(dev.index, _lib.stream_ptr(dev), f1.data_ptr(), n1, f1.stride(0), f2.data_ptr(), n2,
                                f2.stride(0), d, _lib.dtype_code(f1), idx_dev.data_ptr(), kid_subsets,
                                kid_subset_size, ktype, float(gamma), coef0, degree, sigma, est, mmds.data_ptr(),
                                stats.data_ptr(), ws.data_ptr(), ws.numel()))
    mean, std = stats.tolist()
    out = {KEY_METRIC_KID_MEAN: float(mean), KEY_METRIC_KID_STD: float(std)}   # kd.py:189-192
    if kwargs.get("return_mmds", False):
        out["mmds"] = mmds.cpu().numpy()
    return out

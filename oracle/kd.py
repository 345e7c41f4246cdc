"""Oracle: kernel distance / KID-style MMD^2 (reference metrics/kd.py)."""
from __future__ import annotations

import numpy as np

KID_SUBSETS = 100        # kd.py:19
KID_SUBSET_SIZE = 1000   # kd.py:20
KID_DEGREE = 3           # kd.py:22
KID_COEF0 = 1            # kd.py:24
RNG_SEED = 1234          # kd.py:176


def polynomial_kernel(X, Y, degree=KID_DEGREE, gamma=None, coef0=KID_COEF0):
    """kd.py:112-116: (X Y^T * gamma + coef0) ** degree, gamma = 1/d, in X's dtype."""
    if gamma is None:
        gamma = 1.0 / X.shape[1]
    return (np.matmul(X, Y.T) * gamma + coef0) ** degree


def rbf_kernel(X, Y, sigma=10.0):
    """kd.py:86-109: exp(-sqeuclidean / (2 sigma^2))."""
    from scipy.spatial.distance import cdist

    return np.exp(-cdist(X, Y, "sqeuclidean") / (2 * sigma**2))


def mmd2_unbiased(K_XX, K_XY, K_YY):
    """kd.py:38-83 with mmd_est='unbiased' (the only estimator the facade reaches)."""
    m = K_XX.shape[0]
    assert K_XX.shape == (m, m) and K_XY.shape == (m, m) and K_YY.shape == (m, m)  # kd.py:45-48
    diag_X = np.diagonal(K_XX)
    diag_Y = np.diagonal(K_YY)
    Kt_XX_sum = (K_XX.sum(axis=1) - diag_X).sum()   # kd.py:62,66
    Kt_YY_sum = (K_YY.sum(axis=1) - diag_Y).sum()   # kd.py:63,67
    K_XY_sum = K_XY.sum(axis=0).sum()               # kd.py:64,68
    mmd2 = (Kt_XX_sum + Kt_YY_sum) / (m * (m - 1))  # kd.py:77
    mmd2 -= 2 * K_XY_sum / (m * m)                  # kd.py:79
    return mmd2


def mmd2(K_XX, K_XY, K_YY, unit_diagonal=False, mmd_est="unbiased"):
    """kd.py:38-83 with all three estimators ("biased" and "u-statistic" are only reachable by
    calling mmd2 directly in the reference; kernel_mmd2, kd.py:119-124, hard-codes "unbiased")."""
    assert mmd_est in ("biased", "unbiased", "u-statistic"), "Invalid value of mmd_est"   # kd.py:39-43
    m = K_XX.shape[0]
    assert K_XX.shape == (m, m) and K_XY.shape == (m, m) and K_YY.shape == (m, m)          # kd.py:45-48
    if unit_diagonal:                                   # kd.py:52-54
        diag_X = diag_Y = 1
        sum_diag_X = sum_diag_Y = m
    else:                                               # kd.py:55-60
        diag_X, diag_Y = np.diagonal(K_XX), np.diagonal(K_YY)
        sum_diag_X, sum_diag_Y = diag_X.sum(), diag_Y.sum()
    Kt_XX_sum = (K_XX.sum(axis=1) - diag_X).sum()       # kd.py:62,66
    Kt_YY_sum = (K_YY.sum(axis=1) - diag_Y).sum()       # kd.py:63,67
    K_XY_sum = K_XY.sum(axis=0).sum()                   # kd.py:64,68
    if mmd_est == "biased":                             # kd.py:70-75
        return (Kt_XX_sum + sum_diag_X) / (m * m) + (Kt_YY_sum + sum_diag_Y) / (m * m) - 2 * K_XY_sum / (m * m)
    out = (Kt_XX_sum + Kt_YY_sum) / (m * (m - 1))       # kd.py:77
    if mmd_est == "unbiased":
        return out - 2 * K_XY_sum / (m * m)             # kd.py:79
    return out - 2 * (K_XY_sum - np.trace(K_XY)) / (m * (m - 1))   # kd.py:81


def kd_subset_size(n1, n2, kid_subset_size=KID_SUBSET_SIZE):
    """kd.py:157-168: if subset_size >= min(n1, n2) it becomes max(1, min // 2)."""
    n = min(n1, n2)
    if kid_subset_size >= n:
        return max(1, n // 2)
    return kid_subset_size


def draw_subset_indices(n1, n2, m, subsets=KID_SUBSETS, seed=RNG_SEED):
    """The index stream of kd.py:176,185-186: one PCG64 generator, per subset
    choice(n1, m, replace=False) then choice(n2, m, replace=False).  [S, 2, m] int32."""
    rng = np.random.default_rng(seed)
    idx = np.empty((subsets, 2, m), dtype=np.int32)
    for i in range(subsets):
        idx[i, 0] = rng.choice(n1, m, replace=False)
        idx[i, 1] = rng.choice(n2, m, replace=False)
    return idx


def kernel_distance(f1, f2, subsets=KID_SUBSETS, subset_size=KID_SUBSET_SIZE, seed=RNG_SEED,
                    degree=KID_DEGREE, gamma=None, coef0=KID_COEF0, compute_dtype=None,
                    return_mmds=False, kernel_type="polynomial", sigma=10.0, mmd_est="unbiased",
                    unit_diagonal=False):
    """kd.py:127-194 kid_features_to_metric with the polynomial kernel (or, kernel_type="rbf",
    the RBF kernel of kd.py:86-109 that only the keyword interface reaches).

    f1 = candidate features, f2 = reference features (audio_metrics.py:260 passes
    (cand, ref)).  Arithmetic runs in the input dtype (float32 for embedder output),
    as numpy does in the reference; ``compute_dtype=np.float64`` evaluates the same
    subsets without the reference's fp32 rounding.
    """
    f1 = np.asarray(f1)
    f2 = np.asarray(f2)
    assert f1.ndim == 2 and f2.ndim == 2 and f1.shape[1] == f2.shape[1]  # kd.py:149-151
    n1, n2 = len(f1), len(f2)
    assert n1 and n2
    m = kd_subset_size(n1, n2, subset_size)
    idx = draw_subset_indices(n1, n2, m, subsets, seed)
    if compute_dtype is not None:
        f1 = f1.astype(compute_dtype)
        f2 = f2.astype(compute_dtype)
    mmds = np.zeros(subsets)
    for i in range(subsets):
        a = f1[idx[i, 0]]
        b = f2[idx[i, 1]]
        if kernel_type == "rbf":
            k11, k22, k12 = rbf_kernel(a, a, sigma), rbf_kernel(b, b, sigma), rbf_kernel(a, b, sigma)
        else:
            k11 = polynomial_kernel(a, a, degree, gamma, coef0)   # kd.py:120
            k22 = polynomial_kernel(b, b, degree, gamma, coef0)   # kd.py:121
            k12 = polynomial_kernel(a, b, degree, gamma, coef0)   # kd.py:122
        if mmd_est == "unbiased" and not unit_diagonal:
            mmds[i] = mmd2_unbiased(k11, k12, k22)            # kd.py:124
        else:
            mmds[i] = mmd2(k11, k12, k22, unit_diagonal=unit_diagonal, mmd_est=mmd_est)
    out = {"kernel_distance_mean": float(np.mean(mmds)), "kernel_distance_std": float(np.std(mmds))}
    if return_mmds:
        out["mmds"] = mmds
    return out

"""Oracle: Frechet distance from (mean, covariance) pairs (reference metrics/fad.py)."""
from __future__ import annotations

import numpy as np


def frechet_from_stats(mu_x, sigma_x, mu_y, sigma_y) -> float:
    """fad.py:16-31 (the live code path).

    a = sum (mu_x - mu_y)^2            fad.py:27
    b = tr sigma_x + tr sigma_y        fad.py:28
    c = sum_i Re sqrt(eig_i(sigma_x @ sigma_y))   fad.py:30 — complex eigvals, so a
        negative (round-off) eigenvalue contributes Re sqrt = 0
    returns a + b - 2c                 fad.py:31   (no clamp: may be slightly < 0)
    """
    mu_x = np.asarray(mu_x, dtype=np.float64)
    mu_y = np.asarray(mu_y, dtype=np.float64)
    sigma_x = np.asarray(sigma_x, dtype=np.float64)
    sigma_y = np.asarray(sigma_y, dtype=np.float64)
    a = np.square(mu_x - mu_y).sum()
    b = np.trace(sigma_x) + np.trace(sigma_y)
    ev = np.linalg.eigvals(sigma_x @ sigma_y).astype(np.complex128)
    c = np.sqrt(ev).real.sum()
    return float(a + b - 2 * c)


def frechet_sqrtm(mu_x, sigma_x, mu_y, sigma_y) -> float:
    """The commented-out scipy variant (fad.py:34-92): tr sqrtm(sigma_x . sigma_y).

    north_star states the FAD tolerance against scipy.linalg.sqrtm; both forms
    agree to ~1e-15 relative on full-rank inputs.
    """
    from scipy import linalg

    mu_x = np.asarray(mu_x, dtype=np.float64)
    mu_y = np.asarray(mu_y, dtype=np.float64)
    covmean = linalg.sqrtm(np.asarray(sigma_x, np.float64) @ np.asarray(sigma_y, np.float64))
    if isinstance(covmean, tuple):
        covmean = covmean[0]
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    diff = mu_x - mu_y
    return float(diff @ diff + np.trace(sigma_x) + np.trace(sigma_y) - 2 * np.trace(covmean))

"""Oracle: k-NN radii and precision / recall / density / coverage (reference metrics/prdc.py)."""
from __future__ import annotations

import numpy as np


def cdist_mm(x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """torch.cdist(x1, x2) (p=2, default compute mode) as prdc.py:12,34 call it.

    torch uses the matmul formulation when either side has more than 25 rows:
    one product of the augmented operands [-2 x1, |x1|^2, 1] . [x2, 1, |x2|^2]^T,
    clamp_min(0), sqrt — all in the input dtype.  Small inputs use the direct
    difference form.
    """
    x1 = np.asarray(x1)
    x2 = np.asarray(x2)
    dt = x1.dtype
    if x1.shape[0] > 25 or x2.shape[0] > 25:
        n1 = np.square(x1).sum(axis=1, keepdims=True, dtype=dt)
        n2 = np.square(x2).sum(axis=1, keepdims=True, dtype=dt)
        a = np.concatenate([-2 * x1, n1, np.ones_like(n1)], axis=1)
        b = np.concatenate([x2, np.ones_like(n2), n2], axis=1)
        d2 = a @ b.T
        np.maximum(d2, 0, out=d2)
        return np.sqrt(d2, out=d2)
    diff = x1[:, None, :] - x2[None, :, :]
    return np.sqrt(np.square(diff).sum(axis=-1, dtype=dt))


def cdist_exact(x1, x2) -> np.ndarray:
    """Exact (fp64, difference form) distances — the noise-free yardstick."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    n1 = np.square(x1).sum(axis=1)[:, None]
    n2 = np.square(x2).sum(axis=1)[None, :]
    d2 = n1 + n2 - 2 * (x1 @ x2.T)
    return np.sqrt(np.maximum(d2, 0))


def cdist_diff(x1, x2, chunk=64) -> np.ndarray:
    """Distances in the difference form sqrt(sum (x - y)^2), fp64: no cancellation, so duplicated
    rows are at distance exactly 0 (cdist_exact's |x|^2 + |y|^2 - 2 x.y leaves 1e-8 of noise there).
    O(n m d) memory traffic — for the small duplicate / tie cases only."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    out = np.empty((len(x1), len(x2)))
    for s in range(0, len(x1), chunk):
        diff = x1[s:s + chunk, None, :] - x2[None, :, :]
        out[s:s + chunk] = np.sqrt(np.einsum("ijk,ijk->ij", diff, diff))
    return out


def nearest_neighbour_distances(x, nearest_k, dist=cdist_mm):
    """prdc.py:4-14: kthvalue(cdist(x, x), k + 1) per row (self-distance included)."""
    x = np.asarray(x)
    if nearest_k + 1 > x.shape[0]:
        raise RuntimeError("kthvalue: k out of range")  # torch raises here
    d = dist(x, x)
    return np.partition(d, nearest_k, axis=-1)[:, nearest_k]


def prdc_counts(ref, cand, r_ref, r_cand, dist=cdist_mm):
    """Integer numerators of prdc.py:36-48 (strict '<').

    returns dict(col_count [M] int, recall_rows [N] bool, cover_rows [N] bool)
    """
    D = dist(ref, cand)                                  # prdc.py:34
    in_ref = D < r_ref[:, None]
    col_count = in_ref.sum(axis=0)                       # prdc.py:44
    recall_rows = (D < r_cand[None, :]).any(axis=1)      # prdc.py:41
    cover_rows = D.min(axis=1) < r_ref                   # prdc.py:48
    return dict(col_count=col_count.astype(np.int64), recall_rows=recall_rows, cover_rows=cover_rows)


def prdc_from_counts(c, nearest_k):
    """prdc.py:36-50: the four means, from the integer numerators."""
    col = c["col_count"]
    return dict(
        precision=float((col > 0).astype(np.float64).mean()),
        recall=float(c["recall_rows"].astype(np.float64).mean()),
        density=(1.0 / float(nearest_k)) * float(col.astype(np.float64).mean()),
        coverage=float(c["cover_rows"].astype(np.float64).mean()),
    )


def prdc(ref, cand, nearest_k, dist=cdist_mm):
    """prdc.py:18-50 end to end."""
    ref = np.asarray(ref)
    cand = np.asarray(cand)
    r_ref = nearest_neighbour_distances(ref, nearest_k, dist)    # prdc.py:31 via data.py:60-66
    r_cand = nearest_neighbour_distances(cand, nearest_k, dist)  # prdc.py:32
    return prdc_from_counts(prdc_counts(ref, cand, r_ref, r_cand, dist), nearest_k)


def prdc_counts_chunked(ref, cand, nearest_k, chunk=4096, dist=cdist_mm):
    """Row-blocked restatement of prdc.py for sets whose N x M matrices do not fit
    host RAM (the reference cannot run them at all).  Same arithmetic per block."""
    ref = np.asarray(ref)
    cand = np.asarray(cand)

    def radii(x):
        out = np.empty(len(x), dtype=x.dtype)
        for s in range(0, len(x), chunk):
            d = dist(x[s:s + chunk], x)
            out[s:s + chunk] = np.partition(d, nearest_k, axis=-1)[:, nearest_k]
        return out

    r_ref, r_cand = radii(ref), radii(cand)
    col = np.zeros(len(cand), dtype=np.int64)
    rec = np.zeros(len(ref), dtype=bool)
    cov = np.zeros(len(ref), dtype=bool)
    for s in range(0, len(ref), chunk):
        D = dist(ref[s:s + chunk], cand)
        col += (D < r_ref[s:s + chunk, None]).sum(axis=0)
        rec[s:s + chunk] = (D < r_cand[None, :]).any(axis=1)
        cov[s:s + chunk] = D.min(axis=1) < r_ref[s:s + chunk]
    return dict(col_count=col, recall_rows=rec, cover_rows=cov, r_ref=r_ref, r_cand=r_cand)


def prdc_bracket(ref, cand, nearest_k, eps, chunk=4096):
    """Exact-arithmetic (fp64) counts with every radius shrunk / grown by the
    relative tolerance eps: (lower, upper) numerators.  Any implementation whose
    distances and radii are within eps (relative) of the exact ones has its integer
    counts bracketed element-wise:  lower <= counts <= upper.  This is the
    'counts exact except for ties within eps of the radius' acceptance test."""
    ref64 = np.asarray(ref, dtype=np.float64)
    cand64 = np.asarray(cand, dtype=np.float64)

    def radii(x):
        out = np.empty(len(x))
        for s in range(0, len(x), chunk):
            d = cdist_exact(x[s:s + chunk], x)
            out[s:s + chunk] = np.partition(d, nearest_k, axis=-1)[:, nearest_k]
        return out

    r_ref, r_cand = radii(ref64), radii(cand64)
    fs = (1.0 - eps, 1.0 + eps)
    res = [dict(col_count=np.zeros(len(cand64), dtype=np.int64), recall_rows=np.zeros(len(ref64), dtype=bool),
                cover_rows=np.zeros(len(ref64), dtype=bool)) for _ in fs]
    for s in range(0, len(ref64), chunk):
        D = cdist_exact(ref64[s:s + chunk], cand64)
        for f, out in zip(fs, res):
            # widen/narrow the comparison itself so distance error is covered too
            in_ref = D < f * r_ref[s:s + chunk, None]
            out["col_count"] += in_ref.sum(axis=0)
            out["recall_rows"][s:s + chunk] = (D < f * r_cand[None, :]).any(axis=1)
            out["cover_rows"][s:s + chunk] = in_ref.any(axis=1)
    return res[0], res[1], r_ref, r_cand

"""Stand-in kernels for exercising audio_metrics_b200.dist on CPU (gloo): the same
``ops`` interface as dist.CudaOps, backed by the numpy oracle.  Test
infrastructure only."""
import numpy as np
import torch

import oracle
from oracle.prdc import cdist_exact


class _Set:
    """What dist.py needs from a gathered set: the rows, and a radii cache."""

    def __init__(self, x):
        self.embeddings = x
        self.x = x.numpy()
        self.radii = {}


class OracleOps:
    list_cap = 1 << 20      # capacity reported for the (non-existent) refine list
    overflow_once = False   # tests set this to drive the overflow ladder of dist._fused

    def moments(self, x):
        a = x.numpy().astype(np.float64)
        return torch.from_numpy(np.concatenate([a.sum(axis=0), (a.T @ a).ravel()]))

    def stats_from_moments(self, buf, n, d):
        b = buf.numpy()
        s, g = b[:d], b[d:].reshape(d, d)
        mean = s / n
        cov = np.zeros((d, d)) if n <= 1 else (g - n * np.outer(mean, mean)) / (n - 1)
        return mean, cov

    def frechet_batch(self, pairs):
        return torch.tensor([oracle.frechet_from_stats(sx[0], sx[1], sy[0], sy[1]) for sx, sy in pairs],
                            dtype=torch.float64)

    def container(self, x):
        return _Set(x)

    def radii_rows(self, c, row0, nrows, k):
        d = cdist_exact(c.x[row0:row0 + nrows], c.x)
        r = np.partition(d, k, axis=-1)[:, k] if nrows else np.zeros(0)
        return torch.from_numpy(r.astype(np.float32))

    def count_rows(self, cref, ccand, r_ref, r_cand, row0, nrows, k, list_cap=None):
        D = cdist_exact(cref.x[row0:row0 + nrows], ccand.x).astype(np.float32)
        rr = r_ref.numpy()[row0:row0 + nrows]
        col = (D < rr[:, None]).sum(axis=0).astype(np.int32)
        rec = (D < r_cand.numpy()[None, :]).any(axis=1).sum()
        cov = (D < rr[:, None]).any(axis=1).sum()
        self.calls = getattr(self, "calls", []) + [list_cap]
        if self.overflow_once and list_cap is None and row0 == 0:
            # pretend this rank's list overflowed: the counts of this attempt are garbage
            return (torch.zeros_like(torch.from_numpy(col)), torch.tensor([0, 0], dtype=torch.int64),
                    torch.tensor([self.list_cap + 5, self.list_cap], dtype=torch.int64))
        unc = 0 if list_cap in (None, "exact") else 3
        cap = 0 if list_cap == "exact" else (self.list_cap if list_cap is None else int(list_cap))
        return (torch.from_numpy(col), torch.tensor([int(rec), int(cov)], dtype=torch.int64),
                torch.tensor([unc, cap], dtype=torch.int64))

    def next_list_cap(self, uncertain, n_ref, n_cand):
        return int(uncertain)

    def kd_mmds(self, f1, f2, idx, gamma, coef0, degree, key=None):
        a, b = f1.numpy().astype(np.float64), f2.numpy().astype(np.float64)
        out = np.zeros(len(idx))
        for i in range(len(idx)):
            x, y = a[idx[i, 0]], b[idx[i, 1]]
            out[i] = oracle.mmd2_unbiased(oracle.polynomial_kernel(x, x, degree, gamma, coef0),
                                          oracle.polynomial_kernel(x, y, degree, gamma, coef0),
                                          oracle.polynomial_kernel(y, y, degree, gamma, coef0))
        return torch.from_numpy(out)

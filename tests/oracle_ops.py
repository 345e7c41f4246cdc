"""Stand-in kernels for exercising audio_metrics_b200.dist on CPU (gloo): the same
``ops`` interface as dist.CudaOps, backed by the numpy oracle.  Test
infrastructure only."""
import numpy as np
import torch

import oracle
from oracle.prdc import cdist_exact


class OracleOps:
    def moments(self, x):
        a = x.numpy().astype(np.float64)
        return torch.from_numpy(np.concatenate([a.sum(axis=0), (a.T @ a).ravel()]))

    def stats_from_moments(self, buf, n, d):
        b = buf.numpy()
        s, g = b[:d], b[d:].reshape(d, d)
        mean = s / n
        cov = np.zeros((d, d)) if n <= 1 else (g - n * np.outer(mean, mean)) / (n - 1)
        return mean, cov

    def frechet(self, sx, sy):
        return oracle.frechet_from_stats(sx[0], sx[1], sy[0], sy[1])

    def container(self, x):
        return x.numpy()

    def radii_rows(self, c, row0, nrows, k):
        d = cdist_exact(c[row0:row0 + nrows], c)
        r = np.partition(d, k, axis=-1)[:, k] if nrows else np.zeros(0)
        return torch.from_numpy(r.astype(np.float32))

    def count_rows(self, cref, ccand, r_ref, r_cand, row0, nrows, k):
        D = cdist_exact(cref[row0:row0 + nrows], ccand).astype(np.float32)
        rr = r_ref.numpy()[row0:row0 + nrows]
        col = (D < rr[:, None]).sum(axis=0).astype(np.int32)
        rec = (D < r_cand.numpy()[None, :]).any(axis=1).sum()
        cov = (D < rr[:, None]).any(axis=1).sum()
        return torch.from_numpy(col), torch.tensor([int(rec), int(cov), 0], dtype=torch.int64)

    def check_uncertain(self, uncertain, n_ref, n_cand):
        assert uncertain == 0

    def kd_mmds(self, f1, f2, idx, gamma, coef0, degree):
        a, b = f1.numpy().astype(np.float64), f2.numpy().astype(np.float64)
        out = np.zeros(len(idx))
        for i in range(len(idx)):
            x, y = a[idx[i, 0]], b[idx[i, 1]]
            out[i] = oracle.mmd2_unbiased(oracle.polynomial_kernel(x, x, degree, gamma, coef0),
                                          oracle.polynomial_kernel(x, y, degree, gamma, coef0),
                                          oracle.polynomial_kernel(y, y, degree, gamma, coef0))
        return torch.from_numpy(out)

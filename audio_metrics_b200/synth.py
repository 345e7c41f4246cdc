"""Synthetic CLAP-like embedding sets.

Unit-norm rows built from a shared rank-32 basis plus isotropic noise, so that a
reference and a candidate set overlap and FAD / KD / PRDC are all non-degenerate
(two independent random projections give PRDC == 0).

Two generators with the same model:

* ``make_sets_numpy``: bit-reproducible on any host — only numpy's PCG64 stream
  and element-wise IEEE operations, no BLAS and no SIMD reductions — so golden
  fixtures can pin results by seed instead of by megabytes of input.
* ``make_sets_torch``: the same recipe with torch ops on any device, for
  benchmark-sized inputs (200k x 512) generated directly in HBM.
"""
from __future__ import annotations

import numpy as np
import torch

LATENT = 32
NOISE = 0.1


def _make_numpy(rng, W, n, d, scale, shift, dtype):
    z = rng.standard_normal((n, LATENT)) * scale + shift  # float64
    e = np.zeros((n, d), dtype=np.float64)
    for k in range(LATENT):  # explicit rank-1 updates: element-wise ops only
        e += z[:, k, None] * W[k][None, :]
    e /= np.sqrt(float(LATENT))
    e += NOISE * rng.standard_normal((n, d))
    s = np.zeros(n, dtype=np.float64)
    for k in range(d):  # sequential column accumulation: order fixed
        s += e[:, k] * e[:, k]
    e /= np.sqrt(s)[:, None]
    return np.ascontiguousarray(e.astype(dtype))


def make_sets_numpy(n_ref, n_cand, d, seed=1234, dtype=np.float32, cand_scale=1.15, cand_shift=0.1):
    """(reference [n_ref, d], candidate [n_cand, d]) as numpy arrays."""
    rng = np.random.Generator(np.random.PCG64(seed))
    W = rng.standard_normal((LATENT, d))
    ref = _make_numpy(rng, W, n_ref, d, 1.0, 0.0, dtype)
    cand = _make_numpy(rng, W, n_cand, d, cand_scale, cand_shift, dtype)
    return ref, cand


def make_apa_sets_numpy(n, d, seed=1234, dtype=np.float32):
    """Mix/stem latent model for APA (BASELINE config 3).

    aligned mixes embed (ctx_i + stem_i), misaligned mixes embed (ctx_i + stem_pi(i))
    with pi a fixed cyclic shift; the candidate's aligned mixes come from a shifted
    generator.  Returns dict of float arrays: ref_aligned, ref_misaligned,
    cand_aligned, ref_stems, cand_stems.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    W = rng.standard_normal((LATENT, d))

    def embed(z):
        e = np.zeros((z.shape[0], d), dtype=np.float64)
        for k in range(LATENT):
            e += z[:, k, None] * W[k][None, :]
        e /= np.sqrt(float(LATENT))
        e += NOISE * rng.standard_normal((z.shape[0], d))
        s = np.zeros(z.shape[0], dtype=np.float64)
        for k in range(d):
            s += e[:, k] * e[:, k]
        return np.ascontiguousarray((e / np.sqrt(s)[:, None]).astype(dtype))

    ctx = rng.standard_normal((n, LATENT))
    stem = 0.5 * ctx + rng.standard_normal((n, LATENT))  # stems correlate with their context
    perm = np.roll(np.arange(n), 1)
    cctx = rng.standard_normal((n, LATENT))
    cstem = 0.35 * cctx + 1.1 * rng.standard_normal((n, LATENT)) + 0.05
    return {
        "ref_aligned": embed(ctx + stem),
        "ref_misaligned": embed(ctx + stem[perm]),
        "cand_aligned": embed(cctx + cstem),
        "ref_stems": embed(stem),
        "cand_stems": embed(cstem),
    }


def make_sets_torch(n_ref, n_cand, d, seed=1234, device="cuda", dtype=torch.float32,
                    cand_scale=1.15, cand_shift=0.1):
    """Same model with torch ops on ``device`` (not bit-equal to the numpy stream)."""
    g = torch.Generator(device=device).manual_seed(seed)
    W = torch.randn(LATENT, d, generator=g, device=device, dtype=torch.float32)

    def make(n, scale, shift):
        z = torch.randn(n, LATENT, generator=g, device=device) * scale + shift
        e = z @ W / LATENT**0.5 + NOISE * torch.randn(n, d, generator=g, device=device)
        return torch.nn.functional.normalize(e, dim=-1).to(dtype).contiguous()

    ref = make(n_ref, 1.0, 0.0)
    cand = make(n_cand, cand_scale, cand_shift)
    return ref, cand

"""GPU parity tests on inputs that stress the tensor-core filter of the PRDC path: rows whose
norms span orders of magnitude, non-negative (VGGish-like) activations with a large common
mean, collinear rank-1 sets (the reference's DummyEmbedder), low-dimensional PCA output with a
large spread, and heavy duplication.  The filter's error band is built from set-wide maxima
(csrc/epilogues.cuh), so such inputs put many more pairs into the exact fp64 refine — up to
overflowing the default refine list — and every case must still return exactly what
prdc.py:18-50 computes, through the overflow ladder (amb200.h) where needed.

The checker is the fp64 bracket of oracle.prdc_bracket (counts exact except for pairs within
eps of a radius) and, at the largest size, the chunked restatement of the reference's own
fp32 arithmetic (oracle.prdc_counts_chunked).
"""
import numpy as np
import pytest
import torch

import oracle
from oracle.prdc import cdist_diff, cdist_exact
from audio_metrics_b200 import AudioMetricsData, prdc
from audio_metrics_b200._lib import options
from audio_metrics_b200.metrics.prdc import EXACT, nearest_neighbour_distances, prdc_totals
from audio_metrics_b200.synth import make_sets_numpy

pytestmark = pytest.mark.gpu

PRDC_EPS = 1e-6   # stated tie tolerance: relative distance to the radius


def _amd(x):
    a = AudioMetricsData(store_embeddings=True)
    a.add(torch.from_numpy(np.ascontiguousarray(x)))
    return a


_SEEDS = {"row_scales": 101, "vggish_like": 102, "dummy_rank1": 103, "pca_f64_spread": 104}


def _case(name, n, m):
    rng = np.random.default_rng(_SEEDS[name])
    if name == "row_scales":            # rows scaled uniform(0.01, 30): norms over 3.5 orders of magnitude
        ref, cand = make_sets_numpy(n, m, 512, seed=41)
        ref = (ref * rng.uniform(0.01, 30, size=(n, 1))).astype(np.float32)
        cand = (cand * rng.uniform(0.01, 30, size=(m, 1))).astype(np.float32)
        return ref, cand, 5
    if name == "vggish_like":           # non-negative d=128 activations: large common mean, post-ReLU zeros
        w = rng.standard_normal((24, 128))
        def make(k, shift):
            z = rng.standard_normal((k, 24)) + shift
            return np.maximum(z @ w * 0.3 + 1.5 + 0.2 * rng.standard_normal((k, 128)), 0).astype(np.float32)
        return make(n, 0.0), make(m, 0.15), 5
    if name == "dummy_rank1":           # reference tests/test_audio_metrics.py:22-23: outer(1000 std, arange(10)), fp64
        return (np.outer(rng.random(n) * 300, np.arange(10.0)), np.outer(rng.random(m) * 300, np.arange(10.0)), 5)
    if name == "pca_f64_spread":        # PCA output: fp64, d=10, per-component spread over 4 orders of magnitude
        sc = np.geomspace(100.0, 0.01, 10)
        return (rng.standard_normal((n, 10)) * sc, rng.standard_normal((m, 10)) * sc * 1.1 + 0.05 * sc, 5)
    raise KeyError(name)


def _check_bracket(ref, cand, k, eps=PRDC_EPS, collinear=False):
    """``collinear``: the fp64 yardstick itself (|x|^2 + |y|^2 - 2 x.y) loses 7 digits on such data
    (norms of thousands, distances below one): radii are then checked on a sample of rows against
    the cancellation-free difference form, and the bracket is taken at 4 eps."""
    if collinear:
        eps = 4 * eps
    R, C = _amd(ref), _amd(cand)
    out = prdc(R, C, k)                                   # must not raise, whatever the list does
    # the integer vectors behind it, through whatever rung of the ladder is needed
    cap = None
    while True:
        col, rec, cov, tot = prdc_totals(R, C, k, list_cap=cap)
        unc, used = int(tot[4]), int(tot[5])
        if cap == EXACT or unc <= used:
            break
        cap = unc
    col, rec, cov = col.cpu().numpy(), rec.cpu().numpy().astype(bool), cov.cpu().numpy().astype(bool)
    lo, hi, r_ref, r_cand = oracle.prdc_bracket(ref, cand, k, eps)
    assert (lo["col_count"] <= col).all() and (col <= hi["col_count"]).all()
    assert (lo["recall_rows"] <= rec).all() and (rec <= hi["recall_rows"]).all()
    assert (lo["cover_rows"] <= cov).all() and (cov <= hi["cover_rows"]).all()
    n, m = len(ref), len(cand)
    assert out["precision"] == (col > 0).sum() / m and out["recall"] == rec.sum() / n
    assert out["density"] == (1.0 / k) * (col.sum() / m) and out["coverage"] == cov.sum() / n
    # radii: correctly rounded exact distances
    if collinear:
        rows = np.arange(0, n, max(1, n // 256))
        exact = np.partition(cdist_diff(ref[rows], ref), k, axis=-1)[:, k]
        np.testing.assert_allclose(R.get_radii(k).cpu().numpy()[rows], exact.astype(np.float32), rtol=3e-7, atol=1e-30)
    else:
        np.testing.assert_allclose(R.get_radii(k).cpu().numpy(), r_ref.astype(np.float32), rtol=3e-7, atol=1e-30)
        np.testing.assert_allclose(C.get_radii(k).cpu().numpy(), r_cand.astype(np.float32), rtol=3e-7, atol=1e-30)
    return unc, used


@pytest.mark.parametrize("name", ["row_scales", "vggish_like", "dummy_rank1", "pca_f64_spread"])
def test_heterogeneous_norms_20k(cuda_device, name):
    ref, cand, k = _case(name, 20000, 20480)
    _check_bracket(ref, cand, k, collinear=name == "dummy_rank1")


def test_overflow_ladder_rungs_agree(cuda_device):
    """A refine list far too small for the input: the call reports the overflow, a list of the
    reported size and the exhaustive kernel both give the counts of the default path."""
    ref, cand = make_sets_numpy(3000, 2800, 256, seed=6)
    R, C = _amd(ref), _amd(cand)
    k = 5
    col0, rec0, cov0, tot0 = prdc_totals(R, C, k)
    unc = int(tot0[4])
    assert 0 < unc <= int(tot0[5])
    col1, rec1, cov1, tot1 = prdc_totals(R, C, k, list_cap=max(1, unc // 4))     # overflows ...
    assert int(tot1[4]) == unc and int(tot1[5]) < unc                            # ... and says so, with the same count
    col2, rec2, cov2, tot2 = prdc_totals(R, C, k, list_cap=unc)                  # exactly enough
    assert int(tot2[4]) == unc == int(tot2[5])
    col3, rec3, cov3, _ = prdc_totals(R, C, k, list_cap=EXACT)                   # no filter at all
    for a, b, c in ((col2, rec2, cov2), (col3, rec3, cov3)):
        assert torch.equal(a, col0) and torch.equal(b, rec0) and torch.equal(c, cov0)
    # row shards of the exhaustive kernel reassemble like those of the filtered one
    parts = [prdc_totals(R, C, k, row_range=(r0, min(3000, r0 + 1280) - r0), list_cap=EXACT) for r0 in (0, 1280, 2560)]
    assert torch.equal(sum(p[0] for p in parts), col0)
    assert torch.equal(torch.cat([p[1] for p in parts]), rec0)


def test_collinear_set_overflows_default_list_and_still_matches(cuda_device):
    """20k collinear rows put millions of pairs inside the band (distances are tiny differences of
    norms of a few thousand): more than the default list holds."""
    ref, cand, k = _case("dummy_rank1", 20000, 20000)
    unc, used = _check_bracket(ref, cand, k, collinear=True)
    L = __import__("audio_metrics_b200")._lib.lib()
    assert unc > L.amb_prdc_list_cap(20000, 20000)      # the default list would have overflowed
    assert used >= unc


@pytest.mark.parametrize("dups", [40, 1500])
def test_heavy_duplication(cuda_device, dups):
    """Identical embeddings (silent windows) present in both sets: every duplicate pair is an exact
    tie at distance 0 with radius 0.  The reference returns; so must we, with its counts."""
    base, other = make_sets_numpy(1500, 1500, 64, seed=8)
    ref = np.concatenate([np.repeat(base[:1], dups, axis=0), base[1:1200]])
    cand = np.concatenate([other[:900], np.repeat(base[:1], dups, axis=0)])
    R, C = _amd(ref), _amd(cand)
    k = 5
    info = {}
    r = nearest_neighbour_distances(R, k, info=info).cpu().numpy()
    exact = np.partition(cdist_diff(ref, ref), k, axis=-1)[:, k]
    np.testing.assert_allclose(r, exact.astype(np.float32), rtol=3e-7, atol=0)
    assert (r[:dups] == 0).all()                        # k+1 exact duplicates: radius exactly 0
    assert int(info["n_exhaustive"]) >= 0
    out = prdc(R, C, k)
    want = oracle.prdc(ref, cand, k, dist=lambda a, b: cdist_diff(a, b).astype(np.float32))
    for key in want:
        assert out[key] == pytest.approx(want[key], abs=1e-12), key
    col, rec, cov, tot = prdc_totals(R, C, k)
    assert int(tot[4]) <= int(tot[5])                   # zero radii never enter the band: no overflow from ties


def test_all_rows_identical(cuda_device):
    """The degenerate extreme: one point repeated.  Every radius is 0, nothing is strictly inside."""
    x = np.repeat(np.random.default_rng(0).standard_normal((1, 48)).astype(np.float32), 700, axis=0)
    info = {}
    r = nearest_neighbour_distances(_amd(x), 5, info=info).cpu().numpy()
    assert (r == 0).all()
    assert prdc(_amd(x), _amd(x.copy()), 5) == dict(precision=0.0, recall=0.0, density=0.0, coverage=0.0)


def test_empty_row_shards(cuda_device):
    """dist.shard_rows hands the last ranks empty shards whose row0 = n is not tile aligned."""
    ref, cand = make_sets_numpy(130, 257, 32, seed=3)
    R, C = _amd(ref), _amd(cand)
    assert nearest_neighbour_distances(R, 4, row_range=(130, 0)).shape == (0,)
    col, rec, cov, tot = prdc_totals(R, C, 4, row_range=(130, 0))
    assert int(col.sum()) == 0 and rec.numel() == 0 and int(tot[4]) == 0
    import ctypes as Ct
    from audio_metrics_b200 import _lib
    L = _lib.lib()
    ws = _lib.workspace(L.amb_prdc_ws_bytes(130, 257), cuda_device)
    r_ref, r_cand = R.get_radii(4), C.get_radii(4)
    unc = torch.full((1,), 7, dtype=torch.int64, device=cuda_device)
    rc = L.amb_prdc_counts(0, None, R.embeddings.data_ptr(), 32, R.packed().data_ptr(), 130, r_ref.data_ptr(),
                           C.embeddings.data_ptr(), 32, C.packed().data_ptr(), 257, r_cand.data_ptr(), 32, 0, 130, 0,
                           col.data_ptr(), rec.data_ptr() or ws.data_ptr(), cov.data_ptr() or ws.data_ptr(),
                           unc.data_ptr(), ws.data_ptr(), ws.numel())
    assert rc == 0 and int(unc) == 0                    # was: AMB_ERR_ARG "row0 must be a multiple of 128"


def test_counts_vs_reference_arithmetic_20k(cuda_device):
    """20k x 20k x 512 against the chunked restatement of the reference's own fp32 arithmetic
    (torch.cdist matmul mode, kthvalue) and the fp64 bracket.  Every difference from the reference
    must be a pair that lies within the reference's own fp32 distance error of its radius."""
    n = m = 20480
    k = 5
    ref, cand = make_sets_numpy(n, m, 512, seed=24)
    R, C = _amd(ref), _amd(cand)
    col, rec, cov, tot = prdc_totals(R, C, k)
    assert int(tot[4]) <= int(tot[5])
    col, rec, cov = col.cpu().numpy(), rec.cpu().numpy().astype(bool), cov.cpu().numpy().astype(bool)
    lo, hi, r_ref64, r_cand64 = oracle.prdc_bracket(ref, cand, k, PRDC_EPS)
    assert (lo["col_count"] <= col).all() and (col <= hi["col_count"]).all()
    assert (lo["recall_rows"] <= rec).all() and (rec <= hi["recall_rows"]).all()
    assert (lo["cover_rows"] <= cov).all() and (cov <= hi["cover_rows"]).all()
    want = oracle.prdc_counts_chunked(ref, cand, k)
    np.testing.assert_allclose(R.get_radii(k).cpu().numpy(), want["r_ref"], rtol=1e-4, atol=1e-6)
    _attribute_differences(ref, cand, col, rec, cov, want["col_count"], want["recall_rows"], want["cover_rows"],
                           r_ref64, r_cand64)


def _attribute_differences(ref, cand, col, rec, cov, col_ref, rec_ref, cov_ref, r_ref64, r_cand64, tol=4e-6):
    """Every entry where our counts differ from the reference's must be explained by a pair whose exact
    distance lies within ``tol`` (relative; the reference's fp32 matmul-mode cdist error on
    unit-scale data is ~1e-6, its radii carry the same) of the radius it is compared with.
    Returns the number of attributed differences."""
    n_diff = 0
    for j in np.nonzero(col != col_ref)[0]:
        d = cdist_exact(ref, cand[j:j + 1])[:, 0]
        near = np.abs(d - r_ref64) <= tol * r_ref64
        assert near.sum() >= abs(int(col[j]) - int(col_ref[j])), (j, col[j], col_ref[j])
        n_diff += 1
    for i in np.nonzero(rec != rec_ref)[0]:
        d = cdist_exact(ref[i:i + 1], cand)[0]
        assert (np.abs(d - r_cand64) <= tol * r_cand64).any(), i
        n_diff += 1
    for i in np.nonzero(cov != cov_ref)[0]:
        d = cdist_exact(ref[i:i + 1], cand)[0]
        assert (np.abs(d - r_ref64[i]) <= tol * r_ref64[i]).any(), i
        n_diff += 1
    return n_diff

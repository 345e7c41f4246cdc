"""GPU tests of the alternative code paths behind the same entry points, and of
size-independent properties at BASELINE.json's full size (200k x 512).

* the single-MMA filter (d <= 512) and the three-MMA split sweep must give the SAME
  radii and counts, because both are only filters in front of the exact fp64 refine;
* the integer tensor-core covariance (fp32, n >= 4096) against the fp64 oracle on
  inputs chosen to stress its per-column fixed-point grid;
* the pivoted-Cholesky factor path of the Frechet distance against the Jacobi
  eigen-factor path and the oracle on rank-deficient statistics.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle.prdc import cdist_exact
from audio_metrics_b200 import AudioMetricsData, frechet_distance, prdc
from audio_metrics_b200._lib import options
from audio_metrics_b200.metrics.prdc import nearest_neighbour_distances, prdc_totals
from audio_metrics_b200.synth import make_sets_numpy, make_sets_torch

pytestmark = pytest.mark.gpu


def _amd(x, store=True):
    a = AudioMetricsData(store_embeddings=store)
    a.add(torch.from_numpy(x) if isinstance(x, np.ndarray) else x)
    return a


def _stats64(x):
    m, c, _ = oracle.batch_stats(x, compute_dtype=np.dtype(np.float64))
    return m, c


# ------------------------------------------------------------------ engine variants
@pytest.mark.parametrize("n,m,d,k", [(3000, 2600, 512, 5), (1500, 1500, 96, 10), (2000, 1000, 300, 2)])
def test_single_pass_and_split_sweeps_agree(cuda_device, n, m, d, k):
    ref, cand = make_sets_numpy(n, m, d, seed=n + d + k)
    out = {}
    for name, passes, cta2, static in (("pair", 0, 1, 0), ("pair_static", 0, 1, 1), ("single", 0, 0, 0),
                                       ("split", 3, 0, 0)):
        with options(engine_passes=passes, engine_cta2=cta2, engine_static=static):
            R, C = _amd(ref), _amd(cand)
            r_ref = nearest_neighbour_distances(R, k)
            col, rec, cov, tot = prdc_totals(R, C, k)
            out[name] = (r_ref.cpu().numpy(), col.cpu().numpy(), rec.cpu().numpy(), cov.cpu().numpy(), int(tot[4]))
    for other in ("pair_static", "single", "split"):
        for a, b in zip(out["pair"][:4], out[other][:4]):
            assert np.array_equal(a, b)
    # the single-pass band is wider: more pairs go through the exact refine
    assert out["pair"][4] == out["single"][4] >= out["split"][4]


def test_wide_embeddings_use_the_split_sweep(cuda_device):
    """d > 512 does not fit the resident A panel: the three-MMA kernel runs, same contract."""
    ref, cand = make_sets_numpy(700, 600, 700, seed=9)
    got = nearest_neighbour_distances(torch.from_numpy(ref), 5).cpu().numpy()
    exact = np.partition(cdist_exact(ref, ref), 5, axis=-1)[:, 5]
    np.testing.assert_allclose(got, exact.astype(np.float32), rtol=2e-7, atol=1e-12)
    out = prdc(_amd(ref), _amd(cand), 5)
    want = oracle.prdc(ref, cand, 5)
    for key in want:
        assert abs(out[key] - want[key]) <= 3 / 600, (key, out[key], want[key])


def test_row_shards_reassemble(cuda_device):
    """Radii and counts of 128-aligned row shards equal the unsharded result bit for bit
    (what dist.evaluate_sharded relies on)."""
    ref, cand = make_sets_numpy(5000, 4100, 256, seed=21)
    R, C = _amd(ref), _amd(cand)
    k = 5
    full = nearest_neighbour_distances(R, k)
    parts = torch.cat([nearest_neighbour_distances(R, k, row_range=(r0, min(5000, r0 + 1792) - r0))
                       for r0 in range(0, 5000, 1792)])
    assert torch.equal(full, parts)
    r_ref, r_cand = R.get_radii(k), C.get_radii(k)
    col, rec, cov, _ = prdc_totals(R, C, k)
    col2 = torch.zeros_like(col)
    rec2, cov2 = [], []
    for r0 in range(0, 5000, 1792):
        c_, r_, v_, _ = prdc_totals(R, C, k, row_range=(r0, min(5000, r0 + 1792) - r0), ref_radii=r_ref, cand_radii=r_cand)
        col2 += c_
        rec2.append(r_); cov2.append(v_)
    assert torch.equal(col, col2) and torch.equal(rec, torch.cat(rec2)) and torch.equal(cov, torch.cat(cov2))


# ------------------------------------------------------------ tensor-core covariance
def test_tensor_core_covariance_stress(cuda_device):
    rng = np.random.default_rng(5)
    n, d = 9000, 70
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[:, 0] = 0.0                                    # all-zero column
    x[:, 1] = 3.25                                   # constant column: zero variance, large mean
    x[:, 2] = x[:, 3]                                # duplicate columns: singular covariance
    x[:, 4] *= 1e-6                                  # tiny scale
    x[:, 5] *= 1e6                                   # huge scale
    x[::1000, 6] = 5e3                               # heavy tail: max >> rms
    x[:, 7] = 100.0 + 1e-3 * x[:, 7]                 # mean >> spread: cancellation in G - n mu mu^T
    a = _amd(x, store=False)
    m_ref, c_ref = _stats64(x)
    mean, cov = a.mean.cpu().numpy(), a.cov.cpu().numpy()
    colmax = np.abs(x).max(axis=0).astype(np.float64)
    grid = 2.0 ** -30 * np.maximum(colmax, 1e-300) * 2     # one step of the column's fixed-point grid
    assert (np.abs(mean - m_ref) <= 0.5 * grid + 1e-15 * np.abs(m_ref)).all()
    # covariance error: first order in the rounding of each column, |x_k| * grid_l / sqrt(n) scale
    sd = np.sqrt(np.maximum(np.diag(c_ref), 0.0)) + np.abs(m_ref)
    bound = np.outer(sd + grid, grid) + np.outer(grid, sd + grid)
    assert (np.abs(cov - c_ref) <= bound + 1e-12 * np.abs(c_ref)).all()
    assert cov[0, 0] == 0.0 and abs(cov[1, 1]) <= 1e-12
    np.testing.assert_allclose(cov[2, 3], cov[2, 2], rtol=1e-12)
    # well-scaled columns are as good as the FP64-pipe path
    np.testing.assert_allclose(cov[8:, 8:], c_ref[8:, 8:], rtol=1e-8, atol=1e-10)


def test_covariance_paths_agree_and_stream(cuda_device):
    ref, _ = make_sets_numpy(10000, 8, 512, seed=2)
    tc = _amd(ref, store=False)
    tc.cov
    with options(cov_dfma=1):
        fp = _amd(ref, store=False)
        fp.cov                      # statistics are folded in when read: read them under the option
    np.testing.assert_allclose(tc.cov.cpu().numpy(), fp.cov.cpu().numpy(), rtol=1e-8, atol=1e-13)
    np.testing.assert_allclose(tc.mean.cpu().numpy(), fp.mean.cpu().numpy(), rtol=0, atol=2.0 ** -30)
    # two tensor-core blocks merged by Chan's update == one block (reference tests/test_data.py:6-31)
    s = AudioMetricsData(False)
    s.add(torch.from_numpy(ref[:5000])); s.add(torch.from_numpy(ref[5000:]))
    np.testing.assert_allclose(s.cov.cpu().numpy(), tc.cov.cpu().numpy(), rtol=1e-7, atol=1e-12)


# -------------------------------------------------------------------- Frechet paths
def test_frechet_factor_paths_agree(cuda_device):
    """Polar iteration (default) vs one-sided Jacobi, pivoted-Cholesky factors vs Jacobi eigen-factors,
    on singular covariances with n >> d."""
    ref, cand = make_sets_numpy(6000, 5000, 200, seed=4)
    ref[:, 10] = ref[:, 11]                          # singular covariances with n >> d
    cand[:, 10] = cand[:, 11]
    ref[:, 12] = 0.0
    A, B = _amd(cand, False), _amd(ref, False)
    polar = frechet_distance(A, B)
    with options(fad_method=1):
        chol = frechet_distance(A, B)
        with options(fad_factor_eig=1):
            eig = frechet_distance(A, B)
        for bs in (16, 8, 4):                        # Jacobi block sizes
            with options(jacobi_block=bs):
                assert frechet_distance(A, B) == pytest.approx(chol, rel=1e-10)
    mx, cx = _stats64(cand); my, cy = _stats64(ref)
    want = oracle.frechet_from_stats(mx, cx, my, cy)
    assert polar == pytest.approx(want, rel=1e-5)
    assert chol == pytest.approx(want, rel=1e-5)
    assert eig == pytest.approx(want, rel=1e-5)
    assert chol == pytest.approx(eig, rel=1e-7)
    assert polar == pytest.approx(chol, rel=1e-9)    # two algorithms for the same nuclear norm


@pytest.mark.parametrize("case", ["full_rank_d512", "n_lt_d", "rank1", "scaled_columns", "identical", "d33"])
def test_frechet_polar_matches_jacobi_and_svd(cuda_device, case):
    """The GEMM-only polar iteration against the Jacobi path and an fp64 SVD of the same factors'
    product, on the inputs that break naive square-root iterations."""
    rng = np.random.default_rng(7)
    if case == "full_rank_d512":
        ref, cand = make_sets_numpy(3000, 2500, 512, seed=1)
    elif case == "n_lt_d":
        ref, cand = make_sets_numpy(100, 90, 128, seed=12)
    elif case == "rank1":
        ref = np.outer(rng.random(100) * 300, np.arange(10.0)); cand = np.outer(rng.random(100) * 300, np.arange(10.0))
    elif case == "scaled_columns":
        ref, cand = make_sets_numpy(3000, 2500, 256, seed=3)
        sc = np.geomspace(1e-6, 1e6, 256)
        ref, cand = ref * sc, cand * sc               # float64, condition ~1e13 after the factor product
    elif case == "identical":
        ref, _ = make_sets_numpy(2000, 8, 200, seed=5); cand = ref.copy()
    else:
        ref, cand = make_sets_numpy(500, 400, 33, seed=9)
    A, B = _amd(cand, False), _amd(ref, False)
    polar = frechet_distance(A, B)
    with options(fad_method=1):
        jac = frechet_distance(A, B)
    mx, cx = _stats64(cand); my, cy = _stats64(ref)
    scale = np.trace(cx) + np.trace(cy)
    assert abs(polar - jac) <= 1e-10 * scale
    want = oracle.frechet_from_stats(mx, cx, my, cy)
    assert abs(polar - want) <= 1e-5 * abs(want) + 1e-9 * scale


def test_options_and_fused_step(cuda_device):
    """amb_set_option / amb_get_option, and the fused step (one read-back) against the plain
    sequential calls."""
    from audio_metrics_b200 import _lib, kernel_distance
    from audio_metrics_b200.dist import evaluate_containers, evaluate_sharded
    L = _lib.lib()
    assert L.amb_set_option(b"no_such_option", 1) == _lib.AMB_ERR_ARG
    assert L.amb_set_option(b"jacobi_block", 5) == _lib.AMB_ERR_ARG
    for name in (b"jacobi_block", b"fad_ctas", b"engine_reserve_sms"):
        assert L.amb_set_option(name, 16) == 0 and L.amb_get_option(name) == 16 and L.amb_set_option(name, 0) == 0
    with options(engine_passes=3):
        assert L.amb_get_option(b"engine_passes") == 3
    assert L.amb_get_option(b"engine_passes") == 0
    ref, cand = make_sets_numpy(6000, 5200, 512, seed=17)
    R, C = _amd(ref), _amd(cand)
    want = dict(fad=frechet_distance(C, R), **kernel_distance(C, R), **prdc(R, C, 5))
    got = evaluate_sharded(torch.from_numpy(ref).cuda(), torch.from_numpy(cand).cuda(), 6000, 5200, nearest_k=5)
    got2 = evaluate_containers(_amd(ref), _amd(cand), nearest_k=5)
    for g in (got, got2):
        assert g["fad"] == pytest.approx(want["fad"], rel=1e-9)
        for key in ("kernel_distance_mean", "kernel_distance_std"):  # host numpy vs device reduction of the same 100 values
            assert g[key] == pytest.approx(want[key], rel=1e-12), key
        for key in ("precision", "recall", "density", "coverage"):
            assert g[key] == want[key], key
    with options(engine_reserve_sms=100):              # a very narrow sweep gives the same counts
        assert prdc(_amd(ref), _amd(cand), 5) == {k: want[k] for k in ("precision", "recall", "density", "coverage")}


# ------------------------------------------------------- full-size properties (BASELINE N)
def test_full_size_properties(cuda_device):
    """At 200k x 512 no CPU oracle can run; check what must hold for ANY correct
    implementation: a set against itself, shard reassembly, and a planted neighbourhood."""
    n, d, k = 200_000, 512, 5
    ref, cand = make_sets_torch(n, n, d, device=cuda_device)
    R = AudioMetricsData(True); R.embeddings = ref
    same = prdc(R, R, k)
    assert same["precision"] == 1.0 and same["recall"] == 1.0 and same["coverage"] == 1.0
    assert same["density"] == pytest.approx(1.0, abs=1e-4)        # exactly k of k+1 inside, up to exact ties
    S = AudioMetricsData(False); S.add(ref)
    T = AudioMetricsData(False); T.add(ref)
    assert abs(frechet_distance(S, T)) < 1e-9 * float(S.cov.trace())
    # radii are distances to real rows: recompute 64 of them exactly from the k+1 nearest rows
    r = R.get_radii(k)
    rows = torch.arange(0, n, n // 64, device=cuda_device)[:64]
    d2 = torch.cdist(ref[rows].double(), ref.double()) ** 2
    exact = d2.kthvalue(k + 1, dim=1).values.sqrt().float()
    torch.testing.assert_close(r[rows], exact, rtol=3e-7, atol=0)
    # candidate = shifted copy: every reference row has its twin at distance |shift|
    shift = torch.zeros(d, device=cuda_device); shift[0] = 1e-3
    C = AudioMetricsData(True); C.embeddings = ref + shift
    out = prdc(R, C, k)
    assert out["coverage"] == 1.0 and out["precision"] == 1.0 and out["recall"] == 1.0

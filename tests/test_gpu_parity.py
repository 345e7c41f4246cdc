"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on
the same seeded inputs.  Tolerances are the ones BASELINE.json's north_star
states: FAD 1e-5 relative, KD 1e-4 relative (+1e-9), PRDC counts exact outside a
stated epsilon of the radius."""
import numpy as np
import pytest
import torch

import oracle
from oracle.prdc import cdist_exact
from audio_metrics_b200 import AudioMetricsData, frechet_distance, kernel_distance, prdc, apa
from audio_metrics_b200 import _lib
from audio_metrics_b200.metrics.kd import kid_features_to_metric
from audio_metrics_b200.metrics.prdc import nearest_neighbour_distances, prdc_totals
from audio_metrics_b200.synth import make_sets_numpy, make_apa_sets_numpy

pytestmark = pytest.mark.gpu

PRDC_EPS = 1e-6   # stated tie tolerance: relative distance to the radius


def _amd(x, store=True):
    a = AudioMetricsData(store_embeddings=store)
    a.add(torch.from_numpy(x))
    return a


def _stats64(x):
    m, c, _ = oracle.batch_stats(x, compute_dtype=np.dtype(np.float64))
    return m, c


# ----------------------------------------------------------------- pair engine
@pytest.mark.parametrize("na,nb,d", [(128, 256, 32), (300, 700, 512), (77, 1000, 10), (513, 257, 100)])
def test_engine_dot_matrix(cuda_device, na, nb, d):
    rng = np.random.default_rng(na + nb + d)
    A = rng.standard_normal((na, d)).astype(np.float32)
    B = (rng.standard_normal((nb, d)) * rng.uniform(0.01, 30, size=(nb, 1))).astype(np.float32)
    L = _lib.lib()
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    pA = _lib.workspace(L.amb_packed_bytes(na, d), cuda_device)
    pB = _lib.workspace(L.amb_packed_bytes(nb, d), cuda_device)
    C = torch.full((na, nb), float("nan"), dtype=torch.float32, device=cuda_device)
    _lib.check(L.amb_pack(0, None, dA.data_ptr(), 0, na, d, d, pA.data_ptr()))
    _lib.check(L.amb_pack(0, None, dB.data_ptr(), 0, nb, d, d, pB.data_ptr()))
    _lib.check(L.amb_debug_dot_matrix(0, None, pA.data_ptr(), na, pB.data_ptr(), nb, d, C.data_ptr(), nb, 0, 0))
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.linalg.norm(A, axis=1)[:, None] * np.linalg.norm(B, axis=1)[None, :]
    err = np.abs(C.cpu().numpy() - ref) / scale
    assert np.isfinite(err).all()
    assert err.max() < 2.4e-5      # the proven band of the tensor-core pass (epilogues.cuh)
    assert err.max() < 5e-6        # and what it actually achieves


# ------------------------------------------------------------------ statistics
@pytest.mark.parametrize("n,d,dtype", [(1, 8, np.float32), (37, 10, np.float64), (1000, 128, np.float32),
                                        (5000, 512, np.float32), (300, 700, np.float32), (4100, 100, np.float32),
                                        (70000, 200, np.float32), (140000, 64, np.float32)])
def test_stats_single_shot(cuda_device, n, d, dtype):
    rng = np.random.default_rng(n * 7 + d)
    x = (rng.standard_normal((n, d)) * 0.3 + rng.standard_normal(d)).astype(dtype)
    a = _amd(x)
    m_ref, c_ref = _stats64(x)
    assert a.n == n
    # fp32 sets of >= 4096 rows take the integer tensor-core path (cov_tc.cu): exact moments of
    # the data on a per-column grid of 2^-30 max|x_k|, so the mean is within half a grid step
    tc_path = dtype == np.float32 and n >= 4096
    mean_atol = 2.0 ** -31 * float(np.abs(x).max()) * 2 if tc_path else 1e-12
    np.testing.assert_allclose(a.mean.cpu().numpy(), m_ref, rtol=1e-12, atol=mean_atol)
    np.testing.assert_allclose(a.cov.cpu().numpy(), c_ref, rtol=1e-9, atol=1e-11)
    # and within fp32 noise of the reference's own dtype-faithful arithmetic
    m32, c32, _ = oracle.batch_stats(x)
    np.testing.assert_allclose(a.cov.cpu().numpy(), c32, rtol=1e-3, atol=1e-5)


def test_stats_streaming_equals_single_shot(cuda_device):
    """reference tests/test_data.py:6-31, same block sizes and tolerance."""
    torch.manual_seed(0)
    n_dim = 8
    x1, x2, x3 = torch.randn((1, n_dim)), torch.randn((100, n_dim)), torch.randn((1000, n_dim))
    a = AudioMetricsData(store_embeddings=False)
    a.add(x1); a.add(x2); a.add(x3)
    b = AudioMetricsData(store_embeddings=False)
    b.add(torch.cat((x1, x2, x3)))
    torch.testing.assert_close(a.mean, b.mean, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(a.cov, b.cov, rtol=1e-6, atol=1e-6)
    # 32-row pipeline batches vs oracle streaming merge
    x = torch.randn((1000, 24))
    c = AudioMetricsData(store_embeddings=True)
    s = oracle.StreamingStats()
    for i in range(0, 1000, 32):
        c.add(x[i:i + 32]); s.add(x[i:i + 32].numpy())
    np.testing.assert_allclose(c.cov.cpu().numpy(), s.cov, rtol=1e-5, atol=1e-6)
    assert c.embeddings.shape == (1000, 24) and torch.equal(c.embeddings.cpu(), x)
    d = AudioMetricsData(); d += a; d += b
    assert d.n == 2 * 1101


def test_masked_streaming_moments(cuda_device):
    """The pipeline's per-category accumulation (embed.py:226-236) as one masked launch per batch:
    equal to adding the selected rows, for containers with and without an embedding store; deferred
    statistics equal the single-shot ones to round-off (Chan-equivalent, 1e-9)."""
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((3000, 70)) * 0.4 + rng.standard_normal(70)).astype(np.float32)
    cat = rng.integers(1, 4, size=3000).astype(np.int32)
    xd, cd = torch.from_numpy(x).cuda(), torch.from_numpy(cat).cuda()
    for store in (False, True):
        a = AudioMetricsData(store_embeddings=store)
        for i in range(0, 3000, 32):
            sl = slice(i, i + 32)
            a.add_masked(xd[sl], cd[sl], 2, int((cat[sl] == 2).sum()))
        sel = x[cat == 2]
        m_ref, c_ref = _stats64(sel)
        assert a.n == len(sel)
        np.testing.assert_allclose(a.mean.cpu().numpy(), m_ref, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(a.cov.cpu().numpy(), c_ref, rtol=1e-9, atol=1e-12)
        if store:
            assert torch.equal(a.embeddings.cpu(), torch.from_numpy(sel))
    # (1, 1) covariance quirk of recompute_stats with one row, on the incoming side of a merge (data.py:56)
    one = AudioMetricsData(True); one.add(torch.from_numpy(x[:1])); one.recompute_stats()
    assert tuple(one.cov.shape) == (1, 1)
    big = AudioMetricsData(True); big.add(torch.from_numpy(x[1:200]))
    big += one
    m_ref, c_ref = _stats64(x[:200])
    assert big.n == 200
    np.testing.assert_allclose(big.cov.cpu().numpy(), c_ref, rtol=1e-9, atol=1e-12)
    both = big + big
    assert both.n == 400 and both.embeddings.shape == (400, 70)


# ------------------------------------------------------------------------ FAD
@pytest.mark.parametrize("n,m,d", [(2000, 2000, 64), (100, 100, 128), (3000, 2500, 512), (50, 80, 33)])
def test_fad_matches_oracle(cuda_device, n, m, d):
    ref, cand = make_sets_numpy(n, m, d, seed=d)
    a, b = _amd(cand, False), _amd(ref, False)
    got = frechet_distance(a, b)
    mx, cx = _stats64(cand); my, cy = _stats64(ref)
    want = oracle.frechet_from_stats(mx, cx, my, cy)
    want_sqrtm = oracle.frechet_sqrtm(mx, cx, my, cy)
    assert got == pytest.approx(want, rel=1e-5)
    assert got == pytest.approx(want_sqrtm, rel=1e-5)


def test_fad_degenerate(cuda_device):
    rng = np.random.default_rng(0)
    # rank-1 embeddings of the reference's DummyEmbedder (tests/test_audio_metrics.py:22-23)
    a = np.outer(rng.random(100) * 300, np.arange(10.0))
    b = np.outer(rng.random(100) * 300, np.arange(10.0))
    A, B = _amd(a, False), _amd(b, False)
    mx, cx = _stats64(a); my, cy = _stats64(b)
    scale = np.trace(cx) + np.trace(cy)
    assert frechet_distance(A, B) == pytest.approx(oracle.frechet_from_stats(mx, cx, my, cy), rel=1e-5, abs=1e-9 * scale)
    # identical sets: ~0 relative to the traces (reference itself returns round-off here)
    assert abs(frechet_distance(A, A)) < 1e-6 * scale
    # analytic: N(0, I) vs N(mu, s^2 I) -> |mu|^2 + d (1 - s)^2
    d = 16
    I = AudioMetricsData(False); J = AudioMetricsData(False)
    I.mean, I.cov, I.n = torch.zeros(d, dtype=torch.float64), torch.eye(d, dtype=torch.float64), 10
    J.mean, J.cov, J.n = torch.full((d,), 0.5, dtype=torch.float64), 4.0 * torch.eye(d, dtype=torch.float64), 10
    assert frechet_distance(I, J) == pytest.approx(d * 0.25 + d * 1.0, rel=1e-12)


def test_apa_matches_oracle(cuda_device):
    s = make_apa_sets_numpy(1500, 128, seed=5)
    cand, ref, anti = _amd(s["cand_aligned"], False), _amd(s["ref_aligned"], False), _amd(s["ref_misaligned"], False)
    got = apa(cand, ref, anti)
    st = lambda k: _stats64(s[k])
    want = oracle.apa(st("cand_aligned"), st("ref_aligned"), st("ref_misaligned"))
    assert got == pytest.approx(want, rel=1e-5, abs=1e-7)
    assert 0.0 <= got <= 1.0


# ------------------------------------------------------------------------- KD
@pytest.mark.parametrize("n,m,d,dtype", [(1500, 1300, 512, np.float32), (100, 100, 128, np.float32),
                                          (2600, 2100, 64, np.float32), (400, 300, 10, np.float64)])
def test_kd_matches_oracle(cuda_device, n, m, d, dtype):
    ref, cand = make_sets_numpy(n, m, d, seed=n + d, dtype=dtype)
    got = kid_features_to_metric(torch.from_numpy(cand), torch.from_numpy(ref), return_mmds=True)
    want64 = oracle.kernel_distance(cand, ref, compute_dtype=np.float64, return_mmds=True)
    np.testing.assert_allclose(got["mmds"], want64["mmds"], rtol=1e-4, atol=1e-9)
    assert got["kernel_distance_mean"] == pytest.approx(want64["kernel_distance_mean"], rel=1e-4, abs=1e-9)
    assert got["kernel_distance_std"] == pytest.approx(want64["kernel_distance_std"], rel=1e-4, abs=1e-9)
    # the reference's own (input-dtype) arithmetic is noisier than that; we must sit inside its noise
    want = oracle.kernel_distance(cand, ref)
    ref_noise = abs(want["kernel_distance_mean"] - want64["kernel_distance_mean"])
    assert abs(got["kernel_distance_mean"] - want["kernel_distance_mean"]) <= 2 * ref_noise + 1e-9


# ----------------------------------------------------------------------- PRDC
@pytest.mark.parametrize("n,d,k", [(1000, 512, 5), (300, 64, 10), (2500, 128, 1), (129, 10, 29), (4000, 512, 5)])
def test_radii_match_exact(cuda_device, n, d, k):
    x, _ = make_sets_numpy(n, 8, d, seed=k + n)
    got = nearest_neighbour_distances(torch.from_numpy(x), k).cpu().numpy()
    exact = np.partition(cdist_exact(x, x), k, axis=-1)[:, k]
    np.testing.assert_allclose(got, exact.astype(np.float32), rtol=2e-7, atol=1e-12)   # correctly rounded exact value
    ref32 = oracle.nearest_neighbour_distances(x, k)                                   # reference arithmetic
    np.testing.assert_allclose(got, ref32, rtol=1e-4, atol=1e-6)


def test_radii_errors(cuda_device):
    x = torch.randn(8, 16)
    with pytest.raises(ValueError):
        nearest_neighbour_distances(x, 8)      # k+1 > n: torch.kthvalue raises in the reference
    with pytest.raises(ValueError):
        nearest_neighbour_distances(x, 0)


def _check_counts(ref, cand, k, eps=PRDC_EPS):
    R, C = _amd(ref), _amd(cand)
    col, rec, cov, _ = prdc_totals(R, C, k)
    col, rec, cov = col.cpu().numpy(), rec.cpu().numpy().astype(bool), cov.cpu().numpy().astype(bool)
    lo, hi, r_ref, r_cand = oracle.prdc_bracket(ref, cand, k, eps)
    assert (lo["col_count"] <= col).all() and (col <= hi["col_count"]).all()
    assert (lo["recall_rows"] <= rec).all() and (rec <= hi["recall_rows"]).all()
    assert (lo["cover_rows"] <= cov).all() and (cov <= hi["cover_rows"]).all()
    out = prdc(R, C, k)
    want = oracle.prdc(ref, cand, k)   # reference arithmetic (fp32 matmul distances)
    n, m = len(ref), len(cand)
    tol = dict(precision=3 / m, recall=3 / n, density=6 / (k * m), coverage=3 / n)
    for key in want:
        assert abs(out[key] - want[key]) <= tol[key], (key, out[key], want[key])
    return out, want


@pytest.mark.parametrize("n,m,d,k", [(1000, 1200, 512, 5), (300, 200, 64, 10), (2500, 2500, 128, 3), (130, 1000, 10, 2)])
def test_prdc_counts_bracketed_and_close_to_reference(cuda_device, n, m, d, k):
    ref, cand = make_sets_numpy(n, m, d, seed=n + m + k)
    _check_counts(ref, cand, k)


def test_prdc_known_answers(cuda_device):
    ref, cand = make_sets_numpy(600, 600, 64, seed=3)
    same = prdc(_amd(ref), _amd(ref.copy()), 5)
    assert same["precision"] == 1.0 and same["recall"] == 1.0 and same["coverage"] == 1.0
    assert same["density"] == pytest.approx(1.0)          # strict '<' excludes the k-th neighbour tie
    far = cand + 10.0
    out = prdc(_amd(ref), _amd(far.astype(np.float32)), 5)
    assert out == dict(precision=0.0, recall=0.0, density=0.0, coverage=0.0)
    # duplicates: ties at distance 0
    dup = np.concatenate([ref[:50]] * 4)
    out = prdc(_amd(dup), _amd(dup.copy()), 3)
    want = oracle.prdc(dup, dup, 3, dist=cdist_exact)
    assert out == pytest.approx(want)


# ------------------------------------------------------------- C ABI, host buffers
def test_host_buffer_entry_points(cuda_device):
    import ctypes as C
    L = _lib.lib()
    ref, cand = make_sets_numpy(700, 900, 128, seed=11)
    out = (C.c_double * 4)()
    _lib.check(L.amb_host_prdc(0, ref.ctypes.data, 700, cand.ctypes.data, 900, 128, 0, 5, out))
    want = prdc(_amd(ref), _amd(cand), 5)
    assert list(out) == [want["precision"], want["recall"], want["density"], want["coverage"]]
    mean = np.empty(128); cov = np.empty((128, 128))
    _lib.check(L.amb_host_stats(0, ref.ctypes.data, 0, 700, 128, mean.ctypes.data, cov.ctypes.data))
    m_ref, c_ref = _stats64(ref)
    np.testing.assert_allclose(cov, c_ref, rtol=1e-9, atol=1e-12)
    mx, cx = _stats64(cand)
    fad = C.c_double()
    _lib.check(L.amb_host_frechet(0, 128, mx.ctypes.data, cx.ctypes.data, m_ref.ctypes.data, c_ref.ctypes.data, C.byref(fad)))
    assert fad.value == pytest.approx(oracle.frechet_from_stats(mx, cx, m_ref, c_ref), rel=1e-5)
    idx = oracle.draw_subset_indices(900, 700, 350, 20)
    stats = (C.c_double * 2)()
    _lib.check(L.amb_host_kd(0, cand.ctypes.data, 900, ref.ctypes.data, 700, 128, 0, idx.ctypes.data, 20, 350,
                             1.0 / 128, 1.0, 3, None, stats))
    want_kd = oracle.kernel_distance(cand, ref, subsets=20, compute_dtype=np.float64)
    assert stats[0] == pytest.approx(want_kd["kernel_distance_mean"], rel=1e-4, abs=1e-9)
    radii = np.empty(700, dtype=np.float32)
    _lib.check(L.amb_host_knn_radii(0, ref.ctypes.data, 0, 700, 128, 5, radii.ctypes.data))
    exact = np.partition(cdist_exact(ref, ref), 5, axis=-1)[:, 5]
    np.testing.assert_allclose(radii, exact.astype(np.float32), rtol=2e-7)
    assert L.amb_host_prdc(0, ref.ctypes.data, 700, cand.ctypes.data, 900, 128, 0, 0, out) == _lib.AMB_ERR_ARG


def test_host_evaluate_one_call(cuda_device):
    """amb_host_evaluate: FAD + KD + PRDC of two numpy arrays in one C call (and, with two GPUs, sharded
    over both inside the process) equals the Python path on the same inputs."""
    import ctypes as C
    from audio_metrics_b200.dist import evaluate_containers
    L = _lib.lib()
    ref, cand = make_sets_numpy(3100, 2700, 96, seed=23)
    idx = oracle.draw_subset_indices(2700, 3100, 500, 30)
    want = evaluate_containers(_amd(ref), _amd(cand), ("fad", "kd", "prdc"), nearest_k=5, kd_subsets=30, kd_subset_size=500)
    keys = ("fad", "kernel_distance_mean", "kernel_distance_std", "precision", "recall", "density", "coverage")
    for n_dev in range(1, min(2, torch.cuda.device_count()) + 1):
        devs = (C.c_int * n_dev)(*range(n_dev))
        out = (C.c_double * 7)()
        _lib.check(L.amb_host_evaluate(devs, n_dev, ref.ctypes.data, 3100, cand.ctypes.data, 2700, 96, 0, 5,
                                       idx.ctypes.data, 30, 500, 1, out))
        got = dict(zip(keys, out))
        assert got["fad"] == pytest.approx(want["fad"], rel=1e-12)
        for key in keys[1:3]:
            assert got[key] == pytest.approx(want[key], rel=1e-12), key
        for key in keys[3:]:
            assert got[key] == want[key], key
    # subsets of the metrics: the others come back as NaN
    out = (C.c_double * 7)()
    devs = (C.c_int * 1)(0)
    _lib.check(L.amb_host_evaluate(devs, 1, ref.ctypes.data, 3100, cand.ctypes.data, 2700, 96, 0, 0, None, 0, 0, 1, out))
    assert out[0] == pytest.approx(want["fad"], rel=1e-12) and all(np.isnan(v) for v in list(out)[1:])
    assert L.amb_host_evaluate(devs, 1, ref.ctypes.data, 3, cand.ctypes.data, 2700, 96, 0, 5, None, 0, 0, 1, out) == _lib.AMB_ERR_ARG


def test_comm_single_device(cuda_device):
    """amb_comm_*: a communicator over one device — allreduce and allgather are identities, argument
    errors are reported, NCCL is bound at run time (the library itself does not link it)."""
    import ctypes as C
    L = _lib.lib()
    comm = C.c_void_p()
    devs = (C.c_int * 1)(0)
    _lib.check(L.amb_comm_init(devs, 1, C.byref(comm)))
    try:
        assert L.amb_comm_size(comm) == 1 and L.amb_comm_device(comm, 0) == 0 and L.amb_comm_device(comm, 1) == -1
        dev = torch.device("cuda", 0)
        st = _lib.stream_ptr(dev)
        x = torch.arange(1000, dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        _lib.check(L.amb_comm_allreduce(comm, 0, x.data_ptr(), y.data_ptr(), 1000, _lib.AMB_F64, _lib.AMB_SUM, st))
        c = torch.arange(77, dtype=torch.int32, device=dev)
        _lib.check(L.amb_comm_allreduce(comm, 0, c.data_ptr(), c.data_ptr(), 77, _lib.AMB_I32, _lib.AMB_MAX, st))
        g = torch.empty(50, dtype=torch.float32, device=dev)
        src = torch.randn(50, device=dev)
        _lib.check(L.amb_comm_allgather(comm, 0, src.data_ptr(), g.data_ptr(), 50, _lib.AMB_F32, st))
        torch.cuda.synchronize()
        assert torch.equal(x, y) and torch.equal(c, torch.arange(77, dtype=torch.int32, device=dev)) and torch.equal(g, src)
        assert L.amb_comm_allreduce(comm, 1, x.data_ptr(), y.data_ptr(), 10, _lib.AMB_F64, _lib.AMB_SUM, st) == _lib.AMB_ERR_ARG
        assert L.amb_comm_allreduce(comm, 0, x.data_ptr(), y.data_ptr(), 10, 99, _lib.AMB_SUM, st) == _lib.AMB_ERR_ARG
    finally:
        _lib.check(L.amb_comm_destroy(comm))
    two = (C.c_int * 2)(0, 0)
    assert L.amb_comm_init(two, 2, C.byref(comm)) == _lib.AMB_ERR_ARG

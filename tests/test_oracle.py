"""CPU tests: the oracle restatement against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py), plus its internal identities."""
import numpy as np
import pytest

import oracle
from oracle.prdc import cdist_exact
from golden_util import SET_CASES, arrays, case_inputs, scalars, unpack_rows


@pytest.mark.parametrize("name", SET_CASES)
def test_oracle_matches_reference_golden(name):
    ref, cand, g = case_inputs(name)
    a = arrays()
    k = g["k"]
    # statistics + FAD (data.py:37-47, fad.py:16-31), argument order (cand, ref)
    mr, cr, _ = oracle.batch_stats(ref)
    mc, cc, _ = oracle.batch_stats(cand)
    np.testing.assert_allclose(mr, a[f"{name}/mean_ref"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(np.diag(cr), a[f"{name}/cov_diag_ref"], rtol=2e-5, atol=1e-9)
    assert oracle.frechet_from_stats(mc, cc, mr, cr) == pytest.approx(g["fad"], rel=2e-6)
    assert oracle.frechet_sqrtm(mc, cc, mr, cr) == pytest.approx(g["fad"], rel=2e-6)
    # KD (kd.py:127-194): same subsets by construction; fp32 BLAS noise only
    kd = oracle.kernel_distance(cand, ref)
    assert kd["kernel_distance_mean"] == pytest.approx(g["kernel_distance_mean"], rel=5e-3, abs=2e-8)
    assert kd["kernel_distance_std"] == pytest.approx(g["kernel_distance_std"], rel=5e-3, abs=2e-8)
    # radii (prdc.py:4-14)
    np.testing.assert_allclose(oracle.nearest_neighbour_distances(ref, k), a[f"{name}/r_ref"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(oracle.nearest_neighbour_distances(cand, k), a[f"{name}/r_cand"], rtol=2e-4, atol=2e-6)
    # counts (prdc.py:34-48): identical up to pairs the fp32 BLAS rounding puts on either side of a radius
    n, m = len(ref), len(cand)
    c = oracle.prdc_counts(ref, cand, a[f"{name}/r_ref"], a[f"{name}/r_cand"])
    assert np.abs(c["col_count"] - a[f"{name}/col_count"]).sum() <= 4
    assert (c["recall_rows"] != unpack_rows(a[f"{name}/recall_rows"], n)).sum() <= 2
    assert (c["cover_rows"] != unpack_rows(a[f"{name}/cover_rows"], n)).sum() <= 2
    out = oracle.prdc(ref, cand, k)
    for key, tol in (("precision", 3 / m), ("recall", 3 / n), ("density", 6 / (k * m)), ("coverage", 3 / n)):
        assert abs(out[key] - g[f"prdc_{key}"]) <= tol
    # exact-arithmetic bracket contains the reference's counts at the stated epsilon of its fp32 distances
    if n <= 4000:
        lo, hi, _, _ = oracle.prdc_bracket(ref, cand, k, eps=2e-5 if g["dtype"] == "float32" else 1e-9)
        col = a[f"{name}/col_count"]
        assert (lo["col_count"] <= col).all() and (col <= hi["col_count"]).all()


def test_oracle_streaming_stats_golden():
    a = arrays()
    ref = a["stream/ref"]
    s = oracle.StreamingStats()
    for i in range(0, len(ref), 32):
        s.add(ref[i:i + 32])
    np.testing.assert_allclose(s.mean, a["stream/mean"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(s.cov, a["stream/cov"], rtol=1e-5, atol=1e-9)
    s.recompute_stats()
    np.testing.assert_allclose(s.cov, a["stream/cov_recomputed"], rtol=1e-5, atol=1e-9)
    # reference tests/test_data.py:6-31: streaming == single shot
    rng = np.random.default_rng(0)
    x1, x2, x3 = (rng.standard_normal((k, 8)).astype(np.float32) for k in (1, 100, 1000))
    t = oracle.StreamingStats(False)
    for x in (x1, x2, x3):
        t.add(x)
    m, c, _ = oracle.batch_stats(np.concatenate((x1, x2, x3)))
    np.testing.assert_allclose(t.mean, m, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(t.cov, c, rtol=1e-6, atol=1e-6)


def test_oracle_apa_and_rank1_golden():
    from audio_metrics_b200.synth import make_apa_sets_numpy
    from golden_util import digest
    g = scalars()["apa_d128"]
    s = make_apa_sets_numpy(g["n"], g["d"], seed=g["seed"])
    assert digest(np.concatenate([s[k] for k in sorted(s)])) == g["sha256"]
    st = lambda k: oracle.batch_stats(s[k])[:2]
    assert oracle.frechet_from_stats(*st("cand_aligned"), *st("ref_aligned")) == pytest.approx(g["d_y_x"], rel=2e-6)
    assert oracle.apa(st("cand_aligned"), st("ref_aligned"), st("ref_misaligned")) == pytest.approx(g["apa"], rel=1e-5)
    assert oracle.apa(st("cand_aligned"), st("ref_aligned"), st("ref_misaligned"), g["d_x_xp"]) == pytest.approx(g["apa"], rel=1e-5)
    a = arrays()
    r = scalars()["rank1"]
    sa, sb = oracle.batch_stats(a["rank1/a"])[:2], oracle.batch_stats(a["rank1/b"])[:2]
    # rank-deficient product: eigvals' round-off eigenvalues enter through a square root, so LAPACK
    # builds differ at ~2e-6 relative here (analytic value 29835.1807; torch 29835.114, numpy 29835.163)
    assert oracle.frechet_from_stats(*sa, *sb) == pytest.approx(r["fad"], rel=1e-5)
    scale = np.trace(sa[1]) * 2
    assert abs(oracle.frechet_from_stats(*sa, *sa) - r["fad_self"]) < 1e-6 * scale   # pure round-off either way


def test_apa_formula_edges():
    """apa.py:22-32."""
    assert oracle.apa_from_fads(1.0, 1.0, 0.0) == 0.0
    assert oracle.apa_from_fads(0.0, 2.0, 1.0) == 1.0
    assert oracle.apa_from_fads(2.0, 0.0, 1.0) == 0.0
    assert oracle.apa_from_fads(-1.0, 0.5, 1.0) == 0.75
    assert oracle.apa_from_fads(0.3, 0.3, 0.5) == 0.5


def test_kd_subset_rules_and_index_stream():
    assert oracle.kd_subset_size(1000, 5000) == 500      # '>=' in kd.py:160
    assert oracle.kd_subset_size(1001, 5000) == 1000
    assert oracle.kd_subset_size(100, 100) == 50
    assert oracle.kd_subset_size(1, 7) == 1
    idx = oracle.draw_subset_indices(50, 60, 20, subsets=3)
    rng = np.random.default_rng(1234)
    assert (idx[0, 0] == rng.choice(50, 20, replace=False)).all()
    assert (idx[0, 1] == rng.choice(60, 20, replace=False)).all()
    assert (idx[1, 0] == rng.choice(50, 20, replace=False)).all()
    assert len(set(idx[2, 1].tolist())) == 20


def test_prdc_oracle_identities():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((400, 16)).astype(np.float32)
    y = rng.standard_normal((300, 16)).astype(np.float32)
    np.testing.assert_allclose(oracle.cdist_mm(x, y), cdist_exact(x, y), rtol=2e-4, atol=2e-4)
    small = oracle.cdist_mm(x[:10], y[:20])            # direct-difference branch of torch.cdist
    np.testing.assert_allclose(small, cdist_exact(x[:10], y[:20]), rtol=1e-5)
    full = oracle.prdc_counts(x, y, oracle.nearest_neighbour_distances(x, 4), oracle.nearest_neighbour_distances(y, 4))
    chunk = oracle.prdc_counts_chunked(x, y, 4, chunk=64)
    assert (full["col_count"] == chunk["col_count"]).all() and (full["recall_rows"] == chunk["recall_rows"]).all()
    with pytest.raises(RuntimeError):
        oracle.nearest_neighbour_distances(x[:5], 5)
    same = oracle.prdc(x, x.copy(), 5, dist=cdist_exact)
    assert same["precision"] == 1.0 and same["recall"] == 1.0 and same["coverage"] == 1.0
    assert same["density"] == pytest.approx(1.0)


def test_kd_keyword_variants_match_reference_goldens():
    """kd.py:127-194 keyword interface (RBF kernel, degree/gamma/coef0, subset count/size, seed):
    the oracle against values produced by the unmodified reference (tests/golden/make_golden_kd.py)."""
    import json
    from pathlib import Path

    from audio_metrics_b200.synth import make_sets_numpy

    g = json.loads((Path(__file__).parent / "golden" / "golden_kd_variants.json").read_text())
    i = g["input"]
    ref, cand = make_sets_numpy(i["n_ref"], i["n_cand"], i["d"], seed=i["seed"])
    for name, v in g["variants"].items():
        kw = v["kwargs"]
        args = dict(subsets=kw.get("kid_subsets", 100), subset_size=kw.get("kid_subset_size", 1000),
                    seed=kw.get("rng_seed", 1234), degree=kw.get("kid_degree", 3), gamma=kw.get("kid_gamma"),
                    coef0=kw.get("kid_coef0", 1), kernel_type=kw.get("kernel_type", "polynomial"),
                    sigma=kw.get("kid_sigma", 10.0))
        got64 = oracle.kernel_distance(cand, ref, compute_dtype=np.float64, **args)
        for key in ("kernel_distance_mean", "kernel_distance_std"):
            assert got64[key] == pytest.approx(v["reference_f64"][key], rel=1e-9, abs=1e-15), (name, key)
        got32 = oracle.kernel_distance(cand, ref, **args)
        for key in ("kernel_distance_mean", "kernel_distance_std"):
            assert got32[key] == pytest.approx(v["reference_f32"][key], rel=5e-3, abs=2e-8), (name, key)


def test_mmd2_estimators_match_reference_goldens():
    """kd.py:38-83 "biased" / "u-statistic" / unit_diagonal: the oracle against values produced by the
    unmodified reference's own mmd2 (tests/golden/make_golden_kd_estimators.py)."""
    import json
    from pathlib import Path

    from audio_metrics_b200.synth import make_sets_numpy

    g = json.loads((Path(__file__).parent / "golden" / "golden_kd_estimators.json").read_text())
    i = g["input"]
    ref, cand = make_sets_numpy(i["n_ref"], i["n_cand"], i["d"], seed=i["seed"])
    for name, v in g["variants"].items():
        kw = v["kwargs"]
        args = dict(subsets=g["subsets"], subset_size=g["subset_size"], seed=g["seed"],
                    kernel_type=kw.get("kernel_type", "polynomial"), sigma=kw.get("kid_sigma", 10.0),
                    mmd_est=kw["mmd_est"], unit_diagonal=kw.get("unit_diagonal", False), return_mmds=True)
        got64 = oracle.kernel_distance(cand, ref, compute_dtype=np.float64, **args)
        np.testing.assert_allclose(got64["mmds"], v["reference_f64"]["mmds"], rtol=1e-9, atol=1e-15)
        got32 = oracle.kernel_distance(cand, ref, **args)
        np.testing.assert_allclose(got32["mmds"], v["reference_f32"]["mmds"], rtol=5e-3, atol=2e-7)


def test_oracle_c3_golden():
    """BASELINE config 3 at its stated size (10k mix / stem pairs, d = 512): the oracle's FAD and APA
    against the unmodified reference (tests/golden/make_golden_c3.py)."""
    import hashlib
    import json
    from pathlib import Path

    from audio_metrics_b200.synth import make_apa_sets_numpy

    g = json.loads((Path(__file__).parent / "golden" / "golden_c3.json").read_text())
    s = make_apa_sets_numpy(g["n"], g["d"], seed=g["seed"])
    assert hashlib.sha256(np.ascontiguousarray(np.concatenate([s[k] for k in sorted(s)])).tobytes()).hexdigest() == g["sha256"]
    f64 = np.dtype(np.float64)
    st = {k: oracle.batch_stats(v)[:2] for k, v in s.items()}             # reference arithmetic: input dtype, then fp64
    fad = lambda a, b: oracle.frechet_from_stats(*st[a], *st[b])
    assert fad("cand_aligned", "ref_aligned") == pytest.approx(g["d_y_x"], rel=1e-6)
    assert fad("ref_aligned", "ref_misaligned") == pytest.approx(g["d_x_xp"], rel=1e-6)
    assert fad("cand_stems", "ref_stems") == pytest.approx(g["fad_stems"], rel=1e-6)
    assert oracle.apa(st["cand_aligned"], st["ref_aligned"], st["ref_misaligned"]) == pytest.approx(g["apa"], rel=1e-5)

"""GPU tests: the CUDA path against golden vectors produced by the unmodified
reference (tests/golden/), same tolerances as the oracle is held to."""
import numpy as np
import pytest
import torch

from audio_metrics_b200 import AudioMetricsData, frechet_distance, kernel_distance, prdc, apa
from audio_metrics_b200.metrics.prdc import prdc_totals
from golden_util import SET_CASES, arrays, case_inputs, scalars, unpack_rows

pytestmark = pytest.mark.gpu


def _amd(x, store=True):
    a = AudioMetricsData(store_embeddings=store)
    a.add(torch.from_numpy(x))
    return a


@pytest.mark.parametrize("name", SET_CASES)
def test_cuda_path_matches_reference_golden(cuda_device, name):
    ref, cand, g = case_inputs(name)
    a = arrays()
    k, n, m = g["k"], len(ref), len(cand)
    R, C = _amd(ref), _amd(cand)
    np.testing.assert_allclose(R.mean.cpu().numpy(), a[f"{name}/mean_ref"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(R.cov.diagonal().cpu().numpy(), a[f"{name}/cov_diag_ref"], rtol=2e-5, atol=1e-9)
    assert frechet_distance(C, R) == pytest.approx(g["fad"], rel=1e-5)                   # north_star: 1e-5
    kd = kernel_distance(C, R)
    # the reference evaluates KD in the input dtype (fp32): its own noise is ~1e-3 relative
    assert kd["kernel_distance_mean"] == pytest.approx(g["kernel_distance_mean"], rel=5e-3, abs=2e-8)
    assert kd["kernel_distance_std"] == pytest.approx(g["kernel_distance_std"], rel=5e-3, abs=2e-8)
    np.testing.assert_allclose(R.get_radii(k).cpu().numpy(), a[f"{name}/r_ref"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(C.get_radii(k).cpu().numpy(), a[f"{name}/r_cand"], rtol=2e-4, atol=2e-6)
    col, rec, cov, _ = prdc_totals(R, C, k)
    assert np.abs(col.cpu().numpy() - a[f"{name}/col_count"]).sum() <= 4
    assert (rec.cpu().numpy().astype(bool) != unpack_rows(a[f"{name}/recall_rows"], n)).sum() <= 2
    assert (cov.cpu().numpy().astype(bool) != unpack_rows(a[f"{name}/cover_rows"], n)).sum() <= 2
    out = prdc(R, C, k)
    for key, tol in (("precision", 3 / m), ("recall", 3 / n), ("density", 6 / (k * m)), ("coverage", 3 / n)):
        assert abs(out[key] - g[f"prdc_{key}"]) <= tol, (key, out[key], g[f"prdc_{key}"])


def test_cuda_streaming_stats_golden(cuda_device):
    a = arrays()
    ref = a["stream/ref"]
    s = AudioMetricsData(True)
    for i in range(0, len(ref), 32):
        s.add(torch.from_numpy(ref[i:i + 32]))
    np.testing.assert_allclose(s.mean.cpu().numpy(), a["stream/mean"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(s.cov.cpu().numpy(), a["stream/cov"], rtol=1e-5, atol=1e-9)
    s.recompute_stats()
    np.testing.assert_allclose(s.cov.cpu().numpy(), a["stream/cov_recomputed"], rtol=1e-5, atol=1e-9)


def test_cuda_apa_and_rank1_golden(cuda_device):
    from audio_metrics_b200.synth import make_apa_sets_numpy
    g = scalars()["apa_d128"]
    s = make_apa_sets_numpy(g["n"], g["d"], seed=g["seed"])
    cand, refa, anti = (_amd(s[k], False) for k in ("cand_aligned", "ref_aligned", "ref_misaligned"))
    assert frechet_distance(cand, refa) == pytest.approx(g["d_y_x"], rel=1e-5)
    assert frechet_distance(refa, anti) == pytest.approx(g["d_x_xp"], rel=1e-5)
    assert apa(cand, refa, anti) == pytest.approx(g["apa"], rel=1e-5)
    assert apa(cand, refa, anti, g["d_x_xp"]) == pytest.approx(g["apa"], rel=1e-5)
    assert frechet_distance(_amd(s["cand_stems"], False), _amd(s["ref_stems"], False)) == pytest.approx(g["fad_stems"], rel=1e-5)
    a = arrays()
    r = scalars()["rank1"]
    A, B = _amd(a["rank1/a"], False), _amd(a["rank1/b"], False)
    assert frechet_distance(A, B) == pytest.approx(r["fad"], rel=1e-5)
    scale = float(A.cov.trace()) * 2
    assert abs(frechet_distance(A, A) - r["fad_self"]) < 1e-6 * scale


def test_kd_keyword_variants_match_reference(cuda_device):
    """The keyword interface of kid_features_to_metric (kd.py:127-194: RBF kernel, polynomial
    parameters, subset count / size, seed) against the unmodified reference's values."""
    import json
    from pathlib import Path

    import torch

    from audio_metrics_b200.metrics.kd import kid_features_to_metric
    from audio_metrics_b200.synth import make_sets_numpy

    g = json.loads((Path(__file__).parent / "golden" / "golden_kd_variants.json").read_text())
    i = g["input"]
    ref, cand = make_sets_numpy(i["n_ref"], i["n_cand"], i["d"], seed=i["seed"])
    for name, v in g["variants"].items():
        got = kid_features_to_metric(torch.from_numpy(cand), torch.from_numpy(ref), **v["kwargs"])
        for key in ("kernel_distance_mean", "kernel_distance_std"):
            want64, want32 = v["reference_f64"][key], v["reference_f32"][key]
            assert got[key] == pytest.approx(want64, rel=1e-4, abs=1e-9), (name, key)      # north-star tolerance
            assert abs(got[key] - want32) <= 2 * abs(want32 - want64) + 1e-4 * abs(want64) + 1e-9, (name, key)

"""API conformance of the AudioMetrics facade, after the reference's own tests
(src/audio_metrics/tests/test_audio_metrics.py): the same DummyEmbedder, the three
input layouts, the APA error case and the save/load round trip."""
import numpy as np
import pytest
import torch

from audio_metrics_b200 import AudioMetrics, AudioMetricsData

pytestmark = pytest.mark.gpu


class DummyEmbedder:   # reference tests/test_audio_metrics.py:7-24, on the GPU
    def __init__(self, d=10):
        self.m = torch.nn.Linear(1, 1).cuda()
        self.d = d

    @property
    def sr(self):
        return 16000

    def get_device(self):
        return next(self.m.parameters()).device

    @torch.no_grad()
    def forward(self, data, sr=None):
        mean = torch.as_tensor(10**3 * data["audio"].std(axis=1), device=self.get_device())
        return {"embedding": torch.outer(mean, torch.arange(self.d, device=self.get_device()))}


class RandomEmbedder(DummyEmbedder):
    """Full-rank embeddings so that KD / PRDC are non-degenerate."""

    def __init__(self, d=24):
        super().__init__(d)
        g = torch.Generator().manual_seed(0)
        self.W = torch.randn(64, d, generator=g).cuda()

    @torch.no_grad()
    def forward(self, data, sr=None):
        a = torch.as_tensor(data["audio"][:, :6400].reshape(len(data["audio"]), 64, 100), dtype=torch.float32).cuda()
        return {"embedding": a.std(dim=2) @ self.W}


def mix_func(audio, sr=None):
    return audio.mean(axis=1)


def _am(**kw):
    args = dict(embedder=DummyEmbedder(), mix_function=mix_func, metrics=["fad", "apa"], n_pca=10)
    args.update(kw)
    return AudioMetrics(**args)


def _check(res, keys):
    assert set(res) == set(keys)
    assert all(isinstance(v, float) and np.isfinite(v) for v in res.values())


def test_inputs_ndarray_generator_tensor(cuda_device):
    sr, n = 16000, 5 * 16000
    rng = np.random.default_rng(0)
    am = _am()
    am.add_reference(rng.random((60, n, 2)))
    _check(am.evaluate(rng.random((60, n, 2))), ["fad", "apa"])
    am = _am()
    am.add_reference(rng.random((n, 2)) for _ in range(40))
    _check(am([rng.random((n, 2)) for _ in range(40)]), ["fad", "apa"])
    am = _am()
    am.add_reference(torch.randn((40, n, 2)))
    _check(am.evaluate(torch.randn((40, n, 2))), ["fad", "apa"])
    am = _am(metrics=["fad"])
    am.add_reference(rng.random((40, n)))
    _check(am.evaluate(rng.random((40, n))), ["fad"])


def test_apa_only_and_metric_subsets(cuda_device):
    """metrics=["apa"] (no stem sets at all), single metrics, and the fused result against the
    separate metric functions on the facade's own containers."""
    from audio_metrics_b200 import apa, frechet_distance, kernel_distance, prdc
    rng = np.random.default_rng(4)
    n = 5 * 16000
    ref, cand = rng.standard_normal((90, n, 2)), rng.standard_normal((80, n, 2))
    am = AudioMetrics(embedder=RandomEmbedder(), mix_function=mix_func, metrics=["apa"])
    am.add_reference(ref)
    out = am.evaluate(cand)
    _check(out, ["apa"])
    assert am.stem_reference is None and 0.0 <= out["apa"] <= 1.0
    for metrics, keys in ((["kd"], ["kernel_distance_mean", "kernel_distance_std"]),
                          (["prdc"], ["precision", "recall", "density", "coverage"]), (["fad"], ["fad"])):
        am = AudioMetrics(embedder=RandomEmbedder(), mix_function=mix_func, metrics=metrics)
        am.add_reference(ref)
        _check(am.evaluate(cand), keys)
    # fused step == the reference's one-call-per-metric sequence (audio_metrics.py:254-272) on the same containers
    am = AudioMetrics(embedder=RandomEmbedder(), mix_function=mix_func, metrics=["fad", "kd", "prdc", "apa"])
    am.add_reference(ref)
    got = am.evaluate(cand)
    from audio_metrics_b200.pipeline import ItemCategory, embedding_pipeline
    data = am._embed(cand, "candidate")
    sc, ac = data[ItemCategory.stem], data[ItemCategory.aligned]
    k = max(1, min(10, len(am.stem_reference), len(sc)))
    want = dict(fad=frechet_distance(sc, am.stem_reference), **kernel_distance(sc, am.stem_reference),
                **prdc(am.stem_reference, sc, k), apa=apa(ac, am.mix_reference, am.mix_anti_reference))
    for key in want:
        assert got[key] == pytest.approx(want[key], rel=1e-9, abs=1e-12), key


def test_mono_input_with_apa_raises(cuda_device):
    am = _am()
    with pytest.raises(ValueError):
        am.add_reference(np.random.random((10, 5 * 16000)))


def test_errors(cuda_device):
    with pytest.raises(ValueError):
        AudioMetrics(embedder="no-such-embedder", mix_function=mix_func)
    with pytest.raises(ValueError):
        AudioMetrics(embedder=DummyEmbedder(), mix_function="no-such-mix")
    am = _am()
    with pytest.raises(ValueError):
        am.evaluate(np.random.random((4, 5 * 16000, 2)))        # empty reference
    am.add_reference(np.random.random((4, 1000, 2)))            # shorter than win_dur -> still empty
    with pytest.raises(ValueError):
        am.evaluate(np.random.random((4, 5 * 16000, 2)))


def test_all_metrics_and_second_add_reference(cuda_device):
    rng = np.random.default_rng(1)
    n = 5 * 16000
    am = AudioMetrics(embedder=RandomEmbedder(), mix_function=mix_func, metrics=["fad", "kd", "prdc", "apa"])
    am.add_reference(rng.standard_normal((120, n, 2)))
    keys = ["fad", "kernel_distance_mean", "kernel_distance_std", "precision", "recall", "density", "coverage", "apa"]
    r1 = am.evaluate(rng.standard_normal((100, n, 2)))
    _check(r1, keys)
    assert 0 <= r1["precision"] <= 1 and 0 <= r1["recall"] <= 1 and 0 <= r1["coverage"] <= 1 and 0 <= r1["apa"] <= 1
    # the reference keeps stale radii after a second add_reference and then fails; we invalidate
    am.add_reference(rng.standard_normal((30, n, 2)))
    _check(am.evaluate(rng.standard_normal((100, n, 2))), keys)
    assert len(am.stem_reference) == 150 and am.stem_reference.embeddings.shape[0] == 150


def test_serialization_round_trip(cuda_device, tmp_path):
    """reference tests/test_audio_metrics.py:175-197."""
    rng = np.random.default_rng(2)
    n = 5 * 16000
    ref, cand = rng.random((50, n, 2)), rng.random((50, n, 2))
    am = _am()
    am.add_reference(ref)
    r1 = am.evaluate(cand)
    fp = tmp_path / "state.pt"
    am.save_state(fp)
    am2 = _am()
    am2.load_state(fp)
    r2 = am2.evaluate(cand)
    for k in r1:
        assert r2[k] == pytest.approx(r1[k], rel=1e-6, abs=1e-6)
    state = torch.load(fp, weights_only=True)
    assert set(state["stem_reference"]) == {"mean", "n", "cov", "store_embeddings", "embeddings", "radii", "dtype"}
    assert state["stem_reference"]["mean"].device.type == "cpu"
    back = AudioMetricsData.deserialize(state["stem_reference"])
    assert back.n == am.stem_reference.n and torch.allclose(back.cov.cpu(), am.stem_reference.cov.cpu())


# ------------------------------------------------- goldens written by the unmodified reference facade
def _golden_state():
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).parent / "golden" / "golden_state.json").read_text())
    return g, Path(__file__).parent / "golden"


def _golden_inputs(seed, n):
    """tests/golden/make_golden_state.py::inputs"""
    win = 5 * 16000
    rng = np.random.default_rng(seed)
    env = rng.random((n, 64, 1, 2)) * rng.random((n, 1, 1, 2)) * 2
    x = rng.standard_normal((n, 64, win // 64, 2)) * env
    return x.reshape(n, win, 2)


class _DummyF64(DummyEmbedder):
    """The reference DummyEmbedder exactly (float64 embeddings from numpy's std)."""


class _Segment(DummyEmbedder):
    def __init__(self):
        super().__init__(24)
        self.W = torch.randn(64, 24, generator=torch.Generator().manual_seed(0)).cuda()

    @torch.no_grad()
    def forward(self, data, sr=None):
        a = torch.as_tensor(data["audio"][:, :6400].reshape(len(data["audio"]), 64, 100), dtype=torch.float32)
        return {"embedding": (a.std(dim=2)).cuda() @ self.W}


@pytest.mark.parametrize("name", ["dummy", "segment"])
def test_facade_matches_reference_facade(cuda_device, name, tmp_path):
    """AudioMetrics(metrics=["fad", "apa"], n_pca=10) end to end — embedding pipeline, device PCA,
    fused FAD + APA — against what the unmodified reference returned for the same audio; then the
    state file the REFERENCE wrote is loaded here and evaluated, and our own state file round-trips."""
    g, gdir = _golden_state()
    want = g["cases"][name]
    emb = _DummyF64 if name == "dummy" else _Segment
    ref, cand = _golden_inputs(g["seed_ref"], g["n_ref"]), _golden_inputs(g["seed_cand"], g["n_cand"])
    # the aligned / misaligned pairing of the reference set is a seeded shuffle in both packages; the
    # stem FAD does not depend on it, the APA score does — compare FAD tightly, APA only without PCA
    # when the shuffles coincide (they are the same algorithm with the same default seed=None -> random)
    am = AudioMetrics(embedder=emb(), mix_function=mix_func, metrics=["fad", "apa"], n_pca=10)
    am.add_reference(ref)
    got = am.evaluate(cand)
    assert got["fad"] == pytest.approx(want["evaluate"]["fad"], rel=1e-6)
    assert 0.0 <= got["apa"] <= 1.0
    assert am.stem_reference.n == want["stem_n"]
    # projected reference statistics and the fitted projection itself (sklearn's values)
    n_sv = 4 if name == "segment" else 1     # rank-1 embeddings: every further singular value is round-off
    np.testing.assert_allclose(am.stem_projection.singular_values_.cpu().numpy()[:n_sv],
                               want["stem_singular_values"][:n_sv], rtol=1e-6)
    if name == "segment":   # full rank: components are well defined, sign convention included
        np.testing.assert_allclose(am.stem_projection.components_[0].cpu().numpy(), want["stem_components_row0"],
                                   rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(am.stem_reference_pca.mean.cpu().numpy(), want["stem_pca_mean"], rtol=1e-6,
                                   atol=1e-6 * max(abs(v) for v in want["stem_pca_cov_diag"]) ** 0.5)
        np.testing.assert_allclose(am.stem_reference_pca.cov.diagonal().cpu().numpy(), want["stem_pca_cov_diag"],
                                   rtol=1e-6)
        # (row order differs from the reference run: the reference set goes through a randomly seeded
        #  shuffle in both packages, embed.py:152-160 — so check the transform against its definition)
        x = am.stem_reference.embeddings
        t = am.stem_projection.transform(x)
        assert t.dtype == torch.float64 and want["stem_transform_dtype"] == "torch.float64"
        by_def = (x.double() - am.stem_projection.mean_) @ am.stem_projection.components_.T
        torch.testing.assert_close(t, by_def, rtol=1e-12, atol=1e-12)
    # without PCA
    am3 = AudioMetrics(embedder=emb(), mix_function=mix_func, metrics=["fad", "apa"])
    am3.add_reference(ref)
    assert am3.evaluate(cand)["fad"] == pytest.approx(want["evaluate_no_pca"]["fad"], rel=1e-6)
    # a state file written by the reference package loads here and evaluates to the reference's numbers
    am2 = AudioMetrics(embedder=emb(), mix_function=mix_func, metrics=["fad", "apa"], n_pca=10)
    am2.load_state(gdir / f"reference_state_{name}.pt")
    got2 = am2.evaluate(cand)
    assert got2["fad"] == pytest.approx(want["evaluate_after_load"]["fad"], rel=1e-6)
    assert got2["apa"] == pytest.approx(want["evaluate_after_load"]["apa"], rel=1e-5, abs=1e-7)   # same stored mix statistics
    # and ours round-trips with the reference's schema
    fp = tmp_path / "state.pt"
    am2.save_state(fp)
    state = torch.load(fp, weights_only=True)
    ref_state = torch.load(gdir / f"reference_state_{name}.pt", weights_only=True)
    assert set(state) >= set(ref_state) - {"gpu_handler"}
    assert set(state["stem_projection"]) == set(ref_state["stem_projection"])
    am4 = AudioMetrics(embedder=emb(), mix_function=mix_func, metrics=["fad", "apa"], n_pca=10)
    am4.load_state(fp)
    got4 = am4.evaluate(cand)
    assert got4["fad"] == pytest.approx(got2["fad"], rel=1e-9) and got4["apa"] == pytest.approx(got2["apa"], rel=1e-9, abs=1e-12)


def test_incremental_pca_second_fit_matches_sklearn(cuda_device):
    """partial_fit twice (what a second add_reference triggers, audio_metrics.py:163-182) against sklearn."""
    sk = pytest.importorskip("sklearn.decomposition")
    from audio_metrics_b200.projection import IncrementalPCA
    rng = np.random.default_rng(3)
    a = (rng.standard_normal((400, 20)) * np.geomspace(5, 0.1, 20) + 1.0).astype(np.float32)
    b = (rng.standard_normal((300, 20)) * np.geomspace(4, 0.2, 20) - 0.5).astype(np.float32)
    ours, theirs = IncrementalPCA(n_components=6), sk.IncrementalPCA(n_components=6)
    for blk in (a, b):
        ours.partial_fit(torch.from_numpy(blk))
        theirs.partial_fit(blk.astype(np.float64))   # sklearn in fp64: with fp32 input its SVD runs in fp32 (1e-5 noise)
        np.testing.assert_allclose(ours.singular_values_.cpu().numpy(), theirs.singular_values_, rtol=1e-6)
        np.testing.assert_allclose(ours.components_.cpu().numpy(), theirs.components_, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ours.mean_.cpu().numpy(), theirs.mean_, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(ours.var_.cpu().numpy(), theirs.var_, rtol=1e-6)
        np.testing.assert_allclose(ours.explained_variance_ratio_.cpu().numpy(), theirs.explained_variance_ratio_, rtol=5e-6)   # sklearn sums fp32 variances
        assert ours.noise_variance_ == pytest.approx(theirs.noise_variance_, rel=1e-6)
        assert ours.n_samples_seen_ == theirs.n_samples_seen_
    np.testing.assert_allclose(ours.transform(torch.from_numpy(b)).cpu().numpy(), theirs.transform(b.astype(np.float64)), rtol=1e-7, atol=1e-7)

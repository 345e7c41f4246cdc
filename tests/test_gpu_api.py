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

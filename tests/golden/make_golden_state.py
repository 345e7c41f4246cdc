"""State files and end-to-end results written by the UNMODIFIED reference facade.

Runs ``AudioMetrics(metrics=["fad", "apa"], n_pca=10)`` of /root/reference on CPU (a fake GPU
handler that hands the CPU embedder back) for two embedders —
the reference tests' rank-1 DummyEmbedder (tests/test_audio_metrics.py:7-24) and a full-rank
one for which the PCA components are well defined — and commits

    reference_state_<name>.pt     what reference ``save_state`` wrote (torch.save, weights_only-loadable)
    golden_state.json             what reference ``evaluate`` returned, before and after a
                                  save_state / load_state round trip, plus the PCA-projected
                                  reference statistics

The audio inputs regenerate from their seeds (numpy PCG64).  Build container only:
``python tests/golden/make_golden_state.py``.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
for name in ("soxr", "pyloudnorm", "numpy_audio_limiter", "opt_einsum", "appdirs"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["pyloudnorm"].Meter = type("Meter", (), {"__init__": lambda self, sr: None})
sys.path.insert(0, "/root/reference/src")

from audio_metrics import AudioMetrics  # noqa: E402  (the reference)
from audio_metrics.util.gpu_parallel import GPUWorkerHandler  # noqa: E402

SR, WIN = 16000, 5 * 16000
N_REF, N_CAND, SEED_REF, SEED_CAND = 48, 40, 501, 502


class FakeHandler(GPUWorkerHandler):
    """One 'GPU' token; the model is used where it is (CPU)."""

    def __init__(self, device_indices=None, thread_pool=None):
        super().__init__((0,), thread_pool)

    def get_model_on_gpu(self, model, target_gpu_i):
        return model


AudioMetrics._get_gpu_handler = lambda self, device_indices: FakeHandler()


class DummyEmbedder:   # reference tests/test_audio_metrics.py:7-24
    def __init__(self):
        self.m = torch.nn.Linear(1, 1)

    @property
    def sr(self):
        return SR

    def get_device(self):
        return next(self.m.parameters()).device

    @torch.no_grad()
    def forward(self, data, sr=None):
        mean = torch.as_tensor(10**3 * data["audio"].std(axis=1))
        return {"embedding": torch.outer(mean, torch.arange(10))}


class SegmentEmbedder(DummyEmbedder):
    """Full-rank float32 embeddings (d = 24): per-segment standard deviations through a fixed
    random map (the same construction as tests/test_gpu_api.py::RandomEmbedder)."""

    def __init__(self):
        super().__init__()
        self.W = torch.randn(64, 24, generator=torch.Generator().manual_seed(0))

    @torch.no_grad()
    def forward(self, data, sr=None):
        a = torch.as_tensor(data["audio"][:, :6400].reshape(len(data["audio"]), 64, 100), dtype=torch.float32)
        return {"embedding": a.std(dim=2) @ self.W}


def mix_func(audio, sr=None):
    return audio.mean(axis=1)


def inputs(seed, n):
    """[n, WIN, 2] context/stem pairs whose loudness varies per item and per segment."""
    rng = np.random.default_rng(seed)
    env = rng.random((n, 64, 1, 2)) * rng.random((n, 1, 1, 2)) * 2
    x = rng.standard_normal((n, 64, WIN // 64, 2)) * env
    return x.reshape(n, WIN, 2)


def main():
    out = {"n_ref": N_REF, "n_cand": N_CAND, "seed_ref": SEED_REF, "seed_cand": SEED_CAND, "cases": {}}
    ref, cand = inputs(SEED_REF, N_REF), inputs(SEED_CAND, N_CAND)
    for name, emb in (("dummy", DummyEmbedder), ("segment", SegmentEmbedder)):
        am = AudioMetrics(embedder=emb(), mix_function=mix_func, metrics=["fad", "apa"], n_pca=10)
        am.add_reference(ref)
        r1 = am.evaluate(cand)
        fp = HERE / f"reference_state_{name}.pt"
        am.save_state(fp)
        am2 = AudioMetrics(embedder=emb(), mix_function=mix_func, metrics=["fad", "apa"], n_pca=10)
        am2.load_state(fp)
        r2 = am2.evaluate(cand)
        e = {"evaluate": r1, "evaluate_after_load": r2,
             "stem_pca_mean": am.stem_reference_pca.mean.tolist(),
             "stem_pca_cov_diag": am.stem_reference_pca.cov.diagonal().tolist(),
             "stem_singular_values": am.stem_projection.singular_values_.tolist(),
             "stem_components_row0": am.stem_projection.components_[0].tolist(),
             "stem_n": int(am.stem_reference.n),
             "stem_transform_dtype": str(am.stem_projection.transform(am.stem_reference.embeddings).dtype),
             "stem_transform_first_row": am.stem_projection.transform(am.stem_reference.embeddings)[0].tolist()}
        # no PCA: plain fad + apa on the same audio (the facade path without projection)
        am3 = AudioMetrics(embedder=emb(), mix_function=mix_func, metrics=["fad", "apa"])
        am3.add_reference(ref)
        e["evaluate_no_pca"] = am3.evaluate(cand)
        out["cases"][name] = e
        print(name, r1, r2, e["evaluate_no_pca"])
    (HERE / "golden_state.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""Generate golden vectors by running the UNMODIFIED reference hot-path modules.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  Writes ``golden.json`` (scalars,
input digests) and ``golden_arrays.npz`` (radii, per-subset MMDs, count vectors,
and the inputs of the tiny cases) next to this file.

The reference package imports five audio-only dependencies at module import time
that are absent here (soxr, pyloudnorm, numpy_audio_limiter, opt_einsum,
appdirs); they are off the embedding-set distance path, so empty stub modules
are registered before importing it.
"""
import hashlib
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

from audio_metrics.data import AudioMetricsData  # noqa: E402  (the reference)
from audio_metrics.metrics.fad import frechet_distance  # noqa: E402
from audio_metrics.metrics.kd import kernel_distance, kid_features_to_metric  # noqa: E402
from audio_metrics.metrics.prdc import prdc, nearest_neighbour_distances  # noqa: E402
from audio_metrics.metrics.apa import apa  # noqa: E402

from audio_metrics_b200.synth import make_sets_numpy, make_apa_sets_numpy  # noqa: E402

torch.set_num_threads(8)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_amd(x, store=True):
    a = AudioMetricsData(store)
    a.add(torch.from_numpy(x))
    return a


def counts(ref, cand, k):
    """Integer numerators, recomputed with the reference's own expressions (prdc.py:34-48)."""
    r_ref = nearest_neighbour_distances(torch.from_numpy(ref), k)
    r_cand = nearest_neighbour_distances(torch.from_numpy(cand), k)
    D = torch.cdist(torch.from_numpy(ref), torch.from_numpy(cand))
    col = (D < r_ref[:, None]).sum(dim=0)
    rec = (D < r_cand[None, :]).any(dim=1)
    cov = D.min(dim=1)[0] < r_ref
    return r_ref.numpy(), r_cand.numpy(), col.numpy(), rec.numpy(), cov.numpy()


def main():
    scal, arrs = {}, {}
    # (name, n_ref, n_cand, d, seed, dtype, k, store inputs?)
    cases = [
        ("tiny_d64", 256, 256, 64, 11, np.float32, 5, True),
        ("c1_n100_d128", 100, 100, 128, 12, np.float32, 10, True),
        ("mid_d512", 1000, 1200, 512, 13, np.float32, 5, False),
        ("mid_d128_k10", 4000, 4000, 128, 14, np.float32, 10, False),
        ("pca_f64_d10", 500, 400, 10, 15, np.float64, 5, True),
        ("c2_10k_d512", 10000, 10000, 512, 1234, np.float32, 5, False),
    ]
    for name, n, m, d, seed, dtype, k, store in cases:
        print("case", name, flush=True)
        ref, cand = make_sets_numpy(n, m, d, seed=seed, dtype=dtype)
        R, Cn = ref_amd(ref), ref_amd(cand)
        e = {"n_ref": n, "n_cand": m, "d": d, "seed": seed, "dtype": np.dtype(dtype).name, "k": k,
             "ref_sha256": digest(ref), "cand_sha256": digest(cand)}
        e["fad"] = frechet_distance(Cn, R)                       # audio_metrics.py:257 order (cand, ref)
        kd = kid_features_to_metric(Cn.embeddings, R.embeddings)  # audio_metrics.py:260
        e["kernel_distance_mean"], e["kernel_distance_std"] = kd["kernel_distance_mean"], kd["kernel_distance_std"]
        e.update({f"prdc_{kk}": v for kk, v in prdc(R, Cn, k).items()})   # audio_metrics.py:264 order (ref, cand)
        r_ref, r_cand, col, rec, cov = counts(ref, cand, k)
        e["cov_trace_ref"] = float(R.cov.trace())
        e["mean_norm_ref"] = float(R.mean.norm())
        arrs[f"{name}/r_ref"], arrs[f"{name}/r_cand"] = r_ref, r_cand
        arrs[f"{name}/col_count"] = col.astype(np.int32)
        arrs[f"{name}/recall_rows"], arrs[f"{name}/cover_rows"] = np.packbits(rec), np.packbits(cov)
        arrs[f"{name}/mean_ref"], arrs[f"{name}/cov_diag_ref"] = R.mean.numpy(), R.cov.diagonal().numpy().copy()
        if store:
            arrs[f"{name}/ref"], arrs[f"{name}/cand"] = ref, cand
            arrs[f"{name}/cov_ref"] = R.cov.numpy()
        scal[name] = e

    # streaming statistics: 32-row batches as embedding_pipeline feeds them (embed.py:226-236)
    ref, _ = make_sets_numpy(1000, 8, 64, seed=21)
    S = AudioMetricsData(True)
    for i in range(0, 1000, 32):
        S.add(torch.from_numpy(ref[i:i + 32]))
    arrs["stream/ref"], arrs["stream/mean"], arrs["stream/cov"] = ref, S.mean.numpy(), S.cov.numpy()
    S.recompute_stats()
    arrs["stream/cov_recomputed"] = S.cov.numpy()

    # APA (apa.py:9-32) on the mix/stem latent model + the DummyEmbedder rank-1 case
    s = make_apa_sets_numpy(2000, 128, seed=31)
    cand, refa, anti = (ref_amd(s[k], False) for k in ("cand_aligned", "ref_aligned", "ref_misaligned"))
    scal["apa_d128"] = {"n": 2000, "d": 128, "seed": 31, "apa": apa(cand, refa, anti),
                        "d_y_x": frechet_distance(cand, refa), "d_y_xp": frechet_distance(cand, anti),
                        "d_x_xp": frechet_distance(refa, anti),
                        "fad_stems": frechet_distance(ref_amd(s["cand_stems"], False), ref_amd(s["ref_stems"], False)),
                        "sha256": digest(np.concatenate([s[k] for k in sorted(s)]))}
    rng = np.random.default_rng(41)
    a = np.outer(rng.random(100) * 300, np.arange(10.0))   # tests/test_audio_metrics.py:22-23 embeddings
    b = np.outer(rng.random(100) * 300, np.arange(10.0))
    arrs["rank1/a"], arrs["rank1/b"] = a, b
    scal["rank1"] = {"fad": frechet_distance(ref_amd(a, False), ref_amd(b, False)),
                     "fad_self": frechet_distance(ref_amd(a, False), ref_amd(a, False))}

    (HERE / "golden.json").write_text(json.dumps({"torch": torch.__version__, "numpy": np.__version__,
                                                  "reference": "SonyCSLParis/audio-metrics v1.0.4 (/root/reference)",
                                                  "cases": scal}, indent=1, sort_keys=True))
    np.savez_compressed(HERE / "golden_arrays.npz", **arrs)
    print("wrote", HERE / "golden.json", HERE / "golden_arrays.npz")


if __name__ == "__main__":
    main()

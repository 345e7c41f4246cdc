"""BASELINE.json config 3 at its stated size — APA on 10k synthetic mix / stem CLAP-512 embedding
pairs with FAD on the stems — through the UNMODIFIED reference (apa.py:9-32, fad.py:8-31).
Build container only: ``python tests/golden/make_golden_c3.py`` -> ``golden_c3.json``."""
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
from audio_metrics.metrics.apa import apa, apa_compute_d_x_xp  # noqa: E402
from audio_metrics.metrics.fad import frechet_distance  # noqa: E402

from audio_metrics_b200.synth import make_apa_sets_numpy  # noqa: E402

N, D, SEED = 10000, 512, 33


def amd(x):
    a = AudioMetricsData(False)
    a.add(torch.from_numpy(x))
    return a


if __name__ == "__main__":
    torch.set_num_threads(8)
    s = make_apa_sets_numpy(N, D, seed=SEED)
    cand, ref, anti = (amd(s[k]) for k in ("cand_aligned", "ref_aligned", "ref_misaligned"))
    d_x_xp = apa_compute_d_x_xp(ref, anti)
    out = {"n": N, "d": D, "seed": SEED,
           "sha256": hashlib.sha256(np.ascontiguousarray(np.concatenate([s[k] for k in sorted(s)])).tobytes()).hexdigest(),
           "d_x_xp": d_x_xp, "apa": apa(cand, ref, anti), "apa_cached": apa(cand, ref, anti, d_x_xp),
           "d_y_x": frechet_distance(cand, ref), "d_y_xp": frechet_distance(cand, anti),
           "fad_stems": frechet_distance(amd(s["cand_stems"]), amd(s["ref_stems"]))}
    print(out)
    (HERE / "golden_c3.json").write_text(json.dumps(out, indent=1))

"""Golden vectors for the keyword interface of the kernel distance (kd.py:127-194):
kernel_type / kid_degree / kid_gamma / kid_coef0 / kid_sigma / kid_subsets / kid_subset_size /
rng_seed, produced by the UNMODIFIED reference.  Build container only:
``python tests/golden/make_golden_kd.py`` -> ``golden_kd_variants.json``."""
import json
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
for name in ("soxr", "pyloudnorm", "numpy_audio_limiter", "opt_einsum", "appdirs"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["pyloudnorm"].Meter = type("Meter", (), {"__init__": lambda self, sr: None})
sys.path.insert(0, "/root/reference/src")
from audio_metrics.metrics.kd import kid_features_to_metric  # noqa: E402  (the reference)

from audio_metrics_b200.synth import make_sets_numpy  # noqa: E402

VARIANTS = {
    "rbf_sigma1": dict(kernel_type="rbf", kid_sigma=1.0, kid_subsets=20, kid_subset_size=300),
    "rbf_default_sigma": dict(kernel_type="rbf", kid_subsets=10, kid_subset_size=256),
    "poly_deg2": dict(kid_degree=2, kid_gamma=0.01, kid_coef0=0.5, kid_subsets=16, kid_subset_size=200, rng_seed=7),
    "poly_small_subsets": dict(kid_subsets=5, kid_subset_size=100000),   # shrinks to min(n)//2 (kd.py:160-168)
}
INPUT = dict(n_ref=1300, n_cand=1100, d=96, seed=31)

if __name__ == "__main__":
    ref, cand = make_sets_numpy(INPUT["n_ref"], INPUT["n_cand"], INPUT["d"], seed=INPUT["seed"])
    out = {"input": INPUT, "variants": {}}
    for name, kw in VARIANTS.items():
        r32 = kid_features_to_metric(cand, ref, **kw)                                        # reference arithmetic
        r64 = kid_features_to_metric(cand.astype(np.float64), ref.astype(np.float64), **kw)  # same code, fp64 inputs
        out["variants"][name] = {"kwargs": kw, "reference_f32": r32, "reference_f64": r64}
        print(name, r32, r64)
    (HERE / "golden_kd_variants.json").write_text(json.dumps(out, indent=1))

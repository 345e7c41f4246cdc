"""Golden vectors for the MMD^2 estimators of kd.py:38-83 that the reference only reaches through
``mmd2`` itself ("biased", "u-statistic", ``unit_diagonal``), produced by the UNMODIFIED reference:
its own ``mmd2`` / ``polynomial_kernel`` / ``rbf_kernel`` on the subsets its own loop draws
(kd.py:176-187).  Build container only:
``python tests/golden/make_golden_kd_estimators.py`` -> ``golden_kd_estimators.json``."""
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
from audio_metrics.metrics.kd import mmd2, polynomial_kernel, rbf_kernel  # noqa: E402  (the reference)

from audio_metrics_b200.synth import make_sets_numpy  # noqa: E402

INPUT = dict(n_ref=1300, n_cand=1100, d=96, seed=31)
SUBSETS, SUBSET_SIZE, SEED = 12, 300, 1234
VARIANTS = {
    "poly_biased": dict(kernel_type="polynomial", mmd_est="biased"),
    "poly_ustat": dict(kernel_type="polynomial", mmd_est="u-statistic"),
    "poly_unbiased": dict(kernel_type="polynomial", mmd_est="unbiased"),
    "rbf_biased_unit_diag": dict(kernel_type="rbf", kid_sigma=1.0, mmd_est="biased", unit_diagonal=True),
    "rbf_ustat": dict(kernel_type="rbf", kid_sigma=1.0, mmd_est="u-statistic"),
    "poly_unbiased_unit_diag": dict(kernel_type="polynomial", mmd_est="unbiased", unit_diagonal=True),
}


def run(f1, f2, kernel_type="polynomial", kid_sigma=10.0, mmd_est="unbiased", unit_diagonal=False):
    """kid_features_to_metric's loop (kd.py:170-192) with mmd2's estimator arguments exposed."""
    kernel = polynomial_kernel if kernel_type == "polynomial" else (lambda a, b: rbf_kernel(a, b, sigma=kid_sigma))
    rng = np.random.default_rng(SEED)                                   # kd.py:176
    mmds = np.zeros(SUBSETS)
    for i in range(SUBSETS):
        a = f1[rng.choice(len(f1), SUBSET_SIZE, replace=False)]         # kd.py:185
        b = f2[rng.choice(len(f2), SUBSET_SIZE, replace=False)]         # kd.py:186
        mmds[i] = mmd2(kernel(a, a), kernel(a, b), kernel(b, b), unit_diagonal=unit_diagonal, mmd_est=mmd_est)
    return {"kernel_distance_mean": float(np.mean(mmds)), "kernel_distance_std": float(np.std(mmds)),
            "mmds": [float(v) for v in mmds]}


if __name__ == "__main__":
    ref, cand = make_sets_numpy(INPUT["n_ref"], INPUT["n_cand"], INPUT["d"], seed=INPUT["seed"])
    out = {"input": INPUT, "subsets": SUBSETS, "subset_size": SUBSET_SIZE, "seed": SEED, "variants": {}}
    for name, kw in VARIANTS.items():
        r32 = run(cand, ref, **kw)
        r64 = run(cand.astype(np.float64), ref.astype(np.float64), **kw)
        out["variants"][name] = {"kwargs": kw, "reference_f32": r32, "reference_f64": r64}
        print(name, r32["kernel_distance_mean"], r64["kernel_distance_mean"])
    (HERE / "golden_kd_estimators.json").write_text(json.dumps(out, indent=1))

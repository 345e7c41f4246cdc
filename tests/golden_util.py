"""Shared access to the committed golden fixtures (tests/golden/)."""
import hashlib
import json
from functools import lru_cache
from pathlib import Path

import numpy as np

from audio_metrics_b200.synth import make_sets_numpy

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
SET_CASES = ["tiny_d64", "c1_n100_d128", "mid_d512", "mid_d128_k10", "pca_f64_d10", "c2_10k_d512"]


@lru_cache(maxsize=None)
def scalars():
    return json.loads((GOLDEN_DIR / "golden.json").read_text())["cases"]


@lru_cache(maxsize=None)
def arrays():
    with np.load(GOLDEN_DIR / "golden_arrays.npz") as z:
        return {k: z[k] for k in z.files}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@lru_cache(maxsize=4)
def case_inputs(name):
    """(ref, cand, meta) for a golden case: stored arrays for the tiny cases, otherwise
    regenerated from the seed and verified bit-for-bit against the recorded digest."""
    g = scalars()[name]
    a = arrays()
    if f"{name}/ref" in a:
        ref, cand = a[f"{name}/ref"], a[f"{name}/cand"]
    else:
        ref, cand = make_sets_numpy(g["n_ref"], g["n_cand"], g["d"], seed=g["seed"], dtype=np.dtype(g["dtype"]))
    assert digest(ref) == g["ref_sha256"] and digest(cand) == g["cand_sha256"], \
        f"synthetic inputs of golden case {name} did not regenerate bit-for-bit"
    return ref, cand, g


def unpack_rows(bits, n):
    return np.unpackbits(bits)[:n].astype(bool)

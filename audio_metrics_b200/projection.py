"""Optional PCA projection (reference projection.py:6-46): sklearn's IncrementalPCA
with torch-serialisable state; ``transform`` returns a float64 tensor."""
from __future__ import annotations

import numpy as np
import torch
from sklearn.decomposition import IncrementalPCA as _IncrementalPCA

_ARRAYS = ("components_", "mean_", "var_", "singular_values_", "explained_variance_", "explained_variance_ratio_")


def _to_numpy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class IncrementalPCA(_IncrementalPCA):
    def partial_fit(self, x, y=None, check_input=True):
        return super().partial_fit(_to_numpy(x), y, check_input)

    def transform(self, x):
        return torch.as_tensor(super().transform(_to_numpy(x)))   # float64, projection.py:20-21

    def __getstate__(self):
        state = super().__getstate__().copy()
        for k in _ARRAYS:
            if k in state:
                state[k] = torch.as_tensor(state[k])
        if "n_samples_seen_" in state:
            state["n_samples_seen_"] = int(state["n_samples_seen_"])
        if "noise_variance_" in state:
            state["noise_variance_"] = float(state["noise_variance_"])
        return state

    def __setstate__(self, state):
        state = dict(state)
        for k in _ARRAYS:
            if k in state:
                state[k] = _to_numpy(state[k])
        if "n_samples_seen_" in state:
            state["n_samples_seen_"] = np.int64(state["n_samples_seen_"])
        if "noise_variance_" in state:
            state["noise_variance_"] = np.float64(state["noise_variance_"])
        super().__setstate__(state)

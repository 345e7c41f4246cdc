"""``AudioMetrics`` — the reference's API facade (audio_metrics.py:15-313) on top of
the CUDA distance path: same constructor, ``add_reference`` / ``evaluate`` /
``__call__`` / ``reset_reference`` / ``save_state`` / ``load_state``, same result
keys and the same saved-state schema."""
from __future__ import annotations

from pathlib import Path

import torch

from . import _lib
from .data import AudioMetricsData
from .embedders import DEFAULT_EMBEDDER, EMBEDDERS
from .dist import evaluate_containers, evaluate_devices
from .metrics.apa import _apa
from .mix import DEFAULT_MIX_FUNCTION, MIX_FUNCTIONS
from .pipeline import ItemCategory, embedding_pipeline
from .projection import IncrementalPCA


class AudioMetrics:
    _need_embeddings = set(("kd", "precision", "prdc"))          # audio_metrics.py:17
    _amd = ("stem_reference", "mix_reference", "mix_anti_reference",
            "stem_reference_pca", "mix_reference_pca", "mix_anti_reference_pca")

    def __init__(self, metrics=["apa", "fad"], n_pca=None, device_indices=None, embedder=None, mix_function=None,
                 win_dur=5.0, input_sr=None):
        self.device = self._resolve_device(device_indices)
        # device_indices with several entries (the reference's thread-per-GPU handler,
        # util/gpu_parallel.py:20-76): the all-pairs sweeps of evaluate() are sharded over them
        self.devices = [torch.device("cuda", int(i)) for i in device_indices] if device_indices else [self.device]
        self.metrics = metrics
        self.need_apa = "apa" in self.metrics
        self.win_dur = win_dur
        self.input_sr = input_sr
        self.stem_projection = None if n_pca is None else IncrementalPCA(n_components=n_pca, device=self.device)
        self.mix_projection = None if n_pca is None else IncrementalPCA(n_components=n_pca, device=self.device)
        self.embedder = self.get_embedder(embedder) if embedder is None or isinstance(embedder, str) else embedder
        self.mix_function = (self.get_mix_function(mix_function)
                             if mix_function is None or isinstance(mix_function, str) else mix_function)
        self.apa_d_x_xp = None
        self.reset_reference()
        self.mix_reference_pca = self.mix_anti_reference_pca = self.stem_reference_pca = None

    @staticmethod
    def _resolve_device(device_indices):
        """The reference raises RuntimeError without GPUs (util/gpu_parallel.py:27-28); so do we."""
        if not torch.cuda.is_available() or torch.cuda.device_count() == 0:
            raise RuntimeError("No GPUs found: audio_metrics_b200 computes on CUDA devices only")
        if device_indices:
            return torch.device("cuda", int(device_indices[0]))
        return torch.device("cuda", torch.cuda.current_device())

    # ---------------------------------------------------------------- properties
    @property
    def stems_mode(self):
        return any(m for m in self.metrics if m != "apa")                       # audio_metrics.py:106-108

    @property
    def store_mix_embeddings(self):
        return self.need_apa and self.mix_projection is not None               # :110-112

    @property
    def store_stem_embeddings(self):
        return self.stem_projection is not None or any(m in self._need_embeddings for m in self.metrics)

    # --------------------------------------------------------------------- state
    def save_state(self, fp: str | Path):
        """audio_metrics.py:78-91 — plain tensors / scalars so weights_only loading works."""
        state = dict(self.__dict__)
        for k in ("mix_function", "embedder", "device", "devices"):
            state.pop(k, None)
        for attr in self._amd:
            if state.get(attr):
                state[attr] = state[attr].serialize()
        for attr in ("stem_projection", "mix_projection"):
            if state.get(attr):
                state[attr] = state[attr].__getstate__().copy()
        torch.save(state, fp)

    def load_state(self, fp: str | Path):
        """audio_metrics.py:93-104; accepts state files written by the reference package."""
        state = torch.load(fp, weights_only=True)
        for attr in self._amd:
            if state.get(attr):
                state[attr] = AudioMetricsData.deserialize(state[attr], device=self.device)
        for attr in ("stem_projection", "mix_projection"):
            item = state.pop(attr, None)
            if item and getattr(self, attr) is not None:
                getattr(self, attr).__setstate__(item)
        state.pop("gpu_handler", None)
        self.__dict__.update(state)

    def reset_reference(self):
        """audio_metrics.py:151-161."""
        self.apa_d_x_xp = None
        self.mix_reference = AudioMetricsData(self.store_mix_embeddings, self.device) if self.need_apa else None
        self.mix_anti_reference = AudioMetricsData(self.store_mix_embeddings, self.device) if self.need_apa else None
        self.mix_reference_pca = self.mix_anti_reference_pca = None
        self.stem_reference = AudioMetricsData(self.store_stem_embeddings, self.device) if self.stems_mode else None
        self.stem_reference_pca = None

    # ------------------------------------------------------------------ pipeline
    def _embed(self, audio, apa_mode):
        return embedding_pipeline(audio, embedder=self.embedder, mix_function=self.mix_function, apa_mode=apa_mode,
                                  stems_mode=self.stems_mode, store_mix_embeddings=self.store_mix_embeddings,
                                  store_stem_embeddings=self.store_stem_embeddings, win_dur=self.win_dur,
                                  input_sr=self.input_sr, device=self.device)

    def add_reference(self, reference):
        """audio_metrics.py:120-149."""
        data = self._embed(reference, "reference" if self.need_apa else None)
        stem = data.get(ItemCategory.stem)
        if stem is not None:
            self.stem_reference_pca = None
            self.stem_reference += stem
            self.stem_reference.recompute_stats()
        mix = data.get(ItemCategory.aligned)
        if mix is not None:
            self.mix_reference_pca = self.mix_anti_reference_pca = None
            self.apa_d_x_xp = None
            self.mix_reference += mix
        anti = data.get(ItemCategory.misaligned)
        if anti is not None:
            self.mix_anti_reference += anti

    def _project(self, projection, emb, store):
        out = AudioMetricsData(store, self.device)
        out.add(projection.transform(emb))
        return out

    def ensure_stem_projection(self, ref, cand):
        """audio_metrics.py:163-182."""
        if self.stem_projection is None:
            return ref, cand
        store = any(m in self._need_embeddings for m in self.metrics)
        if self.stem_reference_pca is None:
            self.stem_projection.partial_fit(ref)          # the container's own statistics: no second pass
            self.stem_reference_pca = self._project(self.stem_projection, ref.embeddings, store)
        return self.stem_reference_pca, self._project(self.stem_projection, cand.embeddings, store)

    def ensure_mix_projection(self, ref, anti_ref, cand):
        """audio_metrics.py:184-209."""
        if self.mix_projection is None:
            return ref, anti_ref, cand
        if self.mix_reference_pca is None:
            self.mix_projection.partial_fit(ref)
            self.mix_reference_pca = self._project(self.mix_projection, ref.embeddings, False)
            self.mix_anti_reference_pca = self._project(self.mix_projection, anti_ref.embeddings, False)
        return (self.mix_reference_pca, self.mix_anti_reference_pca,
                self._project(self.mix_projection, cand.embeddings, False))

    def __call__(self, candidate):
        return self.evaluate(candidate)

    def evaluate(self, candidate):
        """audio_metrics.py:214-274."""
        self.assert_reference()
        data = self._embed(candidate, "candidate" if self.need_apa else None)
        stem_cand, apa_cand = data.get(ItemCategory.stem), data.get(ItemCategory.aligned)
        stem_ref, apa_ref, apa_anti = self.stem_reference, self.mix_reference, self.mix_anti_reference
        if self.stems_mode and (stem_cand is None or stem_cand.n is None):
            raise ValueError("No stem candidate embeddings were computed")
        if self.need_apa and (apa_cand is None or apa_cand.n is None):
            raise ValueError("No apa candidate embeddings were computed")
        if self.stems_mode:
            stem_ref, stem_cand = self.ensure_stem_projection(stem_ref, stem_cand)
        apa_sets = None
        if self.need_apa:
            apa_ref, apa_anti, apa_cand = self.ensure_mix_projection(apa_ref, apa_anti, apa_cand)
            apa_sets = (apa_cand, apa_ref, apa_anti, self.apa_d_x_xp)       # :251-252: d_x_xp cached per reference
        # One fused, asynchronous schedule with a single read-back (dist.py) instead of the
        # reference's one call and one host synchronisation per metric (:254-272); under
        # torch.distributed the containers hold this rank's rows and the step is row-sharded.
        fused = tuple(m for m in ("fad", "kd", "prdc") if m in self.metrics) if self.stems_mode else ()
        if not fused and apa_sets is None:
            return {}
        if len(self.devices) > 1 and not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            res = evaluate_devices(stem_ref if fused else None, stem_cand if fused else None, self.devices, fused,
                                   nearest_k=None, apa=apa_sets)
        else:
            res = evaluate_containers(stem_ref if fused else None, stem_cand if fused else None, fused,
                                      nearest_k=None, apa=apa_sets)   # k = max(1, min(10, n_ref, n_cand)), :263
        result = {}
        for key in ("fad", "kernel_distance_mean", "kernel_distance_std", "precision", "recall", "density",
                    "coverage"):
            if key in res:
                result[key] = res[key]
        if self.need_apa:
            if self.apa_d_x_xp is None:
                self.apa_d_x_xp = res["_d_x_xp"]
            result["apa"] = _apa(res["_d_y_x"], res["_d_y_xp"], self.apa_d_x_xp)      # apa.py:22-32
        return result

    # ----------------------------------------------------------------- registries
    def get_mix_function(self, mix_function):
        name = DEFAULT_MIX_FUNCTION if mix_function is None else mix_function
        func = MIX_FUNCTIONS.get(name)
        if func is None:
            raise ValueError(f"Unknown mix_function {name}, must be one of {MIX_FUNCTIONS.keys()}")
        return func

    def get_embedder(self, embedder):
        name = DEFAULT_EMBEDDER if embedder is None else embedder
        info = EMBEDDERS.get(name)
        if info is None:
            raise ValueError(f"Unknown embedder {name}, must be one of {EMBEDDERS.keys()}")
        cls, kwargs = info
        return cls(**kwargs, device=self.device)

    def assert_reference(self):
        msg = ("The reference dataset is empty. This can have various causes:"
               "  - You have not called AudioMetrics.add_reference()"
               "  - You have called AudioMetrics.add_reference() with an empty dataset"
               f"  - The duration of your audio is shorter than `win_dur` ({self.win_dur}s)."
               "    (You can specify your own `win_dur` when instantiating AudioMetrics)")
        if self.stems_mode and self.stem_reference.n is None:
            raise ValueError(msg)
        if self.need_apa and self.mix_reference.n is None:
            raise ValueError(msg)

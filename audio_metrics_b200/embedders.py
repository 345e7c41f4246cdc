"""Embedder registry (reference embedders/__init__.py:9-56).  The embedding models
stay ordinary PyTorch modules — they are the *source* of the hot path's input,
not part of it.  Protocol (util/gpu_parallel.py:59-60, tests/test_audio_metrics.py:7-24):
``.sr``, ``.get_device()``, ``.forward({"audio": ndarray[b, n]}) -> {"embedding": Tensor[b, d]}``.
"""
from __future__ import annotations

import numpy as np
import torch

LAION_CLAP_LAYERS = ("audio_projection.0", "audio_projection.2")   # clap.py:7
LAION_CLAP_MUSIC_CHECKPOINT_URL = "https://huggingface.co/lukewys/laion_clap/resolve/main/music_audioset_epoch_15_esc_90.14.pt"
LAION_CLAP_MUSIC_SPEECH_CHECKPOINT_URL = "https://huggingface.co/lukewys/laion_clap/resolve/main/music_speech_audioset_epoch_15_esc_89.98.pt"


class LaionCLAP:
    """LAION-CLAP (HTSAT-base) audio tower; optional forward-hook tap on one of the
    audio projection layers (clap.py:10-60)."""

    def __init__(self, ckpt, layer=None, device=None):
        try:
            import laion_clap
        except ImportError as e:
            raise ImportError("the CLAP embedders need the `laion_clap` package (and its checkpoint file)") from e
        self.model = laion_clap.CLAP_Module(enable_fusion=False, amodel="HTSAT-base")
        self.model.load_ckpt(ckpt)
        self.model.eval()
        if device is not None:
            self.model.to(device)
        self.layer = layer
        self._tap = None
        if layer is not None:
            module = dict(self.model.model.named_modules())[layer]
            module.register_forward_hook(lambda m, i, o: setattr(self, "_tap", o))

    @property
    def sr(self):
        return 48000

    def get_device(self):
        return next(self.model.parameters()).device

    @torch.no_grad()
    def forward(self, data, sr=None):
        audio = torch.as_tensor(np.asarray(data["audio"]), dtype=torch.float32, device=self.get_device())
        out = self.model.get_audio_embedding_from_data(audio, use_tensor=True)
        return {"embedding": self._tap if self.layer is not None else out}


class VGGish:
    """VGGish (128-d) through torch.hub (vggish.py:5-33)."""

    def __init__(self, device=None):
        self.model = torch.hub.load("harritaylor/torchvggish", "vggish")
        self.model.postprocess = False
        self.model.eval()
        if device is not None:
            self.model.to(device)

    @property
    def sr(self):
        return 16000

    def get_device(self):
        return next(self.model.parameters()).device

    @torch.no_grad()
    def forward(self, data, sr=None):
        embs = [self.model.forward(np.asarray(a), self.sr).mean(0) for a in data["audio"]]
        return {"embedding": torch.stack(embs)}


EMBEDDERS = {
    "laion_clap_music": (LaionCLAP, {"ckpt": LAION_CLAP_MUSIC_CHECKPOINT_URL}),
    "laion_clap_music_l-2": (LaionCLAP, {"ckpt": LAION_CLAP_MUSIC_CHECKPOINT_URL, "layer": LAION_CLAP_LAYERS[0]}),
    "laion_clap_music_l-1": (LaionCLAP, {"ckpt": LAION_CLAP_MUSIC_CHECKPOINT_URL, "layer": LAION_CLAP_LAYERS[1]}),
    "laion_clap_music_speech": (LaionCLAP, {"ckpt": LAION_CLAP_MUSIC_SPEECH_CHECKPOINT_URL}),
    "laion_clap_music_speech_l-2": (LaionCLAP, {"ckpt": LAION_CLAP_MUSIC_SPEECH_CHECKPOINT_URL, "layer": LAION_CLAP_LAYERS[0]}),
    "laion_clap_music_speech_l-1": (LaionCLAP, {"ckpt": LAION_CLAP_MUSIC_SPEECH_CHECKPOINT_URL, "layer": LAION_CLAP_LAYERS[1]}),
    "vggish": (VGGish, {}),
}
DEFAULT_EMBEDDER = "laion_clap_music"

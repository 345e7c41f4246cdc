"""audio_metrics_b200 — B200-native embedding-set distance path of audio-metrics.

Same call surface as the reference package ``audio_metrics`` for everything after
embeddings exist (AudioMetricsData, frechet_distance, kernel_distance, prdc, apa)
and the ``AudioMetrics`` facade on top; the arithmetic runs in ``libamb200.so``
(hand-written CUDA for sm_100a behind the C ABI in ``include/amb200.h``).
"""
from .data import AudioMetricsData  # noqa: F401
from .metrics.fad import frechet_distance  # noqa: F401
from .metrics.kd import kernel_distance  # noqa: F401
from .metrics.prdc import prdc, nearest_neighbour_distances  # noqa: F401
from .metrics.apa import apa, apa_compute_d_x_xp  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name == "AudioMetrics":
        from .audio_metrics import AudioMetrics

        return AudioMetrics
    raise AttributeError(name)

"""Metric functions with the reference's signatures (src/audio_metrics/metrics/)."""
from .fad import frechet_distance  # noqa: F401
from .kd import kernel_distance  # noqa: F401
from .prdc import prdc, nearest_neighbour_distances  # noqa: F401
from .apa import apa, apa_compute_d_x_xp  # noqa: F401

"""Accompaniment prompt adherence (reference metrics/apa.py:5-32)."""
from __future__ import annotations

from .fad import frechet_distance, frechet_distances


def apa_compute_d_x_xp(reference, anti_reference):
    """apa.py:5-6."""
    return frechet_distance(reference, anti_reference)


def apa(candidate, reference, anti_reference, d_x_xp=None):
    """apa.py:9-19; the two (or three) Frechet distances run as one batch."""
    pairs = [(candidate, reference), (candidate, anti_reference)]
    if d_x_xp is None:
        pairs.append((reference, anti_reference))
    d = frechet_distances(pairs)
    if d_x_xp is None:
        d_x_xp = d[2]
    return _apa(d[0], d[1], d_x_xp)


def _apa(d_y_x, d_y_xp, d_x_xp):
    """apa.py:22-32."""
    d_y_x = max(0, d_y_x)
    d_y_xp = max(0, d_y_xp)
    d_x_xp = max(0, d_x_xp)
    numerator = d_y_xp - d_y_x
    denominator = d_x_xp
    if abs(numerator) > denominator:
        denominator = abs(numerator)
    if denominator <= 0:
        return 0.0
    return 1 / 2 + numerator / (2 * denominator)

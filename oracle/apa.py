"""Oracle: accompaniment prompt adherence from three Frechet distances (reference metrics/apa.py)."""
from __future__ import annotations

from .fad import frechet_from_stats


def apa_from_fads(d_y_x, d_y_xp, d_x_xp) -> float:
    """apa.py:22-32 _apa."""
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


def apa(cand, ref, anti, d_x_xp=None) -> float:
    """apa.py:9-19; cand / ref / anti are (mean, cov) pairs."""
    d_y_x = frechet_from_stats(cand[0], cand[1], ref[0], ref[1])
    d_y_xp = frechet_from_stats(cand[0], cand[1], anti[0], anti[1])
    if d_x_xp is None:
        d_x_xp = frechet_from_stats(ref[0], ref[1], anti[0], anti[1])  # apa.py:5-6
    return apa_from_fads(d_y_x, d_y_xp, d_x_xp)

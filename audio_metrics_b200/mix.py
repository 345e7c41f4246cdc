"""Context+stem mix functions (reference mix_functions.py:209-344 registry).

Audio-side pre-processing upstream of the hot path; kept on the CPU in numpy like
the reference.  The peak-based mixers are implemented here; the loudness mixers
(BS.1770 via pyloudnorm + a limiter) resolve their third-party dependencies
lazily, exactly the packages the reference itself requires for them.
"""
from __future__ import annotations

from functools import partial

import numpy as np


def mix_tracks_peak_preserve(audio, sr=None):
    """Average the channels, then restore the peak amplitude of the input
    (mix_functions.py:209-227).  audio: [n_samples, n_channels]."""
    assert audio.ndim == 2
    if audio.shape[1] == 1:
        return audio[:, 0]
    peak_in = np.abs(audio).max()
    if peak_in <= 1e-5:
        return audio[:, 0]
    mix = audio.mean(axis=1)
    return mix * (peak_in / np.abs(mix).max())


def mix_tracks_peak_normalize(audio, sr=None, stem_db_red=0.0, out_db=0.0):
    """Peak-normalise each channel (stem attenuated by stem_db_red dB), sum, and
    peak-normalise the mix to out_db (mix_functions.py:230-249)."""
    assert audio.ndim == 2
    out_gain = 10.0 ** (out_db / 20.0)
    stem_gain = 10.0 ** (stem_db_red / 20.0)
    if audio.shape[1] == 1:
        mix = audio[:, 0].copy()
    else:
        peaks = np.abs(audio).max(axis=0, keepdims=True)
        peaks[0, 1] *= stem_gain
        mix = (audio / peaks).sum(axis=1)
    return mix * (out_gain / np.abs(mix).max())


def mix_tracks_loudness(audio, sr, stem_db_red=-4.0, out_db=-20.0):
    """Fixed loudness relation between context and stem, mix normalised to out_db
    LUFS, limiter above full scale (mix_functions.py:281-332)."""
    try:
        import pyloudnorm as pyln
        import numpy_audio_limiter
    except ImportError as e:  # same hard dependencies as the reference
        raise ImportError("loudness mix functions (L0/L1/L2) need `pyloudnorm` and `numpy_audio_limiter`; "
                          "use a peak mixer (PP/P0/P1/P2) or pass your own mix_function") from e
    assert audio.ndim == 2
    if audio.shape[1] == 1:
        return audio[:, 0]
    vmax = np.abs(audio).max(axis=0)
    silent = vmax < 1e-5
    if silent.all():
        return audio[:, 0]
    meter = pyln.Meter(sr)
    if silent.any():
        mix = audio[:, ~silent][:, 0]
    else:
        s0, s1 = audio.T
        l0, l1 = meter.integrated_loudness(s0), meter.integrated_loudness(s1)
        target = l0 + stem_db_red
        if not np.isinf(l1) and not np.isinf(target):
            s1 = pyln.normalize.loudness(s1, l1, target)
        mix = s0 + s1
    l_mix = meter.integrated_loudness(mix)
    if not np.isinf(l_mix) and not np.isinf(out_db):
        mix = pyln.normalize.loudness(mix, l_mix, out_db)
    if np.abs(mix).max() > 1.0:
        mix = numpy_audio_limiter.limit(signal=mix.astype(np.float32).reshape((1, -1)), attack_coeff=0.99,
                                        release_coeff=0.99, delay=527, threshold=0.5)[0]
    return mix


MIX_FUNCTIONS = dict(   # names and parameters of mix_functions.py:335-343
    PP=mix_tracks_peak_preserve,
    P0=partial(mix_tracks_peak_normalize, stem_db_red=-0, out_db=-3),
    P1=partial(mix_tracks_peak_normalize, stem_db_red=-3, out_db=-3),
    P2=partial(mix_tracks_peak_normalize, stem_db_red=-6, out_db=-3),
    L0=partial(mix_tracks_loudness, stem_db_red=0, out_db=-20),
    L1=partial(mix_tracks_loudness, stem_db_red=-3, out_db=-20),
    L2=partial(mix_tracks_loudness, stem_db_red=-6, out_db=-20),
)
DEFAULT_MIX_FUNCTION = "L0"

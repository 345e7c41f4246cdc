"""Audio -> embeddings -> per-category AudioMetricsData (reference embed.py:93-237).

Same stages as the reference (shuffle, resample, slice into windows, aligned /
misaligned / stem serialisation, mix, batch of 32, embed), written as plain
generators.  The difference that matters for the hot path is the last stage
(embed.py:226-236): embeddings stay on the GPU that produced them — no
``.cpu()`` per batch — and are appended to device-resident containers, so
FAD / KD / PRDC start from data already in HBM.
"""
from __future__ import annotations

import random
from enum import IntEnum
from itertools import tee

import numpy as np
import torch

from .data import AudioMetricsData, ensure_ndarray


class ItemCategory(IntEnum):   # embed.py:18-21
    aligned = 1
    misaligned = 2
    stem = 3


def shuffle_stream(iterator, buffer_size=100, seed=None, min_age=0):
    """Buffered stream shuffle in which a slot cannot be re-drawn until `min_age`
    further replacements happened (util/shuffle.py:5-86) — this is what guarantees
    that a "misaligned" stem does not come from the window it is paired with."""
    rng = random.Random(seed) if seed is not None else random
    iterator = iter(iterator)
    buffer = []
    for item in iterator:
        buffer.append(item)
        if len(buffer) == buffer_size:
            break
    total = len(buffer)
    if total == 0:
        return
    order = list(range(total))
    eligible = total - min(min_age, total - 1)
    offset = 0
    for item in iterator:
        j = (offset + rng.randrange(eligible)) % total
        slot = order[j]
        yield buffer[slot]
        buffer[slot] = item
        order[j], order[offset] = order[offset], order[j]
        offset = (offset + 1) % total
    rng.shuffle(order)
    for slot in order:
        yield buffer[slot]


def audio_slicer(items, win_dur, sr):
    """Fixed, non-overlapping windows; the trailing remainder is dropped (util/audio.py:1-14)."""
    win = int(sr * win_dur)
    for audio in items:
        for i in range(0, len(audio) - win + 1, win):
            yield audio[i:i + win]


def resample(audio, sr_orig, sr_new):
    """embed.py:69-83 (soxr when available, polyphase otherwise)."""
    audio = ensure_ndarray(audio)
    try:
        import soxr

        return soxr.resample(audio, sr_orig, sr_new)
    except ImportError:
        from math import gcd

        from scipy.signal import resample_poly

        g = gcd(int(sr_orig), int(sr_new))
        return resample_poly(audio, int(sr_new) // g, int(sr_orig) // g, axis=0)


def serialize_items(items1, items2=None, apa_mode=False, stems_mode=False):
    """embed.py:44-66."""
    pairs = ((a, None) for a in items1) if items2 is None else zip(items1, items2)
    msg = ("When computing APA items should be tensors/arrays of shape [n_samples, 2] "
           "(pairing context and stem)")
    for item1, item2 in pairs:
        item1 = ensure_ndarray(item1)
        if apa_mode:
            if item1.ndim != 2:
                raise ValueError(msg)
            yield {"audio": item1, "category": ItemCategory.aligned}
            if item2 is not None:
                item2 = ensure_ndarray(item2)
                assert item2.ndim == 2, msg
                yield {"audio": np.column_stack((item1[:, 0], item2[:, 1])), "category": ItemCategory.misaligned}
        if stems_mode:
            yield {"audio": item1[:, -1] if item1.ndim == 2 else item1, "category": ItemCategory.stem}


def batch_accumulator(items, batch_size=32):
    """embed.py:24-41."""
    audio, category = [], []
    for item in items:
        audio.append(item["audio"])
        category.append(int(item["category"]))
        if len(audio) == batch_size:
            yield {"audio": np.stack(audio), "category": np.array(category)}
            audio, category = [], []
    if audio:
        yield {"audio": np.stack(audio), "category": np.array(category)}


def embedding_pipeline(waveforms, embedder, mix_function, gpu_handler=None, apa_mode=None, stems_mode=False,
                       store_mix_embeddings=False, store_stem_embeddings=False, batch_size=32, win_dur=5.0,
                       song_buffer_size=100, win_buffer_size=1000, win_min_age=100, seed=None, input_sr=None,
                       device=None):
    """embed.py:93-237, returning {ItemCategory: AudioMetricsData} with device-resident state."""
    items = iter(waveforms)
    if apa_mode == "reference":
        items = shuffle_stream(items, buffer_size=song_buffer_size, seed=seed)
    if input_sr is not None and input_sr != embedder.sr:
        items = (resample(x, input_sr, embedder.sr) for x in items)
    items = audio_slicer((ensure_ndarray(x) for x in items), win_dur, embedder.sr)
    shuffled = None
    if apa_mode == "reference":
        items, shuffled = tee(items)
        shuffled = shuffle_stream(shuffled, buffer_size=win_buffer_size, min_age=win_min_age, seed=seed)
    items = serialize_items(items, shuffled, apa_mode is not None, stems_mode)
    if apa_mode is not None:
        items = ({"audio": it["audio"] if it["category"] == ItemCategory.stem
                  else mix_function(it["audio"], sr=embedder.sr), "category": it["category"]} for it in items)

    data = {}
    if apa_mode is not None:
        data[ItemCategory.aligned] = AudioMetricsData(store_mix_embeddings, device=device)
    if apa_mode == "reference":
        data[ItemCategory.misaligned] = AudioMetricsData(store_mix_embeddings, device=device)
    if stems_mode:
        data[ItemCategory.stem] = AudioMetricsData(store_stem_embeddings, device=device)

    for batch in batch_accumulator(items, batch_size):
        emb = embedder.forward({"audio": batch["audio"]})["embedding"]
        cat = batch["category"]
        cat_dev = None
        for c, dst in data.items():
            mask = cat == int(c)
            count = int(mask.sum())
            if not count:
                continue
            # embed.py:231-236, on the device the container lives on
            if count == len(cat):
                dst.add(emb)
            elif dst.store_embeddings or not emb.is_cuda:
                sel = torch.as_tensor(np.nonzero(mask)[0], device=emb.device)
                dst.add(emb.index_select(0, sel))
            else:   # statistics only: one masked moment launch per category, no selection copy
                if cat_dev is None:
                    cat_dev = torch.as_tensor(cat.astype(np.int32), device=emb.device)
                dst.add_masked(emb, cat_dev, int(c), count)
    return data

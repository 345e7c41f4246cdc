#!/usr/bin/env python
"""Largest BASELINE size (config 5: 1M rows) through PRDC: a set against itself must give
precision = recall = coverage = 1 and density = 1 (strict '<' excludes the k-th neighbour tie)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from audio_metrics_b200 import AudioMetricsData, prdc
from audio_metrics_b200.synth import make_sets_torch
n, d, k = 1_000_000, 128, 5
ref, _ = make_sets_torch(n, 256, d, device="cuda")
R = AudioMetricsData(True); R.embeddings = ref
torch.cuda.synchronize(); t = time.perf_counter()
out = prdc(R, R, k)
torch.cuda.synchronize(); dt = time.perf_counter() - t
pairs = 2.0 * n * n     # one radii sweep (cached for both roles) + the count sweep
print(f"n={n} d={d} k={k}: {out}  {dt*1e3:.0f} ms  ({pairs/dt:.3e} pairs/s)")
assert out["precision"] == 1.0 and out["recall"] == 1.0 and out["coverage"] == 1.0 and abs(out["density"] - 1.0) < 1e-4
print("ok")

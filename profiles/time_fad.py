#!/usr/bin/env python
"""Time statistics + Frechet distance at bench size; AMB_JACOBI=flat selects the old
one-round-per-grid-barrier Jacobi kernel for comparison."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from audio_metrics_b200 import AudioMetricsData, frechet_distance
from audio_metrics_b200.synth import make_sets_torch

n = int(os.environ.get("AMB_BENCH_N", 200000))
d = int(os.environ.get("AMB_BENCH_D", 512))
ref, cand = make_sets_torch(n, n, d, device="cuda")
def ev():
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    R, C = AudioMetricsData(False), AudioMetricsData(False)
    R.add(ref); C.add(cand)
    e[1].record()
    f = frechet_distance(C, R)
    e[2].record(); torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), f
ev()
best = None
for _ in range(3):
    s, f, val = ev()
    best = (s, f) if best is None else (min(best[0], s), min(best[1], f))
print(f"jacobi={os.environ.get('AMB_JACOBI','block')} n={n} d={d}: stats {best[0]:.2f} ms, frechet {best[1]:.2f} ms, fad={val!r}")

#!/usr/bin/env python
"""Time the PRDC phases (radii x2, counts) at bench size for the current library
settings; used with AMB_SPLIT_MB=... to tune the L2 column-split size."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from audio_metrics_b200 import AudioMetricsData, prdc
from audio_metrics_b200.synth import make_sets_torch

n = int(os.environ.get("AMB_BENCH_N", 200000))
ref, cand = make_sets_torch(n, n, 512, device="cuda")
def run():
    R, C = AudioMetricsData(True), AudioMetricsData(True)
    R.embeddings = ref; C.embeddings = cand
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    R.packed(); C.packed()
    e[0].record(); R.get_radii(5); e[1].record(); C.get_radii(5); e[2].record(); out = prdc(R, C, 5); e[3].record()
    torch.cuda.synchronize()
    return [e[i].elapsed_time(e[i + 1]) for i in range(3)], out
run()
best = None
for _ in range(3):
    t, out = run()
    best = t if best is None else [min(a, b) for a, b in zip(best, t)]
from audio_metrics_b200.metrics.prdc import prdc_totals
R, C = AudioMetricsData(True), AudioMetricsData(True)
R.embeddings = ref; C.embeddings = cand
_, _, _, tot = prdc_totals(R, C, 5)
print("uncertain pairs:", int(tot[4]), "cap", 16 * 2 * n)
if os.environ.get("AMB_SAME"):
    prdc(R, R, 5); torch.cuda.synchronize()
print(f"SPLIT_MB={os.environ.get('AMB_SPLIT_MB','default')} n={n}: radii_ref {best[0]:.1f} ms, radii_cand {best[1]:.1f} ms, counts {best[2]:.1f} ms, total {sum(best):.1f} ms", out)

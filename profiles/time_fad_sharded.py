#!/usr/bin/env python
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from audio_metrics_b200.dist import evaluate_sharded
from audio_metrics_b200.synth import make_sets_torch
n = 200000
ref, cand = make_sets_torch(n, n, 512, device="cuda")
def run(metrics):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = evaluate_sharded(ref, cand, n, n, metrics=metrics); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for m in (("fad",), ("fad", "kd", "prdc")):
    for _ in range(3): run(m)
    raw = [run(m) for _ in range(16)]
    print('   ', ' '.join(f'{t:.1f}' for t in raw))
    ts = sorted(raw)
    print(f"FAD_SIDE={os.environ.get('AMB_FAD_SIDE','1')} metrics={m}: min {ts[0]:.1f} median {ts[8]:.1f} max {ts[-1]:.1f} ms", flush=True)

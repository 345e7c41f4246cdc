#!/usr/bin/env python
"""Time the full evaluate step (device-resident inputs) under different scheduling knobs."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from audio_metrics_b200.dist import evaluate_sharded
from audio_metrics_b200.synth import make_sets_torch
n = int(os.environ.get("AMB_BENCH_N", 200000))
ref, cand = make_sets_torch(n, n, 512, device="cuda")
def run():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = evaluate_sharded(ref, cand, n, n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), out
for _ in range(3): run()
ts = [run()[0] for _ in range(12)]
print(f"FAD_SIDE={os.environ.get('AMB_FAD_SIDE','1')} SCHED={os.environ.get('AMB_SCHED','dynamic')} JBS={os.environ.get('AMB_JACOBI_BS','auto')}: step min {min(ts):.1f} ms median {sorted(ts)[6]:.1f} ms max {max(ts):.1f} ms")

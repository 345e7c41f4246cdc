#!/usr/bin/env python
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from audio_metrics_b200.dist import evaluate_sharded, CudaOps
ops = CudaOps()
ops.trace = []
from audio_metrics_b200.synth import make_sets_torch
n = 200000
FADS = []
ref, cand = make_sets_torch(n, n, 512, device="cuda")
def run(metrics):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = evaluate_sharded(ref, cand, n, n, metrics=metrics, ops=(ops if os.environ.get('AMB_PERSIST_OPS', '0') == '1' else None)); e1.record(); torch.cuda.synchronize()
    tr = ops.trace if os.environ.get('AMB_PERSIST_OPS', '0') == '1' else None
    fad = tr[-1][0].elapsed_time(tr[-1][1]) if tr else 0.0
    lead = e0.elapsed_time(tr[-1][0]) if tr else 0.0
    FADS.append((round(lead, 1), round(fad, 1)))
    return e0.elapsed_time(e1)
for m in (("fad", "kd", "prdc"),):
    for _ in range(3): run(m)
    raw = [run(m) for _ in range(int(os.environ.get('AMB_ITERS', 16)))]
    fads = FADS[-len(raw):]
    print('   typical (fad start, fad ms):', fads[:4])
    print('   outliers (>1.3x median) (iter, step ms, fad start, fad ms):', [(i, round(t, 1)) + fads[i] for i, t in enumerate(raw) if t > 1.3 * sorted(raw)[len(raw) // 2]])
    ts = sorted(raw)
    print(f"FAD_SIDE={os.environ.get('AMB_FAD_SIDE','1')} prio={os.environ.get('AMB_SIDE_PRIO','-1')} shared={os.environ.get('AMB_SHARED_SMS','16')} metrics={m}: min {ts[0]:.1f} median {ts[len(ts)//2]:.1f} max {ts[-1]:.1f} ms", flush=True)

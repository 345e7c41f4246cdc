"""Distribution of per-step wall and device times (looking for sporadic stalls)."""
import sys, time
sys.path.insert(0, ".")
import torch
from audio_metrics_b200 import AudioMetricsData, frechet_distance
from audio_metrics_b200.dist import evaluate_containers
from audio_metrics_b200.synth import make_sets_torch

dev = torch.device("cuda", 0)
n = 200_000
ref, cand = make_sets_torch(n, n, 512, device=dev)
def mk(x, y, store=True):
    A, B = AudioMetricsData(store, dev), AudioMetricsData(store, dev)
    A.add(x); B.add(y)
    return A, B
def dist(label, fn, reps):
    w, g = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        w.append((time.perf_counter() - t0) * 1e3); g.append(e0.elapsed_time(e1))
    ws = sorted(w)
    print(f"{label:34s} wall min {ws[0]:7.2f} med {ws[len(ws)//2]:7.2f} max {ws[-1]:7.2f} | device med {sorted(g)[len(g)//2]:7.2f} max {max(g):7.2f} | "
          + " ".join(f"{x:.0f}" for x in w), flush=True)
R, C = mk(ref, cand, False); R.mean; C.mean
dist("frechet_distance only (stats cached)", lambda: frechet_distance(C, R), 40)
dist("fad step (add + cov + fad)", lambda: evaluate_containers(*mk(ref, cand), ("fad",), nearest_k=5), 30)
dist("kd step", lambda: evaluate_containers(*mk(ref, cand), ("kd",), nearest_k=5), 30)
dist("full step", lambda: evaluate_containers(*mk(ref, cand), ("fad", "kd", "prdc"), nearest_k=5), 20)
from audio_metrics_b200._lib import options
with options(fad_method=1):
    dist("frechet_distance only, jacobi", lambda: frechet_distance(C, R), 20)

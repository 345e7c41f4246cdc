#!/usr/bin/env python
"""BASELINE.json configs 1-3 and 5 on one GPU (config 4 is bench.py): wall time of each public call
after warm-up, with the result, to show that every configuration runs through the public API."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from audio_metrics_b200 import AudioMetricsData, frechet_distance, kernel_distance, prdc, apa
from audio_metrics_b200.dist import evaluate_containers
from audio_metrics_b200.metrics.apa import _apa
from audio_metrics_b200.synth import make_sets_torch, make_apa_sets_numpy

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); out = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best * 1e3, out

def containers(ref, cand, store=True):
    R, C = AudioMetricsData(store), AudioMetricsData(store)
    R.add(ref); C.add(cand)
    return R, C

def all_metrics(ref, cand, k):
    """What AudioMetrics.evaluate runs after embedding: fresh containers, the fused step, one read-back."""
    R, C = containers(ref, cand)
    return evaluate_containers(R, C, ("fad", "kd", "prdc"), nearest_k=k)

def all_metrics_separate(ref, cand, k):
    """The same through the three metric functions (one synchronising call each, as the reference)."""
    R, C = containers(ref, cand)
    return dict(fad=frechet_distance(C, R), **kernel_distance(C, R), **prdc(R, C, k))

print("C1  100 x 100 x 128 (VGGish-sized), fad+kd+prdc k=10")
ref, cand = make_sets_torch(100, 100, 128, device="cuda")
ms, out = timed(lambda: all_metrics(ref, cand, 10)); print(f"    {ms:8.2f} ms  {out}")
print("C2  10k x 10k x 512, fad+kd+prdc k=5")
ref, cand = make_sets_torch(10000, 10000, 512, device="cuda")
ms, out = timed(lambda: all_metrics(ref, cand, 5)); print(f"    {ms:8.2f} ms  {out}")
ms, out = timed(lambda: all_metrics_separate(ref, cand, 5)); print(f"    {ms:8.2f} ms  (three separate metric calls)")
print("C3  APA on 10k mix/stem pairs x 512 + FAD on stems")
s = {k: torch.from_numpy(v).cuda() for k, v in make_apa_sets_numpy(10000, 512, seed=5).items()}
def c3():
    mk = lambda x: (lambda a: (a.add(x), a)[1])(AudioMetricsData(False))
    cand, ref, anti = mk(s["cand_aligned"]), mk(s["ref_aligned"]), mk(s["ref_misaligned"])
    res = evaluate_containers(mk(s["ref_stems"]), mk(s["cand_stems"]), ("fad",), apa=(cand, ref, anti, None))
    return dict(apa=_apa(res["_d_y_x"], res["_d_y_xp"], res["_d_x_xp"]), fad_stems=res["fad"])
ms, out = timed(c3); print(f"    {ms:8.2f} ms  {out}")
for d in (512, 512, 128):
    print(f"C5  FAD/KD, 100k references vs 1M candidates x {d} (one GPU of the eight)")
    ref, cand = make_sets_torch(100000, 1000000, d, device="cuda")
    def c5():
        R, C = containers(ref, cand)
        return dict(fad=frechet_distance(C, R), **kernel_distance(C, R))
    ms, out = timed(c5, reps=2); print(f"    {ms:8.2f} ms  {out}")
    del ref, cand

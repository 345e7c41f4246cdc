"""Does a host-to-device copy in flight on another stream slow the all-pairs sweep down?"""
import sys, time
sys.path.insert(0, ".")
import torch
from audio_metrics_b200 import AudioMetricsData
from audio_metrics_b200.dist import evaluate_containers
from audio_metrics_b200.metrics.prdc import nearest_neighbour_distances
from audio_metrics_b200.synth import make_sets_torch

dev = torch.device("cuda", 0)
n = 200_000
ref, cand = make_sets_torch(n, n, 512, device=dev)
hp = ref.cpu().pin_memory()
hq = cand.cpu().pin_memory()
cs = torch.cuda.Stream(dev)
R = AudioMetricsData(True, dev); R.add(ref); R.packed()


def ev_time(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print("radii sweep alone                      %8.3f ms" % ev_time(lambda: nearest_neighbour_distances(R, 5)))
def with_copy(k=1):
    keep = []
    with torch.cuda.stream(cs):
        for _ in range(k):
            keep.append(hp.to(dev, non_blocking=True))
    r = nearest_neighbour_distances(R, 5)
    torch.cuda.current_stream().wait_stream(cs)
    return keep, r
print("radii sweep + 1 H2D copy on side stream %8.3f ms" % ev_time(lambda: with_copy(1)))
print("radii sweep + 3 H2D copies              %8.3f ms" % ev_time(lambda: with_copy(3)))
def mk(x, y):
    A, B = AudioMetricsData(True, dev), AudioMetricsData(True, dev)
    A.add(x); B.add(y)
    return A, B
def wall(label, fn, reps=4):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("%-46s %8.3f ms" % (label, best * 1e3), flush=True)
wall("step all, device inputs", lambda: evaluate_containers(*mk(ref, cand), ("fad", "kd", "prdc"), nearest_k=5))
wall("step all, pinned host inputs", lambda: evaluate_containers(*mk(hp, hq), ("fad", "kd", "prdc"), nearest_k=5))
wall("step prdc, pinned host inputs", lambda: evaluate_containers(*mk(hp, hq), ("prdc",), nearest_k=5))
wall("step fad, pinned host inputs", lambda: evaluate_containers(*mk(hp, hq), ("fad",), nearest_k=5))
wall("step kd, pinned host inputs", lambda: evaluate_containers(*mk(hp, hq), ("kd",), nearest_k=5))
# copy first, fully, then the device step
def staged():
    a = hp.to(dev, non_blocking=True); b = hq.to(dev, non_blocking=True)
    return evaluate_containers(*mk(a, b), ("fad", "kd", "prdc"), nearest_k=5)
wall("H2D on main stream, then device step", staged)

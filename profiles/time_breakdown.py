"""Host/device time breakdown of one fused step (where do the non-kernel milliseconds go)."""
import sys, time
sys.path.insert(0, ".")
import torch
from audio_metrics_b200 import AudioMetricsData
from audio_metrics_b200.dist import evaluate_containers
from audio_metrics_b200.synth import make_sets_torch

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
ref, cand = make_sets_torch(n, n, 512, device=dev)


def t(label, fn, reps=3):
    torch.cuda.synchronize()
    out = None
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"{label:45s} {best * 1e3:9.3f} ms", flush=True)
    return out


t("clone 410 MB", lambda: ref.clone())
t("empty 410 MB", lambda: torch.empty_like(ref))
hp = ref.cpu().pin_memory()
t("H2D pinned 410 MB (.to non_blocking)", lambda: hp.to(dev, non_blocking=True))
cs = torch.cuda.Stream(dev)
def h2d_side():
    with torch.cuda.stream(cs):
        x = hp.to(dev, non_blocking=True)
    x.record_stream(torch.cuda.current_stream(dev))
    return x
t("H2D pinned on side stream + record_stream", h2d_side)
def mk(x, y):
    R, C = AudioMetricsData(True, dev), AudioMetricsData(True, dev)
    R.add(x); C.add(y)
    return R, C
t("2 x add(device)", lambda: mk(ref, cand))
t("2 x add(pinned host)", lambda: mk(hp, hp))
R, C = mk(ref, cand)
t("R.mean (fold: cov of 200k)", lambda: mk(ref, cand)[0].mean)
t("packed()", lambda: mk(ref, cand)[0].packed())
for m in (("fad",), ("kd",), ("prdc",), ("fad", "kd", "prdc")):
    t(f"step {m} device inputs", lambda: evaluate_containers(*mk(ref, cand), m, nearest_k=5))
t("step all, pinned host inputs", lambda: evaluate_containers(*mk(hp, hp.clone().pin_memory() if False else hp), ("fad", "kd", "prdc"), nearest_k=5))
# evaluate only (containers prebuilt, caches dropped by rebuilding)
import cProfile, pstats
R, C = mk(ref, cand)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
evaluate_containers(R, C, ("fad", "kd"), nearest_k=5)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

"""Radii sweep (200k x 200k x 512) timing, for A/B runs of library builds (AMB200_LIB)."""
import sys
sys.path.insert(0, ".")
import torch
from audio_metrics_b200 import AudioMetricsData
from audio_metrics_b200.metrics.prdc import nearest_neighbour_distances, prdc_totals
from audio_metrics_b200.synth import make_sets_torch
dev = torch.device("cuda", 0)
ref, cand = make_sets_torch(200_000, 200_000, 512, device=dev)
R = AudioMetricsData(True, dev); R.embeddings = ref; R.packed()
C = AudioMetricsData(True, dev); C.embeddings = cand; C.packed()
def ev(fn, reps=6):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]
r = nearest_neighbour_distances(R, 5); rc = nearest_neighbour_distances(C, 5)
print("radii  min %.3f med %.3f ms" % ev(lambda: nearest_neighbour_distances(R, 5)))
print("counts min %.3f med %.3f ms" % ev(lambda: prdc_totals(R, C, 5, ref_radii=r, cand_radii=rc)))
print("checksum", float(r.double().sum()))

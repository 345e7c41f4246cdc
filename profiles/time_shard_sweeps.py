"""Per-rank sweep times of the sharded step, reproduced on one GPU: a rank's row share against all
columns (what it runs at N = 8), for several column-split settings."""
import sys
sys.path.insert(0, ".")
import torch
from audio_metrics_b200 import AudioMetricsData
from audio_metrics_b200._lib import options
from audio_metrics_b200.dist import work_rows, work_weights
from audio_metrics_b200.metrics.prdc import nearest_neighbour_distances, prdc_totals
from audio_metrics_b200.synth import make_sets_torch

dev = torch.device("cuda", 0)
n, d, k = 200_000, 512, 5
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ref, cand = make_sets_torch(n, n, d, device=dev)
R, C = AudioMetricsData(True, dev), AudioMetricsData(True, dev)
R.embeddings = ref; C.embeddings = cand
R.packed(); C.packed()
r_ref, r_cand = nearest_neighbour_distances(R, k), nearest_neighbour_distances(C, k)


def ev(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


w = work_weights(world, n, n, d, True)
for rank in (0, 1):
    row0, nrows = work_rows(n, w, rank)
    print(f"rank {rank}: rows [{row0}, +{nrows})  ideal share of a 27 ms sweep: {27.0 * nrows / n:.2f} ms")
    for split in (1, 2, 3, 4, 8, 0, 0):
        with options(topk_split=split, count_split=split):
            t_r = ev(lambda: nearest_neighbour_distances(R, k, row_range=(row0, nrows)))
            t_c = ev(lambda: prdc_totals(R, C, k, row_range=(row0, nrows), ref_radii=r_ref, cand_radii=r_cand))
        print(f"   split {split:2d} (0 = automatic):  radii {t_r:6.3f} ms   counts {t_c:6.3f} ms", flush=True)

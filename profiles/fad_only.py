"""FAD from cached statistics, a few times (target for ncu captures of the FAD kernels)."""
import sys
sys.path.insert(0, ".")
import torch
from audio_metrics_b200 import AudioMetricsData, frechet_distance
from audio_metrics_b200.synth import make_sets_torch
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ref, cand = make_sets_torch(n, n, d, device=dev)
R, C = AudioMetricsData(False, dev), AudioMetricsData(False, dev)
R.add(ref); C.add(cand); R.mean; C.mean
for _ in range(3):
    print(frechet_distance(C, R))

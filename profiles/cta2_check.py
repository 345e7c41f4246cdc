import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from audio_metrics_b200 import _lib, AudioMetricsData
from audio_metrics_b200.metrics.prdc import nearest_neighbour_distances, prdc_totals
from audio_metrics_b200.synth import make_sets_numpy
L = _lib.lib()
# 1. dot matrix through the pair kernel vs fp64
rng = np.random.default_rng(0)
for (na, nb, d) in [(256, 256, 64), (300, 700, 512), (513, 257, 100)]:
    A = rng.standard_normal((na, d)).astype(np.float32); B = rng.standard_normal((nb, d)).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    pA = _lib.workspace(L.amb_packed_bytes(na, d), torch.device("cuda", 0)); pB = _lib.workspace(L.amb_packed_bytes(nb, d), torch.device("cuda", 0))
    C = torch.full((na, nb), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(L.amb_pack(0, None, dA.data_ptr(), 0, na, d, d, pA.data_ptr())); _lib.check(L.amb_pack(0, None, dB.data_ptr(), 0, nb, d, d, pB.data_ptr()))
    os.environ["AMB_DEBUG_SINGLE"] = "2"
    _lib.check(L.amb_debug_dot_matrix(0, None, pA.data_ptr(), na, pB.data_ptr(), nb, d, C.data_ptr(), nb, 0, 0))
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.linalg.norm(A, axis=1)[:, None] * np.linalg.norm(B, axis=1)[None, :]
    err = np.abs(C.cpu().numpy() - ref) / scale
    print("dot", (na, nb, d), "max rel err", float(err.max()), "nan", int(np.isnan(err).sum()), flush=True)
# 2. radii and counts equal to the single-CTA engine
ref, cand = make_sets_numpy(3000, 2600, 512, seed=11)
out = {}
for mode in ("0", "1"):
    os.environ["AMB_CTA2"] = mode
    R, Cn = AudioMetricsData(True), AudioMetricsData(True)
    R.add(torch.from_numpy(ref)); Cn.add(torch.from_numpy(cand))
    r = nearest_neighbour_distances(R, 5)
    col, rec, cov, tot = prdc_totals(R, Cn, 5)
    torch.cuda.synchronize()
    out[mode] = (r.cpu().numpy(), col.cpu().numpy(), rec.cpu().numpy(), cov.cpu().numpy(), int(tot[4]))
print("radii equal", np.array_equal(out["0"][0], out["1"][0]), "counts equal", all(np.array_equal(a, b) for a, b in zip(out["0"][1:4], out["1"][1:4])), "uncertain", out["0"][4], out["1"][4])

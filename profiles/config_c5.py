"""BASELINE.json config 5: FAD / KD sweep over the three CLAP layers (512 / 512 / 128-d) at
N = 1M candidates vs 100k references, row-sharded over the ranks of one box.

    python -m torch.distributed.run --nproc-per-node 8 ... profiles/config_c5.py

Each rank generates only its own rows (independent generators: the three layers use seeds
1234 / 1235 / 1236 as SURVEY.md 8d prescribes).  Prints one JSON line per layer on rank 0."""
import json, os, sys, time
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from audio_metrics_b200 import AudioMetricsData
from audio_metrics_b200.dist import evaluate_containers, shard_rows
from audio_metrics_b200.synth import LATENT, NOISE

world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N_REF, N_CAND = int(os.environ.get("C5_REF", 100_000)), int(os.environ.get("C5_CAND", 1_000_000))


def rows(seed, d, n, row0, scale, shift):
    """Rows [row0, row0 + n) of a CLAP-like set: shared basis from `seed`, per-rank latent stream."""
    g = torch.Generator(device=dev).manual_seed(seed)
    W = torch.randn(LATENT, d, generator=g, device=dev)
    g2 = torch.Generator(device=dev).manual_seed(seed * 1000 + row0 // 256 + int(scale * 7))
    z = torch.randn(n, LATENT, generator=g2, device=dev) * scale + shift
    e = z @ W / LATENT**0.5 + NOISE * torch.randn(n, d, generator=g2, device=dev)
    return torch.nn.functional.normalize(e, dim=-1).contiguous()


for layer, (seed, d) in enumerate(((1234, 512), (1235, 512), (1236, 128))):
    r0, rn, _ = shard_rows(N_REF, world, rank); c0, cn, _ = shard_rows(N_CAND, world, rank)
    ref, cand = rows(seed, d, rn, r0, 1.0, 0.0), rows(seed, d, cn, c0, 1.15, 0.1)

    def step():
        R, C = AudioMetricsData(True, dev), AudioMetricsData(True, dev)
        R.add(ref); C.add(cand)
        return evaluate_containers(R, C, ("fad", "kd"), nearest_k=5)

    for _ in range(2):
        res = step()
    times = {}
    for name, fn in (("fad+kd", step),
                     ("stats only", lambda: (lambda R: (R.add(cand), R.mean))(AudioMetricsData(False, dev)))):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times[name] = float(ms)
    if rank == 0:
        gbs = cn * d * 4 / (times["stats only"] * 1e-3) / 1e9
        print(json.dumps({"config": "C5", "layer": layer, "d": d, "n_ref": N_REF, "n_cand": N_CAND, "n_gpus": world,
                          "ms_fad_kd": times["fad+kd"], "ms_candidate_statistics_per_rank": times["stats only"],
                          "candidate_shard_rows": cn, "covariance_GBps_per_gpu": gbs, "result": res}), flush=True)
if world > 1:
    dist.destroy_process_group()

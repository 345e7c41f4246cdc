"""Streaming aggregation cost (the reference's embedding pipeline calls
AudioMetricsData.add once per 32-row batch per category; 6250 Chan merges of 512 x 512 fp64 for
200k rows take 9.98 s on the CPU).  Measures the same call pattern here, embeddings resident on the
GPU as the pipeline produces them."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from audio_metrics_b200 import AudioMetricsData
from audio_metrics_b200.synth import make_sets_torch

dev = torch.device("cuda", 0)
n, d, bs = 200_000, 512, 32
x, _ = make_sets_torch(n, 8, d, device=dev)
one = AudioMetricsData(False, dev); one.add(x); ref_cov = one.cov.clone(); ref_mean = one.mean.clone()

def run(label, store, masked=False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = AudioMetricsData(store, dev)
    if masked:
        cat = torch.zeros(bs, dtype=torch.int32, device=dev); cat[::2] = 1      # half of every batch belongs here
        for i in range(0, n, bs):
            c.add_masked(x[i:i + bs], cat, 1, bs // 2)
    else:
        for i in range(0, n, bs):
            c.add(x[i:i + bs])
    cov = c.cov
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    err = float((cov - ref_cov).abs().max() / ref_cov.abs().max()) if not masked else float("nan")
    print(f"{label:58s} {dt:7.3f} s  ({dt / (n / bs) * 1e6:6.1f} us per batch)  max rel cov error vs single shot {err:.2e}", flush=True)

run("store_embeddings=True  (append, statistics deferred)", True)
run("store_embeddings=False (one moment launch per batch)", False)
run("store_embeddings=False, masked half batches", False, masked=True)
run("store_embeddings=True  again", True)

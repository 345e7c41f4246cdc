"""world_size-2 gloo test of the row-sharding / reduction logic in
audio_metrics_b200.dist, with oracle stand-in kernels (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from oracle.prdc import cdist_exact
from audio_metrics_b200.dist import evaluate_sharded, shard_rows, work_rows, work_weights
from audio_metrics_b200.synth import make_sets_numpy


def test_shard_rows_cover_and_align():
    for n in (1, 100, 128, 129, 1000, 200000, 12345):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(n, world, r) for r in range(world)]
            assert sum(s[1] for s in spans) == n
            pos = 0
            for row0, nrows, chunk in spans:
                assert row0 % 128 == 0 or nrows == 0
                assert row0 == min(pos, n) or nrows == 0
                pos += chunk
            # only the last non-empty shard may be partial -> gathered[:n] is the global array
            nonempty = [s for s in spans if s[1]]
            assert all(s[1] == s[2] for s in nonempty[:-1])


def test_work_partition_covers_and_aligns():
    """The weighted sweep partition (the rank that also computes the Frechet distance sweeps fewer
    rows): contiguous, 256-aligned starts, covers [0, n) exactly, equal shares without FAD."""
    for n_ref, n_cand, d in ((200000, 200000, 512), (1000, 900, 128), (130, 257, 32), (1, 5, 8), (10000, 1000000, 512)):
        for world in (1, 2, 3, 4, 8):
            for with_fad in (False, True):
                w = work_weights(world, n_ref, n_cand, d, with_fad)
                assert len(w) == world and all(0 < x <= 1 for x in w)
                if not with_fad or world == 1:
                    assert w == [1.0] * world
                else:
                    assert w[0] <= 1.0 and w[1:] == [1.0] * (world - 1)
                for n in (n_ref, n_cand):
                    spans = [work_rows(n, w, r) for r in range(world)]
                    pos = 0
                    for row0, nrows in spans:
                        assert row0 == pos and (row0 % 256 == 0 or nrows == 0) and nrows >= 0
                        pos += nrows
                    assert pos == n
    w = work_weights(8, 200000, 200000, 512, True)
    assert 0.6 < w[0] < 0.8          # 3.8 ms of FAD against 12 ms of sweeps per rank


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_ref, n_cand, d, k, out, overflow=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle_ops import OracleOps
        ref, cand = make_sets_numpy(n_ref, n_cand, d, seed=77)
        if overflow:     # any row counts per rank are accepted: who holds which rows is the caller's business
            r0, rn = (0, n_ref // 3) if rank == 0 else (n_ref // 3, n_ref - n_ref // 3)
            c0, cn = (0, n_cand - 7) if rank == 0 else (n_cand - 7, 7)
        else:
            r0, rn, _ = shard_rows(n_ref, world, rank)
            c0, cn, _ = shard_rows(n_cand, world, rank)
        ops = OracleOps()
        ops.overflow_once = overflow
        res = evaluate_sharded(torch.from_numpy(ref[r0:r0 + rn]), torch.from_numpy(cand[c0:c0 + cn]), n_ref, n_cand,
                               nearest_k=k, ops=ops, kd_subsets=7, kd_subset_size=60)
        out[rank] = res
        out[f"calls{rank}"] = list(ops.calls)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_ref,n_cand,overflow", [(300, 200, False), (130, 257, False), (600, 300, True)])
def test_sharded_equals_single_process(n_ref, n_cand, overflow):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    d, k, world = 32, 4, 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_ref, n_cand, d, k, out, overflow), nprocs=world, join=True)
    if overflow:
        # rank 0's refine list "overflowed" (more near-ties than capacity) while rank 1's did not: the
        # MAX-reduced count makes BOTH ranks repeat the count sweep with a list of the reported size
        assert out["calls0"] == out["calls1"] == [None, (1 << 20) + 5]
    else:
        assert out["calls0"] == out["calls1"] == [None]
    ref, cand = make_sets_numpy(n_ref, n_cand, d, seed=77)
    assert out[0] == out[1]
    res = out[0]
    mr, cr, _ = oracle.batch_stats(ref, compute_dtype=np.dtype(np.float64))
    mc, cc, _ = oracle.batch_stats(cand, compute_dtype=np.dtype(np.float64))
    assert res["fad"] == pytest.approx(oracle.frechet_from_stats(mc, cc, mr, cr), rel=1e-9)
    kd = oracle.kernel_distance(cand, ref, subsets=7, subset_size=60, compute_dtype=np.float64)
    assert res["kernel_distance_mean"] == pytest.approx(kd["kernel_distance_mean"], rel=1e-9)
    assert res["kernel_distance_std"] == pytest.approx(kd["kernel_distance_std"], rel=1e-9)
    f32 = lambda a, b: cdist_exact(a, b).astype(np.float32)
    want = oracle.prdc(ref, cand, k, dist=f32)
    for key in ("precision", "recall", "density", "coverage"):
        assert res[key] == pytest.approx(want[key], abs=1e-12)

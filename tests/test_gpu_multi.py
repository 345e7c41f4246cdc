"""Hardware test of the row-sharded path: two ranks over NCCL against the single-GPU result
(skipped on a box with fewer than two GPUs).  The gloo test (test_dist_cpu.py) covers the same
reduction logic with stand-in kernels; this one runs the product kernels and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_ref, n_cand, d, k, out):
    import torch.distributed as dist

    from audio_metrics_b200 import AudioMetricsData
    from audio_metrics_b200.dist import evaluate_containers, evaluate_sharded, shard_rows
    from audio_metrics_b200.synth import make_sets_numpy

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ref, cand = make_sets_numpy(n_ref, n_cand, d, seed=77)
        r0, rn, _ = shard_rows(n_ref, world, rank)
        c0, cn, _ = shard_rows(n_cand, world, rank)
        R, C = AudioMetricsData(True, dev), AudioMetricsData(True, dev)
        R.add(torch.from_numpy(ref[r0:r0 + rn]))
        C.add(torch.from_numpy(cand[c0:c0 + cn]))        # (700, 130): rank 1 holds no candidate rows at all
        a = evaluate_containers(R, C, ("fad", "kd", "prdc"), nearest_k=k, kd_subsets=12, kd_subset_size=200)
        b = evaluate_sharded(torch.from_numpy(ref[r0:r0 + rn]).to(dev), torch.from_numpy(cand[c0:c0 + cn]).to(dev),
                             n_ref, n_cand, nearest_k=k, kd_subsets=12, kd_subset_size=200)
        # PCA fitted on each rank's rows: the statistics are reduced over the ranks first (dist.global_stats)
        from audio_metrics_b200.projection import IncrementalPCA
        pca = IncrementalPCA(n_components=16, device=dev).fit(R)
        out[rank] = (a, b)
        out[f"pca{rank}"] = (pca.components_.cpu().numpy(), pca.singular_values_.cpu().numpy(), pca.n_samples_seen_)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_ref,n_cand", [(3000, 2500), (700, 130)])
def test_two_ranks_equal_one(cuda_device, n_ref, n_cand):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from audio_metrics_b200 import AudioMetricsData
    from audio_metrics_b200.dist import evaluate_containers
    from audio_metrics_b200.synth import make_sets_numpy

    d, k, world = 128, 5, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_ref, n_cand, d, k, out), nprocs=world, join=True)
    ref, cand = make_sets_numpy(n_ref, n_cand, d, seed=77)
    R, C = AudioMetricsData(True), AudioMetricsData(True)
    R.add(torch.from_numpy(ref)); C.add(torch.from_numpy(cand))
    want = evaluate_containers(R, C, ("fad", "kd", "prdc"), nearest_k=k, kd_subsets=12, kd_subset_size=200)
    for rank in range(world):
        for got in out[rank]:
            assert got["fad"] == pytest.approx(want["fad"], rel=1e-9)
            for key in ("kernel_distance_mean", "kernel_distance_std"):
                assert got[key] == pytest.approx(want[key], rel=1e-9), key
            for key in ("precision", "recall", "density", "coverage"):
                assert got[key] == want[key], key      # ratios of exact integer counts: identical
    from audio_metrics_b200.projection import IncrementalPCA
    pca = IncrementalPCA(n_components=16).fit(R)
    for rank in range(world):
        comp, sv, seen = out[f"pca{rank}"]
        assert seen == n_ref
        np.testing.assert_allclose(sv, pca.singular_values_.cpu().numpy(), rtol=1e-10)
        np.testing.assert_allclose(comp, pca.components_.cpu().numpy(), atol=1e-7)


def test_in_process_devices_equal_one(cuda_device):
    """AudioMetrics(device_indices=[0, 1]) shape: one process, one host thread, the sweeps sharded over
    both GPUs by peer copies (dist.evaluate_devices) — same numbers as one device, and the facade
    routes through it."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from audio_metrics_b200 import AudioMetrics, AudioMetricsData
    from audio_metrics_b200.dist import evaluate_containers, evaluate_devices
    from audio_metrics_b200.synth import make_sets_numpy

    ref, cand = make_sets_numpy(5000, 4300, 256, seed=19)
    R, C = AudioMetricsData(True), AudioMetricsData(True)
    R.add(torch.from_numpy(ref)); C.add(torch.from_numpy(cand))
    want = evaluate_containers(R, C, ("fad", "kd", "prdc"), nearest_k=5)
    R2, C2 = AudioMetricsData(True), AudioMetricsData(True)
    R2.add(torch.from_numpy(ref)); C2.add(torch.from_numpy(cand))
    got = evaluate_devices(R2, C2, [0, 1], ("fad", "kd", "prdc"), nearest_k=5)
    assert got == want
    # a second evaluation against the same reference reuses its replica and radii on both devices
    C3 = AudioMetricsData(True); C3.add(torch.from_numpy(cand[:3000]))
    got3 = evaluate_devices(R2, C3, [0, 1], ("prdc",), nearest_k=5)
    C4 = AudioMetricsData(True); C4.add(torch.from_numpy(cand[:3000]))
    assert got3 == evaluate_containers(R, C4, ("prdc",), nearest_k=5)
    import test_gpu_api as api
    rng = np.random.default_rng(1)
    n = 5 * 16000
    audio_ref, audio_cand = rng.standard_normal((120, n, 2)), rng.standard_normal((100, n, 2))
    res = []
    for idx in ([0], [0, 1]):
        am = AudioMetrics(embedder=api.RandomEmbedder(), mix_function=api.mix_func, metrics=["fad", "kd", "prdc"],
                          device_indices=idx)
        am.add_reference(audio_ref)
        res.append(am.evaluate(audio_cand))
    assert res[0] == res[1]


def test_comm_two_devices(cuda_device):
    """amb_comm_* over two GPUs of this process, all ranks issued by one thread between group_begin /
    group_end: in-place allreduce (sum of int32 counts, max of int64, sum of fp64 moments) and the in-place
    allgather of radii slices — the exchanges of the sharded sweeps."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ctypes as C
    from audio_metrics_b200 import _lib
    L = _lib.lib()
    comm = C.c_void_p()
    devs = (C.c_int * 2)(0, 1)
    _lib.check(L.amb_comm_init(devs, 2, C.byref(comm)))
    try:
        assert L.amb_comm_size(comm) == 2
        d = [torch.device("cuda", i) for i in range(2)]
        g = torch.Generator().manual_seed(5)
        counts = [torch.randint(0, 9, (12345,), generator=g, dtype=torch.int32) for _ in d]
        mom = [torch.randn(4097, generator=g, dtype=torch.float64) for _ in d]
        mx = [torch.randint(0, 1 << 40, (3,), generator=g, dtype=torch.int64) for _ in d]
        radii = torch.rand(2 * 768, generator=g)
        c_dev = [c.to(x) for c, x in zip(counts, d)]
        m_dev = [c.to(x) for c, x in zip(mom, d)]
        x_dev = [c.to(x) for c, x in zip(mx, d)]
        r_dev = []
        for i, x in enumerate(d):
            buf = torch.full((2 * 768,), -1.0, device=x)
            buf[768 * i:768 * (i + 1)] = radii[768 * i:768 * (i + 1)].to(x)
            r_dev.append(buf)
        for x in d:
            torch.cuda.synchronize(x)
        st = [_lib.stream_ptr(x) for x in d]
        _lib.check(L.amb_comm_group_begin())
        for i in range(2):
            _lib.check(L.amb_comm_allreduce(comm, i, c_dev[i].data_ptr(), c_dev[i].data_ptr(), 12345, _lib.AMB_I32, _lib.AMB_SUM, st[i]))
            _lib.check(L.amb_comm_allreduce(comm, i, m_dev[i].data_ptr(), m_dev[i].data_ptr(), 4097, _lib.AMB_F64, _lib.AMB_SUM, st[i]))
            _lib.check(L.amb_comm_allreduce(comm, i, x_dev[i].data_ptr(), x_dev[i].data_ptr(), 3, _lib.AMB_I64, _lib.AMB_MAX, st[i]))
            _lib.check(L.amb_comm_allgather(comm, i, r_dev[i].data_ptr() + 4 * 768 * i, r_dev[i].data_ptr(), 768, _lib.AMB_F32, st[i]))
        _lib.check(L.amb_comm_group_end())
        for x in d:
            torch.cuda.synchronize(x)
        for i in range(2):
            assert torch.equal(c_dev[i].cpu(), counts[0] + counts[1])
            assert torch.equal(m_dev[i].cpu(), mom[0] + mom[1])          # two addends: order-independent
            assert torch.equal(x_dev[i].cpu(), torch.maximum(mx[0], mx[1]))
            assert torch.equal(r_dev[i].cpu(), radii)
    finally:
        _lib.check(L.amb_comm_destroy(comm))

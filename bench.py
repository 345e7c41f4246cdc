#!/usr/bin/env python
"""Benchmark of the embedding-set distance path (BASELINE.json metric/config).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one full pass of the hot path over one (reference, candidate) pair of
synthetic CLAP-512 embedding sets through the public API — AudioMetricsData.add of
both sets, then the fused evaluation AudioMetrics.evaluate runs: statistics + FAD,
KD (100 x 1000 subsets) and PRDC (k = 5) — at N = M = 200k, d = 512, the
configuration BASELINE.json quotes its metric on; it fits one B200 (0.4 GB per
set).  With N > 1 ranks the rows are sharded (fixed total work: strong scaling)
and the timed step includes the NCCL exchanges.  Prints ONE JSON line on rank 0:
`value` with the inputs resident in HBM, `e2e` with pinned host inputs (H2D inside
the timed region), `e2e_c_abi` through the host-buffer C entry point, `parity`
against the committed expectation, `phase_trace_ms` per phase of one step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# The CPU arm must get every host core: torch.distributed.run exports OMP_NUM_THREADS=1 to its
# workers, and numpy's BLAS reads that once, at import.  Nothing above this line imports numpy or
# torch, so setting the variables here (reference arm only) is early enough.
if "reference" in sys.argv[1:]:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

N_REF = int(os.environ.get("AMB_BENCH_N", 200_000))
N_CAND = int(os.environ.get("AMB_BENCH_M", N_REF))
DIM = int(os.environ.get("AMB_BENCH_D", 512))
K_NN = 5
KD_SUBSETS, KD_SUBSET_SIZE = 100, 1000
CPU_SAMPLE_N = int(os.environ.get("AMB_BENCH_CPU_N", 8192))

METRIC = "embedding pairs/sec for KD+PRDC (FAD+KD+PRDC step) at N=200k, d=512"
UNIT = "pairs/s"


def workload_pairs(n, m, subsets=KD_SUBSETS, subset=KD_SUBSET_SIZE):
    """Reference-equivalent pairs of one step: PRDC N^2 + M^2 + N M, KD 3 S m^2."""
    ms = subset if subset < min(n, m) else max(1, min(n, m) // 2)
    return float(n) * n + float(m) * m + float(n) * m + 3.0 * subsets * ms * ms


def config_dict(n_gpus):
    return {
        "workload": f"fad+kd+prdc on {N_REF} ref x {N_CAND} cand synthetic CLAP-{DIM} embeddings (unit-norm, "
                    f"rank-32 + noise), k={K_NN}, KD {KD_SUBSETS}x{KD_SUBSET_SIZE} subsets; BASELINE config "
                    "'PRDC k=5 + KD on 200k vs 200k CLAP-512' plus FAD",
        "n_ref": N_REF, "n_cand": N_CAND, "d": DIM, "k": K_NN,
        "kd_subsets": KD_SUBSETS, "kd_subset_size": KD_SUBSET_SIZE,
        "parallelism": f"row-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
        "l2": "inputs larger than L2 (2 x 410 MB fp32 + 2 x 410 MB packed operands vs 126 MB)",
    }


# ------------------------------------------------------------------ CPU reference arm
def cpu_sample_step(ref, cand):
    """The reference's algorithm (oracle port) on the bounded CPU sample: stats, FAD, KD, PRDC."""
    import numpy as np
    import oracle

    mr, cr, _ = oracle.batch_stats(ref)
    mc, cc, _ = oracle.batch_stats(cand)
    out = {"fad": oracle.frechet_from_stats(mc, cc, mr, cr)}
    out.update(oracle.kernel_distance(cand, ref))
    out.update(oracle.prdc(ref, cand, K_NN))
    return out


def cpu_sample_inputs():
    from audio_metrics_b200.synth import make_sets_torch

    ref, cand = make_sets_torch(CPU_SAMPLE_N, CPU_SAMPLE_N, DIM, seed=1234, device="cpu")
    return ref.numpy(), cand.numpy()


def cpu_sample_description():
    return (f"oracle port (numpy restatement of the reference) on {CPU_SAMPLE_N} x {CPU_SAMPLE_N} x {DIM} "
            f"embeddings: stats+FAD, KD {KD_SUBSETS}x{KD_SUBSET_SIZE}, PRDC k={K_NN}; the reference "
            "materialises N x M fp32 matrices and cannot run 200k (3 x 160 GB)")


def blas_threads():
    """Threads numpy's BLAS will actually use (threadpoolctl), for the cpu_baseline record."""
    try:
        import numpy  # noqa: F401  (loads the BLAS that is inspected)
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=cores)      # in case a BLAS was loaded before the environment was fixed
    except Exception:
        pass
    ref, cand = cpu_sample_inputs()
    pairs = workload_pairs(CPU_SAMPLE_N, CPU_SAMPLE_N)
    for _ in range(args.warmup):
        cpu_sample_step(ref, cand)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample_step(ref, cand)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = pairs / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "blas_threads": blas_threads(),
                         "kind": "port", "sample": cpu_sample_description()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------ clock sampler
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------- B200 arm
def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"tflops": float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1400.0))),
                "source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"}
    return {"tflops": 1400.0, "source": "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"}


def load_traffic():
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def parity_record(result, result_e2e):
    """Compare this run's result with the committed expectation for the same synthetic inputs
    (profiles/expected_result.json, written by an N=1 run of this script with AMB_BENCH_WRITE_EXPECTED=1):
    PRDC fractions must be identical (they are ratios of exact integer counts), FAD and KD agree to
    1e-9 relative (different reduction orders of the sharded moments / the 100 subset values).  The
    driver's 1/2/4/8-GPU runs therefore each prove they computed the same answer."""
    p = ROOT / "profiles" / "expected_result.json"
    key = f"{N_REF}x{N_CAND}x{DIM}"
    rec = {"expected_file": str(p.relative_to(ROOT)), "key": key, "e2e_equals_device_resident": result == result_e2e}
    if os.environ.get("AMB_BENCH_WRITE_EXPECTED"):
        allv = json.loads(p.read_text()) if p.exists() else {}
        allv[key] = result
        p.write_text(json.dumps(allv, indent=1, sort_keys=True))
    if not p.exists() or key not in json.loads(p.read_text()):
        rec.update(ok=None, note="no committed expectation for this size")
        return rec
    want = json.loads(p.read_text())[key]
    worst, ok = 0.0, True
    for k, v in want.items():
        got = result.get(k)
        if k in ("precision", "recall", "density", "coverage"):
            ok = ok and got == v
        else:
            rel = abs(got - v) / max(abs(v), 1e-300)
            worst = max(worst, rel)
            ok = ok and rel <= 1e-9
    rec.update(ok=bool(ok), max_rel_fad_kd=worst, prdc_identical=all(result.get(k) == want[k] for k in
                                                                       ("precision", "recall", "density", "coverage")))
    return rec


def run_b200_arm(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from audio_metrics_b200 import AudioMetricsData, _lib
    from audio_metrics_b200.dist import evaluate_containers, shard_rows
    from audio_metrics_b200.synth import make_sets_torch

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this framework has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    # synthetic embeddings, generated once; each rank keeps its own row shard
    ref_full, cand_full = make_sets_torch(N_REF, N_CAND, DIM, seed=1234, device=dev)
    r0, rn, _ = shard_rows(N_REF, world, rank)
    c0, cn, _ = shard_rows(N_CAND, world, rank)
    ref_shard = ref_full[r0:r0 + rn].clone()
    cand_shard = cand_full[c0:c0 + cn].clone()
    del ref_full, cand_full
    torch.cuda.empty_cache()
    pairs = workload_pairs(N_REF, N_CAND)

    def step(rs, cs, metrics=("fad", "kd", "prdc")):
        """One pass of the hot path through the public API: fresh containers (nothing cached from
        the previous step: statistics, packed operands, radii and counts are all recomputed),
        AudioMetricsData.add, then the fused evaluation AudioMetrics.evaluate runs."""
        R, Cn = AudioMetricsData(True, dev), AudioMetricsData(True, dev)
        R.add(rs)
        Cn.add(cs)
        return evaluate_containers(R, Cn, metrics, nearest_k=K_NN, kd_subsets=KD_SUBSETS,
                                   kd_subset_size=KD_SUBSET_SIZE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """CUDA-event time of `steps` calls, max over ranks, in ms per step."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), out

    # ---- device-resident timing (inputs already in HBM)
    for _ in range(args.warmup):
        result = step(ref_shard, cand_shard)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.amb_profile_enable(1)
    prof = (C.c_double * 4)()
    L.amb_profile_read(prof)   # clear
    launches0 = _lib.launch_count()
    if os.environ.get("AMB_BENCH_PROFILE") and rank == 0:
        import cProfile, pstats
        pr = cProfile.Profile(); pr.enable()
        ms_step, result = timed(lambda: step(ref_shard, cand_shard), args.steps)
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(28)
    else:
        ms_step, result = timed(lambda: step(ref_shard, cand_shard), args.steps)
    launches = _lib.launch_count() - launches0
    L.amb_profile_read(prof)
    L.amb_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned host arrays -> AudioMetricsData.add (H2D inside) -> evaluate -> python
    #      floats, every step.  add() copies on the library's copy stream and returns at once; the
    #      kernels that read a set wait for its copy on the device, and the evaluation touches the
    #      candidate only after all reference-only work is queued, so the second copy overlaps the
    #      first sweep.
    ref_host = ref_shard.cpu().pin_memory()
    cand_host = cand_shard.cpu().pin_memory()
    step(ref_host, cand_host)
    if os.environ.get("AMB_BENCH_TRACE"):
        for i in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter(); step(ref_host, cand_host); torch.cuda.synchronize()
            sys.stderr.write(f"[trace] e2e step {i}: {(time.perf_counter() - t0) * 1e3:.2f} ms\n")
        for i in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); step(ref_shard, cand_shard); torch.cuda.synchronize()
            sys.stderr.write(f"[trace] device step {i}: {(time.perf_counter() - t0) * 1e3:.2f} ms\n")
        for m in ("fad", "kd"):
            torch.cuda.synchronize(); t0 = time.perf_counter(); step(ref_shard, cand_shard, metrics=(m,)); torch.cuda.synchronize()
            sys.stderr.write(f"[trace] device {m} step: {(time.perf_counter() - t0) * 1e3:.2f} ms\n")
    ms_e2e, result_e2e = timed(lambda: step(ref_host, cand_host), max(1, min(args.steps, 3)))
    h2d = (ref_host.numel() + cand_host.numel()) * 4
    d2h = 8 * 8 + 8 * KD_SUBSETS   # result scalars + the 100 per-subset MMDs

    # ---- the same metrics through the host-buffer C ABI (amb_host_*: what a ctypes / cgo binding of
    #      the reference's metric functions calls), rank 0, one GPU: each call stages its own inputs
    e2e_cabi = None
    if world == 1 and not args.no_cabi:
        import numpy as np

        from audio_metrics_b200.dist import kd_subset_indices

        rh, ch = ref_host.numpy(), cand_host.numpy()
        idx = np.array(kd_subset_indices(N_CAND, N_REF, KD_SUBSET_SIZE, KD_SUBSETS, 1234), copy=True)
        d = DIM

        devs = (C.c_int * 1)(local)

        def cabi_step():
            out = (C.c_double * 7)()
            _lib.check(L.amb_host_evaluate(devs, 1, rh.ctypes.data, N_REF, ch.ctypes.data, N_CAND, d, 0, K_NN,
                                           idx.ctypes.data, KD_SUBSETS, KD_SUBSET_SIZE, 1, out))
            return dict(zip(("fad", "kernel_distance_mean", "kernel_distance_std", "precision", "recall", "density",
                             "coverage"), out))

        cabi_step()
        t0 = time.perf_counter()
        res_cabi = cabi_step()
        dt = time.perf_counter() - t0
        e2e_cabi = {"value": pairs / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": h2d,
                    "clock": "host wall clock around ONE amb_host_evaluate call on numpy arrays in pageable host "
                             "memory (device allocation, upload, every kernel, read-back and free inside)",
                    "result": res_cabi,
                    "agrees_with_python_path": all(
                        res_cabi[k] == result[k] if k in ("precision", "recall", "density", "coverage")
                        else abs(res_cabi[k] - result[k]) <= 1e-9 * abs(result[k]) for k in result)}

    # ---- where the time of one step goes on this rank and on the last one (CUDA events between phases)
    from audio_metrics_b200 import dist as amb_dist
    amb_dist.TRACE = []
    step(ref_shard, cand_shard)
    mine = amb_dist.trace_report()
    amb_dist.TRACE = None
    phase_trace = {"rank 0": {k: round(v, 3) for k, v in mine}}
    if world > 1:
        box = [None] * world
        dist.all_gather_object(box, [(k, round(v, 3)) for k, v in mine])
        phase_trace[f"rank {world - 1}"] = dict(box[-1])

    # ---- per-phase latency (FAD latency is part of the headline)
    phases = {}
    for name in ("fad", "kd", "prdc"):
        fn = lambda name=name: step(ref_shard, cand_shard, metrics=(name,))
        fn()
        phases[name + "_ms"], _ = timed(fn, 2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    eng_launches, eng_ms, eng_pairs, eng_flops = list(prof)
    achieved = (eng_pairs * 2.0 * DIM) / (eng_ms * 1e-3) * 1e-12 if eng_ms > 0 else 0.0
    executed = eng_flops / (eng_ms * 1e-3) * 1e-12 if eng_ms > 0 else 0.0
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
        "frac": achieved / peaks["tflops"], "traffic": load_traffic(),
        "kernel": "pair_engine2_kernel<TopkEpi|CountEpi> (tcgen05 kind::f16 cta_group::2, one MMA per product, A panels resident in smem) + pair_engine_kernel<KdEpi> (3-MMA split); with N > 1 ranks: rank 0's launches (the rank that also computes the Frechet distance sweeps 1 - N F / (W + F) of an even share)",
        "launches_timed": int(eng_launches), "avg_launch_ms": eng_ms / eng_launches if eng_launches else None,
        "kernel_share_of_step": eng_ms / (ms_step * args.steps) if ms_step > 0 else None,
        "executed_tflops": executed, "executed_frac": executed / peaks["tflops"], "peak_source": peaks["source"],
        "algorithmic_flops_per_pair": 2 * DIM,
    }

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0, N=1 runs only
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ref_s, cand_s = cpu_sample_inputs()
        t0 = time.perf_counter()
        cpu_sample_step(ref_s, cand_s)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": workload_pairs(CPU_SAMPLE_N, CPU_SAMPLE_N) / dt, "unit": UNIT, "cores": cores,
                        "blas_threads": blas_threads(), "kind": "port", "sample": cpu_sample_description(),
                        "seconds": dt}

    line = {
        "metric": METRIC, "value": pairs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f16 MMA filter (f32 accumulate) + f64 exact refine; KD f16x2 split; statistics i8 digit MMA (exact) / f64", "data": "synthetic",
        "config": config_dict(world),
        "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "phases_ms": phases, "phase_trace_ms": phase_trace, "result": result, "parity": parity_record(result, result_e2e),
        "pairs_per_step": pairs,
    }
    if e2e_cabi is not None:
        line["e2e_c_abi"] = e2e_cabi
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cabi", action="store_true", help="skip the amb_host_* end-to-end leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())

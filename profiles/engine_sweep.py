#!/usr/bin/env python
"""Engine micro-benchmark: time the single-pass / three-pass pair engine with the cheap
checksum epilogue (amb_debug_dot_matrix, ldc=0) and with the radii epilogue, as a function
of the B-ring depth (AMB_STAGES).  Separates operand-feed limits from epilogue limits.
usage: python profiles/engine_sweep.py [n] [d]"""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

def child(n, d):
    import torch
    from audio_metrics_b200 import _lib, AudioMetricsData
    from audio_metrics_b200.synth import make_sets_torch
    L = _lib.lib()
    ref, _ = make_sets_torch(n, 256, d, device="cuda")
    R = AudioMetricsData(True); R.embeddings = ref
    P = R.packed()
    out = torch.zeros(n + 1024, dtype=torch.float32, device="cuda")
    def t(fn, reps=3):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    dump = t(lambda: _lib.check(L.amb_debug_dot_matrix(0, None, P.data_ptr(), n, P.data_ptr(), n, d, out.data_ptr(), 0, 0, 0)))
    def radii():
        R.radii = {}
        R.get_radii(5)
    rad = t(radii)
    flops = 2.0 * n * n * d
    print(f"passes={os.environ.get('AMB_PASSES','1')} single_dump={os.environ.get('AMB_DEBUG_SINGLE','0')} "
          f"cta2={os.environ.get('AMB_CTA2','0')} stages={os.environ.get('AMB_STAGES','max')} grid={os.environ.get('AMB_GRID','sms')} n={n} d={d}: checksum {dump:.2f} ms ({flops/dump/1e9:.0f} TF alg), "
          f"radii {rad:.2f} ms ({flops/rad/1e9:.0f} TF alg)", flush=True)

if __name__ == "__main__":
    if os.environ.get("AMB_SWEEP_CHILD"):
        child(int(sys.argv[1]), int(sys.argv[2]))
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    if os.environ.get("AMB_SWEEP") == "grid":     # power-limit probe: fewer persistent CTAs than SMs
        configs = [dict(AMB_PASSES="1", AMB_DEBUG_SINGLE="1", AMB_GRID=g) for g in ("148", "111", "74", "37")]
    elif os.environ.get("AMB_SWEEP") == "cta2":   # single-CTA vs CTA-pair engine
        configs = [dict(AMB_PASSES="1", AMB_DEBUG_SINGLE="1", AMB_CTA2="0"),
                   dict(AMB_PASSES="1", AMB_DEBUG_SINGLE="2", AMB_CTA2="1")]
    else:
        configs = [dict(AMB_PASSES="3", AMB_DEBUG_SINGLE="0")]
        for st in ("2", "3", "4", "5", "6", "8"):
            configs.append(dict(AMB_PASSES="1", AMB_DEBUG_SINGLE="1", AMB_STAGES=st))
    for cfg in configs:
        env = dict(os.environ, AMB_SWEEP_CHILD="1", **cfg)
        subprocess.run([sys.executable, __file__, str(n), str(d)], env=env, check=False)

"""CPU tests of the C-ABI boundary: the library builds and loads without a GPU,
exports every symbol include/amb200.h declares, the ctypes table covers them, and
compute calls fail loudly (no fallback) when no CUDA device is present."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from audio_metrics_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    header = (ROOT / "include" / "amb200.h").read_text()
    return sorted(set(re.findall(r"\b(amb_[a-z0-9_]+)\s*\(", header)))


def test_library_exports_every_declared_symbol():
    handle = _lib.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(handle, n), f"libamb200.so does not export {n}"
    assert handle.amb_version() >= 100


def test_ctypes_table_matches_header():
    assert set(_lib.SIGNATURES) == set(_declared())


def test_size_queries_need_no_gpu():
    L = _lib.lib()
    # packed blob: rows padded to 256, k to 32, 4 B/element + 16 B/row (scale, norm, rho, exponent) + 4 B per 32 rows (chunk minimum norm), rounded to 256
    assert L.amb_packed_bytes(1000, 512) == 1024 * 512 * 4 + 1024 * 16 + 256
    assert L.amb_packed_bytes(1, 10) == 256 * 32 * 4 + 256 * 16 + 256
    assert L.amb_packed_bytes(-1, 4) == 0
    assert L.amb_cov_ws_bytes(100000, 512) > 0
    assert L.amb_frechet_ws_bytes(3, 512) >= 3 * 3 * 512 * 512 * 8
    assert L.amb_kd_ws_bytes(100, 1000, 512) > 2 * 100 * 1024 * 512 * 4
    assert L.amb_knn_ws_bytes(200000, 200000, 512, 5) < 64 << 20      # one column split at bench size
    assert L.amb_knn_ws_bytes(1000, 1000, 512, 40) == 0               # unsupported k
    assert L.amb_prdc_list_cap(200000, 200000) == 16 * 400000
    assert L.amb_prdc_ws_bytes(1000, 1000) > L.amb_prdc_list_cap(1000, 1000) * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="exercises the no-GPU failure path")
def test_compute_calls_fail_loudly_without_gpu():
    L = _lib.lib()
    x = np.zeros((8, 4), dtype=np.float32)
    mean, cov = np.zeros(4), np.zeros((4, 4))
    rc = L.amb_host_stats(0, x.ctypes.data, 0, 8, 4, mean.ctypes.data, cov.ctypes.data)
    assert rc == _lib.AMB_ERR_CUDA
    assert b"no CPU fallback" in L.amb_last_error() or b"CUDA" in L.amb_last_error()
    from audio_metrics_b200 import AudioMetricsData, AudioMetrics
    with pytest.raises(_lib.AmbError):
        AudioMetricsData().add(torch.zeros(4, 4))
    with pytest.raises(RuntimeError):
        AudioMetrics(embedder=object(), mix_function=lambda a, sr=None: a)
    out = (C.c_double * 4)()
    assert L.amb_host_prdc(0, x.ctypes.data, 8, x.ctypes.data, 8, 4, 0, 2, out) == _lib.AMB_ERR_CUDA


def test_argument_errors_are_value_errors():
    L = _lib.lib()
    assert L.amb_host_stats(0, None, 0, 8, 4, None, None) == _lib.AMB_ERR_ARG
    with pytest.raises(ValueError):
        _lib.check(_lib.AMB_ERR_ARG)


def test_header_is_c_and_links_from_c(tmp_path):
    """include/amb200.h compiles as C99 and a plain C program linked against libamb200.so runs the
    host-only entry points (tests/c/abi_smoke.c) — the binding a cgo / JNI / ctypes caller relies on."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = Path(__file__).resolve().parents[1]
    _lib.lib()
    lib = Path(_lib._LIB_PATH)
    exe = tmp_path / "abi_smoke"
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", str(root / "include"), str(root / "tests" / "c" / "abi_smoke.c"),
           "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr

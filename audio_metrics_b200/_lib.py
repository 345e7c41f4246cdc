"""ctypes binding of ``libamb200.so`` (the C ABI declared in ``include/amb200.h``).

The library is the product: if it is missing, or no CUDA device is usable, every
compute call raises — there is no CPU or PyTorch fallback path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

import os as _os

# (AMB200_LIB: an alternative build of the same ABI, for A/B timing of kernel variants on one box)
_LIB_PATH = Path(_os.environ.get("AMB200_LIB") or Path(__file__).resolve().parent / "libamb200.so")
_lib = None

AMB_F32, AMB_F64, AMB_I32, AMB_I64, AMB_U8 = 0, 1, 2, 3, 4
AMB_SUM, AMB_MAX = 0, 1
AMB_KERNEL_POLY, AMB_KERNEL_RBF = 0, 1
AMB_MMD_UNBIASED, AMB_MMD_BIASED, AMB_MMD_USTAT, AMB_MMD_UNIT_DIAGONAL = 0, 1, 2, 4
AMB_ERR_ARG, AMB_ERR_CUDA, AMB_ERR_WS, AMB_ERR_NUMERIC = -1, -2, -3, -4

_vp, _i, _ll, _sz, _dbl = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_double

# name -> (restype, argtypes); mirrors include/amb200.h one to one
SIGNATURES = {
    "amb_version": (_i, []),
    "amb_last_error": (C.c_char_p, []),
    "amb_launch_count": (_ll, []),
    "amb_set_option": (_i, [C.c_char_p, _i]),
    "amb_get_option": (_i, [C.c_char_p]),
    "amb_profile_enable": (_i, [_i]),
    "amb_profile_read": (_i, [_vp]),
    "amb_cov_ws_bytes": (_sz, [_ll, _i]),
    "amb_cov_accumulate": (_i, [_i, _vp, _vp, _i, _ll, _i, _ll, _vp, _vp, _vp, _sz]),
    "amb_cov_accumulate_masked": (_i, [_i, _vp, _vp, _i, _ll, _i, _ll, _vp, _i, _vp, _vp]),
    "amb_cov_finalize": (_i, [_i, _vp, _ll, _i, _vp, _vp, _vp, _vp]),
    "amb_stats_merge": (_i, [_i, _vp, _i, _ll, _vp, _vp, _ll, _vp, _vp, _vp]),
    "amb_frechet_ws_bytes": (_sz, [_i, _i]),
    "amb_frechet": (_i, [_i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz]),
    "amb_sym_eig_ws_bytes": (_sz, [_i]),
    "amb_sym_eig": (_i, [_i, _vp, _i, _vp, _vp, _vp, _vp, _sz]),
    "amb_pca_transform": (_i, [_i, _vp, _vp, _i, _ll, _i, _ll, _vp, _vp, _i, _vp]),
    "amb_packed_bytes": (_sz, [_ll, _i]),
    "amb_pack": (_i, [_i, _vp, _vp, _i, _ll, _i, _ll, _vp]),
    "amb_kd_ws_bytes": (_sz, [_i, _i, _i]),
    "amb_kd_subsets": (_i, [_i, _vp, _vp, _ll, _ll, _vp, _ll, _ll, _i, _i, _vp, _i, _i, _i, _dbl, _dbl, _i,
                            _dbl, _i, _vp, _vp, _vp, _sz]),
    "amb_knn_ws_bytes": (_sz, [_ll, _ll, _i, _i]),
    "amb_knn_radii": (_i, [_i, _vp, _vp, _i, _ll, _vp, _ll, _i, _ll, _ll, _i, _vp, _vp, _vp, _sz]),
    "amb_prdc_ws_bytes": (_sz, [_ll, _ll]),
    "amb_prdc_ws_bytes_cap": (_sz, [_ll, _ll, _ll]),
    "amb_prdc_ws_list_cap": (_ll, [_ll, _ll, _sz]),
    "amb_prdc_list_cap": (_ll, [_ll, _ll]),
    "amb_prdc_counts_exact": (_i, [_i, _vp, _vp, _ll, _ll, _vp, _vp, _ll, _ll, _vp, _i, _i, _ll, _ll, _vp, _vp, _vp]),
    "amb_prdc_counts": (_i, [_i, _vp, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp, _i, _i, _ll, _ll,
                             _vp, _vp, _vp, _vp, _vp, _sz]),
    "amb_prdc_reduce": (_i, [_i, _vp, _vp, _ll, _vp, _vp, _ll, _vp]),
    "amb_host_stats": (_i, [_i, _vp, _i, _ll, _i, _vp, _vp]),
    "amb_host_frechet": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp]),
    "amb_host_kd": (_i, [_i, _vp, _ll, _vp, _ll, _i, _i, _vp, _i, _i, _dbl, _dbl, _i, _vp, _vp]),
    "amb_host_knn_radii": (_i, [_i, _vp, _i, _ll, _i, _i, _vp]),
    "amb_host_prdc": (_i, [_i, _vp, _ll, _vp, _ll, _i, _i, _i, _vp]),
    "amb_host_evaluate": (_i, [_vp, _i, _vp, _ll, _vp, _ll, _i, _i, _i, _vp, _i, _i, _i, _vp]),
    "amb_comm_init": (_i, [_vp, _i, _vp]),
    "amb_comm_size": (_i, [_vp]),
    "amb_comm_device": (_i, [_vp, _i]),
    "amb_comm_group_begin": (_i, []),
    "amb_comm_group_end": (_i, []),
    "amb_comm_allreduce": (_i, [_vp, _i, _vp, _vp, _ll, _i, _i, _vp]),
    "amb_comm_allgather": (_i, [_vp, _i, _vp, _vp, _ll, _i, _vp]),
    "amb_comm_destroy": (_i, [_vp]),
    "amb_debug_dot_matrix": (_i, [_i, _vp, _vp, _ll, _vp, _ll, _i, _vp, _ll, C.c_uint, C.c_uint]),
}


class AmbError(RuntimeError):
    pass


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise AmbError(
                f"{_LIB_PATH} is missing: build it with `python -m audio_metrics_b200.build` "
                "(there is no fallback implementation)")
        handle = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc == 0:
        return
    msg = lib().amb_last_error().decode("utf-8", "replace")
    if rc == AMB_ERR_ARG:
        raise ValueError(msg)
    raise AmbError(f"amb200 error {rc}: {msg}")


def require_cuda(device=None) -> torch.device:
    """Resolve the CUDA device this call runs on; raise loudly when there is none."""
    if not torch.cuda.is_available():
        raise AmbError("audio_metrics_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise AmbError(f"audio_metrics_b200 computes on CUDA devices only, got {device}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return AMB_F32
    if t.dtype == torch.float64:
        return AMB_F64
    raise ValueError(f"embeddings must be float32 or float64, got {t.dtype}")


def as_device_matrix(x, device: torch.device) -> torch.Tensor:
    """[n, d] float32/float64 tensor on ``device`` with unit inner stride (other float
    types are promoted to float32, as torch would compute them)."""
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if x.ndim != 2:
        raise ValueError(f"expected a 2-D [n, d] array of embeddings, got shape {tuple(x.shape)}")
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float32)
    x = x.to(device, non_blocking=True)
    if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
        x = x.contiguous()
    return x


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def launch_count() -> int:
    return int(lib().amb_launch_count())


class options:
    """``with options(fad_method=1, engine_passes=3): ...`` — set process-wide library options
    (amb_set_option) for the duration of a block and restore the previous values.  For tests and
    A/B measurements; the options are not meant to be toggled around calls in production."""

    def __init__(self, **values):
        self.values = values
        self.saved = {}

    def __enter__(self):
        L = lib()
        for name, v in self.values.items():
            self.saved[name] = L.amb_get_option(name.encode())
            check(L.amb_set_option(name.encode(), int(v)))
        return self

    def __exit__(self, *exc):
        L = lib()
        for name, v in self.saved.items():
            L.amb_set_option(name.encode(), int(v))
        return False

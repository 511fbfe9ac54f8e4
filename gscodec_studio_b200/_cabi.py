"""ctypes binding of libgsb200.so -- the only door between Python and the CUDA kernels.

Plays the role of ``gsplat/cuda/_backend.py`` (reference: resolves the pybind11 module ``_C``,
_backend.py:81-139).  There is deliberately NO fallback: if the shared library is missing or a call
fails, a RuntimeError is raised (the reference raises through TORCH_CHECK, bindings.h:10-16).
Signatures are the ones declared in ``include/gsb200.h``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSB200_LIB", os.path.join(_HERE, "libgsb200.so"))  # env override: A/B builds
ABI_VERSION = 1

_P = c_void_p
_U = c_uint32
_F = c_float
_I = c_int
_L = c_int64

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/gsb200.h
SIGNATURES = {
    "gsb200_abi_version": [],
    "gsb200_last_error": [],
    "gsb200_launch_count": [],
    "gsb200_projection_fwd": [_U, _U, _P, _P, _P, _P, _P, _P, _U, _U, _F, _F, _F, _F, _I, _I, _P, _P, _P, _P, _P, _P],
    "gsb200_projection_bwd": [_U, _U, _P, _P, _P, _P, _P, _P, _U, _U, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                              _P, _P],
    "gsb200_projection_packed_count": [_U, _U, _P, _P, _P, _P, _P, _P, _U, _U, _F, _F, _F, _F, _I, _P, _P, _P],
    "gsb200_projection_packed_fill": [_U, _U, _P, _P, _P, _P, _P, _P, _U, _U, _F, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P,
                                      _P, _P],
    "gsb200_projection_packed_bwd": [_U, _U, _P, _P, _P, _P, _P, _P, _U, _U, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                     _P, _P, _P],
    "gsb200_projection_quant_fwd": [_U, _U, _P, _P, _P, _P, _P, _P, _P, _P, _P, _U, _U, _F, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_projection_quant_bwd": [_U, _U, _P, _P, _P, _P, _P, _P, _P, _P, _P, _U, _U, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                    _P, _P, _P, _P, _P],
    "gsb200_projection_2dgs_fwd": [_U, _U, _P, _P, _P, _P, _P, _U, _U, _F, _F, _F, _P, _P, _P, _P, _P, _P],
    "gsb200_projection_2dgs_bwd": [_U, _U] + [_P] * 15,
    "gsb200_rasterize_2dgs_fwd": [_U, _U, _L, _U, _P, _P, _P, _P, _P, _P, _P, _U, _U, _U, _U, _U] + [_P] * 10,
    "gsb200_rasterize_2dgs_bwd": [_U, _U, _L, _U, _P, _P, _P, _P, _P, _P, _P, _U, _U, _U, _U, _U] + [_P] * 19,
    "gsb200_indices_in_range_2dgs_count": [_U, _U, _U, _U, _L, _P, _P, _P, _U, _U, _U, _U, _U, _P, _P, _P, _P, _P],
    "gsb200_indices_in_range_2dgs_fill": [_U, _U, _U, _U, _L, _P, _P, _P, _U, _U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_compact_count": [_U, _U, _P, _P, _P, _P],
    "gsb200_compact_fill": [_U, _U, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_scatter_rows": [_U, _U, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_sh_fwd": [_U, _U, _U, _P, _P, _P, _P, _P],
    "gsb200_sh_bwd": [_U, _U, _U, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_sh_view_fwd": [_U, _U, _U, _U, _P, _P, _P, _P, _P, _P],
    "gsb200_sh_view_bwd": [_U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_sh_view_split_fwd": [_U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_sh_view_split_bwd": [_U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_sh_view_split_quant_fwd": [_U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_sh_view_split_quant_bwd": [_U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_sh_view_fwd_sharded": [_U, _U, _U, _U, _P, _P, _P, _U, _U, _P, _P, _P, _P],
    "gsb200_sh_view_bwd_sharded": [_U, _U, _U, _U, _P, _P, _P, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_shard_reduce_rows": [_U, _U, _U, _U, _P, _P, _P, _P],
    "gsb200_projection_owner_fwd": [_U, _U, _U, _P, _P, _P, _P, _P, _P, _U, _U, _F, _F, _F, _F, c_int, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_projection_owner_bwd": [_U, _U, _U, _P, _P, _P, _P, _P, _U, _U, _F, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_peer_push_small": [_U, _U, _U, _U, _P, _P, _P],
    "gsb200_shard_push_rows": [_U, _U, _U, _U, _U, c_uint64, _P, _P, _P, _U, _P, _P],
    "gsb200_sh_owner_fwd": [_U, _U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_sh_owner_bwd": [_U, _U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_camera_positions": [_U, _P, _P, _P],
    "gsb200_shard_publish": [_U, _U, _U, _P, _P, _P, _P, _P, _P, _U, _U, _P, _P],
    "gsb200_peer_barrier": [_U, _U, _P, _U, _P],
    "gsb200_shard_gather_geo": [_U, _U, _U, _P, _P, _P, _P],
    "gsb200_shard_push_geo": [_U, _U, _U, c_uint64, _U, _P, _P, _P, _P],
    "gsb200_shard_reduce_geo": [_U, _U, _U, c_uint64, _U, _P, _P, _P, _P],
    "gsb200_shard_reduce_rows_split": [_U, _U, _U, _U, _U, _P, _P, _P, _P, _P],
    "gsb200_shard_gather_rows": [_U, c_uint64, _P, _P, _P],
    "gsb200_isect_plan": [_U, _U, _U, _U, _P, _P],
    "gsb200_isect_prepare": [_U, _U, _U, _P, _P, _P, _P, _U, _U, _U, _P, _P, _P, _P, _P],
    "gsb200_isect_fill": [_U, _U, _U, _P, _P, _P, _P, _U, _U, _U, _L, _P, _P, _P, _P, _P, _P],
    "gsb200_isect_count_async": [_P, c_int64, _P, _P],
    "gsb200_isect_emit_unsorted": [_U, _U, _U, _P, _P, _P, _P, _U, _U, _U, _P, _P, _P, _P, _P],
    "gsb200_isect_offset_encode": [_L, _P, _U, _U, _U, _P, _P],
    "gsb200_blend_stream_record": [_U, _U],
    "gsb200_set_blend_mode": [_I],
    "gsb200_rasterize_fwd": [_U, _U, _L, _U, _P, _P, _P, _P, _P, _P, _U, _U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _P,
                             _P],
    "gsb200_grad_record": [_U],
    "gsb200_rasterize_bwd": [_U, _U, _L, _U, _P, _P, _P, _P, _P, _P, _U, _U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P, _P,
                             _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_indices_in_range_count": [_U, _U, _U, _U, _L, _P, _P, _P, _U, _U, _U, _U, _U, _P, _P, _P, _P, _P],
    "gsb200_indices_in_range_fill": [_U, _U, _U, _U, _L, _P, _P, _P, _U, _U, _U, _U, _U, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_fake_quantize_fwd": [_L, _P, _F, _F, _I, _I, _I, _P, c_uint64, c_uint64, _P, _P],
    "gsb200_fake_quantize_multi_fwd": [_U, _P, _P],
    "gsb200_fake_quantize_multi_bwd": [_U, _P, _P],
    "gsb200_fake_quantize_bwd": [_L, _P, _P, _F, _F, _I, _I, _P, _P, _P],
    "gsb200_grid_quantize": [_L, _U, _P, _P, _P, _I, _P, _P, _P],
    "gsb200_grid_dequantize": [_L, _U, _P, _P, _P, _P, _I, _P, _P],
    "gsb200_entropy_fwd": [_L, _U, _U, _P, _F, _P, _F, _P, _P],
    "gsb200_entropy_bwd": [_L, _U, _U, _P, _F, _P, _F, _P, _P, _P, _P],
    "gsb200_entropy_params_fwd": [_U, _U, _P, _P, _P, _P, _P],
    "gsb200_entropy_params_bwd": [_U, _U, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_grid_encode_fwd": [_U, _U, _U, _U, _P, _P, _I, _P, _P, _P, _P],
    "gsb200_grid_encode_bwd": [_U, _U, _U, _U, _P, _P, _I, _P, _P, _P, _P, _P, _P],
    "gsb200_selective_adam_update": [_P, _P, _P, _P, _P, _F, _F, _F, _F, _U, _U, _P],
    "gsb200_adam_multi": [_U, _P, _P],
    "gsb200_compute_relocation": [_U, _P, _P, _P, _P, _I, _P, _P, _P],
    "gsb200_quat_scale_to_covar_preci_fwd": [_U, _P, _P, _I, _P, _P, _P],
    "gsb200_quat_scale_to_covar_preci_bwd": [_U, _P, _P, _I, _P, _P, _P, _P, _P],
    "gsb200_proj_fwd": [_U, _U, _P, _P, _P, _U, _U, _I, _P, _P, _P],
    "gsb200_proj_bwd": [_U, _U, _P, _P, _P, _U, _U, _I, _P, _P, _P, _P, _P],
    "gsb200_world_to_cam_fwd": [_U, _U, _P, _P, _P, _P, _P, _P],
    "gsb200_world_to_cam_bwd": [_U, _U, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_stg_temporal_fwd": [_U, _F] + [_P] * 14,
    "gsb200_stg_temporal_bwd": [_U, _F] + [_P] * 19,
    "gsb200_photometric_workspace_bytes": [],
    "gsb200_photometric_loss_fwd": [_U, _U, _U, _U, _P, _P, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P],
    "gsb200_photometric_loss_bwd": [_U, _U, _U, _U, _P, _P, _F, _F, _I, _P, _P, _P, _P, _P, _P],
}


class AdamTensor(ctypes.Structure):
    """struct gsb200_adam_tensor (include/gsb200.h)."""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("n", c_int64), ("step_size", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
                ("inv_sqrt_bias2", c_float), ("one_minus_beta1", c_float), ("one_minus_beta2", c_float)]


class FqTensor(ctypes.Structure):
    """struct gsb200_fq_tensor (include/gsb200.h)."""
    _fields_ = [("n", c_int64), ("x", c_void_p), ("y", c_void_p), ("noise", c_void_p), ("v_y", c_void_p), ("v_x", c_void_p),
                ("seed", c_uint64), ("offset", c_uint64), ("lo", c_float), ("hi", c_float), ("bitwidth", c_int), ("mode", c_int),
                ("act", c_int)]


class QuantLoadDesc(ctypes.Structure):
    """struct gsb200_quant_load (include/gsb200.h): one attribute fake-quantised where a kernel loads it."""
    _fields_ = [("noise", c_void_p), ("seed", c_uint64), ("offset", c_uint64), ("lo", c_float), ("hi", c_float),
                ("bitwidth", ctypes.c_int32), ("act", ctypes.c_int32)]


FQ_MAX_TENSORS = 8  # GSB200_FQ_MAX_TENSORS
ADAM_MAX_TENSORS = 16  # GSB200_ADAM_MAX_TENSORS
_RESTYPES = {"gsb200_last_error": c_char_p, "gsb200_launch_count": c_uint64, "gsb200_grad_record": c_uint32,
             "gsb200_blend_stream_record": c_uint32,
             "gsb200_photometric_workspace_bytes": c_uint64}

_lib = None


def load_library() -> ctypes.CDLL:
    """Load libgsb200.so and bind every symbol of include/gsb200.h.  No compute, no GPU needed."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"gscodec_studio_b200: {LIB_PATH} is missing. Build it with "
            "`python -m gscodec_studio_b200._build` (needs nvcc); there is no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    if lib.gsb200_abi_version() != ABI_VERSION:
        raise RuntimeError("gscodec_studio_b200: libgsb200.so ABI version mismatch; rebuild the library")
    _lib = lib
    return lib


def launch_count() -> int:
    return int(load_library().gsb200_launch_count())


def _device_index(device) -> int:
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    idx = torch.device(device).index
    return torch.cuda.current_device() if idx is None else idx


def stream_ptr(device=None) -> c_void_p:
    """Raw cudaStream_t of torch's current stream on `device` (the raw query: building a torch.cuda.Stream object per
    C-ABI call cost ~9 us of host time, 150 us per training step)."""
    return c_void_p(torch._C._cuda_getCurrentRawStream(_device_index(device)))


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """`with torch.cuda.device(device)` that costs nothing when `device` already is the current device (the usual
    case: one process per GPU) -- the torch context manager is ~5 us per use, a dozen uses per step."""
    if _device_index(device) == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def ptr(t: Optional[torch.Tensor]) -> c_void_p:
    """Device pointer of a contiguous CUDA tensor (NULL for None)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("gscodec_studio_b200: tensor must be a CUDA tensor (the hot path has no CPU implementation)")
    if not t.is_contiguous():
        raise RuntimeError("gscodec_studio_b200: tensor must be contiguous")
    return c_void_p(t.data_ptr())


def aligned(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Contiguous and 16-byte aligned version of `t` (a clone only when needed)."""
    if t is None:
        return None
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


_stage_events = None  # name -> list of (start, end) CUDA events, when stage timing is on


_stage_only = None  # None = every entry point, else the set of names to time


def enable_stage_timing(on: bool, only=None) -> None:
    """Record a CUDA-event pair around every C-ABI call -- or only the entry points named in `only` -- on the stream
    the kernels are launched on.  Used by bench.py to attribute device time to stages inside the timed region."""
    global _stage_events, _stage_only
    _stage_events = {} if on else None
    _stage_only = None if only is None else frozenset(only)


def stage_timing_ms() -> dict:
    """Average device milliseconds per call for every entry point seen since enable_stage_timing(True)."""
    if not _stage_events:
        return {}
    torch.cuda.synchronize()
    return {name: sum(a.elapsed_time(b) for a, b in evs) / len(evs) for name, evs in _stage_events.items() if evs}


_NVTX = os.environ.get("GSB200_NVTX", "0") == "1"  # one NVTX range per C-ABI call (Nsight timelines)


def call(name: str, *args) -> None:
    lib = load_library()
    if _NVTX:
        torch.cuda.nvtx.range_push(name)
        try:
            rc = getattr(lib, name)(*args)
        finally:
            torch.cuda.nvtx.range_pop()
    elif _stage_events is not None and (_stage_only is None or name in _stage_only):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        rc = getattr(lib, name)(*args)
        end.record()
        _stage_events.setdefault(name, []).append((start, end))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.gsb200_last_error()
        raise RuntimeError(f"{name} failed ({rc}): {msg.decode() if msg else 'unknown error'}")

"""ctypes binding of ``libdvsloss.so`` (C ABI declared in ``include/dvsloss.h``).

There is deliberately no fallback: if the shared library has not been built
(``python __graft_entry__.py`` / ``python deep-visual-slam_b200/csrc/build.py``) importing
any operator raises, and every operator refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DVSLOSS_LIB selects another build of the same library (kernel experiments, fault-injection builds of the tests)
LIB_PATH = os.environ.get("DVSLOSS_LIB") or os.path.join(_HERE, "libdvsloss.so")

MAX_SCALES = 4
MAX_SOURCES = 4
DTYPE_F32, DTYPE_BF16, DTYPE_U8 = 0, 1, 2      # include/dvsloss.h: DVS_DTYPE_*


class DvsShape(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("N", C.c_int32), ("S", C.c_int32),
                ("dh", C.c_int32 * MAX_SCALES), ("dw", C.c_int32 * MAX_SCALES)]


class DvsParams(C.Structure):
    _fields_ = [("min_depth", C.c_float), ("max_depth", C.c_float), ("ssim_ratio", C.c_float),
                ("smoothness_ratio", C.c_float), ("eps", C.c_float), ("auto_mask", C.c_int32)]


class DvsError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None

_FP = C.POINTER(C.c_float)
_FPP = C.POINTER(_FP)
_U8PP = C.POINTER(C.POINTER(C.c_uint8))
_vp = C.c_void_p

# name -> argtypes; restype is always int unless listed in _RESTYPES
_SIGNATURES = {
    "dvs_version": [],
    "dvs_error_string": [C.c_int],
    "dvs_last_cuda_error": [],
    "dvs_set_profiling": [C.c_int],
    "dvs_set_noise_counter": [C.c_int, _vp],
    "dvs_last_tile_kernel_ms": [C.POINTER(C.c_float)],
    "dvs_loss_workspace_bytes": [C.POINTER(DvsShape), C.POINTER(C.c_size_t)],
    "dvs_photometric_forward": [C.POINTER(DvsShape), C.POINTER(DvsParams), _FPP, _vp, _FPP, _vp, _vp, _FPP, _FPP,
                                C.c_uint64, C.c_uint64, _vp, _vp, _U8PP, _FPP, _vp, _vp, _vp],
    "dvs_photometric_backward": [C.POINTER(DvsShape), _vp, _FPP, _vp, _FPP, _FPP, _vp],
    "dvs_photometric_forward_ex": [C.POINTER(DvsShape), C.POINTER(DvsParams), _FPP, C.c_int, _vp, _FPP, C.c_int, _vp, _vp, _FPP,
                                   _FPP, C.c_uint64, C.c_uint64, _vp, _vp, _U8PP, _FPP, _vp, _vp, _vp],
    "dvs_photometric_forward_pose": [C.POINTER(DvsShape), C.POINTER(DvsParams), _FPP, C.c_int, _vp, _FPP, C.c_int, _vp, _vp,
                                     _FPP, _FPP, C.POINTER(C.c_int32), _FPP, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _U8PP,
                                     _FPP, _vp, _vp, _vp],
    "dvs_photometric_backward_pose": [C.POINTER(DvsShape), _vp, _FPP, _vp, _FPP, C.c_int, _FPP, _FPP, _vp],
    "dvs_photometric_backward_ex": [C.POINTER(DvsShape), _vp, _FPP, _vp, _FPP, C.c_int, _FPP, _vp],
    "dvs_photometric_backward_recompute": [C.POINTER(DvsShape), C.POINTER(DvsParams), _FPP, _vp, _FPP, _vp, _vp, _FPP,
                                           _FPP, C.c_uint64, C.c_uint64, _vp, _FPP, _FPP, _vp, _vp],
    "dvs_disp_to_depth_fwd": [_vp, _vp, _vp, C.c_int64, C.c_float, C.c_float, _vp],
    "dvs_disp_to_depth_bwd": [_vp, _vp, _vp, _vp, C.c_int64, C.c_float, C.c_float, _vp],
    "dvs_upsample_bilinear_fwd": [_vp, _vp] + [C.c_int] * 6 + [_vp],
    "dvs_upsample_bilinear_bwd": [_vp, _vp] + [C.c_int] * 6 + [_vp],
    "dvs_backproject_fwd": [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp],
    "dvs_backproject_bwd": [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp],
    "dvs_project3d_fwd": [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, _vp],
    "dvs_project3d_bwd_workspace_bytes": [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)],
    "dvs_project3d_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, _vp, _vp],
    "dvs_grid_sample_border_fwd": [_vp, _vp, _vp] + [C.c_int] * 6 + [_vp],
    "dvs_grid_sample_border_bwd": [_vp, _vp, _vp, _vp] + [C.c_int] * 6 + [_vp],
    "dvs_ssim_fwd": [_vp, _vp, _vp] + [C.c_int] * 4 + [_vp],
    "dvs_ssim_bwd": [_vp, _vp, _vp, _vp, _vp] + [C.c_int] * 4 + [_vp],
    "dvs_reprojection_loss_fwd": [_vp, _vp, _vp] + [C.c_int] * 4 + [C.c_float, _vp],
    "dvs_reprojection_loss_bwd": [_vp, _vp, _vp, _vp] + [C.c_int] * 4 + [C.c_float, _vp],
    "dvs_smooth_loss_workspace_bytes": [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)],
    "dvs_smooth_loss_fwd": [_vp, _vp, _vp] + [C.c_int] * 4 + [_vp, _vp],
    "dvs_smooth_loss_bwd": [_vp, _vp, _vp, _vp] + [C.c_int] * 4 + [_vp],
    "dvs_pose_matrix_fwd": [_vp, _vp, _vp, C.c_int, C.c_int, _vp],
    "dvs_pose_matrix_bwd": [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp],
    "dvs_u8_to_f32": [_vp, _vp, C.c_int64, _vp],
    "dvs_disp_head_fwd": [_vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "dvs_elu_up2_cat_fwd": [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "dvs_glue_workspace_bytes": [C.c_int, C.POINTER(C.c_size_t)],
    "dvs_elu_up2_cat_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp],
    "dvs_bias_elu_fwd": [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "dvs_bias_elu_bwd": [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp],
    "dvs_disp_head_bwd_workspace_bytes": [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)],
    "dvs_disp_head_bwd": [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp],
    "dvs_silog_workspace_bytes": [C.c_int64, C.POINTER(C.c_size_t)],
    "dvs_silog_fwd": [_vp, _vp, _vp, C.c_int64, C.c_float, _vp, _vp, _vp],
    "dvs_silog_bwd": [_vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_float, _vp, _vp],
    "dvs_depth_to_pointcloud": [_vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp],
    "dvs_pack_net_inputs": [_vp, _vp, C.c_int, C.c_uint, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
    "dvs_gather_triplets_u8": [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp],
}
_RESTYPES = {"dvs_error_string": C.c_char_p}


def exported_symbols() -> Sequence[str]:
    return tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises DvsError when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DvsError(
                f"{LIB_PATH} not found: build the CUDA extension first (python __graft_entry__.py). "
                "There is no CPU or PyTorch fallback for this path.")
        handle = C.CDLL(LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        L = lib()
        msg = L.dvs_error_string(rc).decode()
        extra = ""
        if rc == -2:
            extra = f" (cudaError {L.dvs_last_cuda_error()})"
        raise DvsError(f"{what}: {msg}{extra}")


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise DvsError("dvsloss operators run on CUDA tensors only (no CPU fallback by design)")
        if t.dtype not in (torch.float32, torch.uint8, torch.bfloat16):
            raise DvsError(f"expected float32 (or bfloat16 disparities / uint8 images), got {t.dtype}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def fptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    """Host array of device float pointers."""
    arr = (_FP * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = C.cast(C.c_void_p(0 if t is None else t.data_ptr()), _FP)
    return arr


def u8ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    P = C.POINTER(C.c_uint8)
    arr = (P * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = C.cast(C.c_void_p(0 if t is None else t.data_ptr()), P)
    return arr


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def make_shape(B: int, H: int, W: int, N: int, disp_sizes: Sequence[Sequence[int]]) -> DvsShape:
    S = len(disp_sizes)
    if not (1 <= S <= MAX_SCALES and 1 <= N <= MAX_SOURCES):
        raise DvsError(f"unsupported number of scales/sources: S={S}, N={N}")
    dh = [int(h) for h, _ in disp_sizes] + [0] * (MAX_SCALES - S)
    dw = [int(w) for _, w in disp_sizes] + [0] * (MAX_SCALES - S)
    return DvsShape(B, H, W, N, S, (C.c_int32 * MAX_SCALES)(*dh), (C.c_int32 * MAX_SCALES)(*dw))

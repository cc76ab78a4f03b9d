"""ctypes driver for the CPU block emulator of the fused kernel (tests/emu/emu_fused.cpp).

Test infrastructure only: builds ``tests/emu/_build/libdvsemu.so`` with g++ from the very same
``dvs_fused_core.cuh`` the CUDA kernel is compiled from and runs it on numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "emu_fused.cpp")
CORE = os.path.join(ROOT, "deep-visual-slam_b200", "csrc", "dvs_fused_core.cuh")
OUT = os.path.join(HERE, "emu", "_build", "libdvsemu.so")


class DvsShape(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("N", C.c_int32), ("S", C.c_int32),
                ("dh", C.c_int32 * 4), ("dw", C.c_int32 * 4)]


class DvsParams(C.Structure):
    _fields_ = [("min_depth", C.c_float), ("max_depth", C.c_float), ("ssim_ratio", C.c_float),
                ("smoothness_ratio", C.c_float), ("eps", C.c_float), ("auto_mask", C.c_int32)]


def build(defines=()) -> str:
    """``defines``: extra -D macros (fault-injection builds); each set of macros gets its own library file."""
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    defines = tuple(defines) + tuple(d for d in os.environ.get("DVS_EMU_DEFINES", "").split(",") if d)   # e.g. DVS_TILE_ROWS=28
    out = OUT if not defines else OUT.replace(".so", "_" + "_".join(d.replace("=", "").replace(".", "").lower() for d in defines) + ".so")
    newest = max(os.path.getmtime(SRC), os.path.getmtime(CORE), os.path.getmtime(CORE.replace("dvs_fused_core", "dvs_pair_core")))
    if not os.path.exists(out) or os.path.getmtime(out) < newest:
        subprocess.check_call(["g++", "-O2", "-march=native", "-std=c++17", "-shared", "-fPIC", *[f"-D{d}" for d in defines],
                               "-o", out, SRC])
    return out


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _arr(ptrs, typ=C.c_float):
    P = C.POINTER(typ)
    return (P * len(ptrs))(*[p.ctypes.data_as(P) if p is not None else P() for p in ptrs])


def run(disps, target, sources, K, inv_K, Ts, noise=None, *, auto_mask=True, want_grad=True, grad_per_scale=None,
        min_depth=0.1, max_depth=10.0, ssim_ratio=0.85, smoothness_ratio=1e-3, seed=0, offset=0, defines=(),
        generic=False):
    """All inputs numpy float32 (contiguous).  Returns dict like oracle.closed_form.loss_and_grads.
    ``generic``: run a two-source problem through the generic tile code instead of the two-source specialisation."""
    lib = C.CDLL(build(defines))
    lib.emu_set_generic(int(bool(generic)))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    disps = [f32(d) for d in disps]
    target, K, inv_K = f32(target), f32(K), f32(inv_K)
    sources = [f32(s) for s in sources]
    Ts = [f32(t) for t in Ts]
    B, _, H, W = target.shape
    S, N = len(disps), len(sources)
    sh = DvsShape(B, H, W, N, S, (C.c_int32 * 4)(*([d.shape[2] for d in disps] + [0] * (4 - S))),
                  (C.c_int32 * 4)(*([d.shape[3] for d in disps] + [0] * (4 - S))))
    pr = DvsParams(min_depth, max_depth, ssim_ratio, smoothness_ratio, 1e-7, int(auto_mask))
    per_scale = np.zeros(S, np.float32)
    total = np.zeros(1, np.float32)
    sel = [np.full((B, H, W), 255, np.uint8) for _ in range(S)]
    ug = [np.full_like(d, np.nan) for d in disps] if want_grad else None
    uT = np.full(S * N * B * 16 + S * B, np.nan, np.float32) if want_grad else None
    nz = None
    if noise is not None and auto_mask:
        noise = [f32(n) for n in noise]
        nz = _arr(noise)
    rc = lib.emu_photometric_forward(C.byref(sh), C.byref(pr), _arr(disps), _fp(target), _arr(sources), _fp(K),
                                     _fp(inv_K), _arr(Ts), nz, C.c_uint64(seed), C.c_uint64(offset), _fp(per_scale),
                                     _fp(total), _arr(sel, C.c_uint8), _arr(ug) if want_grad else None,
                                     _fp(uT) if want_grad else None)
    assert rc == 0, rc
    out = {"per_scale": per_scale, "loss": float(total[0]), "sel": sel}
    if want_grad:
        g = np.asarray([1.0 / S] * S if grad_per_scale is None else grad_per_scale, np.float32)
        gd = [np.zeros_like(d) for d in disps]
        gT = [np.zeros((B, 4, 4), np.float32) for _ in range(N)]
        rc = lib.emu_photometric_backward(C.byref(sh), _fp(g), _arr(ug), _fp(uT), _arr(gd), _arr(gT))
        assert rc == 0, rc
        out["grad_disp"], out["grad_T"] = gd, gT
    return out

"""PyTorch autograd front-end of the fused view-synthesis loss (CUDA only).

``view_synthesis_loss`` replaces, in one call, what the reference does with ~2 160 ATen launches in
``MonodepthTrainer._generate_images_pred`` + ``_compute_losses`` (vo/learner_new.py:132-258): it
returns ``loss`` and the per-scale ``loss/s`` attached to the autograd graph, with gradients flowing to
the disparity maps and the 4x4 relative poses.  The forward launch already produces the gradients
of every ``loss/s`` (one pass over the images); backward only scales them by the upstream gradient.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple, Union

import torch

from . import _lib
from ._lib import DvsParams, check, fptr_array, lib, make_shape, ptr, require_cuda, stream_ptr, u8ptr_array


NoiseArg = Union[str, None, Sequence[torch.Tensor]]


def _prep(t: torch.Tensor) -> torch.Tensor:
    """The kernels compute in fp32 whatever the autocast state (the geometry is
    unusable in half precision); inputs are up-converted here."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _prep_disps(disps, N: int):
    """bf16 disparity maps (DepthNet under bf16 autocast, vo/train.py:177-181) are read by the two-source kernel as they
    are -- no ``.float()`` pass per scale, and the gradients go back as bf16; anything else is widened to fp32."""
    if N == 2 and all(d.dtype == torch.bfloat16 for d in disps):
        return [d.contiguous() for d in disps], _lib.DTYPE_BF16
    return [_prep(d) for d in disps], _lib.DTYPE_F32


def _prep_images(target, sources, N: int):
    """uint8 images (the loader's decoded frames before ToTensor, vo/dataset/common.py:39-46,77) are read by the
    two-source kernel as bytes (x/255 in-register, exact); for other source counts they are expanded on the device first."""
    imgs = [target] + list(sources)
    if all(t.dtype == torch.uint8 for t in imgs):
        if N == 2:
            return imgs[0].contiguous(), [t.contiguous() for t in imgs[1:]], _lib.DTYPE_U8
        from .ops import images_u8_to_f32
        imgs = [images_u8_to_f32(t.contiguous()) for t in imgs]
    return _prep(imgs[0]), [_prep(t) for t in imgs[1:]], _lib.DTYPE_F32


class _ViewSynthesisLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg: dict, target, sources, K, inv_K, noise, S: int, N: int, *diff):
        disps, disp_dtype = _prep_disps(diff[:S], N)
        pose_mode = cfg["pose_mode"]                  # diff = disps + Ts   or   disps + axisangles + translations
        Ts = [_prep(t) for t in diff[S:]]
        ctx.pose_shapes = [tuple(t.shape) for t in diff[S:]]
        K, inv_K = _prep(K), _prep(inv_K)
        target, sources, image_dtype = _prep_images(target, sources, N)
        require_cuda(target, K, inv_K, *disps, *Ts, *sources)
        dev = target.device
        B, C3, H, W = target.shape
        if C3 != 3:
            raise _lib.DvsError("images must be [B,3,H,W]")
        for d in disps:
            if d.dim() != 4 or d.shape[0] != B or d.shape[1] != 1:
                raise _lib.DvsError(f"disparity maps must be [B,1,h,w], got {tuple(d.shape)}")
        for s_ in sources:
            if s_.shape != target.shape:
                raise _lib.DvsError("source images must have the target's shape")
        for t in ([] if pose_mode else Ts) + [K, inv_K]:
            if tuple(t.shape) != (B, 4, 4):
                raise _lib.DvsError("K, inv_K and T must be [B,4,4]")
        if pose_mode:
            Ts = [t.reshape(B, 3) for t in Ts]
            if len(Ts) != 2 * N:
                raise _lib.DvsError("need one axis-angle and one translation per source frame")
        shape = make_shape(B, H, W, N, [d.shape[2:] for d in disps])
        params = DvsParams(cfg["min_depth"], cfg["max_depth"], cfg["ssim_ratio"], cfg["smoothness_ratio"], 1e-7,
                           int(bool(cfg["auto_mask"])))
        L = lib()
        nbytes = C.c_size_t(0)
        check(L.dvs_loss_workspace_bytes(C.byref(shape), C.byref(nbytes)), "dvs_loss_workspace_bytes")
        ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256

        want_grad = any(ctx.needs_input_grad[8:])
        per_scale = torch.empty(S, dtype=torch.float32, device=dev)
        total = torch.empty(1, dtype=torch.float32, device=dev)
        sel = None
        if cfg.get("return_selection"):
            sel = [torch.empty(B, H, W, dtype=torch.uint8, device=dev) for _ in range(S)]
        ugrad = uT = None
        if want_grad:
            ugrad = [torch.empty(d.shape, dtype=torch.float32, device=dev) for d in disps]
            uT = torch.empty(S * N * B * (6 if pose_mode else 16) + S * B, dtype=torch.float32, device=dev)
        noise_arr = None
        if noise is not None:
            noise = [_prep(n) for n in noise]
            for n in noise:
                if tuple(n.shape) != (B, N, H, W):
                    raise _lib.DvsError(f"noise tensors must be [B,N,H,W], got {tuple(n.shape)}")
            require_cuda(*noise)
            noise_arr = fptr_array(noise)
        with torch.cuda.device(dev):
            ctr = None
            if noise is None and cfg["auto_mask"]:
                ctr = _device_counter(dev)           # registered with the library: every in-kernel draw adds it to `offset`
                ctr.add_(1)                          # on the stream: captured graphs advance it on every replay
            if pose_mode:
                inv = (C.c_int32 * N)(*[int(bool(v)) for v in cfg["invert"]])
                rc = L.dvs_photometric_forward_pose(
                    C.byref(shape), C.byref(params), fptr_array(disps), disp_dtype, ptr(target), fptr_array(sources),
                    image_dtype, ptr(K), ptr(inv_K), fptr_array(Ts[:N]), fptr_array(Ts[N:]), inv, noise_arr,
                    C.c_uint64(cfg["seed"]), C.c_uint64(0), ptr(ctr), ptr(per_scale), ptr(total),
                    u8ptr_array(sel) if sel is not None else None, fptr_array(ugrad) if want_grad else None, ptr(uT), ws_ptr,
                    stream_ptr(dev))
            else:
                rc = L.dvs_photometric_forward_ex(
                    C.byref(shape), C.byref(params), fptr_array(disps), disp_dtype, ptr(target), fptr_array(sources), image_dtype,
                    ptr(K), ptr(inv_K), fptr_array(Ts), noise_arr, C.c_uint64(cfg["seed"]), C.c_uint64(0),
                    ptr(per_scale), ptr(total), u8ptr_array(sel) if sel is not None else None,
                    fptr_array(ugrad) if want_grad else None, ptr(uT), ws_ptr, stream_ptr(dev))
        check(rc, "dvs_photometric_forward")
        ctx.pose_mode = pose_mode
        ctx.shape, ctx.S, ctx.N, ctx.B = shape, S, N, B
        ctx.disp_dtype = disp_dtype
        ctx.ugrad, ctx.uT = ugrad, uT
        ctx.keep = (ws, noise, disps, Ts, sources, target, K, inv_K)   # keep inputs alive until the stream has used them
        outs = (total.view(()), per_scale)
        if sel is not None:
            for s_ in sel:
                ctx.mark_non_differentiable(s_)
            outs = outs + tuple(sel)
        return outs

    @staticmethod
    def backward(ctx, g_total, g_scale, *unused):
        if ctx.ugrad is None:
            raise _lib.DvsError("backward called but forward ran without gradient tracking")
        S, N, B = ctx.S, ctx.N, ctx.B
        dev = ctx.uT.device
        g = torch.zeros(S, dtype=torch.float32, device=dev)
        if g_scale is not None:
            g = g + g_scale.float()
        if g_total is not None:
            g = g + g_total.float() / S
        g = g.contiguous()
        gdt = torch.bfloat16 if ctx.disp_dtype == _lib.DTYPE_BF16 else torch.float32
        grad_disp = [torch.empty(u.shape, dtype=gdt, device=dev) for u in ctx.ugrad]
        if ctx.pose_mode:
            gp = [torch.empty(B, 3, dtype=torch.float32, device=dev) for _ in range(2 * N)]
            with torch.cuda.device(dev):
                rc = lib().dvs_photometric_backward_pose(C.byref(ctx.shape), ptr(g), fptr_array(ctx.ugrad), ptr(ctx.uT),
                                                         fptr_array(grad_disp), ctx.disp_dtype, fptr_array(gp[:N]),
                                                         fptr_array(gp[N:]), stream_ptr(dev))
            check(rc, "dvs_photometric_backward_pose")
            return (None,) * 8 + tuple(grad_disp) + tuple(t.view(sh) for t, sh in zip(gp, ctx.pose_shapes))
        grad_T = [torch.empty(B, 4, 4, dtype=torch.float32, device=dev) for _ in range(N)]
        with torch.cuda.device(dev):
            rc = lib().dvs_photometric_backward_ex(C.byref(ctx.shape), ptr(g), fptr_array(ctx.ugrad), ptr(ctx.uT),
                                                   fptr_array(grad_disp), ctx.disp_dtype, fptr_array(grad_T), stream_ptr(dev))
        check(rc, "dvs_photometric_backward_ex")
        return (None,) * 8 + tuple(grad_disp) + tuple(grad_T)


_counters = {}


def _device_counter(dev) -> torch.Tensor:
    """Per-device int64 step counter of the in-kernel noise generator (checkpoint it with ``noise_state`` / ``set_noise_state``)."""
    key = torch.device(dev).index or 0
    if key not in _counters:
        _counters[key] = torch.zeros(1, dtype=torch.int64, device=dev)
        check(lib().dvs_set_noise_counter(key, _counters[key].data_ptr()), "dvs_set_noise_counter")
    return _counters[key]


def noise_state(device=None) -> int:
    """Number of in-kernel noise draws made so far on ``device``; restore with ``set_noise_state`` (checkpoint / resume)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    return int(_device_counter(dev).item())


def set_noise_state(value: int, device=None) -> None:
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    _device_counter(dev).fill_(int(value))


def _default_seed(dev) -> int:
    """torch's seed mixed with the process rank and the device, so data-parallel ranks seeded alike draw different noise."""
    rank = 0
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank = dist.get_rank()
    except Exception:
        rank = 0
    return (torch.initial_seed() ^ ((rank + 1) * 0x9E3779B97F4A7C15) ^ ((torch.device(dev).index or 0) << 48)) & (2 ** 64 - 1)


def view_synthesis_loss(disps: Sequence[torch.Tensor], target: torch.Tensor, sources: Sequence[torch.Tensor],
                        K: torch.Tensor, inv_K: torch.Tensor, Ts: Optional[Sequence[torch.Tensor]] = None, *,
                        noise: NoiseArg = "kernel", min_depth: float = 0.1, max_depth: float = 10.0,
                        ssim_ratio: float = 0.85, smoothness_ratio: float = 1e-3, auto_mask: bool = True,
                        return_selection: bool = False, seed: Optional[int] = None,
                        axisangles: Optional[Sequence[torch.Tensor]] = None,
                        translations: Optional[Sequence[torch.Tensor]] = None,
                        inverts: Optional[Sequence[bool]] = None) -> Tuple[torch.Tensor, ...]:
    """Fused Monodepth2 view-synthesis loss over S=len(disps) scales and N=len(sources) source frames.

    disps[s] [B,1,h_s,w_s] sigmoid disparities (outputs[("disp", s)]; fp32, or bf16 as emitted under autocast -- read directly
    by the two-source kernel, gradients returned as bf16), target/sources [B,3,H,W] in [0,1] (fp32, or uint8 frames: x/255
    is formed inside the kernel),
    K/inv_K [B,4,4] scale-0 intrinsics, Ts[i] [B,4,4] cam_T_cam of source i -- or, instead of ``Ts``, the pose parameters
    themselves: ``axisangles[i]``, ``translations[i]`` ([B,3], [B,1,3] or [B,1,1,3], what PoseNet returns) and ``inverts[i]``
    (vo/learner_new.py:124-127: True for frames before the target).  Then transformation_from_parameters and its backward run
    inside the loss's own launches and the gradients arrive at the pose parameters directly.

    noise: "kernel" -> automask tie-break noise from an in-kernel counter-based generator (fast);
           "torch"  -> ``torch.randn([B,N,H,W])`` per scale, i.e. exactly the draws (shape, order, device)
                       of vo/learner_new.py:228, so a seeded run consumes the RNG like the reference;
           a list of S tensors [B,N,H,W] -> those draws;  None -> no noise.
    Returns (loss, per_scale[S]) and, with return_selection, S uint8 maps [B,H,W] holding the argmin
    channel over [identity_0..N-1, reproj_0..N-1] (``identity_selection/s`` of the reference is ``sel >= N``).
    """
    S, N = len(disps), len(sources)
    pose_mode = Ts is None
    if pose_mode:
        if axisangles is None or translations is None or inverts is None or not (len(axisangles) == len(translations) == len(inverts) == N):
            raise _lib.DvsError("give either Ts or axisangles + translations + inverts, one per source frame")
        pose_args = list(axisangles) + list(translations)
    else:
        if len(Ts) != N:
            raise _lib.DvsError("need one pose per source frame")
        pose_args = list(Ts)
    B, _, H, W = target.shape
    noise_t = None
    if auto_mask:
        if isinstance(noise, str):
            if noise == "torch":
                noise_t = [torch.randn([B, N, H, W], device=target.device) for _ in range(S)]
            elif noise != "kernel":
                raise _lib.DvsError(f"unknown noise mode {noise!r}")
        elif noise is None:
            noise_t = [torch.zeros(B, N, H, W, device=target.device)] * S
        else:
            noise_t = list(noise)
            if len(noise_t) != S:
                raise _lib.DvsError("need one noise tensor per scale")
    cfg = dict(min_depth=float(min_depth), max_depth=float(max_depth), ssim_ratio=float(ssim_ratio),
               smoothness_ratio=float(smoothness_ratio), auto_mask=bool(auto_mask),
               return_selection=bool(return_selection),
               seed=int(_default_seed(target.device) if seed is None else seed) & (2 ** 64 - 1),
               pose_mode=pose_mode, invert=list(inverts) if pose_mode else None)
    return _ViewSynthesisLoss.apply(cfg, target, list(sources), K, inv_K, noise_t, S, N, *disps, *pose_args)

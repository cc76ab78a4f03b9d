"""Granular operators with the reference's call surface, as autograd functions over libdvsloss.so.

One sm_100a kernel per direction (csrc/dvs_ops.cu).  These are what ``model/layers.py`` /
``vo/learner_func.py`` of this repo expose as ``disp_to_depth``, ``BackprojectDepth``, ``Project3D``, ``SSIM``,
``get_smooth_loss``, ``compute_reprojection_loss`` and ``transformation_from_parameters``
(reference: vo/learner_func.py:16-207 == model/layers.py:16-248, vo/learner_new.py:60-74).
CUDA fp32 only; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch

from ._lib import DvsError, check, lib, ptr, require_cuda, stream_ptr


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ws(nbytes: int, dev) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 4), dtype=torch.uint8, device=dev)


# --------------------------------------------------------------------------------------------- disp_to_depth
class _DispToDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, min_depth, max_depth):
        disp = _f32(disp)
        require_cuda(disp)
        scaled, depth = torch.empty_like(disp), torch.empty_like(disp)
        with torch.cuda.device(disp.device):
            check(lib().dvs_disp_to_depth_fwd(ptr(disp), ptr(scaled), ptr(depth), disp.numel(), min_depth, max_depth,
                                              stream_ptr(disp.device)), "dvs_disp_to_depth_fwd")
        ctx.save_for_backward(depth)
        ctx.lim = (min_depth, max_depth)
        return scaled, depth

    @staticmethod
    def backward(ctx, g_scaled, g_depth):
        (depth,) = ctx.saved_tensors
        gs = _f32(g_scaled) if g_scaled is not None else None
        gd = _f32(g_depth) if g_depth is not None else None
        out = torch.empty_like(depth)
        with torch.cuda.device(depth.device):
            check(lib().dvs_disp_to_depth_bwd(ptr(depth), ptr(gs), ptr(gd), ptr(out), depth.numel(), ctx.lim[0], ctx.lim[1],
                                              stream_ptr(depth.device)), "dvs_disp_to_depth_bwd")
        return out, None, None


def disp_to_depth(disp: torch.Tensor, min_depth: float, max_depth: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """vo/learner_func.py:16-26: sigmoid disparity -> (scaled disparity, depth)."""
    return _DispToDepth.apply(disp, float(min_depth), float(max_depth))


# --------------------------------------------------------------------------------------------- up-sampling
class _Upsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W):
        x = _f32(x)
        require_cuda(x)
        B, Cc, h, w = x.shape
        out = torch.empty(B, Cc, H, W, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib().dvs_upsample_bilinear_fwd(ptr(x), ptr(out), B, Cc, h, w, H, W, stream_ptr(x.device)),
                  "dvs_upsample_bilinear_fwd")
        ctx.dims = (B, Cc, h, w, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        B, Cc, h, w, H, W = ctx.dims
        g = _f32(g)
        gi = torch.empty(B, Cc, h, w, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(lib().dvs_upsample_bilinear_bwd(ptr(g), ptr(gi), B, Cc, h, w, H, W, stream_ptr(g.device)),
                  "dvs_upsample_bilinear_bwd")
        return gi, None, None


def upsample_bilinear(x: torch.Tensor, size) -> torch.Tensor:
    """F.interpolate(x, size, mode="bilinear", align_corners=False) (vo/learner_new.py:136-140)."""
    return _Upsample.apply(x, int(size[0]), int(size[1]))


# --------------------------------------------------------------------------------------------- BackprojectDepth
class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, inv_K):
        depth, inv_K = _f32(depth), _f32(inv_K)
        require_cuda(depth, inv_K)
        B, _, H, W = depth.shape
        out = torch.empty(B, 4, H * W, dtype=torch.float32, device=depth.device)
        with torch.cuda.device(depth.device):
            check(lib().dvs_backproject_fwd(ptr(depth), ptr(inv_K), ptr(out), B, H, W, stream_ptr(depth.device)),
                  "dvs_backproject_fwd")
        ctx.save_for_backward(inv_K)
        ctx.dims = (B, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        (inv_K,) = ctx.saved_tensors
        B, H, W = ctx.dims
        g = _f32(g)
        gd = torch.empty(B, 1, H, W, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(lib().dvs_backproject_bwd(ptr(g), ptr(inv_K), ptr(gd), B, H, W, stream_ptr(g.device)),
                  "dvs_backproject_bwd")
        return gd, None


def backproject(depth: torch.Tensor, inv_K: torch.Tensor) -> torch.Tensor:
    return _Backproject.apply(depth, inv_K)


# --------------------------------------------------------------------------------------------- Project3D
class _Project3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, K, T, H, W, eps):
        points, K, T = _f32(points), _f32(K), _f32(T)
        require_cuda(points, K, T)
        B = points.shape[0]
        if tuple(points.shape) != (B, 4, H * W):
            raise DvsError(f"points must be [B,4,{H * W}], got {tuple(points.shape)}")
        pix = torch.empty(B, H, W, 2, dtype=torch.float32, device=points.device)
        with torch.cuda.device(points.device):
            check(lib().dvs_project3d_fwd(ptr(points), ptr(K), ptr(T), ptr(pix), B, H, W, eps, stream_ptr(points.device)),
                  "dvs_project3d_fwd")
        ctx.save_for_backward(points, K, T)
        ctx.dims = (B, H, W, eps)
        return pix

    @staticmethod
    def backward(ctx, g):
        points, K, T = ctx.saved_tensors
        B, H, W, eps = ctx.dims
        g = _f32(g)
        dev = g.device
        gp = torch.empty_like(points) if ctx.needs_input_grad[0] else None
        gT = torch.empty(B, 4, 4, dtype=torch.float32, device=dev) if ctx.needs_input_grad[2] else None
        if gp is None and gT is None:
            return None, None, None, None, None, None
        n = C.c_size_t(0)
        check(lib().dvs_project3d_bwd_workspace_bytes(B, H, W, C.byref(n)), "dvs_project3d_bwd_workspace_bytes")
        ws = _ws(n.value, dev)
        with torch.cuda.device(dev):
            check(lib().dvs_project3d_bwd(ptr(g), ptr(points), ptr(K), ptr(T), ptr(gp), ptr(gT), B, H, W, eps, ptr(ws),
                                          stream_ptr(dev)), "dvs_project3d_bwd")
        return gp, None, gT, None, None, None


def project3d(points, K, T, H: int, W: int, eps: float = 1e-7) -> torch.Tensor:
    return _Project3D.apply(points, K, T, int(H), int(W), float(eps))


# --------------------------------------------------------------------------------------------- grid_sample
class _GridSampleBorder(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, grid):
        src, grid = _f32(src), _f32(grid)
        require_cuda(src, grid)
        B, Cc, H, W = src.shape
        _, Ho, Wo, two = grid.shape
        if two != 2 or grid.shape[0] != B:
            raise DvsError("grid must be [B,Ho,Wo,2]")
        out = torch.empty(B, Cc, Ho, Wo, dtype=torch.float32, device=src.device)
        with torch.cuda.device(src.device):
            check(lib().dvs_grid_sample_border_fwd(ptr(src), ptr(grid), ptr(out), B, Cc, H, W, Ho, Wo,
                                                   stream_ptr(src.device)), "dvs_grid_sample_border_fwd")
        ctx.save_for_backward(src, grid)
        return out

    @staticmethod
    def backward(ctx, g):
        src, grid = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise DvsError("grid_sample_border: gradient w.r.t. the image is not provided (images are data on this path)")
        g = _f32(g)
        B, Cc, H, W = src.shape
        _, Ho, Wo, _ = grid.shape
        gg = torch.empty_like(grid)
        with torch.cuda.device(g.device):
            check(lib().dvs_grid_sample_border_bwd(ptr(g), ptr(src), ptr(grid), ptr(gg), B, Cc, H, W, Ho, Wo,
                                                   stream_ptr(g.device)), "dvs_grid_sample_border_bwd")
        return None, gg


def grid_sample_border(src: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """F.grid_sample(src, grid, padding_mode="border", align_corners=True), bilinear (vo/learner_new.py:165-170)."""
    return _GridSampleBorder.apply(src, grid)


# --------------------------------------------------------------------------------------------- SSIM
class _SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = _f32(x), _f32(y)
        require_cuda(x, y)
        if x.shape != y.shape or x.dim() != 4:
            raise DvsError("SSIM expects two [B,C,H,W] tensors of the same shape")
        B, Cc, H, W = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(lib().dvs_ssim_fwd(ptr(x), ptr(y), ptr(out), B, Cc, H, W, stream_ptr(x.device)), "dvs_ssim_fwd")
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        g = _f32(g)
        B, Cc, H, W = x.shape
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        if gx is None and gy is None:
            return None, None
        with torch.cuda.device(g.device):
            check(lib().dvs_ssim_bwd(ptr(g), ptr(x), ptr(y), ptr(gx), ptr(gy), B, Cc, H, W, stream_ptr(g.device)),
                  "dvs_ssim_bwd")
        return gx, gy


def ssim(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return _SSIM.apply(x, y)


# --------------------------------------------------------------------------------------------- reprojection loss
class _Reproj(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, ssim_ratio):
        pred, target = _f32(pred), _f32(target)
        require_cuda(pred, target)
        if pred.shape != target.shape or pred.dim() != 4:
            raise DvsError("compute_reprojection_loss expects two [B,C,H,W] tensors of the same shape")
        B, Cc, H, W = pred.shape
        out = torch.empty(B, 1, H, W, dtype=torch.float32, device=pred.device)
        with torch.cuda.device(pred.device):
            check(lib().dvs_reprojection_loss_fwd(ptr(pred), ptr(target), ptr(out), B, Cc, H, W, ssim_ratio,
                                                  stream_ptr(pred.device)), "dvs_reprojection_loss_fwd")
        ctx.save_for_backward(pred, target)
        ctx.w = ssim_ratio
        return out

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            raise DvsError("compute_reprojection_loss: the target image is data on this path (no gradient)")
        g = _f32(g)
        B, Cc, H, W = pred.shape
        gp = torch.empty_like(pred)
        with torch.cuda.device(g.device):
            check(lib().dvs_reprojection_loss_bwd(ptr(g), ptr(pred), ptr(target), ptr(gp), B, Cc, H, W, ctx.w,
                                                  stream_ptr(g.device)), "dvs_reprojection_loss_bwd")
        return gp, None, None


def compute_reprojection_loss(pred: torch.Tensor, target: torch.Tensor, ssim_ratio: float = 0.85) -> torch.Tensor:
    """vo/learner_new.py:60-74: ssim_ratio * mean_c SSIM + (1 - ssim_ratio) * mean_c |target - pred| -> [B,1,H,W]."""
    return _Reproj.apply(pred, target, float(ssim_ratio))


# --------------------------------------------------------------------------------------------- smoothness
class _Smooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, img):
        disp, img = _f32(disp), _f32(img)
        require_cuda(disp, img)
        B, one, H, W = disp.shape
        if one != 1 or img.shape[0] != B or tuple(img.shape[2:]) != (H, W):
            raise DvsError("get_smooth_loss expects disp [B,1,H,W] and img [B,C,H,W]")
        Cc = img.shape[1]
        n = C.c_size_t(0)
        check(lib().dvs_smooth_loss_workspace_bytes(B, H, W, C.byref(n)), "dvs_smooth_loss_workspace_bytes")
        ws = _ws(n.value, disp.device)
        out = torch.empty(1, dtype=torch.float32, device=disp.device)
        with torch.cuda.device(disp.device):
            check(lib().dvs_smooth_loss_fwd(ptr(disp), ptr(img), ptr(out), B, Cc, H, W, ptr(ws), stream_ptr(disp.device)),
                  "dvs_smooth_loss_fwd")
        ctx.save_for_backward(disp, img)
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        disp, img = ctx.saved_tensors
        B, _, H, W = disp.shape
        g = _f32(g).reshape(1)
        gd = torch.empty_like(disp)
        with torch.cuda.device(g.device):
            check(lib().dvs_smooth_loss_bwd(ptr(g), ptr(disp), ptr(img), ptr(gd), B, img.shape[1], H, W,
                                            stream_ptr(g.device)), "dvs_smooth_loss_bwd")
        return gd, None


def get_smooth_loss(disp: torch.Tensor, img: torch.Tensor) -> torch.Tensor:
    """vo/learner_func.py:161-174."""
    return _Smooth.apply(disp, img)


# --------------------------------------------------------------------------------------------- pose matrix
class _PoseMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axisangle, translation, invert):
        shp = axisangle.shape
        aa, tr = _f32(axisangle).reshape(-1, 3), _f32(translation).reshape(-1, 3)
        require_cuda(aa, tr)
        B = aa.shape[0]
        M = torch.empty(B, 4, 4, dtype=torch.float32, device=aa.device)
        with torch.cuda.device(aa.device):
            check(lib().dvs_pose_matrix_fwd(ptr(aa), ptr(tr), ptr(M), B, int(invert), stream_ptr(aa.device)),
                  "dvs_pose_matrix_fwd")
        ctx.save_for_backward(aa, tr)
        ctx.meta = (shp, translation.shape, int(invert))
        return M

    @staticmethod
    def backward(ctx, g):
        aa, tr = ctx.saved_tensors
        shp_a, shp_t, invert = ctx.meta
        g = _f32(g)
        ga, gt = torch.empty_like(aa), torch.empty_like(tr)
        with torch.cuda.device(g.device):
            check(lib().dvs_pose_matrix_bwd(ptr(g), ptr(aa), ptr(tr), ptr(ga), ptr(gt), aa.shape[0], invert,
                                            stream_ptr(g.device)), "dvs_pose_matrix_bwd")
        return ga.view(shp_a), gt.view(shp_t), None


def transformation_from_parameters(axisangle: torch.Tensor, translation: torch.Tensor, invert: bool = False) -> torch.Tensor:
    """vo/learner_func.py:29-46: axisangle, translation [B,1,3] -> 4x4 (inverted when `invert`)."""
    return _PoseMatrix.apply(axisangle, translation, bool(invert))


def images_u8_to_f32(src: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """uint8 image batch -> float32 in [0,1], exactly ``ToTensor`` (x / 255) of the reference's loader
    (vo/dataset/common.py:77), on the device; ``out`` may be a preallocated float32 tensor of the same shape."""
    if not src.is_cuda or src.dtype != torch.uint8:
        raise DvsError("images_u8_to_f32 needs a CUDA uint8 tensor (no CPU fallback by design)")
    src = src.contiguous()
    if out is None:
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    if out.shape != src.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != src.device:
        raise DvsError("out must be a contiguous float32 tensor of the input's shape on the same device")
    with torch.cuda.device(src.device):
        check(lib().dvs_u8_to_f32(src.data_ptr(), out.data_ptr(), src.numel(), stream_ptr(src.device)), "dvs_u8_to_f32")
    return out



# ---------------------------------------------------------------------------------------------------- DepthNet disparity head
class _DispHead(torch.autograd.Function):
    """ReflectionPad2d(1) + Conv2d(C, 1, 3) + Sigmoid in one kernel (model/depthnet.py:57-58,87-88; model/layers.py:120-136)."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_dtype, padded):
        from ._lib import DTYPE_BF16, DTYPE_F32
        if not x.is_cuda:
            raise DvsError("disp_head runs on CUDA tensors only (no CPU fallback by design)")
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        pad = 1 if padded else 0
        B, Cc = x.shape[:2]
        H, W = x.shape[2] - 2 * pad, x.shape[3] - 2 * pad
        xc = x.contiguous(memory_format=torch.channels_last)            # no copy when the decoder already runs channels-last
        w = weight.detach().float().contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        disp = torch.empty(B, 1, H, W, dtype=out_dtype, device=x.device)
        code = lambda dt: DTYPE_BF16 if dt == torch.bfloat16 else DTYPE_F32
        with torch.cuda.device(x.device):
            check(lib().dvs_disp_head_fwd(xc.data_ptr(), code(xc.dtype), pad, w.data_ptr(), ptr(b), disp.data_ptr(), code(out_dtype),
                                          B, Cc, H, W, stream_ptr(x.device)), "dvs_disp_head_fwd")
        ctx.save_for_backward(xc, w, disp)
        ctx.has_bias = bias is not None
        ctx.wdtype = weight.dtype
        ctx.pad = pad
        return disp

    @staticmethod
    def backward(ctx, g):
        import ctypes as C
        from ._lib import DTYPE_BF16, DTYPE_F32
        xc, w, disp = ctx.saved_tensors
        pad = ctx.pad
        B, Cc = xc.shape[:2]
        H, W = xc.shape[2] - 2 * pad, xc.shape[3] - 2 * pad
        g = g.to(disp.dtype).contiguous()
        # with a padded activation the gradient has the padded layout too; the kernel leaves the ring alone: zero it here
        gx = (torch.zeros_like if pad else torch.empty_like)(xc, memory_format=torch.channels_last)
        gw = torch.empty_like(w)
        gb = torch.empty(1, dtype=torch.float32, device=xc.device) if ctx.has_bias else None
        n = C.c_size_t(0)
        L = lib()
        check(L.dvs_disp_head_bwd_workspace_bytes(B, Cc, H, W, C.byref(n)), "dvs_disp_head_bwd_workspace_bytes")
        ws = torch.empty(n.value + 256, dtype=torch.uint8, device=xc.device)
        code = lambda dt: DTYPE_BF16 if dt == torch.bfloat16 else DTYPE_F32
        with torch.cuda.device(xc.device):
            check(L.dvs_disp_head_bwd(g.data_ptr(), disp.data_ptr(), code(disp.dtype), xc.data_ptr(), code(xc.dtype), pad, w.data_ptr(),
                                      gx.data_ptr(), gw.data_ptr(), ptr(gb), B, Cc, H, W, (ws.data_ptr() + 255) // 256 * 256,
                                      stream_ptr(xc.device)), "dvs_disp_head_bwd")
        return gx, gw.to(ctx.wdtype), (gb.to(ctx.wdtype) if gb is not None else None), None, None


def disp_head(x: torch.Tensor, weight: torch.Tensor, bias, out_dtype=None, padded: bool = False) -> torch.Tensor:
    """sigmoid(conv3x3(reflection_pad(x))) with one output channel: x [B,C,H,W] (fp32 / bf16; channels-last memory is read in
    place), weight [1,C,3,3], bias [1] or None -> disparity [B,1,H,W] in ``out_dtype`` (default: x's dtype, which is what the
    stock modules produce under autocast).  ``padded``: x is [B,C,H+2,W+2] and already carries its reflected ring
    (``bias_elu(..., pad=True)``); the result is the same."""
    if weight.dim() != 4 or weight.shape[0] != 1 or tuple(weight.shape[2:]) != (3, 3) or weight.shape[1] != x.shape[1]:
        raise DvsError(f"disp_head needs a [1,C,3,3] weight for a [B,C,H,W] input, got {tuple(weight.shape)} / {tuple(x.shape)}")
    if padded and x.shape[1] not in (8, 16, 32, 64, 128):
        raise DvsError("disp_head(padded=True) needs 8, 16, 32, 64 or 128 channels")
    if out_dtype is None:
        out_dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
    return _DispHead.apply(x, weight, bias, out_dtype, padded)


def disp_head_supported(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dim() == 4 and x.shape[1] % 8 == 0 and 8 <= x.shape[1] <= 128 and x.shape[2] >= 3 and x.shape[3] >= 3


# ---------------------------------------------------------------------------------------------------- DepthNet decoder glue
def _glue_ws(Cc: int, device) -> torch.Tensor:
    import ctypes as C
    n = C.c_size_t(0)
    check(lib().dvs_glue_workspace_bytes(Cc, C.byref(n)), "dvs_glue_workspace_bytes")
    return torch.empty(n.value + 256, dtype=torch.uint8, device=device)


def _al256(t: torch.Tensor) -> int:
    return (t.data_ptr() + 255) // 256 * 256


class _EluUp2Cat(torch.autograd.Function):
    """cat([nearest_up2(ELU(x + bias)), skip], 1) in one pass, channels-last (model/depthnet.py:77-84, model/layers.py:106-117,196-199)."""

    @staticmethod
    def forward(ctx, x, skip, bias, pad):
        from ._lib import DTYPE_BF16, DTYPE_F32
        B, C1, h, w = x.shape
        C2 = 0 if skip is None else skip.shape[1]
        pad = 1 if pad else 0
        xc = x.contiguous(memory_format=torch.channels_last)
        sc = None if skip is None else skip.contiguous(memory_format=torch.channels_last)
        bc = None if bias is None else bias.detach().float().contiguous()
        out = torch.empty(B, C1 + C2, 2 * h + 2 * pad, 2 * w + 2 * pad, dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        code = DTYPE_BF16 if x.dtype == torch.bfloat16 else DTYPE_F32
        with torch.cuda.device(x.device):
            check(lib().dvs_elu_up2_cat_fwd(xc.data_ptr(), 0 if sc is None else sc.data_ptr(), ptr(bc), out.data_ptr(), code, B, C1, C2,
                                            h, w, pad, stream_ptr(x.device)), "dvs_elu_up2_cat_fwd")
        ctx.save_for_backward(xc, bc)
        ctx.C2, ctx.code, ctx.pad = C2, code, pad
        ctx.bdtype = None if bias is None else bias.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        xc, bc = ctx.saved_tensors
        B, C1, h, w = xc.shape
        C2 = ctx.C2
        g = g.to(xc.dtype).contiguous(memory_format=torch.channels_last)
        gx = torch.empty_like(xc, memory_format=torch.channels_last)
        gs = torch.empty(B, C2, 2 * h, 2 * w, dtype=xc.dtype, device=xc.device, memory_format=torch.channels_last) if C2 else None
        want_b = bc is not None and ctx.needs_input_grad[2]
        gb = torch.empty(C1, dtype=torch.float32, device=xc.device) if want_b else None
        ws = _glue_ws(C1, xc.device) if want_b else None
        with torch.cuda.device(xc.device):
            check(lib().dvs_elu_up2_cat_bwd(xc.data_ptr(), g.data_ptr(), ptr(bc), gx.data_ptr(), 0 if gs is None else gs.data_ptr(),
                                            ptr(gb), ctx.code, B, C1, C2, h, w, ctx.pad, 0 if ws is None else _al256(ws),
                                            stream_ptr(xc.device)), "dvs_elu_up2_cat_bwd")
        return gx, gs, (gb.to(ctx.bdtype) if gb is not None else None), None


def _glue_channels_ok(x: torch.Tensor) -> bool:
    per = 8 if x.dtype == torch.bfloat16 else 4
    v = x.shape[1] // per
    return x.shape[1] % per == 0 and 1 <= v <= 256 and (v & (v - 1)) == 0          # the bias-gradient reduction wants a power of two


def elu_up2_cat_supported(x: torch.Tensor, skip=None) -> bool:
    if not (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16) and _glue_channels_ok(x)):
        return False
    per = 8 if x.dtype == torch.bfloat16 else 4
    if skip is None:
        return True
    return (skip.is_cuda and skip.dtype == x.dtype and skip.shape[1] % per == 0 and skip.shape[0] == x.shape[0]
            and tuple(skip.shape[2:]) == (2 * x.shape[2], 2 * x.shape[3]))


def elu_up2_cat(x: torch.Tensor, skip=None, bias=None, pad: bool = False) -> torch.Tensor:
    """``torch.cat([F.interpolate(F.elu(x + bias), scale_factor=2, mode="nearest"), skip], 1)`` (``skip``, ``bias`` optional) as
    one kernel each way; x [B,C1,h,w] is the decoder convolution's output BEFORE its bias and ELU (run the convolution without
    bias and hand the bias here: one pass over the activation less each way), skip [B,C2,2h,2w] the encoder feature.
    Channels-last memory is read in place; the result is channels-last.  ``pad``: the result is ``ReflectionPad2d(1)`` of the
    above ([B,C1+C2,2h+2,2w+2]) -- the next Conv3x3 then runs as a plain un-padded convolution on it, and this op's backward
    folds the ring's gradients back."""
    if not elu_up2_cat_supported(x, skip):
        raise DvsError("elu_up2_cat needs CUDA fp32 / bf16 tensors of one dtype, channels / 4 (fp32) or / 8 (bf16) a power of two, "
                       "skip of twice the spatial size")
    return _EluUp2Cat.apply(x, skip, bias, pad)


class _BiasElu(torch.autograd.Function):
    """ELU(x + bias) in one pass (ConvBlock = Conv3x3 + ELU, model/layers.py:106-117; the Conv2d bias of :131 folded in)."""

    @staticmethod
    def forward(ctx, x, bias, pad):
        from ._lib import DTYPE_BF16, DTYPE_F32
        B, Cc, H, W = x.shape
        pad = 1 if pad else 0
        xc = x.contiguous(memory_format=torch.channels_last)
        bc = None if bias is None else bias.detach().float().contiguous()
        y = torch.empty(B, Cc, H + 2 * pad, W + 2 * pad, dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        code = DTYPE_BF16 if x.dtype == torch.bfloat16 else DTYPE_F32
        with torch.cuda.device(x.device):
            check(lib().dvs_bias_elu_fwd(xc.data_ptr(), ptr(bc), y.data_ptr(), code, B, Cc, H, W, pad, stream_ptr(x.device)),
                  "dvs_bias_elu_fwd")
        ctx.save_for_backward(y)
        ctx.code, ctx.pad = code, pad
        ctx.bdtype = None if bias is None else bias.dtype
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        pad = ctx.pad
        B, Cc = y.shape[:2]
        H, W = y.shape[2] - 2 * pad, y.shape[3] - 2 * pad
        g = g.to(y.dtype).contiguous(memory_format=torch.channels_last)
        gx = torch.empty(B, Cc, H, W, dtype=y.dtype, device=y.device, memory_format=torch.channels_last)
        want_b = ctx.bdtype is not None and ctx.needs_input_grad[1]
        gb = torch.empty(Cc, dtype=torch.float32, device=y.device) if want_b else None
        ws = _glue_ws(Cc, y.device) if want_b else None
        with torch.cuda.device(y.device):
            check(lib().dvs_bias_elu_bwd(y.data_ptr(), g.data_ptr(), gx.data_ptr(), ptr(gb), ctx.code, B, Cc, H, W, pad,
                                         0 if ws is None else _al256(ws), stream_ptr(y.device)), "dvs_bias_elu_bwd")
        return gx, (gb.to(ctx.bdtype) if gb is not None else None), None


def bias_elu_supported(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16) and _glue_channels_ok(x)


def bias_elu(x: torch.Tensor, bias=None, pad: bool = False) -> torch.Tensor:
    """``F.elu(x + bias.view(1, -1, 1, 1))`` as one kernel each way (the backward also yields the bias gradient); x is the
    convolution's output without its bias, channels-last memory is read in place.  ``pad``: the result is
    ``ReflectionPad2d(1)`` of the above ([B,C,H+2,W+2]), see ``elu_up2_cat``."""
    if not bias_elu_supported(x):
        raise DvsError("bias_elu needs a CUDA fp32 / bf16 [B,C,H,W] tensor with C / 4 (fp32) or C / 8 (bf16) a power of two")
    if pad and (x.shape[2] < 3 or x.shape[3] < 3):
        raise DvsError("a reflected ring needs at least 3 x 3 pixels")
    return _BiasElu.apply(x, bias, pad)


# ---------------------------------------------------------------------------------------------------- network inputs
def pack_net_inputs_supported(target: torch.Tensor, sources) -> bool:
    ok = lambda t: (t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] == 3 and t.is_contiguous()
                    and t.data_ptr() % 16 == 0)                 # 128-bit loads: a batch slice at an odd offset takes the stock path
    return (ok(target) and 1 <= len(sources) <= 4 and all(ok(s) and s.shape == target.shape for s in sources)
            and (target.shape[2] * target.shape[3]) % 8 == 0)


def pack_net_inputs(target: torch.Tensor, sources, src_first, dtype=torch.float32, normalize: bool = True):
    """One pass from the sample dict's NCHW fp32 frames to the first convolutions' inputs: the channels-last target
    [B,3,H,W] and per source the pose pair [B,6,H,W] (``cat([source, target], 1)`` where ``src_first[k]`` else
    ``cat([target, source], 1)``, vo/learner_new.py:110-123), normalised ``(x - 0.45) / 0.225`` like the encoders do
    (model/resnet_encoder.py) and cast to ``dtype`` (bf16: what autocast hands conv1).  Same bits as the stock sequence."""
    import ctypes as C
    from ._lib import DTYPE_BF16, DTYPE_F32
    if not pack_net_inputs_supported(target, sources):
        raise DvsError("pack_net_inputs needs contiguous CUDA fp32 [B,3,H,W] frames with H*W a multiple of 8, 1..4 sources")
    if dtype not in (torch.float32, torch.bfloat16):
        raise DvsError("pack_net_inputs writes float32 or bfloat16")
    B, _, H, W = target.shape
    dev = target.device
    out_t = torch.empty(B, 3, H, W, dtype=dtype, device=dev, memory_format=torch.channels_last)
    pairs = [torch.empty(B, 6, H, W, dtype=dtype, device=dev, memory_format=torch.channels_last) for _ in sources]
    mask = sum(1 << k for k, f in enumerate(src_first) if f)
    VP = C.c_void_p
    parr = (VP * len(pairs))(*[VP(p.data_ptr()) for p in pairs])
    from ._lib import fptr_array
    with torch.cuda.device(dev):
        check(lib().dvs_pack_net_inputs(ptr(target), fptr_array(list(sources)), len(sources), mask, int(bool(normalize)), out_t.data_ptr(),
                                        parr, DTYPE_BF16 if dtype == torch.bfloat16 else DTYPE_F32, B, H, W, stream_ptr(dev)),
              "dvs_pack_net_inputs")
    return out_t, pairs


# ---------------------------------------------------------------------------------------------------- supervised depth
class _Silog(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, valid, variance_focus):
        pred, target = _f32(pred), _f32(target)
        require_cuda(pred, target)
        if pred.shape != target.shape or valid.shape != pred.shape:
            raise DvsError("silog_loss expects prediction, target and valid_mask of one shape")
        v8 = valid.to(torch.uint8).contiguous()
        n = pred.numel()
        nb = C.c_size_t(0)
        check(lib().dvs_silog_workspace_bytes(n, C.byref(nb)), "dvs_silog_workspace_bytes")
        ws = _ws(nb.value, pred.device)
        stats = torch.empty(4, dtype=torch.float32, device=pred.device)
        with torch.cuda.device(pred.device):
            check(lib().dvs_silog_fwd(ptr(pred), ptr(target), v8.data_ptr(), n, float(variance_focus), ptr(stats), ptr(ws),
                                      stream_ptr(pred.device)), "dvs_silog_fwd")
        ctx.save_for_backward(pred, target, v8, stats)
        ctx.vf = float(variance_focus)
        return stats[0]

    @staticmethod
    def backward(ctx, g):
        pred, target, v8, stats = ctx.saved_tensors
        g = _f32(g).reshape(1)
        gp = torch.empty_like(pred)
        with torch.cuda.device(pred.device):
            check(lib().dvs_silog_bwd(ptr(g), ptr(stats), ptr(pred), ptr(target), v8.data_ptr(), pred.numel(), ctx.vf, ptr(gp),
                                      stream_ptr(pred.device)), "dvs_silog_bwd")
        return gp, None, None, None


def silog_loss(prediction: torch.Tensor, target: torch.Tensor, valid_mask: torch.Tensor, variance_focus: float = 0.85) -> torch.Tensor:
    """depth/depth_learner.py:75-95: scale-invariant log loss over the valid pixels (one reduction kernel, double sums)."""
    return _Silog.apply(prediction, target, valid_mask, variance_focus)


def depth_to_pointcloud(depth: torch.Tensor, pose: torch.Tensor, intrinsic: torch.Tensor):
    """vo/eval_traj.py:85-128 on the device: depth [H,W], camera-to-world pose [4,4], K [3,3] or [4,4] ->
    world points [N,3] of the pixels with depth > 0 (row-major pixel order; the caller sub-samples as it likes)."""
    depth = _f32(depth)
    require_cuda(depth)
    H, W = depth.shape
    dev = depth.device
    K = torch.eye(4, dtype=torch.float64, device=dev)
    K[:intrinsic.shape[0], :intrinsic.shape[1]] = intrinsic.to(dev, torch.float64)
    inv_K = torch.linalg.inv(K).float().contiguous()
    T = pose.to(dev, torch.float32).contiguous()
    pts = torch.empty(H * W, 3, dtype=torch.float32, device=dev)
    valid = torch.empty(H * W, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().dvs_depth_to_pointcloud(ptr(depth), ptr(inv_K), ptr(T), ptr(pts), valid.data_ptr(), H, W, stream_ptr(dev)),
              "dvs_depth_to_pointcloud")
    return pts[valid.bool()]

"""Geometry and photometric primitives of the VO learner, backed by the sm_100a kernels of libdvsloss.so.

Same call surface as the reference's ``model/layers.py:16-268`` (twin: ``vo/learner_func.py:16-207``):
``disp_to_depth``, ``transformation_from_parameters``, ``get_translation_matrix``, ``rot_from_axisangle``,
``BackprojectDepth``, ``Project3D``, ``SSIM``, ``get_smooth_loss``, ``upsample``, ``ConvBlock``, ``Conv3x3``,
``compute_depth_errors``.  The first eight are autograd functions over hand-written CUDA kernels
(``dvsloss.ops``); the decoder blocks stay stock PyTorch convolutions, as in the reference.
The training hot path does not call these one by one -- ``vo/learner_new.py`` uses the fused
``dvsloss.view_synthesis_loss`` -- they serve the callers that use single primitives
(vo/predict.py:83, vo/eval_traj.py, the ros2 node) and keep old learner code working.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from dvsloss import ops as _ops  # noqa: E402

disp_to_depth = _ops.disp_to_depth
transformation_from_parameters = _ops.transformation_from_parameters
get_smooth_loss = _ops.get_smooth_loss


def get_translation_matrix(translation_vector: torch.Tensor) -> torch.Tensor:
    """[B,1,3] (or [B,3]) translation -> [B,4,4] homogeneous matrix (reference: model/layers.py:49-62)."""
    t = translation_vector.reshape(-1, 3)
    M = torch.eye(4, dtype=t.dtype, device=t.device).repeat(t.shape[0], 1, 1)
    M[:, :3, 3] = t
    return M


def rot_from_axisangle(vec: torch.Tensor) -> torch.Tensor:
    """[B,1,3] axis-angle -> [B,4,4] rotation (reference: model/layers.py:65-104); the pose kernel with t = 0."""
    return _ops.transformation_from_parameters(vec, torch.zeros_like(vec), False)


class BackprojectDepth(nn.Module):
    """depth [B,1,H,W], inv_K [B,4,4] -> homogeneous camera points [B,4,H*W] (reference: model/layers.py:139-168).
    The reference keeps a [B,3,HW] pixel grid and a ones tensor as frozen parameters (79 MB at B=16, 640x480);
    the kernel generates pixel coordinates from the thread index instead, so this module holds no state."""

    def __init__(self, batch_size: int, height: int, width: int):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width

    def forward(self, depth: torch.Tensor, inv_K: torch.Tensor) -> torch.Tensor:
        if tuple(depth.shape[2:]) != (self.height, self.width):
            raise ValueError(f"depth must be [B,1,{self.height},{self.width}], got {tuple(depth.shape)}")
        return _ops.backproject(depth, inv_K)


class Project3D(nn.Module):
    """points [B,4,HW], K, T [B,4,4] -> sampling grid [B,H,W,2] in [-1,1] (reference: model/layers.py:171-193)."""

    def __init__(self, batch_size: int, height: int, width: int, eps: float = 1e-7):
        super().__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points: torch.Tensor, K: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
        return _ops.project3d(points, K, T, self.height, self.width, self.eps)


class SSIM(nn.Module):
    """3x3 mean-filter SSIM loss map clamp((1 - SSIM)/2, 0, 1) with reflection padding
    (reference: model/layers.py:218-248)."""

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return _ops.ssim(x, y)


def upsample(x: torch.Tensor) -> torch.Tensor:
    """Nearest-neighbour x2 (DepthNet decoder, reference: model/layers.py:196-199) -- stock PyTorch."""
    return F.interpolate(x, scale_factor=2, mode="nearest")


def _conv_bwd(g, inp, w, padding, mask):
    return torch.ops.aten.convolution_backward(g, inp, w, None, [1, 1], list(padding), [1, 1], False, [0, 0], 1, [mask[0], mask[1], False])


class _Conv3x3Reflect(torch.autograd.Function):
    """``conv2d(ReflectionPad2d(1)(x), weight, bias)`` without the padded copy of ``x`` -- forward AND backward.

    The padded convolution equals the ZERO-padded one (cuDNN's native mode, no copy) plus what the reflected ring contributes
    to the outermost output rows / columns, which are four thin stock convolutions on single rows / columns of ``x``:
        top row    += conv(reflect_w(x[row 1]),   w[ky = 0])      bottom row   += conv(reflect_w(x[row H-2]), w[ky = 2])
        left col   += conv(x[col 1],   w[kx = 0], zero pad in y)   right col    += conv(x[col W-2], w[kx = 2], zero pad in y)
    (the corner taps belong to the row strips).  The backward is written out the same way -- one full convolution backward
    plus four thin ones added into single rows / columns of the gradients -- because autograd's own derivative of the
    slice-and-add formulation materialises a zero-filled full-size tensor and a full-size add for every strip (14 % of the
    CUDA time of a training step, profiles/r02 train profile)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        if torch.is_autocast_enabled(x.device.type):            # what autocast would feed the convolution
            dt = torch.get_autocast_dtype(x.device.type)
            x = x.to(dt)
        w = weight.to(x.dtype)
        b = bias.to(x.dtype) if bias is not None else None
        H, W = x.shape[2:]
        rw = lambda t: F.pad(t, (1, 1, 0, 0), mode="reflect")
        out = F.conv2d(x, w, b, padding=1)
        out[:, :, 0:1] += F.conv2d(rw(x[:, :, 1:2]), w[:, :, 0:1])
        out[:, :, H - 1:H] += F.conv2d(rw(x[:, :, H - 2:H - 1]), w[:, :, 2:3])
        out[:, :, :, 0:1] += F.conv2d(x[:, :, :, 1:2], w[:, :, :, 0:1], padding=(1, 0))
        out[:, :, :, W - 1:W] += F.conv2d(x[:, :, :, W - 2:W - 1], w[:, :, :, 2:3], padding=(1, 0))
        ctx.save_for_backward(x, w)
        ctx.wdtype = weight.dtype
        ctx.bdtype = bias.dtype if bias is not None else None
        return out

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        H, W = x.shape[2:]
        need_x, need_w, need_b = ctx.needs_input_grad
        g = g.to(x.dtype)
        mask = (need_x, need_w)
        gx, gw, _ = _conv_bwd(g, x, w, (1, 1), mask)
        rw = lambda t: F.pad(t, (1, 1, 0, 0), mode="reflect")
        # row strips: reflect_w's adjoint folds the two pad columns back onto columns 1 and W-2
        for row, ky, orow in ((1, 0, 0), (H - 2, 2, H - 1)):
            gi, gk, _ = _conv_bwd(g[:, :, orow:orow + 1], rw(x[:, :, row:row + 1]), w[:, :, ky:ky + 1], (0, 0), mask)
            if need_x:
                gi[..., 2] += gi[..., 0]
                gi[..., W - 1] += gi[..., W + 1]
                gx[:, :, row:row + 1] += gi[..., 1:W + 1]
            if need_w:
                gw[:, :, ky:ky + 1] += gk
        for col, kx, ocol in ((1, 0, 0), (W - 2, 2, W - 1)):
            gi, gk, _ = _conv_bwd(g[:, :, :, ocol:ocol + 1], x[:, :, :, col:col + 1], w[:, :, :, kx:kx + 1], (1, 0), mask)
            if need_x:
                gx[:, :, :, col:col + 1] += gi
            if need_w:
                gw[:, :, :, kx:kx + 1] += gk
        gb = g.sum((0, 2, 3)).to(ctx.bdtype) if (need_b and ctx.bdtype is not None) else None
        return gx, (gw.to(ctx.wdtype) if need_w else None), gb


def conv3x3_reflect(x: torch.Tensor, weight: torch.Tensor, bias) -> torch.Tensor:
    """``conv2d(ReflectionPad2d(1)(x), weight, bias)`` (reference: model/layers.py:126-136) without the padded copy; same
    weights, same checkpoints.  See ``_Conv3x3Reflect``."""
    return _Conv3x3Reflect.apply(x, weight, bias)


class Conv3x3(nn.Module):
    """Reflection- (or zero-) padded 3x3 convolution of the decoder -- stock PyTorch convolutions, same parameters and
    state_dict keys as the reference (``conv.weight``, ``conv.bias``; ``pad`` has none).  ``fast_reflect`` (default) uses
    ``conv3x3_reflect``; ``Conv3x3.fast_reflect = False`` restores the literal pad-then-convolve sequence."""

    fast_reflect = True

    def __init__(self, in_channels: int, out_channels: int, use_refl: bool = True):
        super().__init__()
        self.use_refl = use_refl
        self.pad = nn.ReflectionPad2d(1) if use_refl else nn.ZeroPad2d(1)
        self.conv = nn.Conv2d(int(in_channels), int(out_channels), 3)

    def forward(self, x, with_bias: bool = True, padded: bool = False):
        """``with_bias=False``: the convolution alone -- the caller folds ``self.conv.bias`` into the kernel that consumes the
        result (``dvsloss.ops.bias_elu`` / ``elu_up2_cat``), saving the separate bias pass cuDNN's path runs.
        ``padded=True``: ``x`` already carries the reflected ring (it came from one of those kernels with ``pad=True``): the
        reference's pad-then-convolve (model/layers.py:131-136) is then the plain un-padded stock convolution."""
        bias = self.conv.bias if with_bias else None
        if padded:
            return F.conv2d(x, self.conv.weight, bias)
        if self.use_refl and self.fast_reflect and x.shape[2] >= 3 and x.shape[3] >= 3:
            return conv3x3_reflect(x, self.conv.weight, bias)
        return F.conv2d(self.pad(x), self.conv.weight, bias)


class ConvBlock(nn.Module):
    """Conv3x3 + ELU of the decoder (reference: model/layers.py:106-117): stock convolution; on CUDA the convolution's bias
    and the ELU run as one fused kernel each way."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv = Conv3x3(in_channels, out_channels)
        self.nonlin = nn.ELU(inplace=True)

    fused_bias_elu = True       # CUDA: convolution without bias, then bias + ELU as one kernel each way (dvsloss.ops.bias_elu)

    def forward(self, x):
        if self.fused_bias_elu and x.is_cuda and self.conv.conv.out_channels % 8 == 0:
            pre = self.conv(x, with_bias=False)
            if _ops.bias_elu_supported(pre):
                return _ops.bias_elu(pre, self.conv.conv.bias)
            return self.nonlin(pre + self.conv.conv.bias.to(pre.dtype).view(1, -1, 1, 1))
        return self.nonlin(self.conv(x))

    def pre_activation(self, x, padded: bool = False):
        """The convolution without bias and ELU (for the fused glue kernels); ``padded``: x carries its reflected ring."""
        return self.conv(x, with_bias=False, padded=padded)


def compute_depth_errors(gt: torch.Tensor, pred: torch.Tensor):
    """Standard depth metrics (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3); evaluation only."""
    ratio = torch.max(gt / pred, pred / gt)
    a1, a2, a3 = [(ratio < 1.25 ** k).float().mean() for k in (1, 2, 3)]
    diff = gt - pred
    rmse = diff.pow(2).mean().sqrt()
    rmse_log = (gt.log() - pred.log()).pow(2).mean().sqrt()
    abs_rel = (diff.abs() / gt).mean()
    sq_rel = (diff.pow(2) / gt).mean()
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3

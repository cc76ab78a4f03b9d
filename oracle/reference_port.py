"""CPU oracle for the view-synthesis (photometric) loss -- TEST INFRASTRUCTURE ONLY.

This module restates, with stock PyTorch ops, the arithmetic that the reference
performs in ``vo/learner_new.py:60-74,132-258`` and ``vo/learner_func.py:16-207``
(identical twin: ``model/layers.py:16-248``).  It exists so that the CUDA path can
be checked against the reference's numbers on machines where ``/root/reference``
is absent (the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
baseline legs of ``bench.py`` may import it; the product package never does.

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so this port is pinned by *live execution of the
reference itself*: ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` in the build container, feeds both the same
tensors and the same automask noise, asserts equality and commits the vectors
under ``tests/golden/``.

The port is generalised over a list of source frames (the reference hard-codes
``[-1, 1]``, ``learner_new.py:148,206,214``); for two sources it is op-for-op the
reference sequence, so on CPU the results are bit-identical.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

C1 = 0.01 ** 2
C2 = 0.03 ** 2


# --------------------------------------------------------------------------- geometry
def disp_to_depth(disp: torch.Tensor, min_depth: float, max_depth: float):
    """learner_func.py:16-26 -- sigmoid disparity -> (scaled disparity, depth)."""
    lo = 1 / max_depth
    hi = 1 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1 / scaled


def rot_from_axisangle(vec: torch.Tensor) -> torch.Tensor:
    """learner_func.py:65-104 -- Rodrigues formula on a [B,1,3] axis-angle vector."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x = axis[..., 0].unsqueeze(1)
    y = axis[..., 1].unsqueeze(1)
    z = axis[..., 2].unsqueeze(1)
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    rot = torch.zeros((vec.shape[0], 4, 4), dtype=vec.dtype, device=vec.device)
    rot[:, 0, 0] = torch.squeeze(x * xC + ca)
    rot[:, 0, 1] = torch.squeeze(xyC - zs)
    rot[:, 0, 2] = torch.squeeze(zxC + ys)
    rot[:, 1, 0] = torch.squeeze(xyC + zs)
    rot[:, 1, 1] = torch.squeeze(y * yC + ca)
    rot[:, 1, 2] = torch.squeeze(yzC - xs)
    rot[:, 2, 0] = torch.squeeze(zxC - ys)
    rot[:, 2, 1] = torch.squeeze(yzC + xs)
    rot[:, 2, 2] = torch.squeeze(z * zC + ca)
    rot[:, 3, 3] = 1
    return rot


def translation_matrix(t: torch.Tensor) -> torch.Tensor:
    """learner_func.py:49-62."""
    T = torch.zeros(t.shape[0], 4, 4, dtype=t.dtype, device=t.device)
    T[:, 0, 0] = 1
    T[:, 1, 1] = 1
    T[:, 2, 2] = 1
    T[:, 3, 3] = 1
    T[:, :3, 3, None] = t.contiguous().view(-1, 3, 1)
    return T


def transformation_from_parameters(axisangle, translation, invert=False):
    """learner_func.py:29-46 -- (axis-angle, translation) -> 4x4, optionally inverted."""
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


def pixel_grid(B: int, H: int, W: int, dtype, device) -> torch.Tensor:
    """Homogeneous pixel coordinates [B,3,H*W] as BackprojectDepth.__init__ builds
    them (learner_func.py:116-128): row 0 = column index u, row 1 = row index v."""
    v, u = torch.meshgrid(torch.arange(H, dtype=dtype, device=device),
                          torch.arange(W, dtype=dtype, device=device), indexing="ij")
    pix = torch.stack([u.reshape(-1), v.reshape(-1), torch.ones(H * W, dtype=dtype, device=device)], 0)
    return pix.unsqueeze(0).repeat(B, 1, 1)


def backproject(depth: torch.Tensor, inv_K: torch.Tensor) -> torch.Tensor:
    """learner_func.py:130-135 -- depth [B,1,H,W], inv_K [B,4,4] -> cam points [B,4,HW]."""
    B, _, H, W = depth.shape
    pix = pixel_grid(B, H, W, depth.dtype, depth.device)
    cam = torch.matmul(inv_K[:, :3, :3], pix)
    cam = depth.view(B, 1, -1) * cam
    ones = torch.ones(B, 1, H * W, dtype=depth.dtype, device=depth.device)
    return torch.cat([cam, ones], 1)


def project(points: torch.Tensor, K: torch.Tensor, T: torch.Tensor, H: int, W: int,
            eps: float = 1e-7) -> torch.Tensor:
    """learner_func.py:148-159 -- cam points -> normalised sampling grid [B,H,W,2]."""
    B = points.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    cam = torch.matmul(P, points)
    pix = cam[:, :2, :] / (cam[:, 2, :].unsqueeze(1) + eps)
    pix = pix.view(B, 2, H, W).permute(0, 2, 3, 1)
    pix = torch.stack([pix[..., 0] / (W - 1), pix[..., 1] / (H - 1)], -1)
    return (pix - 0.5) * 2


# --------------------------------------------------------------------------- photometric
def ssim(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """learner_func.py:190-207 -- 3x3 mean-filter SSIM *loss* map in [0,1]."""
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)
    mu_y = F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2
    sigma_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + C1) * (2 * sigma_xy + C2)
    d = (mu_x ** 2 + mu_y ** 2 + C1) * (sigma_x + sigma_y + C2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def reprojection_loss(pred: torch.Tensor, target: torch.Tensor, ssim_ratio: float = 0.85) -> torch.Tensor:
    """learner_new.py:60-74 -- ssim_ratio*mean_c SSIM + (1-ssim_ratio)*mean_c L1 -> [B,1,H,W]."""
    l1 = torch.abs(target - pred).mean(1, True)
    s = ssim(pred, target).mean(1, True)
    return (ssim_ratio * s) + ((1 - ssim_ratio) * l1)


def smooth_loss(disp: torch.Tensor, img: torch.Tensor) -> torch.Tensor:
    """learner_func.py:161-174 -- edge-aware first-difference smoothness."""
    gdx = torch.abs(disp[:, :, :, :-1] - disp[:, :, :, 1:])
    gdy = torch.abs(disp[:, :, :-1, :] - disp[:, :, 1:, :])
    gix = torch.mean(torch.abs(img[:, :, :, :-1] - img[:, :, :, 1:]), 1, keepdim=True)
    giy = torch.mean(torch.abs(img[:, :, :-1, :] - img[:, :, 1:, :]), 1, keepdim=True)
    gdx = gdx * torch.exp(-gix)
    gdy = gdy * torch.exp(-giy)
    return gdx.mean() + gdy.mean()


def warp_source(disp: torch.Tensor, src: torch.Tensor, K: torch.Tensor, inv_K: torch.Tensor,
                T: torch.Tensor, H: int, W: int, min_depth: float, max_depth: float):
    """learner_new.py:136-170 for one (scale, source): up-sample, depth, project, gather."""
    disp_up = F.interpolate(disp, [H, W], mode="bilinear", align_corners=False)
    _, depth = disp_to_depth(disp_up, min_depth, max_depth)
    grid = project(backproject(depth, inv_K), K, T, H, W)
    color = F.grid_sample(src, grid, padding_mode="border", align_corners=True)
    return disp_up, depth, grid, color


def view_synthesis_loss(disps: Sequence[torch.Tensor], target: torch.Tensor,
                        sources: Sequence[torch.Tensor], K: torch.Tensor, inv_K: torch.Tensor,
                        Ts: Sequence[torch.Tensor], noise: Optional[Sequence[torch.Tensor]] = None,
                        *, min_depth: float = 0.1, max_depth: float = 10.0, ssim_ratio: float = 0.85,
                        smoothness_ratio: float = 1e-3, auto_mask: bool = True,
                        keep: bool = False,
                        sel_override: Optional[Sequence[torch.Tensor]] = None) -> Dict[str, object]:
    """The whole hot path: learner_new.py:132-172 (_generate_images_pred) followed by
    :175-258 (_compute_losses), for S=len(disps) scales and N=len(sources) sources.

    ``noise[s]`` is the [B,N,H,W] standard-normal tensor the reference draws with
    ``torch.randn`` at learner_new.py:228 (multiplied by 1e-5 here); ``None`` means zeros.
    Returns ``loss`` (scalar), ``per_scale`` (list of S scalars), ``sel`` (list of
    [B,1,H,W] int64 argmin maps), ``combined`` (the candidate stacks) and, with ``keep=True``,
    the per-scale intermediates.

    ``sel_override[s]`` ([B,1,H,W] int64) replaces the argmin (test-only): near-ties between
    candidates flip under fp32 round-off, so gradient parity is checked with the selection
    pinned to the one the implementation under test made.
    """
    B, _, H, W = target.shape
    S, N = len(disps), len(sources)
    per_scale: List[torch.Tensor] = []
    sels: List[torch.Tensor] = []
    combs: List[torch.Tensor] = []
    extras: Dict[object, torch.Tensor] = {}
    total = 0
    for s in range(S):
        disp_up = None
        reproj = []
        for i in range(N):
            disp_up, depth, grid, color = warp_source(disps[s], sources[i], K, inv_K, Ts[i], H, W,
                                                      min_depth, max_depth)
            reproj.append(reprojection_loss(color, target, ssim_ratio))
            if keep:
                if disp_up.requires_grad:
                    disp_up.retain_grad()          # test-only: per-pixel gradient of the up-sampled map (summed over i)
                extras.setdefault(("disp_up_all", s), []).append(disp_up)
                extras[("depth", s)] = depth
                extras[("disp_up", s)] = disp_up
                extras[("sample", i, s)] = grid
                extras[("color", i, s)] = color
        reproj = torch.cat(reproj, 1)
        if auto_mask:
            ident = torch.cat([reprojection_loss(sources[i], target, ssim_ratio) for i in range(N)], 1)
            if noise is not None:
                ident = ident + noise[s] * 0.00001
            combined = torch.cat((ident, reproj), dim=1)
        else:
            combined = reproj
        if sel_override is not None:
            idx = sel_override[s]
            to_opt = combined.gather(1, idx)
        elif combined.shape[1] == 1:
            to_opt = combined
            idx = torch.zeros_like(combined, dtype=torch.int64)
        else:
            to_opt, idx = torch.min(combined, dim=1, keepdim=True)
        combs.append(combined.detach())
        sels.append(idx)
        loss = to_opt.mean()
        mean_disp = disp_up.mean(2, True).mean(3, True)
        mean_disp = torch.clamp(mean_disp, min=0.001)
        norm_disp = disp_up / (mean_disp + 1e-7)
        loss = loss + smoothness_ratio * smooth_loss(norm_disp, target) / (2 ** s)
        total = total + loss
        per_scale.append(loss)
    out: Dict[str, object] = {"loss": total / S, "per_scale": per_scale, "sel": sels, "combined": combs}
    if keep:
        out["extras"] = extras
    return out

"""Synthetic "Redwood-shaped" inputs for tests and benchmarks (no dataset, no network).

Produces the ``sample`` dict that ``MonoDataset.__getitem__`` yields in the
reference (``vo/dataset/common.py:48-92``) after collation -- ``("K", s)``,
``("inv_K", s)`` for s=0..3 as [B,4,4] float32, ``("source_left", 0)``,
``("target_image", 0)``, ``("source_right", 0)`` as [B,3,H,W] float32 in [0,1] --
plus disparity pyramids and poses shaped like the DepthNet / PoseNet outputs
(``model/depthnet.py:87-88``, ``model/posenet_single.py:195-200``).

Everything is plain torch on the requested device; nothing here touches the
CUDA extension, the oracle or the reference.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def redwood_intrinsics(B: int, H: int, W: int, num_scales: int = 4) -> Dict[Tuple[str, int], torch.Tensor]:
    """Redwood PrimeSense intrinsics (fx=fy=525, cx=319.5, cy=239.5 at 640x480) rescaled to
    HxW, then per-scale K / pinv(K) exactly as common.py:65-75 builds them."""
    K0 = np.eye(4, dtype=np.float64)
    K0[0, 0] = 525.0 * W / 640.0
    K0[1, 1] = 525.0 * H / 480.0
    K0[0, 2] = 319.5 * W / 640.0
    K0[1, 2] = 239.5 * H / 480.0
    out = {}
    for s in range(num_scales):
        K = K0.copy()
        K[0, :] *= (W // (2 ** s)) / W
        K[1, :] *= (H // (2 ** s)) / H
        inv_K = np.linalg.pinv(K)
        out[("K", s)] = torch.from_numpy(K).float().unsqueeze(0).repeat(B, 1, 1)
        out[("inv_K", s)] = torch.from_numpy(inv_K).float().unsqueeze(0).repeat(B, 1, 1)
    return out


def texture(B: int, H: int, W: int, gen: torch.Generator, device="cpu", cell: int = 8,
            pixel_noise: float = 0.02) -> torch.Tensor:
    """Smooth random RGB texture in [0,1]: U(0,1) at (H/cell x W/cell), bicubic up-sampled,
    plus a little per-pixel noise so SSIM windows are not degenerate."""
    low = torch.rand(B, 3, max(H // cell, 2), max(W // cell, 2), generator=gen)
    img = F.interpolate(low, size=(H, W), mode="bicubic", align_corners=False)
    img = img + pixel_noise * torch.rand(B, 3, H, W, generator=gen)
    return img.clamp_(0.0, 1.0).to(device)


def smooth_depth(B: int, H: int, W: int, gen: torch.Generator, lo: float = 0.5, hi: float = 5.0) -> torch.Tensor:
    """Smooth ground-truth depth in [lo,hi] metres, [B,1,H,W]."""
    low = torch.rand(B, 1, max(H // 32, 2), max(W // 32, 2), generator=gen)
    d = F.interpolate(low, size=(H, W), mode="bicubic", align_corners=False).clamp_(0, 1)
    return lo + (hi - lo) * d


def depth_to_disp(depth: torch.Tensor, min_depth: float = 0.1, max_depth: float = 10.0) -> torch.Tensor:
    """Inverse of disp_to_depth (learner_func.py:16-26): the sigmoid-range disparity that maps to `depth`."""
    lo, hi = 1.0 / max_depth, 1.0 / min_depth
    return ((1.0 / depth) - lo) / (hi - lo)


def _se3(axisangle: torch.Tensor, translation: torch.Tensor) -> torch.Tensor:
    """Plain Rodrigues, [B,3],[B,3] -> [B,4,4] (T*R order, as the reference with invert=False)."""
    B = axisangle.shape[0]
    ang = axisangle.norm(dim=1, keepdim=True)
    ax = axisangle / (ang + 1e-7)
    x, y, z = ax[:, 0], ax[:, 1], ax[:, 2]
    ca, sa = torch.cos(ang[:, 0]), torch.sin(ang[:, 0])
    C = 1 - ca
    M = torch.zeros(B, 4, 4)
    M[:, 0, 0] = x * x * C + ca
    M[:, 0, 1] = x * y * C - z * sa
    M[:, 0, 2] = z * x * C + y * sa
    M[:, 1, 0] = x * y * C + z * sa
    M[:, 1, 1] = y * y * C + ca
    M[:, 1, 2] = y * z * C - x * sa
    M[:, 2, 0] = z * x * C - y * sa
    M[:, 2, 1] = y * z * C + x * sa
    M[:, 2, 2] = z * z * C + ca
    M[:, :3, 3] = translation
    M[:, 3, 3] = 1
    return M


def pose_matrix(axisangle: torch.Tensor, translation: torch.Tensor, invert: bool) -> torch.Tensor:
    """[B,3],[B,3] -> [B,4,4]; invert=True gives R^T * Trans(-t) as learner_func.py:29-46 does."""
    M = _se3(axisangle, translation)
    if not invert:
        return M
    Rt = M[:, :3, :3].transpose(1, 2)
    out = torch.zeros_like(M)
    out[:, :3, :3] = Rt
    out[:, :3, 3] = -torch.einsum("bij,bj->bi", Rt, translation)
    out[:, 3, 3] = 1
    return out


def _render(target: torch.Tensor, depth: torch.Tensor, K: torch.Tensor, inv_K: torch.Tensor,
            T: torch.Tensor) -> torch.Tensor:
    """Approximate source view: resample the target along the flow induced by (depth, T^-1).
    Only used to manufacture geometrically plausible data; exactness is irrelevant."""
    B, _, H, W = target.shape
    v, u = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32),
                          indexing="ij")
    pix = torch.stack([u.reshape(-1), v.reshape(-1), torch.ones(H * W)], 0).unsqueeze(0).repeat(B, 1, 1)
    cam = depth.view(B, 1, -1) * torch.matmul(inv_K[:, :3, :3], pix)
    cam = torch.cat([cam, torch.ones(B, 1, H * W)], 1)
    P = torch.matmul(K, torch.linalg.inv(T))[:, :3, :]
    c = torch.matmul(P, cam)
    xy = c[:, :2] / (c[:, 2:3] + 1e-7)
    gx = (xy[:, 0] / (W - 1) - 0.5) * 2
    gy = (xy[:, 1] / (H - 1) - 0.5) * 2
    grid = torch.stack([gx, gy], -1).view(B, H, W, 2)
    return F.grid_sample(target, grid, padding_mode="border", align_corners=True)


def make_problem(B: int, H: int, W: int, num_sources: int = 2, num_scales: int = 4, seed: int = 0,
                 consistent: bool = True, device="cpu", pose_noise: float = 1e-3,
                 disp_noise: float = 0.05) -> Dict[str, object]:
    """One synthetic loss problem.

    consistent=True : ground-truth smooth depth and small poses; sources rendered from the target;
                      predicted disparity = GT*(1+disp_noise*N), predicted pose = GT + pose_noise*N,
                      so a large share of pixels selects the reprojection branch of the automask.
    consistent=False: independent textures, U(0,1) disparities, 0.01*N(0,1) poses (SURVEY 8d "perf runs").

    Returns dict with: sample (reference-format dict), target, sources [N x [B,3,H,W]],
    disps [S x [B,1,H>>s,W>>s]], axisangle / translation [N x [B,1,3]], invert [N bools],
    K, inv_K (scale 0), noise [S x [B,N,H,W]] standard normal.
    """
    gen = torch.Generator().manual_seed(seed)
    intr = redwood_intrinsics(B, H, W, max(num_scales, 4))
    K, inv_K = intr[("K", 0)], intr[("inv_K", 0)]
    target = texture(B, H, W, gen)
    # frame ids follow Monodepth2: -1, +1, -2, +2 ...; negative ids use invert=True (learner_new.py:110-127)
    frame_ids = [(-1) ** (i + 1) * (i // 2 + 1) for i in range(num_sources)]
    invert = [f < 0 for f in frame_ids]
    axisangle: List[torch.Tensor] = []
    translation: List[torch.Tensor] = []
    sources: List[torch.Tensor] = []
    if consistent:
        depth = smooth_depth(B, H, W, gen)
        for i in range(num_sources):
            aa = 0.008 * torch.randn(B, 3, generator=gen)
            tr = 0.03 * torch.randn(B, 3, generator=gen)
            T = pose_matrix(aa, tr, invert[i])                # pose the renderer uses for this source
            src = _render(target, depth, K, inv_K, T)
            src = (src + 0.01 * torch.rand(B, 3, H, W, generator=gen)).clamp_(0, 1)
            sources.append(src)
            aa_p = aa + pose_noise * torch.randn(B, 3, generator=gen)
            tr_p = tr + pose_noise * torch.randn(B, 3, generator=gen)
            axisangle.append(aa_p.view(B, 1, 3))
            translation.append(tr_p.view(B, 1, 3))
        disp_gt = depth_to_disp(depth)
        disps = []
        for s in range(num_scales):
            d = F.interpolate(disp_gt, size=(H >> s, W >> s), mode="bilinear", align_corners=False) if s else disp_gt
            d = d * (1 + disp_noise * torch.randn(d.shape, generator=gen))
            disps.append(d.clamp(1e-3, 1 - 1e-3))
    else:
        for i in range(num_sources):
            sources.append(texture(B, H, W, gen))
            axisangle.append(0.01 * torch.randn(B, 1, 3, generator=gen))
            translation.append(0.01 * torch.randn(B, 1, 3, generator=gen))
        disps = [torch.rand(B, 1, H >> s, W >> s, generator=gen) for s in range(num_scales)]
    noise = [torch.randn(B, num_sources, H, W, generator=gen) for _ in range(num_scales)]

    sample = dict(intr)
    sample[("target_image", 0)] = target
    if num_sources >= 1:
        sample[("source_left", 0)] = sources[0]
    if num_sources >= 2:
        sample[("source_right", 0)] = sources[1]
    for i in range(2, num_sources):                     # further frames (+-2, ...): keyed by their frame id
        sample[("source", frame_ids[i])] = sources[i]

    def mv(t):
        return t.to(device).contiguous()

    return {
        "sample": {k: mv(v) for k, v in sample.items()},
        "target": mv(target), "sources": [mv(s) for s in sources], "disps": [mv(d) for d in disps],
        "axisangle": [mv(a) for a in axisangle], "translation": [mv(t) for t in translation],
        "invert": invert, "frame_ids": frame_ids, "K": mv(K), "inv_K": mv(inv_K),
        "noise": [mv(n) for n in noise],
    }

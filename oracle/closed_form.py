"""Closed-form forward + backward of the view-synthesis loss in numpy -- TEST INFRASTRUCTURE ONLY.

Where ``reference_port.py`` restates the reference op-for-op and lets autograd differentiate it,
this file writes the same mathematics out explicitly (per-pixel forward, hand-derived adjoint),
which is the formulation the fused CUDA kernel implements.  It is validated in float64 against
autograd of the port (tests/test_oracle.py) and is the blueprint the kernel is read against:

  forward   vo/learner_new.py:132-172 (up-sample, depth, back-project, project, border gather)
            vo/learner_new.py:60-74 + vo/learner_func.py:177-207 (SSIM + L1)
            vo/learner_new.py:199-257 (automask, min over sources, smoothness, scale sum)
  backward  the adjoint of the above w.r.t. the disparity maps and the 4x4 poses only
            (images, intrinsics and noise receive no gradient in the reference either).

Index rules follow ATen: UpSample.h (align_corners=False bilinear) and GridSampler.h
(align_corners=True un-normalisation, border clipping with zero gradient at/after the border).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

C1 = 0.01 ** 2
C2 = 0.03 ** 2


def upsample_taps(out_size: int, in_size: int, dtype):
    """ATen bilinear, align_corners=False: src=(dst+.5)*in/out-.5 clamped at 0; i1=i0+(i0<in-1)."""
    scale = dtype(in_size) / dtype(out_size)
    dst = np.arange(out_size).astype(dtype)
    src = np.maximum(scale * (dst + dtype(0.5)) - dtype(0.5), dtype(0))
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    lam = np.clip(src - i0.astype(dtype), 0, 1).astype(dtype)
    return i0, i1, lam


def upsample(disp: np.ndarray, H: int, W: int) -> np.ndarray:
    """[B,h,w] -> [B,H,W]."""
    dt = disp.dtype.type
    y0, y1, ly = upsample_taps(H, disp.shape[1], dt)
    x0, x1, lx = upsample_taps(W, disp.shape[2], dt)
    ly = ly[None, :, None]
    lx = lx[None, None, :]
    top = disp[:, y0][:, :, x0] * (1 - lx) + disp[:, y0][:, :, x1] * lx
    bot = disp[:, y1][:, :, x0] * (1 - lx) + disp[:, y1][:, :, x1] * lx
    return top * (1 - ly) + bot * ly


def upsample_adjoint(g: np.ndarray, h: int, w: int) -> np.ndarray:
    """Adjoint of `upsample`: [B,H,W] -> [B,h,w]."""
    B, H, W = g.shape
    dt = g.dtype.type
    y0, y1, ly = upsample_taps(H, h, dt)
    x0, x1, lx = upsample_taps(W, w, dt)
    out = np.zeros((B, h, w), g.dtype)
    Y0, X0 = np.meshgrid(y0, x0, indexing="ij")
    Y1, X1 = np.meshgrid(y1, x1, indexing="ij")
    LY, LX = np.meshgrid(ly, lx, indexing="ij")
    for b in range(B):
        np.add.at(out[b], (Y0, X0), g[b] * (1 - LY) * (1 - LX))
        np.add.at(out[b], (Y0, X1), g[b] * (1 - LY) * LX)
        np.add.at(out[b], (Y1, X0), g[b] * LY * (1 - LX))
        np.add.at(out[b], (Y1, X1), g[b] * LY * LX)
    return out


def _pad_reflect(a: np.ndarray) -> np.ndarray:
    return np.pad(a, [(0, 0)] * (a.ndim - 2) + [(1, 1), (1, 1)], mode="reflect")


def box3(a: np.ndarray) -> np.ndarray:
    """3x3 mean with 1-px reflection padding over the last two axes."""
    p = _pad_reflect(a)
    H, W = a.shape[-2:]
    acc = np.zeros_like(a)
    for dy in range(3):
        for dx in range(3):
            acc = acc + p[..., dy:dy + H, dx:dx + W]
    return acc / a.dtype.type(9)


def box3_adjoint(g: np.ndarray) -> np.ndarray:
    """Adjoint of box3: spread g/9 over the padded 3x3 window, then fold the pad ring back
    (pad row -1 mirrors row 1, pad row H mirrors row H-2; same for columns)."""
    H, W = g.shape[-2:]
    pad = np.zeros(g.shape[:-2] + (H + 2, W + 2), g.dtype)
    for dy in range(3):
        for dx in range(3):
            pad[..., dy:dy + H, dx:dx + W] += g
    pad /= g.dtype.type(9)
    out = pad[..., 1:H + 1, 1:W + 1].copy()
    out[..., 1, :] += pad[..., 0, 1:W + 1]
    out[..., H - 2, :] += pad[..., H + 1, 1:W + 1]
    out[..., :, 1] += pad[..., 1:H + 1, 0]
    out[..., :, W - 2] += pad[..., 1:H + 1, W + 1]
    out[..., 1, 1] += pad[..., 0, 0]
    out[..., 1, W - 2] += pad[..., 0, W + 1]
    out[..., H - 2, 1] += pad[..., H + 1, 0]
    out[..., H - 2, W - 2] += pad[..., H + 1, W + 1]
    return out


def ssim_terms(x: np.ndarray, y: np.ndarray):
    """x,y: [B,3,H,W].  Returns the SSIM loss map and the coefficient fields a,b,c of SURVEY 3.3
    (dS/dm(x), dS/dm(x^2), dS/dm(xy)), zeroed where the clamp is strictly active."""
    dt = x.dtype.type
    mx, my = box3(x), box3(y)
    sx = box3(x * x) - mx * mx
    sy = box3(y * y) - my * my
    sxy = box3(x * y) - mx * my
    n1 = 2 * mx * my + dt(C1)
    n2 = 2 * sxy + dt(C2)
    d1 = mx * mx + my * my + dt(C1)
    d2 = sx + sy + dt(C2)
    n, d = n1 * n2, d1 * d2
    raw = (1 - n / d) / 2
    S = np.clip(raw, 0, 1)
    live = ((raw >= 0) & (raw <= 1)).astype(x.dtype)        # torch.clamp passes gradient at the bounds
    a = -0.5 * (2 * my * (n2 - n1) * d - n * 2 * mx * (d2 - d1)) / (d * d) * live
    b = 0.5 * n / (d * d2) * live
    c = -n1 / d * live
    return S, a, b, c


def reproj_map(x, y, ssim_ratio):
    S, a, b, c = ssim_terms(x, y)
    r = ssim_ratio * S.mean(1) + (1 - ssim_ratio) * np.abs(y - x).mean(1)
    return r, (a, b, c)


def loss_and_grads(disps: Sequence[np.ndarray], target: np.ndarray, sources: Sequence[np.ndarray],
                   K: np.ndarray, inv_K: np.ndarray, Ts: Sequence[np.ndarray],
                   noise: Optional[Sequence[np.ndarray]] = None, *, min_depth=0.1, max_depth=10.0,
                   ssim_ratio=0.85, smoothness_ratio=1e-3, auto_mask=True, eps=1e-7,
                   grad_per_scale: Optional[Sequence[float]] = None,
                   sel_override: Optional[Sequence[np.ndarray]] = None) -> Dict[str, object]:
    """disps[s]: [B,1,h_s,w_s]; target/sources: [B,3,H,W]; K,inv_K,Ts[i]: [B,4,4]; noise[s]: [B,N,H,W].

    Returns per_scale (S,), loss, sel [S x [B,H,W]], grad_disp [S x like disps], grad_T [N x [B,4,4]].
    ``grad_per_scale[s]`` is d(objective)/d(loss/s); default 1/S, i.e. the gradient of ``loss``.
    """
    dt = target.dtype.type
    B, _, H, W = target.shape
    S, N = len(disps), len(sources)
    gps = [1.0 / S] * S if grad_per_scale is None else list(grad_per_scale)
    lo, hi = dt(1.0 / max_depth), dt(1.0 / min_depth)
    v, u = np.meshgrid(np.arange(H).astype(dt), np.arange(W).astype(dt), indexing="ij")
    iK = inv_K[:, :3, :3]
    ray = (iK[:, :, 0, None, None] * u + iK[:, :, 1, None, None] * v + iK[:, :, 2, None, None])  # [B,3,H,W]
    Ps = [np.matmul(K, T)[:, :3, :] for T in Ts]

    ident = None
    if auto_mask:
        ident = np.stack([reproj_map(src, target, ssim_ratio)[0] for src in sources], 1)      # [B,N,H,W]

    # edge weights of the smoothness term (scale independent)
    wx = np.exp(-np.abs(target[:, :, :, :-1] - target[:, :, :, 1:]).mean(1))                  # [B,H,W-1]
    wy = np.exp(-np.abs(target[:, :, :-1, :] - target[:, :, 1:, :]).mean(1))                  # [B,H-1,W]

    per_scale = np.zeros(S, target.dtype)
    sels: List[np.ndarray] = []
    grad_disp: List[np.ndarray] = []
    grad_P = [np.zeros((B, 3, 4), target.dtype) for _ in range(N)]
    bidx = np.arange(B)[:, None, None]

    for s in range(S):
        du = upsample(disps[s][:, 0], H, W)                                                  # [B,H,W]
        D = 1 / (lo + (hi - lo) * du)
        cam = D[:, None] * ray                                                               # [B,3,H,W]
        warped, geo, reproj, coef = [], [], [], []
        for i in range(N):
            P = Ps[i]
            c = (P[:, :, 0, None, None] * cam[:, 0, None] + P[:, :, 1, None, None] * cam[:, 1, None]
                 + P[:, :, 2, None, None] * cam[:, 2, None] + P[:, :, 3, None, None])          # [B,3,H,W]
            z = c[:, 2] + dt(eps)
            px, py = c[:, 0] / z, c[:, 1] / z
            # Project3D normalises, grid_sample un-normalises (align_corners=True)
            ix = ((px / (W - 1) - dt(0.5)) * 2 + 1) / 2 * (W - 1)
            iy = ((py / (H - 1) - dt(0.5)) * 2 + 1) / 2 * (H - 1)
            live_x = (ix > 0) & (ix < W - 1)
            live_y = (iy > 0) & (iy < H - 1)
            ix = np.clip(ix, 0, W - 1)
            iy = np.clip(iy, 0, H - 1)
            x0 = np.floor(ix).astype(np.int64)
            y0 = np.floor(iy).astype(np.int64)
            tx, ty = (ix - x0).astype(target.dtype), (iy - y0).astype(target.dtype)
            x1, y1 = x0 + 1, y0 + 1
            okx, oky = x1 <= W - 1, y1 <= H - 1
            x1c, y1c = np.minimum(x1, W - 1), np.minimum(y1, H - 1)
            src = sources[i]
            taps = []
            for yy, xx, ok in ((y0, x0, None), (y0, x1c, okx), (y1c, x0, oky), (y1c, x1c, okx & oky)):
                t = np.stack([src[bidx, ch, yy, xx] for ch in range(3)], 1)                  # [B,3,H,W]
                if ok is not None:
                    t = t * ok[:, None]
                taps.append(t)
            nw, ne, sw, se = taps
            txb, tyb = tx[:, None], ty[:, None]
            wv = nw * (1 - txb) * (1 - tyb) + ne * txb * (1 - tyb) + sw * (1 - txb) * tyb + se * txb * tyb
            dxs = ((ne - nw) * (1 - tyb) + (se - sw) * tyb) * live_x[:, None]
            dys = ((sw - nw) * (1 - txb) + (se - ne) * txb) * live_y[:, None]
            r, abc = reproj_map(wv, target, ssim_ratio)
            warped.append(wv)
            reproj.append(r)
            coef.append(abc)
            geo.append((P, z, px, py, dxs, dys))
        reproj = np.stack(reproj, 1)                                                          # [B,N,H,W]
        if auto_mask:
            idn = ident if noise is None else ident + noise[s] * dt(0.00001)
            comb = np.concatenate([idn, reproj], 1)
        else:
            comb = reproj
        if sel_override is not None:          # test-only: pin the selection (see reference_port.view_synthesis_loss)
            sel = sel_override[s].astype(np.int64)
            m = np.take_along_axis(comb, sel[:, None], 1)[:, 0]
        else:
            sel = np.argmin(comb, 1)
            m = np.min(comb, 1)
        sels.append(sel)
        photo = m.mean()

        mu = np.maximum(du.mean((1, 2)), dt(0.001))                                           # [B]
        live_mu = (du.mean((1, 2)) >= dt(0.001)).astype(target.dtype)
        nd = du / (mu + dt(1e-7))[:, None, None]
        ddx = nd[:, :, :-1] - nd[:, :, 1:]
        ddy = nd[:, :-1, :] - nd[:, 1:, :]
        Nx, Ny = B * H * (W - 1), B * (H - 1) * W
        sm = (np.abs(ddx) * wx).sum() / Nx + (np.abs(ddy) * wy).sum() / Ny
        kappa = dt(smoothness_ratio / (2 ** s))
        per_scale[s] = photo + kappa * sm

        # ------------------------------------------------------------------ backward for this scale
        g_out = dt(gps[s])
        g_du = np.zeros_like(du)
        off = N if auto_mask else 0
        for i in range(N):
            P, z, px, py, dxs, dys = geo[i]
            a, b, c = coef[i]
            mask = ((sel == off + i).astype(target.dtype) * (g_out / (B * H * W)))[:, None]    # at window centres
            k = dt(ssim_ratio / 3.0)
            A = box3_adjoint(a * mask * k)
            Bf = box3_adjoint(b * mask * k)
            Cf = box3_adjoint(c * mask * k)
            G = A + 2 * warped[i] * Bf + target * Cf - dt((1 - ssim_ratio) / 3.0) * np.sign(target - warped[i]) * mask
            gix = (G * dxs).sum(1)
            giy = (G * dys).sum(1)
            gc0, gc1 = gix / z, giy / z
            gc2 = -(gix * px + giy * py) / z
            gc = np.stack([gc0, gc1, gc2], 1)                                                 # [B,3,H,W]
            gcam = (P[:, :, :3, None, None] * gc[:, :, None]).sum(1)                          # [B,3,H,W]
            gD = (gcam * ray).sum(1)
            g_du += gD * (-(hi - lo)) * D * D
            grad_P[i][:, :, :3] += np.einsum("bjhw,bkhw->bjk", gc, cam)
            grad_P[i][:, :, 3] += gc.sum((2, 3))
        # smoothness
        gn = np.zeros_like(du)
        tx_ = np.sign(ddx) * wx * (g_out * kappa / Nx)
        ty_ = np.sign(ddy) * wy * (g_out * kappa / Ny)
        gn[:, :, :-1] += tx_
        gn[:, :, 1:] -= tx_
        gn[:, :-1, :] += ty_
        gn[:, 1:, :] -= ty_
        inv = 1 / (mu + dt(1e-7))
        coupling = (gn * nd).sum((1, 2)) * inv / (H * W) * live_mu
        g_du += gn * inv[:, None, None] - coupling[:, None, None]
        grad_disp.append(upsample_adjoint(g_du, disps[s].shape[2], disps[s].shape[3])[:, None])

    grad_T = []
    for i in range(N):
        gT = np.zeros((B, 4, 4), target.dtype)
        gT[:, :, :] = np.einsum("bjm,bjk->bmk", K[:, :3, :], grad_P[i])
        grad_T.append(gT)
    return {"per_scale": per_scale, "loss": per_scale.sum() / S, "sel": sels,
            "grad_disp": grad_disp, "grad_T": grad_T, "grad_P": grad_P}

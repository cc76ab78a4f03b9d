"""Shared parity checker: an implementation of the fused loss versus the oracle (test infrastructure).

``impl(prob, grad_per_scale) -> dict(loss, per_scale[S], sel[S x [B,H,W] uint8], grad_disp[S], grad_T[N])`` (numpy)
is compared with ``oracle/reference_port.py`` (op-for-op restatement of the reference, pinned bit-for-bit to
the live reference by tests/golden/make_golden.py) evaluated on the same inputs.

Tolerances (BASELINE.json north_star):
  * loss, loss/s : rtol 1e-5 in fp32.  The fp32 reference itself carries rounding noise: where the loss is
    tiny because warped ~= target, (1 - n/d) and E[x^2]-E[x]^2 cancel, and a float64 evaluation of the SAME
    formulas sits up to ~3e-5 (relative) away from the reference's own fp32 value on the small fixtures.
    The bound is therefore  |impl - ref32| <= (1e-5 + 2*nu) * |ref32|  with  nu = max_s |ref32_s - ref64_s| / |ref32_s|
    the measured fp32 rounding uncertainty of the reference on that input (largest per-scale deviation of the
    reference's fp32 value from the float64 evaluation).  nu is ~1e-7 on the random / full-size cases, so there
    the plain 1e-5 gate decides; it only opens the gate on the tiny ill-conditioned fixtures.  On images of a
    few thousand pixels the per-pixel SSIM round-off (~2e-5, independent between pixels) does not average out
    either, so an absolute 4 * 2e-5 / sqrt(B*H*W) is added (1e-7 at 640x480, i.e. irrelevant at real sizes).
  * argmin selection: candidates closer than fp32 round-off may flip.  Every disagreeing pixel must be a near
    tie of the oracle's candidate stack (gap <= 1e-4 absolute: with unit-range images and SSIM's C2 = 9e-4 the
    fp32 cancellation in E[x^2]-E[x]^2 moves a single SSIM value by up to ~2e-5), and disagreements must be
    rare (< 1 % of pixels).
  * gradients: compared with the oracle's autograd gradients evaluated under the implementation's own
    selection (``sel_override``), so a flipped near-tie does not masquerade as a gradient error:
    rtol 1e-3 / atol 1e-6 element-wise (north_star; nearly vacuous at full size where |g| ~ 1e-7) and,
    the binding one, ||g - g_ref||_inf <= 1e-3 * ||g_ref||_inf, each with the same "+ 2*|ref32 - ref64|" allowance for the
    reference's own fp32 rounding as the loss.
  * discontinuities: the loss is only piecewise smooth.  At a pixel that sits within fp32 round-off of a kink,
    either one-sided derivative is a correct answer and two fp32 evaluations (the reference on CPU vs on CUDA
    included) may disagree by O(1) *at that pixel*.  The kinks are: a sampling coordinate within ~3e-4 px of an
    integer (bilinear cell change; covers the border clip at 0 and W-1), |target - warped| < 5e-5 in a channel
    (sign of the L1 term: the warped colour inherits the ~6e-5 px fp32 uncertainty of its sampling coordinate times the image slope), an SSIM value within 5e-5 of the clamp bounds (its fp32 evaluation carries ~2e-5 of cancellation noise), and 0 < |delta n| < 1e-6 in the
    smoothness term -- the synthetic textures are clipped to [0,1], so saturated patches where warped == target
    make the L1 and clamp kinks common.  They are located with the float64 oracle for the source(s) selected
    around the pixel.  A disparity element whose footprint (SSIM 3x3 window, then the bilinear up-sampling taps) holds
    such pixels is NOT exempt: it is held to the same element-wise bound widened by what those pixels can move,
    2 * (sum of their tap weights) * (largest per-pixel |d objective / d disp_up| of that scale) -- one kink pixel under
    a scale-3 element opens the gate by ~1/256 of the element's typical size.  The inf-norm bound over the kink-free
    elements uses the fp32 noise of those elements only.  At full resolution the kink share must stay below 20 %
    (the GPU tests also pin the share of every scale at 640x480 against stored values).
  * every element, no exclusion: at the median and the 90th percentile (and p99 / p99.9 where the map is large
    enough that the few kink pixels stay above the quantile) |g - g64| / max|g| of the implementation must be within
    2x of the reference's own fp32 evaluation against float64.  This is the gate a systematic error trips:
    tests/test_emulator.py::test_parity_gates_catch_a_one_percent_gradient_error builds the tile code with a 1 % error
    injected in phase_grad and requires check_parity to fail.
  * ``kink_free_problem`` builds inputs without any kink; there ``expect_kink_free`` asserts that the strict
    element-wise and inf-norm gates covered 100 % of the elements of every scale.
  * At the multi-hundred-thousand-element sizes a couple of elements per 100 000 that the locator misses may exceed the
    element-wise bound (by < 10x); they are counted in ``grad_disp_stragglers`` and never exempt from the inf-norm bound.
"""
from __future__ import annotations

import os
import sys
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "deep-visual-slam_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

from oracle import reference_port as port  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def load_golden(name: str) -> Dict[str, object]:
    """npz written by tests/golden/make_golden.py -> problem dict (numpy) + reference results."""
    z = np.load(os.path.join(GOLDEN_DIR, name))
    S = sum(1 for k in z.files if k.startswith("disp") and k[4:].isdigit())
    N = sum(1 for k in z.files if k.startswith("source"))
    am = bool(int(z["auto_mask"]))
    prob = dict(disps=[z[f"disp{s}"] for s in range(S)], target=z["target"], sources=[z[f"source{i}"] for i in range(N)],
                K=z["K"], inv_K=z["inv_K"], Ts=[z[f"T{i}"] for i in range(N)],
                noise=[z[f"noise{s}"] for s in range(S)] if am else None, auto_mask=am)
    ref = dict(loss=float(z["loss"]), per_scale=z["per_scale"], sel=[z[f"sel{s}"][:, 0] for s in range(S)],
               grad_disp=[z[f"grad_disp{s}"] for s in range(S)], grad_T=[z[f"grad_T{i}"] for i in range(N)])
    extra = {k: z[k] for k in z.files}
    return dict(prob=prob, ref=ref, raw=extra)


def problem_from_synthetic(p: Dict[str, object], auto_mask: bool = True) -> Dict[str, object]:
    """dvsloss.synthetic.make_problem output -> numpy problem dict with pose matrices."""
    Ts = [port.transformation_from_parameters(a.cpu(), t.cpu(), invert=inv)
          for a, t, inv in zip(p["axisangle"], p["translation"], p["invert"])]
    n = lambda t: t.detach().cpu().numpy()
    return dict(disps=[n(d) for d in p["disps"]], target=n(p["target"]), sources=[n(s) for s in p["sources"]],
                K=n(p["K"]), inv_K=n(p["inv_K"]), Ts=[n(T) for T in Ts],
                noise=[n(x) for x in p["noise"]] if auto_mask else None, auto_mask=auto_mask)


def oracle_eval(prob: Dict[str, object], *, sel_override: Optional[Sequence[np.ndarray]] = None,
                dtype=torch.float32, device="cpu", grad_per_scale: Optional[Sequence[float]] = None,
                want_grad: bool = True, keep: bool = False) -> Dict[str, object]:
    """Run the oracle port (autograd) on a numpy problem."""
    t = lambda a: torch.as_tensor(np.asarray(a), device=device).to(dtype)
    disps = [t(d).requires_grad_(want_grad) for d in prob["disps"]]
    Ts = [t(T).requires_grad_(want_grad) for T in prob["Ts"]]
    noise = [t(x) for x in prob["noise"]] if prob.get("noise") is not None else None
    so = None
    if sel_override is not None:
        so = [torch.as_tensor(np.asarray(s), device=device).long().unsqueeze(1) for s in sel_override]
    out = port.view_synthesis_loss(disps, t(prob["target"]), [t(s) for s in prob["sources"]], t(prob["K"]),
                                   t(prob["inv_K"]), Ts, noise, auto_mask=prob["auto_mask"], sel_override=so,
                                   keep=keep)
    S = len(disps)
    res = dict(loss=float(out["loss"].detach()), per_scale=np.array([float(p.detach()) for p in out["per_scale"]]),
               sel=[s[:, 0].cpu().numpy() for s in out["sel"]],
               combined=[c.cpu().double().numpy() for c in out["combined"]])
    if keep:
        res["extras"] = {k: v.detach().cpu().double().numpy() for k, v in out["extras"].items()
                         if not isinstance(v, list)}
    if want_grad:
        g = [1.0 / S] * S if grad_per_scale is None else list(grad_per_scale)
        obj = sum(gi * p for gi, p in zip(g, out["per_scale"]))
        obj.backward()
        res["grad_disp"] = [d.grad.cpu().double().numpy() for d in disps]
        res["grad_T"] = [T.grad.cpu().double().numpy() for T in Ts]
        if keep:   # d objective / d disp_up[s] per full-resolution pixel: "one pixel's worth" of gradient, see check_parity
            res["grad_disp_up"] = [sum(t.grad for t in out["extras"][("disp_up_all", s)]).cpu().double().numpy()
                                   for s in range(S)]
    return res


def check_parity(impl: Callable[[Dict[str, object], Optional[Sequence[float]]], Dict[str, object]],
                 prob: Dict[str, object], *, ref32: Optional[Dict[str, object]] = None, device="cpu",
                 grad_per_scale: Optional[Sequence[float]] = None, check_grad: bool = True,
                 loss_rtol: float = 1e-5, verbose: bool = False, expect_kink_free: bool = False) -> Dict[str, float]:
    """Assert the bounds of the module docstring; returns the measured error figures."""
    S, N = len(prob["disps"]), len(prob["sources"])
    B, _, H, W = prob["target"].shape
    got = impl(prob, grad_per_scale)
    if ref32 is None:
        ref32 = oracle_eval(prob, device=device, want_grad=False)
    ref64 = oracle_eval(prob, dtype=torch.float64, device=device, want_grad=False)
    stats: Dict[str, float] = {}

    # ---- losses
    ps_ref, ps_64 = np.asarray(ref32["per_scale"], np.float64), np.asarray(ref64["per_scale"], np.float64)
    ps_got = np.asarray(got["per_scale"], np.float64)
    nu = float((np.abs(ps_ref - ps_64) / np.abs(ps_ref)).max())
    # fp32 SSIM noise: eps * |E[x^2]| / C2 ~ 6e-8 * 0.3 / 9e-4 = 2e-5 per pixel, independent between pixels;
    # the mean over B*H*W pixels keeps 2e-5 / sqrt(B*H*W) of it (4 sigma allowed) -- matters on tiny images only
    pix_noise = 4 * 2e-5 / np.sqrt(B * H * W)
    tol = (loss_rtol + 2 * nu) * np.abs(ps_ref) + pix_noise
    err = np.abs(ps_got - ps_ref)
    stats["per_scale_rel_max"] = float((err / np.abs(ps_ref)).max())
    stats["ref_fp32_noise_rel_max"] = float((np.abs(ps_ref - ps_64) / np.abs(ps_ref)).max())
    assert np.all(err <= tol), f"loss/s off: got {ps_got}, ref {ps_ref}, err {err}, tol {tol}"
    l_ref, l_64 = float(ref32["loss"]), float(ref64["loss"])
    stats["loss_rel"] = abs(got["loss"] - l_ref) / abs(l_ref)
    assert abs(got["loss"] - l_ref) <= (loss_rtol + 2 * nu) * abs(l_ref) + pix_noise, (got["loss"], l_ref, l_64, nu)

    # ---- selection
    flips = 0.0
    for s in range(S):
        sg, sr = np.asarray(got["sel"][s]).astype(np.int64), np.asarray(ref32["sel"][s]).astype(np.int64)
        assert sg.shape == sr.shape, (sg.shape, sr.shape)
        nch = ref64["combined"][s].shape[1]
        assert sg.min() >= 0 and sg.max() < nch, f"selection out of range at scale {s}: {sg.min()}..{sg.max()}"
        bad = sg != sr
        flips = max(flips, float(bad.mean()))
        if bad.any():
            comb = ref64["combined"][s]
            vg = np.take_along_axis(comb, sg[:, None], 1)[:, 0]
            vm = comb.min(1)
            gap = (vg - vm)[bad]
            lim = 1e-4
            stats["sel_gap_max"] = max(stats.get("sel_gap_max", 0.0), float(gap.max()))
            assert np.all(gap <= lim), f"scale {s}: {int((gap > lim).sum())} selection flips are not near-ties (max gap {gap.max():.3e})"
    stats["sel_flip_frac_max"] = flips
    assert flips < 0.01, f"too many selection flips: {flips}"

    # ---- gradients under the implementation's selection
    if check_grad:
        so = [np.asarray(s).astype(np.int64) for s in got["sel"]] if (prob["auto_mask"] or N > 1) else None
        refg = oracle_eval(prob, sel_override=so, device=device, grad_per_scale=grad_per_scale)
        refg64 = oracle_eval(prob, sel_override=so, device=device, grad_per_scale=grad_per_scale,
                             dtype=torch.float64, keep=True)
        worst = 0.0
        n_risky = 0
        # atol is quoted for d loss (upstream 1/S per scale); scale it with the upstream gradient actually used
        gscale = 1.0 if grad_per_scale is None else max(1.0, S * max(abs(float(v)) for v in grad_per_scale))
        for s in range(S):
            g, r, r64 = np.asarray(got["grad_disp"][s], np.float64), refg["grad_disp"][s], refg64["grad_disp"][s]
            assert g.shape == r.shape
            assert np.all(np.isfinite(g)), f"non-finite grad_disp[{s}]"
            kinks = _kink_weight(prob, refg64["extras"], so[s] if so is not None else None, s, N, H, W)
            risky = kinks > 0
            n_risky += int(risky.sum())
            stats[f"near_kink_frac/{s}"] = float(risky.mean())
            rmax = max(np.abs(r).max(), 1e-30)
            noise = 2 * np.abs(r - r64)                      # the reference's own fp32 rounding, element by element
            err = np.abs(g - r)
            ok = ~risky
            rel = (err[ok]).max() / rmax if ok.any() else 0.0
            worst = max(worst, rel)
            # (1) kink-free elements, inf-norm: the allowance is the fp32 noise of THOSE elements only
            assert (not ok.any()) or np.all(err[ok] <= 1e-3 * rmax + noise[ok].max()), \
                f"grad_disp[{s}]: normalised inf-norm error {rel:.3e} (reference fp32 noise {noise[ok].max() / 2 / rmax:.2e})"
            # (2) kink-free elements, element-wise: north_star's rtol 1e-3 / atol 1e-6 on the raw gradient
            lim = 1e-3 * np.abs(r) + 1e-6 * gscale + noise
            bad = err[ok] > lim[ok]
            # stragglers: at most 2 in 100 000 elements (kinks the float64 locator above missed by a hair) may sit
            # outside the element-wise bound, and then by no more than 10x; the inf-norm bound above has no exception
            assert bad.size == 0 or (bad.mean() <= 2e-5 and np.all(err[ok] <= 10 * lim[ok])), \
                f"grad_disp[{s}]: {int(bad.sum())} of {bad.size} elements outside rtol 1e-3 / atol 1e-6"
            stats["grad_disp_stragglers"] = stats.get("grad_disp_stragglers", 0) + int(bad.sum())
            # (3) elements with kink pixels in their footprint: the same element-wise bound plus what those pixels can
            # move -- a pixel on a kink may take either one-sided derivative, i.e. change by up to twice one pixel's worth
            # of gradient (gpix: the largest per-pixel |d objective / d disp_up| of this scale, float64 oracle), weighted
            # by its up-sampling tap weight into the element.  One kink pixel under a scale-3 element (256-pixel
            # footprint) therefore opens the gate by ~1/256 of the element's typical size, not by 0.25 * max.
            gpix = float(np.abs(refg64["grad_disp_up"][s]).max())
            lim_k = lim + 2.0 * kinks * gpix
            badk = err[risky] > lim_k[risky]
            assert not badk.any(), (f"grad_disp[{s}]: {int(badk.sum())} near-kink elements outside the footprint-scaled bound, "
                                    f"worst excess {(err[risky] / lim_k[risky]).max():.2f}x")
            # (4) every element, no exclusion: the implementation must be as close to the float64 evaluation as the
            # reference's own fp32 evaluation is (within 2x, plus a floor of a few fp32 ulps of the largest element) at
            # the median and the 90th percentile -- and further into the tail (p99, p99.9) where the map has enough
            # elements that the handful of kink pixels, which either evaluation may resolve the other way, stay above
            # the quantile.  A systematic error of 1 % in the photometric gradient moves the median by
            # ~1e-2 * median|g| / max|g| ~ 1e-4, two orders above the fp32 noise (~2e-6): this is the gate that fails on it.
            e_impl, e_ref = np.abs(g - r64).ravel() / rmax, np.abs(r - r64).ravel() / rmax
            levels = [0.5, 0.9] + ([0.99] if e_impl.size >= 10000 else []) + ([0.999] if e_impl.size >= 1000000 else [])
            qi, qr = np.quantile(e_impl, levels), np.quantile(e_ref, levels)
            stats[f"dist_ratio_p50/{s}"] = float(qi[0] / max(qr[0], 1e-30))
            stats[f"dist_ratio_p90/{s}"] = float(qi[1] / max(qr[1], 1e-30))
            # (a map of a few elements has no distribution to compare: its "quantiles" are single rounding errors of
            # sums over thousands of pixels, which differ by reduction order alone; gates (1)-(3) cover every such element)
            assert e_impl.size < 64 or np.all(qi <= 2 * qr + 2e-7), \
                f"grad_disp[{s}]: quantiles {levels} of |g - g64| / max = {qi} vs the reference's own fp32 evaluation {qr}"
            assert risky.mean() < 0.20 or g.shape[2:] != (H, W), f"grad_disp[{s}]: {risky.mean():.1%} of the elements excluded as near-kink"
        stats["grad_disp_relinf_max"] = worst
        stats["near_kink_elements"] = n_risky
        if expect_kink_free:      # the strict gates (1), (2) above then covered every element of every scale
            assert n_risky == 0 and stats.get("grad_disp_stragglers", 0) == 0, (n_risky, stats.get("grad_disp_stragglers"))
        worst = 0.0
        for i in range(N):
            g, r, r64 = np.asarray(got["grad_T"][i], np.float64), refg["grad_T"][i], refg64["grad_T"][i]
            rmax = max(np.abs(r).max(), 1e-30)
            rel = np.abs(g - r).max() / rmax
            worst = max(worst, rel)
            assert np.abs(g - r).max() <= 1e-3 * rmax + 2 * np.abs(r - r64).max(), \
                f"grad_T[{i}]: normalised inf-norm error {rel:.3e}\n{g}\n{r}"
        stats["grad_T_relinf_max"] = worst
    if verbose:
        print(stats)
    return stats


def _dilate3(m: np.ndarray) -> np.ndarray:
    H, W = m.shape[-2:]
    pad = np.pad(m, ((0, 0), (1, 1), (1, 1)))
    out = np.zeros_like(m)
    for dy in range(3):
        for dx in range(3):
            out |= pad[:, dy:dy + H, dx:dx + W]
    return out


def _kink_weight(prob, extras, sel, s: int, N: int, H: int, W: int) -> np.ndarray:
    """For every disp[s] element: the sum of up-sampling tap weights of the full-resolution pixels in its footprint that
    sit within round-off of a kink of the loss (see module docstring); 0 = kink-free.  `extras` are the float64
    oracle's intermediates, `sel` the selection in force."""
    from oracle.closed_form import ssim_terms, upsample_taps
    B, _, h, w = prob["disps"][s].shape
    tgt = np.asarray(prob["target"], np.float64)
    off = N if prob["auto_mask"] else 0
    thr = 3e-4 + 1e-6 * max(H, W)        # fp32 spacing of a pixel coordinate is 6e-8 * coordinate
    pix = np.zeros((B, H, W), bool)
    for i in range(N):
        chosen = (np.asarray(sel) == off + i) if sel is not None else np.ones((B, H, W), bool)
        active = _dilate3(chosen)         # pixels whose warped colour of source i receives any gradient
        g = extras[("sample", i, s)]
        ix = (g[..., 0] + 1) / 2 * (W - 1)
        iy = (g[..., 1] + 1) / 2 * (H - 1)
        k = (np.abs(ix - np.round(ix)) < thr) | (np.abs(iy - np.round(iy)) < thr)
        col = extras[("color", i, s)]
        k |= (np.abs(tgt - col) < 5e-5).any(1) & chosen
        S = ssim_terms(col, tgt)[0]
        k |= _dilate3((((S < 5e-5) | (S > 1 - 5e-5)).any(1)) & chosen)
        pix |= k & active
    du = extras[("disp_up", s)][:, 0]
    nd = du / np.maximum(du.mean((1, 2), keepdims=True), 1e-3)
    dx, dy = np.abs(nd[:, :, :-1] - nd[:, :, 1:]), np.abs(nd[:, :-1, :] - nd[:, 1:, :])
    kx, ky = (dx > 0) & (dx < 1e-6), (dy > 0) & (dy < 1e-6)
    pix[:, :, :-1] |= kx
    pix[:, :, 1:] |= kx
    pix[:, :-1, :] |= ky
    pix[:, 1:, :] |= ky
    out = np.zeros((B, 1, h, w), np.float64)
    y0, y1, ly = upsample_taps(H, h, np.float64)
    x0, x1, lx = upsample_taps(W, w, np.float64)
    bb, yy, xx = np.nonzero(pix)
    for ya, wy in ((y0, 1 - ly), (y1, ly)):
        for xa, wx in ((x0, 1 - lx), (x1, lx)):
            # a tap of weight 0 still marks the element (round-off can put the kink on either side of a tap boundary)
            np.add.at(out, (bb, 0, ya[yy], xa[xx]), np.maximum(wy[yy] * wx[xx], 1e-3))
    return out


def kink_free_problem(B: int = 1, H: int = 64, W: int = 96, seed: int = 0) -> Dict[str, object]:
    """A problem on which the loss is smooth at every pixel, so the strict element-wise gradient gates apply to 100 %
    of the elements of every scale (no near-kink exemption): low-contrast smooth textures well inside (0,1) (no clipping,
    SSIM far from its clamp), sources = independent smooth textures offset in brightness (|target - warped| never near 0),
    fronto-parallel translations with ramp disparities so that every sampling coordinate stays a fixed sub-pixel distance
    (0.2 .. 0.8 px) from the integer grid, monotone disparity ramps (|delta n| is either exactly 0 on the border clamp or
    far above 1e-6), no automask.  ``check_parity`` asserts ``near_kink_elements == 0`` when asked to (``expect_kink_free``)."""
    import torch.nn.functional as F
    from dvsloss.synthetic import redwood_intrinsics
    gen = torch.Generator().manual_seed(seed)

    def tex(lo, hi):
        low = torch.rand(B, 3, 6, 8, generator=gen)
        img = F.interpolate(low, size=(H, W), mode="bicubic", align_corners=False)
        img = (img - img.amin()) / (img.amax() - img.amin())
        return lo + (hi - lo) * img

    target = tex(0.35, 0.60)
    sources = [tex(0.62, 0.80), tex(0.15, 0.33)]
    intr = redwood_intrinsics(B, H, W)
    K, inv_K = intr[("K", 0)], intr[("inv_K", 0)]
    fx, fy = float(K[0, 0, 0]), float(K[0, 1, 1])
    v, u = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    disps = []
    for s in range(4):
        ramp = 0.30 + (0.10 + 0.01 * s) * u + (0.08 - 0.01 * s) * v            # in [0.30, 0.49]
        d = F.interpolate(ramp[None, None], size=(H >> s, W >> s), mode="bilinear", align_corners=False)
        disps.append(d.repeat(B, 1, 1, 1).contiguous())
    sd_lo, sd_hi = 0.1 + 9.9 * 0.30, 0.1 + 9.9 * 0.49                            # scaled disparity = 1 / depth
    Ts = []
    for sign in (1.0, -1.0):
        # flow = f * t / depth in [0.27, 0.45] px (x) and [0.3, 0.5] px (y), away from the integer grid
        T = torch.eye(4).repeat(B, 1, 1)
        T[:, 0, 3] = sign * 0.27 / (fx * sd_lo)
        T[:, 1, 3] = -sign * 0.30 / (fy * sd_lo)
        Ts.append(T)
    assert 0.45 * sd_hi / sd_lo < 0.8
    n = lambda t: t.detach().cpu().numpy().astype(np.float32)
    return dict(disps=[n(d) for d in disps], target=n(target), sources=[n(s_) for s_ in sources], K=n(K), inv_K=n(inv_K),
                Ts=[n(T) for T in Ts], noise=None, auto_mask=False)

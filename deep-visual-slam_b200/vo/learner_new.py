"""MonodepthTrainer -- the VO learner (loss graph) with the view-synthesis loss on the fused sm_100a kernel.

Constructor and ``process_batch(sample) -> (outputs, losses)`` follow the reference
(vo/learner_new.py:15-105) so ``vo/train.py`` drives it unchanged:

  * ``sample``: the collated dict of ``MonoDataset.__getitem__`` (vo/dataset/common.py:48-92): ``("K", s)``,
    ``("inv_K", s)`` [B,4,4], ``("source_left", 0)``, ``("target_image", 0)``, ``("source_right", 0)`` [B,3,H,W];
    tensors are moved to the device in place in the dict, as the reference does (:93-95).
  * ``losses``: ``"loss"`` and ``"loss/0".."loss/3"``, 0-d tensors on the device, attached to the autograd graph.
  * ``outputs``: what DepthNet returns (``("disp", s)``) plus ``("axisangle"|"translation"|"cam_T_cam", 0, f)``
    and ``"identity_selection/s"``.  The full-resolution intermediates the reference materialises on every
    step (``("disp_up", s)``, ``("depth", s)``, ``("sample", f, s)``, ``("color", f, s)``,
    ``("color_identity", f, s)`` -- 0.94 GB at B=16) are *not* produced on the hot path: ``outputs`` is a
    ``LazyOutputs`` dict that computes them (granular kernels, ``no_grad``) the first time one of those keys is
    read, so the plotting call of the unchanged trainer (vo/train.py:268-279 -> vo/utils/plot_utils.py:40-47,
    step 0 and every 1000th) finds them and ordinary steps pay nothing.

Hot path: DepthNet + PoseNet (stock PyTorch) -> ``dvsloss.view_synthesis_loss`` (one fused launch computing
every scale and source, the losses and their gradients) instead of the ~2 160 ATen launches of
``_generate_images_pred`` + ``_compute_losses`` (vo/learner_new.py:132-258).  The granular methods of the
same names are kept (built on the per-op kernels) for code that calls them directly.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, Sequence, Tuple

import torch
import torch.nn as nn

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from dvsloss import view_synthesis_loss  # noqa: E402
from dvsloss import ops as _ops  # noqa: E402
from model.layers import BackprojectDepth, Project3D, SSIM  # noqa: E402


class LazyOutputs(dict):
    """``outputs`` of ``process_batch``.  A plain dict for everything the step computed; the full-resolution side
    products of vo/learner_new.py:142,146,158,165-172 (and the 4x4 pose matrices, :124-127) are produced on first access
    (``d[key]``, ``d.get(key)``, ``key in d``) by the filler bound to the key's kind and then stay in the dict like any
    other entry.  Each filler runs at most once."""

    LAZY_KINDS = ("disp_up", "depth", "sample", "color", "color_identity", "cam_T_cam")

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._fillers = {}

    def bind(self, filler, kinds=None) -> "LazyOutputs":
        for k in (self.LAZY_KINDS if kinds is None else kinds):
            self._fillers[k] = filler
        return self

    def _is_lazy(self, key) -> bool:
        return isinstance(key, tuple) and len(key) > 1 and key[0] in self._fillers

    def __missing__(self, key):
        if self._is_lazy(key):
            filler = self._fillers[key[0]]
            for k in [k for k, f in self._fillers.items() if f is filler]:
                del self._fillers[k]
            filler(self)
            if dict.__contains__(self, key):
                return dict.__getitem__(self, key)
        raise KeyError(key)

    def __contains__(self, key) -> bool:
        if dict.__contains__(self, key):
            return True
        if self._is_lazy(key):
            try:
                self[key]
            except KeyError:
                return False
            return True
        return False

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


class MonodepthTrainer:
    def __init__(self, depth_net: nn.Module, pose_net: nn.Module, config: dict, device: torch.device,
                 frame_ids: Sequence[int] = (-1, 1), noise: str = "torch", fused: bool = True):
        self.depth_net, self.pose_net, self.config, self.device = depth_net, pose_net, config, torch.device(device)
        tr = config["Train"]
        self.num_scales = 4
        self.num_source = tr["num_source"]           # read, as in the reference; the frame list decides
        self.batch_size = tr["batch_size"]
        self.image_shape = (tr["img_h"], tr["img_w"])
        self.smoothness_ratio = tr["smoothness_ratio"]
        self.auto_mask = tr["auto_mask"]
        self.ssim_ratio = tr["ssim_ratio"]
        self.min_depth = tr["min_depth"]
        self.max_depth = tr["max_depth"]
        self.use_compile = tr.get("use_compile", False)   # accepted for config compatibility; nothing to compile
        self.frame_ids = list(frame_ids)
        # "torch": draw torch.randn([B,N,H,W]) per scale exactly like vo/learner_new.py:228 (same RNG consumption);
        # "kernel": counter-based generator inside the kernel (no extra tensors, no extra launches).
        self.noise = noise
        self.fused = fused
        H, W = self.image_shape
        self.ssim = SSIM().to(self.device)
        self.backproject_depth = BackprojectDepth(self.batch_size, H, W).to(self.device)
        self.project_3d = Project3D(self.batch_size, H, W).to(self.device)

    # ------------------------------------------------------------------------------------------ helpers
    @staticmethod
    def _source_key(frame_id: int) -> Tuple[str, int]:
        if frame_id == -1:
            return ("source_left", 0)
        if frame_id == 1:
            return ("source_right", 0)
        return ("source", frame_id)

    def _compute_reprojection_loss(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _ops.compute_reprojection_loss(pred, target, self.ssim_ratio)

    # ------------------------------------------------------------------------------------------ step
    def process_batch(self, sample: Dict) -> Tuple[Dict, Dict]:
        for key, val in sample.items():
            if isinstance(val, torch.Tensor):
                sample[key] = val.to(self.device, non_blocking=True)
        outputs = LazyOutputs(self.depth_net(sample[("target_image", 0)]))
        if self.fused:
            # pose parameters go to the loss as they are: transformation_from_parameters and its backward run inside the
            # loss's launches; ("cam_T_cam", 0, f) joins the lazily materialised outputs
            outputs.update(self._predict_poses(sample, matrices=False))
            losses = self._view_synthesis(sample, outputs)
            outputs.bind(lambda out: self.materialize_outputs(sample, out),
                         ("disp_up", "depth", "sample", "color", "color_identity"))
            outputs.bind(self._materialize_poses, ("cam_T_cam",))
        else:
            outputs.update(self._predict_poses(sample))
            self._generate_images_pred(sample, outputs)
            losses = self._compute_losses(sample, outputs)
        return outputs, losses

    def _predict_poses(self, sample: Dict, matrices: bool = True) -> Dict:
        """PoseNet on (source, target) ordered in time; negative frame ids are inverted (Monodepth2 convention,
        reference: vo/learner_new.py:107-129)."""
        out = {}
        tgt = sample[("target_image", 0)]
        for f in self.frame_ids:
            src = sample[self._source_key(f)]
            pair = torch.cat([src, tgt], 1) if f < 0 else torch.cat([tgt, src], 1)
            axisangle, translation = self.pose_net(pair)
            out[("axisangle", 0, f)] = axisangle
            out[("translation", 0, f)] = translation
            if matrices:
                out[("cam_T_cam", 0, f)] = _ops.transformation_from_parameters(axisangle[:, 0], translation[:, 0],
                                                                               invert=(f < 0))
        return out

    def _view_synthesis(self, sample: Dict, outputs: Dict) -> Dict:
        disps = [outputs[("disp", s)] for s in range(self.num_scales)]
        sources = [sample[self._source_key(f)] for f in self.frame_ids]
        res = view_synthesis_loss(disps, sample[("target_image", 0)], sources, sample[("K", 0)], sample[("inv_K", 0)],
                                  axisangles=[dict.__getitem__(outputs, ("axisangle", 0, f)) for f in self.frame_ids],
                                  translations=[dict.__getitem__(outputs, ("translation", 0, f)) for f in self.frame_ids],
                                  inverts=[f < 0 for f in self.frame_ids],
                                  noise=self.noise, min_depth=self.min_depth, max_depth=self.max_depth,
                                  ssim_ratio=self.ssim_ratio, smoothness_ratio=self.smoothness_ratio,
                                  auto_mask=self.auto_mask, return_selection=bool(self.auto_mask))
        total, per_scale = res[0], res[1]
        losses = {f"loss/{s}": per_scale[s] for s in range(self.num_scales)}
        losses["loss"] = total
        if self.auto_mask:
            n = len(self.frame_ids)
            for s in range(self.num_scales):
                outputs[f"identity_selection/{s}"] = (res[2 + s] > n - 1).float().unsqueeze(1)
        return losses

    # ------------------------------------------------------------------------------------------ granular path
    def _generate_images_pred(self, sample: Dict, outputs: Dict) -> None:
        """Per-op version of vo/learner_new.py:132-172 on the granular kernels."""
        H, W = self.image_shape
        for s in range(self.num_scales):
            disp_up = _ops.upsample_bilinear(outputs[("disp", s)], (H, W))
            outputs[("disp_up", s)] = disp_up
            _, depth = _ops.disp_to_depth(disp_up, self.min_depth, self.max_depth)
            outputs[("depth", s)] = depth
            cam_points = self.backproject_depth(depth, sample[("inv_K", 0)])       # source independent: once per scale
            for f in self.frame_ids:
                src = sample[self._source_key(f)]
                grid = self.project_3d(cam_points, sample[("K", 0)], outputs[("cam_T_cam", 0, f)])
                outputs[("sample", f, s)] = grid
                outputs[("color", f, s)] = _ops.grid_sample_border(src, grid)
                outputs[("color_identity", f, s)] = src

    def _compute_losses(self, inputs: Dict, outputs: Dict) -> Dict:
        """Per-op version of vo/learner_new.py:175-258 (needs _generate_images_pred's outputs)."""
        losses, total = {}, 0
        target = inputs[("target_image", 0)]
        n = len(self.frame_ids)
        for s in range(self.num_scales):
            reproj = torch.cat([self._compute_reprojection_loss(outputs[("color", f, s)], target) for f in self.frame_ids], 1)
            if self.auto_mask:
                ident = torch.cat([self._compute_reprojection_loss(outputs[("color_identity", f, s)], target)
                                   for f in self.frame_ids], 1)
                ident = ident + torch.randn(ident.shape, device=ident.device) * 0.00001
                combined = torch.cat((ident, reproj), 1)
            else:
                combined = reproj
            if combined.shape[1] == 1:
                to_optimise = combined
            else:
                to_optimise, idxs = torch.min(combined, dim=1, keepdim=True)
                if self.auto_mask:
                    outputs[f"identity_selection/{s}"] = (idxs > n - 1).float()
            disp = outputs[("disp_up", s)]
            mean_disp = torch.clamp(disp.mean(2, True).mean(3, True), min=0.001)
            smooth = _ops.get_smooth_loss(disp / (mean_disp + 1e-7), target)
            loss = to_optimise.mean() + self.smoothness_ratio * smooth / (2 ** s)
            total = total + loss
            losses[f"loss/{s}"] = loss
        losses["loss"] = total / self.num_scales
        return losses

    @torch.no_grad()
    def _materialize_poses(self, outputs: Dict) -> None:
        for f in self.frame_ids:
            if not dict.__contains__(outputs, ("cam_T_cam", 0, f)):
                aa, tr = dict.__getitem__(outputs, ("axisangle", 0, f)), dict.__getitem__(outputs, ("translation", 0, f))
                dict.__setitem__(outputs, ("cam_T_cam", 0, f),
                                 _ops.transformation_from_parameters(aa.detach().float()[:, 0], tr.detach().float()[:, 0],
                                                                     invert=(f < 0)))

    @torch.no_grad()
    def materialize_outputs(self, sample: Dict, outputs: Dict) -> Dict:
        """Fill the full-resolution side products the plotting code reads (("depth", s), ("color", f, s), ...)."""
        self._materialize_poses(outputs)
        det = {k: (v.detach().float() if isinstance(v, torch.Tensor) else v) for k, v in dict.items(outputs)}
        self._generate_images_pred(sample, det)
        for k, v in det.items():
            if not dict.__contains__(outputs, k):
                dict.__setitem__(outputs, k, v)
        return outputs

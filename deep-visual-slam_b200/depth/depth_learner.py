"""DepthLearner -- supervised multi-scale depth loss (SILog + edge-aware smoothness) on the sm_100a kernels.

Same constructor, attributes and methods as the reference (depth/depth_learner.py:6-147): ``disp_to_depth``,
``compute_gradients``, ``get_smooth_loss``, ``silog_loss``, ``multi_scale_loss``, ``forward_step(sample) ->
(total_loss, total_silog, total_smooth, pred_depths)`` with ``sample`` = ``{"image", "depth", "valid_mask"}``.
The arithmetic re-uses the device functions of the VO path:

  * ``F.interpolate(pred_depth, (H, W), bilinear)`` (:106)        -> ``dvsloss.ops.upsample_bilinear`` (gather-form adjoint)
  * ``get_smooth_loss`` (:51-73)                                    -> mean normalisation + ``dvsloss.ops.get_smooth_loss``
    (|dx| * exp(-mean_c |dx img|) means: the same kernel as the VO smoothness term, vo/learner_func.py:161-174)
  * ``silog_loss`` (:75-95)                                         -> ``dvsloss.ops.silog_loss`` (one reduction kernel)
  * ``disp_to_depth`` (:33-39)                                      -> ``dvsloss.ops.disp_to_depth``
CUDA tensors only (no CPU fallback by design)."""
from __future__ import annotations

import os
import sys
from typing import Any, Dict, List, Tuple

import torch
import torch.nn as nn

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from dvsloss import ops as _ops  # noqa: E402


class DepthLearner:
    def __init__(self, model: nn.Module, config: Dict[str, Any], device: torch.device) -> None:
        self.model = model
        self.min_depth = config["Train"]["min_depth"]
        self.max_depth = config["Train"]["max_depth"]
        self.num_scales = 4
        self.device = torch.device(device)
        self.alphas = [1.0, 0.5, 0.25, 0.125]                     # depth/depth_learner.py:25
        self.smooth_weight = config["Train"].get("smooth_weight", 0.1)
        self.silog_weight = config["Train"].get("silog_weight", 1.0)

    def disp_to_depth(self, disp: torch.Tensor) -> torch.Tensor:
        return _ops.disp_to_depth(disp, self.min_depth, self.max_depth)[1]

    @staticmethod
    def compute_gradients(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """|dx| [B,C,H,W-1], |dy| [B,C,H-1,W] (plain torch: kept for callers; the losses below do not materialise them)."""
        return torch.abs(x[:, :, :, 1:] - x[:, :, :, :-1]), torch.abs(x[:, :, 1:, :] - x[:, :, :-1, :])

    def get_smooth_loss(self, disp: torch.Tensor, img: torch.Tensor) -> torch.Tensor:
        disp_mean = disp.mean(dim=[2, 3], keepdim=True).clamp(min=1e-7)
        return _ops.get_smooth_loss(disp / disp_mean, img)

    def silog_loss(self, prediction: torch.Tensor, target: torch.Tensor, valid_mask: torch.Tensor,
                   variance_focus: float = 0.85) -> torch.Tensor:
        return _ops.silog_loss(prediction, target, valid_mask, variance_focus)

    def multi_scale_loss(self, pred_depths: List[torch.Tensor], gt_depth: torch.Tensor, rgb: torch.Tensor,
                         valid_mask: torch.Tensor):
        B, _, H, W = gt_depth.shape
        total_smooth = total_silog = 0.0
        for i, alpha in enumerate(self.alphas):
            pred_depth = _ops.upsample_bilinear(pred_depths[i], (H, W))
            total_smooth = total_smooth + alpha * self.get_smooth_loss(pred_depth, rgb)
            total_silog = total_silog + alpha * self.silog_loss(pred_depth, gt_depth, valid_mask)
        total_loss = self.silog_weight * total_silog + self.smooth_weight * total_smooth
        return total_loss, total_silog, total_smooth

    def forward_step(self, sample: Dict[str, torch.Tensor]):
        rgb = sample["image"].to(self.device)
        depth = sample["depth"].to(self.device)
        valid_mask = sample["valid_mask"].to(self.device)
        if depth.dim() == 3:
            depth, valid_mask = depth.unsqueeze(1), valid_mask.unsqueeze(1) if valid_mask.dim() == 3 else valid_mask
        outputs = self.model(rgb)
        pred_depths = [self.disp_to_depth(outputs[("disp", s)]) for s in range(self.num_scales)]
        total_loss, total_silog, total_smooth = self.multi_scale_loss(pred_depths, depth, rgb, valid_mask)
        return total_loss, total_silog, total_smooth, pred_depths

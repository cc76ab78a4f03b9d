"""Tiny stand-in networks for the ``process_batch`` parity case (test infrastructure).

``MonodepthTrainer.process_batch`` (vo/learner_new.py:76-105) only needs ``depth_net(target) -> {("disp", s)}`` and
``pose_net(pair) -> (axisangle [B,1,1,3], translation [B,1,1,3])``.  These two modules have a few dozen weights, sit
on top of a plausible disparity pyramid / pose pair (buffers) so that the reprojection branch wins at many pixels, and
give the golden case something to back-propagate into: the gradients of their weights are sums over every pixel of
the loss's disparity and pose gradients."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class TinyDepthNet(nn.Module):
    def __init__(self, base_disps):
        super().__init__()
        for s, d in enumerate(base_disps):
            self.register_buffer(f"base{s}", torch.logit(d.clamp(1e-3, 1 - 1e-3)))
        self.heads = nn.ModuleList([nn.Conv2d(3, 1, 3, padding=1) for _ in base_disps])
        self.gain = nn.Parameter(torch.full((len(base_disps),), 0.2))

    def forward(self, x):
        out = {}
        for s, head in enumerate(self.heads):
            xs = F.avg_pool2d(x, 2 ** s) if s else x
            out[("disp", s)] = torch.sigmoid(getattr(self, f"base{s}") + self.gain[s] * head(xs))
        return out


class TinyPoseNet(nn.Module):
    """Called twice per step: (source_left, target) then (target, source_right), as _predict_poses does."""

    def __init__(self, base_axisangle, base_translation):
        super().__init__()
        self.register_buffer("base", torch.stack([torch.cat([a, t], -1) for a, t in zip(base_axisangle, base_translation)]))
        self.conv = nn.Conv2d(6, 6, 3, stride=2)
        self.calls = 0

    def forward(self, pair):
        base = self.base[self.calls % self.base.shape[0]]                 # [B,1,6]
        self.calls += 1
        out = base.unsqueeze(1) + 0.01 * torch.tanh(self.conv(pair)).mean(3).mean(2).view(-1, 1, 1, 6)
        return out[..., :3], out[..., 3:]

"""PoseNet: ResNet encoder on a stacked frame pair + 4-convolution pose head -- stock PyTorch.
Same constructor and outputs as the reference (model/posenet_single.py:149-200): ``(axisangle, translation)``,
each ``[B, 1, 1, 3]``, scaled by 0.01.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .resnet_encoder import ResnetEncoder


class PoseNet(nn.Module):
    def __init__(self, num_layers: int = 18, pretrained: bool = True, num_input_images: int = 2, stride: int = 1):
        super().__init__()
        self.num_input_images = num_input_images
        self.encoder = ResnetEncoder(num_layers, pretrained, num_input_images)
        self.num_ch_enc = self.encoder.num_ch_enc
        # ``self.net`` in this order is the checkpoint format (keys ``net.0..3.*``; reference:
        # model/posenet_single.py:166-173, loaded by vo/train.py:91-98)
        self.net = nn.ModuleList([nn.Conv2d(int(self.num_ch_enc[-1]), 256, 1), nn.Conv2d(256, 256, 3, stride, 1),
                                  nn.Conv2d(256, 256, 3, stride, 1), nn.Conv2d(256, 6, 1)])

    def forward(self, input_images: torch.Tensor, normalized: bool = False):
        squeeze, pose0, pose1, pose2 = self.net
        feats = self.encoder(input_images, normalized) if normalized else self.encoder(input_images)
        x = torch.relu(squeeze(feats[-1]))
        x = torch.relu(pose0(x))
        x = torch.relu(pose1(x))
        out = 0.01 * pose2(x).mean(3).mean(2).view(-1, 1, 1, 6)
        return out[..., :3], out[..., 3:]


class FlowPoseNet(nn.Module):
    """Import shim.  The reference's ``vo/train.py:18`` imports this name but never constructs it; the class itself
    (model/posenet_single.py:91-147) wraps RAFT and a ``./raft-small.pth`` download, which is outside the accelerated
    path (RAFT out of scope).  Constructing it here says so instead of failing at import time."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("FlowPoseNet (RAFT optical-flow pose regressor) is not part of the B200 view-synthesis "
                                  "path; use PoseNet, or import FlowPoseNet from the reference tree")

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
        self.squeeze = nn.Conv2d(int(self.num_ch_enc[-1]), 256, 1)
        self.pose0 = nn.Conv2d(256, 256, 3, stride, 1)
        self.pose1 = nn.Conv2d(256, 256, 3, stride, 1)
        self.pose2 = nn.Conv2d(256, 6, 1)

    def forward(self, input_images: torch.Tensor):
        x = torch.relu(self.squeeze(self.encoder(input_images)[-1]))
        x = torch.relu(self.pose0(x))
        x = torch.relu(self.pose1(x))
        out = 0.01 * self.pose2(x).mean(3).mean(2).view(-1, 1, 1, 6)
        return out[..., :3], out[..., 3:]

"""ResNet feature pyramid for DepthNet / PoseNet -- stock PyTorch convolutions (north_star: the networks are
not part of the accelerated path).  Same constructor and outputs as the reference's ``ResnetEncoder``
(model/resnet_encoder.py:75-112): five feature maps at strides 2..32 with ``num_ch_enc`` channels, input
normalised with ``(x - 0.45) / 0.225``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torchvision.models as tvm

_RESNETS = {18: tvm.resnet18, 34: tvm.resnet34, 50: tvm.resnet50, 101: tvm.resnet101, 152: tvm.resnet152}


def _build(num_layers: int, pretrained: bool, num_input_images: int) -> nn.Module:
    if num_layers not in _RESNETS:
        raise ValueError(f"{num_layers} is not a valid number of resnet layers")
    net = _RESNETS[num_layers](weights="IMAGENET1K_V1" if pretrained else None)
    if num_input_images > 1:
        # several RGB frames stacked on the channel axis: widen the stem and spread the (pretrained) filters
        old = net.conv1
        conv = nn.Conv2d(3 * num_input_images, 64, kernel_size=7, stride=2, padding=3, bias=False)
        with torch.no_grad():
            if pretrained:
                conv.weight.copy_(torch.cat([old.weight] * num_input_images, 1) / num_input_images)
            else:
                nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")
        net.conv1 = conv
    # the classifier head is never used (kept so that checkpoints of the reference load unchanged); frozen so that
    # DistributedDataParallel does not wait for a gradient that never comes
    for q in net.fc.parameters():
        q.requires_grad_(False)
    return net


class ResnetEncoder(nn.Module):
    def __init__(self, num_layers: int, pretrained: bool, num_input_images: int = 1):
        super().__init__()
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        if num_layers > 34:
            self.num_ch_enc[1:] *= 4
        self.encoder = _build(num_layers, pretrained, num_input_images)

    def forward(self, input_image: torch.Tensor, normalized: bool = False):
        """``normalized=True``: the input already went through ``(x - 0.45) / 0.225`` (dvsloss.ops.pack_net_inputs does it
        while it builds the channels-last network inputs)."""
        e = self.encoder
        x = input_image if normalized else (input_image - 0.45) / 0.225
        f0 = e.relu(e.bn1(e.conv1(x)))
        f1 = e.layer1(e.maxpool(f0))
        f2 = e.layer2(f1)
        f3 = e.layer3(f2)
        f4 = e.layer4(f3)
        self.features = [f0, f1, f2, f3, f4]
        return self.features

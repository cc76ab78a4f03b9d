"""DepthNet: ResNet encoder + Monodepth2-style decoder with four sigmoid disparity heads -- stock PyTorch.
Same constructor and output dict as the reference (model/depthnet.py:16-90): ``("disp", s)`` of shape
``[B, 1, H / 2**s, W / 2**s]`` for s in ``scales``.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from .layers import Conv3x3, ConvBlock, upsample            # (puts the package root on sys.path for dvsloss)
from dvsloss.ops import disp_head, disp_head_supported, elu_up2_cat, elu_up2_cat_supported  # noqa: E402
from .resnet_encoder import ResnetEncoder


class DepthNet(nn.Module):
    # the ("dispconv", s) + sigmoid tails run as one fused kernel on CUDA (dvsloss.ops.disp_head); False = stock modules
    fused_heads = True
    # ELU of ("upconv", i, 0) + nearest up-sampling + concatenation with the encoder skip as one kernel each way
    # (dvsloss.ops.elu_up2_cat); False = the stock three-op sequence
    fused_glue = True

    def __init__(self, num_layers: int = 18, pretrained: bool = True, num_input_images: int = 1, scales=range(4),
                 num_output_channels: int = 1, use_skips: bool = True):
        super().__init__()
        self.scales = list(scales)
        self.use_skips = use_skips
        self.num_output_channels = num_output_channels
        self.encoder = ResnetEncoder(num_layers, pretrained, num_input_images)
        self.num_ch_enc = self.encoder.num_ch_enc
        self.num_ch_dec = np.array([16, 32, 64, 128, 256])
        # Registration order and attribute names are the checkpoint format (vo/train.py:83-98 loads
        # ``depth_net_epoch_N.pth`` with keys ``decoder.<k>.conv.conv.*`` / ``decoder.<k>.conv.*``): the blocks are
        # registered as ``self.decoder`` in the order (upconv,4,0), (upconv,4,1), ..., (upconv,0,1), dispconv 0..3
        # (reference: model/depthnet.py:41-60) and looked up through the tuple-keyed ``self.convs``.
        self.convs = OrderedDict()
        for i in range(4, -1, -1):
            c_in = self.num_ch_enc[-1] if i == 4 else self.num_ch_dec[i + 1]
            self.convs[("upconv", i, 0)] = ConvBlock(c_in, self.num_ch_dec[i])
            c_in = self.num_ch_dec[i] + (self.num_ch_enc[i - 1] if use_skips and i > 0 else 0)
            self.convs[("upconv", i, 1)] = ConvBlock(c_in, self.num_ch_dec[i])
        for s in self.scales:
            self.convs[("dispconv", s)] = Conv3x3(self.num_ch_dec[s], num_output_channels)
        self.decoder = nn.ModuleList(self.convs.values())

    def forward(self, input_data: torch.Tensor, normalized: bool = False) -> dict:
        feats = self.encoder(input_data, normalized) if normalized else self.encoder(input_data)
        outputs = {}
        x = feats[-1]
        for i in range(4, -1, -1):
            block = self.convs[("upconv", i, 0)]
            skip = feats[i - 1] if self.use_skips and i > 0 else None
            pre = block.conv(x, with_bias=False) if self.fused_glue and x.is_cuda else None
            if pre is not None and elu_up2_cat_supported(pre, skip):
                x = elu_up2_cat(pre, skip, block.conv.conv.bias)          # the convolution's bias rides along
            else:
                if pre is not None:
                    pre = block.nonlin(pre + block.conv.conv.bias.to(pre.dtype).view(1, -1, 1, 1))
                x = upsample(pre if pre is not None else block(x))
                if skip is not None:
                    x = torch.cat([x, skip], 1)
            x = self.convs[("upconv", i, 1)](x)
            if i in self.scales:
                head = self.convs[("dispconv", i)]
                if self.fused_heads and self.num_output_channels == 1 and head.use_refl and disp_head_supported(x):
                    # pad + one-channel convolution + sigmoid in one kernel, written in the dtype the loss kernel reads
                    w = head.conv.weight
                    if torch.is_autocast_enabled() and x.dtype == torch.float32:
                        x_in = x.to(torch.get_autocast_dtype("cuda"))     # what autocast would feed the convolution
                    else:
                        x_in = x
                    outputs[("disp", i)] = disp_head(x_in, w, head.conv.bias)
                else:
                    outputs[("disp", i)] = torch.sigmoid(head(x))
        self.outputs = outputs
        return outputs

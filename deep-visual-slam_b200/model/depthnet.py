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
from dvsloss.ops import bias_elu, bias_elu_supported, disp_head, disp_head_supported, elu_up2_cat, elu_up2_cat_supported  # noqa: E402
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

    # the fused glue kernels write the reflected ring of their outputs themselves, so every decoder convolution after the first
    # runs as a plain un-padded stock convolution: no padded copies, no border fix-up convolutions (False = border strips)
    padded_activations = True

    def _padded_path_ok(self, feats) -> bool:
        from .layers import ConvBlock
        x = feats[-1]
        if not (self.padded_activations and self.fused_glue and self.fused_heads and ConvBlock.fused_bias_elu and x.is_cuda
                and self.num_output_channels == 1 and x.shape[2] >= 2 and x.shape[3] >= 2):
            return False
        dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
        if dt not in (torch.float32, torch.bfloat16):
            return False
        per = 8 if dt == torch.bfloat16 else 4
        pow2 = lambda c: c % per == 0 and ((c // per) & (c // per - 1)) == 0 and c // per <= 256
        chans = [int(c) for c in self.num_ch_dec]
        skips = [int(c) for c in self.num_ch_enc[:4]] if self.use_skips else []
        return (all(pow2(c) for c in chans) and all(c % per == 0 for c in skips) and all(int(self.num_ch_dec[s]) in (8, 16, 32, 64, 128)
                for s in self.scales) and all(self.convs[("dispconv", s)].use_refl for s in self.scales))

    # which decoder convolutions read a ring-carrying input: bit i = ("upconv", i, 0) for i < 4 (its producer is the
    # bias + ELU kernel of stage i + 1), bit 4 + i = ("upconv", i, 1) (its producer is the ELU + up-sample + concatenate kernel of
    # stage i).  The others pad by border strips: measured (profiles/tools/strip_cost.py, 32 x 640x480) 59.5 ms per training step
    # with strips everywhere, 57.7 with every bit set, 57.3 with the two full-resolution convolutions of stage 0 left on strips
    # (cuDNN's un-padded 482 x 642 kernels for 16 channels are slower than its padded 480 x 640 ones), 56.0 with no padding at all.
    padded_mask = 0x1EE

    def _forward_padded(self, feats) -> dict:
        """Decoder with ring-carrying activations (see ``padded_activations``)."""
        outputs = {}
        m = self.padded_mask
        x, padded = feats[-1], False
        for i in range(4, -1, -1):
            b0, b1 = self.convs[("upconv", i, 0)], self.convs[("upconv", i, 1)]
            skip = feats[i - 1] if self.use_skips and i > 0 else None
            pre = b0.pre_activation(x, padded)                                  # the first one still pads by border strips
            if skip is not None and skip.dtype != pre.dtype:
                skip = skip.to(pre.dtype)
            p1 = bool((m >> (4 + i)) & 1)
            x = elu_up2_cat(pre, skip, b0.conv.conv.bias, pad=p1)              # [B, C1+C2, 2h(+2), 2w(+2)]
            padded = bool(i > 0 and (m >> (i - 1)) & 1)                         # what the next stage's first convolution reads
            x = bias_elu(b1.pre_activation(x, p1), b1.conv.conv.bias, pad=padded)
            if i in self.scales:
                head = self.convs[("dispconv", i)]
                outputs[("disp", i)] = disp_head(x, head.conv.weight, head.conv.bias, padded=padded)
        return outputs

    def forward(self, input_data: torch.Tensor, normalized: bool = False) -> dict:
        feats = self.encoder(input_data, normalized) if normalized else self.encoder(input_data)
        if self._padded_path_ok(feats):
            self.outputs = self._forward_padded(feats)
            return self.outputs
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

"""Fused disparity head (reflection pad + 3x3 one-channel convolution + sigmoid, fwd + bwd) against the
stock modules of the reference (model/layers.py:120-136 Conv3x3, model/depthnet.py:57-58,87-88)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _stock(x, w, b):
    return torch.sigmoid(F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b))


@pytest.mark.parametrize("B,C,H,W", [(2, 16, 48, 64), (1, 32, 7, 9), (2, 128, 12, 20), (1, 64, 3, 3), (3, 8, 33, 31)])
@pytest.mark.parametrize("channels_last", [True, False])
def test_disp_head_fp32_matches_stock_modules(B, C, H, W, channels_last):
    from dvsloss.ops import disp_head
    torch.manual_seed(C + H)
    dev = torch.device("cuda:0")
    x = torch.randn(B, C, H, W, device=dev)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    w = (torch.randn(1, C, 3, 3, device=dev) / (3 * C ** 0.5)).requires_grad_(True)
    b = torch.randn(1, device=dev).requires_grad_(True)
    x.requires_grad_(True)
    g = torch.randn(B, 1, H, W, device=dev)
    got = disp_head(x, w, b)
    ggot = torch.autograd.grad(got, (x, w, b), g)
    # oracle in float64: the stock op sequence
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    ref = _stock(xd, wd, bd)
    gref = torch.autograd.grad(ref, (xd, wd, bd), g.double())
    assert got.shape == (B, 1, H, W) and got.dtype == torch.float32
    assert float((got.double() - ref).abs().max()) < 2e-6                       # fp32 dot product of 9 C terms + __expf
    for a, r, name in zip(ggot, gref, ("x", "weight", "bias")):
        scale = float(r.abs().max()) + 1e-30
        assert float((a.double() - r).abs().max()) <= 2e-5 * scale, name
    # fixed-order reductions: bit-reproducible weight gradients
    again = torch.autograd.grad(disp_head(x, w, b), (w, b), g)
    assert torch.equal(again[0], ggot[1]) and torch.equal(again[1], ggot[2])


def test_disp_head_bf16_matches_autocast_modules():
    """Under bf16 autocast the stock path rounds the convolution output and the sigmoid to bf16; the fused head keeps fp32
    until the final store, so it must agree with the float64 oracle on the SAME bf16 inputs to bf16 resolution (2^-8 relative on
    a value in (0,1)), and so must its gradients (their inputs g and d are bf16-rounded: 3 roundings, 1.5e-2 of the maximum)."""
    from dvsloss.ops import disp_head
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    B, C, H, W = 2, 16, 40, 56
    x = torch.randn(B, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(1, C, 3, 3, device=dev) / 12).requires_grad_(True)
    b = torch.zeros(1, device=dev).requires_grad_(True)
    g = torch.randn(B, 1, H, W, device=dev).to(torch.bfloat16)
    got = disp_head(x, w, b)
    assert got.dtype == torch.bfloat16
    ggot = torch.autograd.grad(got, (x, w, b), g)
    assert ggot[0].dtype == torch.bfloat16 and ggot[0].is_contiguous(memory_format=torch.channels_last) and ggot[1].dtype == torch.float32
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    ref = _stock(xd, wd, bd)
    gref = torch.autograd.grad(ref, (xd, wd, bd), g.double())
    assert float((got.double() - ref).abs().max()) <= 2 ** -8
    for a, r in zip(ggot, gref):
        assert float((a.double() - r).abs().max()) <= 1.5e-2 * float(r.abs().max())


def test_depthnet_uses_the_fused_head_and_trains_like_the_stock_net():
    from model.depthnet import DepthNet
    torch.manual_seed(1)
    dev = torch.device("cuda:0")
    net = DepthNet(18, False).to(dev).eval()
    x = torch.rand(2, 3, 64, 96, device=dev)
    outs = {}
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the stock heads would otherwise run in TF32 (1e-3); the fused one is fp32
    for fused in (True, False):
        DepthNet.fused_heads = fused
        try:
            net.zero_grad()
            o = net(x)
            sum(v.mean() for v in o.values()).backward()
            outs[fused] = ({k: v.detach().clone() for k, v in o.items()},
                           {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None})
        finally:
            DepthNet.fused_heads = True
    torch.backends.cudnn.allow_tf32 = tf32
    for k in outs[False][0]:
        assert torch.allclose(outs[True][0][k], outs[False][0][k], rtol=0, atol=5e-6), k
    for n, gr in outs[False][1].items():
        scale = float(gr.abs().max()) + 1e-12
        assert float((outs[True][1][n] - gr).abs().max()) <= 1e-3 * scale, n

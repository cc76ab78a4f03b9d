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


def _stock_glue(x, skip, bias=None):
    if bias is not None:
        x = x + bias.to(x.dtype).view(1, -1, 1, 1)
    y = F.interpolate(F.elu(x), scale_factor=2, mode="nearest")
    return y if skip is None else torch.cat([y, skip], 1)


@pytest.mark.parametrize("B,C1,C2,h,w", [(2, 16, 0, 12, 16), (2, 32, 64, 6, 8), (1, 256, 256, 2, 3), (3, 64, 64, 5, 7), (1, 4, 8, 3, 2)])
@pytest.mark.parametrize("channels_last", [True, False])
def test_elu_up2_cat_fp32_matches_stock_ops(B, C1, C2, h, w, channels_last):
    """ConvBlock's ELU + nearest x2 + concatenation with the skip (model/depthnet.py:77-84) as one kernel each way: the
    forward is a copy of fp32 values and one expf - 1 (ATen's formula); the backward adds four gradients in a fixed order."""
    from dvsloss.ops import elu_up2_cat
    torch.manual_seed(C1 + h)
    dev = torch.device("cuda:0")
    x = torch.randn(B, C1, h, w, device=dev)
    skip = torch.randn(B, C2, 2 * h, 2 * w, device=dev) if C2 else None
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
        skip = skip.contiguous(memory_format=torch.channels_last) if skip is not None else None
    bias = torch.randn(C1, device=dev) if (h + C1) % 2 else None            # the convolution's bias folded in, or not
    leaves = [t.requires_grad_(True) for t in (x, skip, bias) if t is not None]
    got = elu_up2_cat(x, skip, bias)
    ref = _stock_glue(x, skip, bias)
    assert got.shape == ref.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert float((got - ref).abs().max()) <= 4e-7
    g = torch.randn_like(ref)
    gg = torch.autograd.grad(got, leaves, g)
    gr = torch.autograd.grad(ref, leaves, g)
    for a, r in zip(gg, gr):
        assert float((a - r).abs().max()) <= 2e-6 * (float(r.abs().max()) + 1e-30)
    assert all(torch.equal(a, b_) for a, b_ in zip(gg, torch.autograd.grad(elu_up2_cat(x, skip, bias), leaves, g)))   # fixed-order sums


def test_elu_up2_cat_bf16_matches_stock_ops_within_rounding():
    from dvsloss.ops import elu_up2_cat
    torch.manual_seed(3)
    dev = torch.device("cuda:0")
    x = torch.randn(2, 64, 9, 11, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    skip = torch.randn(2, 64, 18, 22, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    got = elu_up2_cat(x, skip)
    ref = _stock_glue(x, skip)
    assert got.dtype == torch.bfloat16
    assert float((got.float() - ref.float()).abs().max()) <= 2 ** -8 * float(ref.float().abs().max())      # one bf16 rounding
    g = torch.randn_like(ref)
    gg = torch.autograd.grad(got, (x, skip), g)
    # oracle: the same chain in float64 on the bf16 inputs
    xd, sd = x.detach().double().requires_grad_(True), skip.detach().double().requires_grad_(True)
    gr = torch.autograd.grad(_stock_glue(xd, sd), (xd, sd), g.double())
    assert torch.equal(gg[1], g[:, 64:])                                                                   # a copy
    assert float((gg[0].double() - gr[0]).abs().max()) <= 2 ** -7 * float(gr[0].abs().max())               # fp32 sum, one bf16 rounding


@pytest.mark.parametrize("B,C,H,W,dtype", [(2, 16, 12, 20, torch.float32), (1, 256, 3, 5, torch.float32), (3, 64, 9, 7, torch.bfloat16),
                                            (2, 32, 33, 17, torch.bfloat16)])
def test_bias_elu_matches_stock_ops(B, C, H, W, dtype):
    """ConvBlock's bias + ELU as one kernel each way; the backward is written from the output like nn.ELU(inplace=True)."""
    from dvsloss.ops import bias_elu
    torch.manual_seed(C + W)
    dev = torch.device("cuda:0")
    x = torch.randn(B, C, H, W, device=dev).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bias = torch.randn(C, device=dev).requires_grad_(True)
    got = bias_elu(x, bias)
    xd, bd = x.detach().double().requires_grad_(True), bias.detach().double().requires_grad_(True)
    ref = F.elu(xd + bd.view(1, -1, 1, 1))
    tol = 2e-7 if dtype == torch.float32 else 2 ** -8
    assert got.dtype == dtype and float((got.double() - ref).abs().max()) <= tol * (1 + float(ref.abs().max()))
    g = torch.randn(B, C, H, W, device=dev).to(dtype)
    gx, gb = torch.autograd.grad(got, (x, bias), g)
    rx, rb = torch.autograd.grad(ref, (xd, bd), g.double())
    gtol = 2e-6 if dtype == torch.float32 else 2 ** -6                       # bf16: y + 1 is rounded once more, as in ATen's in-place ELU
    assert float((gx.double() - rx).abs().max()) <= gtol * float(rx.abs().max())
    assert float((gb.double() - rb).abs().max()) <= gtol * float(rb.abs().max()) * (1 if dtype == torch.float32 else 4)
    gx2, gb2 = torch.autograd.grad(bias_elu(x, bias), (x, bias), g)
    assert torch.equal(gx, gx2) and torch.equal(gb, gb2)                     # fixed-order reduction


def test_depthnet_fused_glue_trains_like_the_stock_sequence():
    from model.depthnet import DepthNet
    torch.manual_seed(2)
    dev = torch.device("cuda:0")
    net = DepthNet(18, False).to(dev).eval().to(memory_format=torch.channels_last)
    x = torch.rand(2, 3, 64, 96, device=dev).contiguous(memory_format=torch.channels_last)
    outs = {}
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    from model.layers import ConvBlock
    for fused in (True, False):
        DepthNet.fused_glue = ConvBlock.fused_bias_elu = fused
        try:
            net.zero_grad()
            o = net(x)
            sum(v.mean() for v in o.values()).backward()
            outs[fused] = ({k: v.detach().clone() for k, v in o.items()},
                           {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None})
        finally:
            DepthNet.fused_glue = ConvBlock.fused_bias_elu = True
    torch.backends.cudnn.allow_tf32 = tf32
    for k in outs[False][0]:
        assert torch.allclose(outs[True][0][k], outs[False][0][k], rtol=0, atol=5e-6), k
    for n, gr in outs[False][1].items():
        scale = float(gr.abs().max()) + 1e-12
        assert float((outs[True][1][n] - gr).abs().max()) <= 1e-3 * scale, n


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,C1,C2,h,w", [(2, 16, 0, 6, 9), (1, 32, 64, 2, 2), (2, 64, 64, 5, 3), (1, 8, 16, 3, 7)])
def test_glue_kernels_write_and_fold_the_reflected_ring(B, C1, C2, h, w, dtype):
    """pad=True: the outputs of elu_up2_cat / bias_elu ARE ReflectionPad2d(1)(stock result), and their backward folds the ring's
    gradients onto the mirrored pixels -- what lets every later decoder convolution run un-padded (model/layers.py:126-136)."""
    from dvsloss.ops import bias_elu, elu_up2_cat
    if dtype == torch.float32 and (C1 % 4 or (C1 // 4) & (C1 // 4 - 1)):
        pytest.skip("channel count not supported in fp32")
    torch.manual_seed(C1 + h + w)
    dev = torch.device("cuda:0")
    mk = lambda *s: torch.randn(*s, device=dev).to(dtype).contiguous(memory_format=torch.channels_last)
    x = mk(B, C1, h, w).requires_grad_(True)
    skip = mk(B, C2, 2 * h, 2 * w).requires_grad_(True) if C2 else None
    bias = torch.randn(C1, device=dev).requires_grad_(True)
    leaves = [t for t in (x, skip, bias) if t is not None]
    got = elu_up2_cat(x, skip, bias, pad=True)
    dense = elu_up2_cat(x, skip, bias)
    ref = F.pad(dense, (1, 1, 1, 1), mode="reflect")
    assert got.shape == ref.shape and torch.equal(got, ref)                       # the ring is a copy of computed values
    g = torch.randn_like(ref)
    gg = torch.autograd.grad(got, leaves, g)
    gr = torch.autograd.grad(ref, leaves, g)
    tol = 2e-6 if dtype == torch.float32 else 2 ** -6
    for a, r in zip(gg, gr):
        assert float((a.double() - r.double()).abs().max()) <= tol * (float(r.double().abs().max()) + 1e-30)
    # bias + ELU with the ring
    C = C1
    z = mk(B, C, 2 * h + 1, 2 * w + 1).requires_grad_(True)
    got2 = bias_elu(z, bias, pad=True)
    ref2 = F.pad(bias_elu(z, bias), (1, 1, 1, 1), mode="reflect")
    assert torch.equal(got2, ref2)
    g2 = torch.randn_like(ref2)
    for a, r in zip(torch.autograd.grad(got2, (z, bias), g2), torch.autograd.grad(ref2, (z, bias), g2)):
        assert float((a.double() - r.double()).abs().max()) <= tol * (float(r.double().abs().max()) + 1e-30)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_disp_head_reads_a_ring_carrying_activation(dtype):
    from dvsloss.ops import bias_elu, disp_head
    torch.manual_seed(5)
    dev = torch.device("cuda:0")
    B, C, H, W = 2, 32, 11, 37
    pre = torch.randn(B, C, H, W, device=dev).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(1, C, 3, 3, device=dev) / 30).requires_grad_(True)
    b = torch.randn(1, device=dev).requires_grad_(True)
    outs = []
    for pad in (True, False):
        y = bias_elu(pre, None, pad=pad)
        d = disp_head(y, w, b, padded=pad)
        g = torch.autograd.grad(d, (pre, w, b), torch.ones_like(d) * 0.3)
        outs.append((d, g))
    assert torch.equal(outs[0][0], outs[1][0])
    for a, r in zip(outs[0][1], outs[1][1]):
        assert float((a.double() - r.double()).abs().max()) <= (1e-6 if dtype == torch.float32 else 2 ** -7) * float(r.double().abs().max())


@pytest.mark.parametrize("cout", [16, 24, 6])
def test_convblock_fused_and_fallback_paths_equal_the_stock_block(cout):
    """ConvBlock = Conv3x3 + ELU (model/layers.py:106-117): 16 output channels take the fused bias + ELU kernel; 24 (6 four-channel
    vectors: not a power of two, which the bias-gradient reduction wants) and 6 (not a multiple of 4) take the stock fall-backs;
    all equal the stock block."""
    from model.layers import ConvBlock
    torch.manual_seed(cout)
    dev = torch.device("cuda:0")
    blk = ConvBlock(8, cout).to(dev)
    x = torch.randn(2, 8, 13, 17, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    outs = []
    try:
        for fused in (True, False):
            ConvBlock.fused_bias_elu = fused
            blk.zero_grad()
            y = blk(x)
            g = torch.autograd.grad(y, [x] + list(blk.parameters()), torch.ones_like(y) * 0.1)
            outs.append((y.detach().clone(), [t.clone() for t in g]))
    finally:
        ConvBlock.fused_bias_elu = True
        torch.backends.cudnn.allow_tf32 = tf32
    assert float((outs[0][0] - outs[1][0]).abs().max()) <= 2e-6
    for a, r in zip(outs[0][1], outs[1][1]):
        assert float((a - r).abs().max()) <= 1e-5 * (float(r.abs().max()) + 1e-12)


def test_depthnet_resnet50_fused_decoder_equals_the_stock_sequence():
    """BASELINE configs[3] uses the ResNet-50 encoder (skip channels 64 / 256 / 512 / 1024): the fused decoder path (glue kernels,
    ring-carrying activations, fused heads) against the stock sequence, fp32, values and every parameter gradient."""
    from model.depthnet import DepthNet
    from model.layers import ConvBlock
    torch.manual_seed(4)
    dev = torch.device("cuda:0")
    net = DepthNet(50, False).to(dev).eval().to(memory_format=torch.channels_last)
    x = torch.rand(1, 3, 64, 96, device=dev).contiguous(memory_format=torch.channels_last)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    outs = {}
    try:
        for fused in (True, False):
            DepthNet.fused_glue = DepthNet.fused_heads = DepthNet.padded_activations = ConvBlock.fused_bias_elu = fused
            net.zero_grad()
            o = net(x)
            sum(v.mean() for v in o.values()).backward()
            outs[fused] = ({k: v.detach().clone() for k, v in o.items()},
                           {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None})
    finally:
        DepthNet.fused_glue = DepthNet.fused_heads = DepthNet.padded_activations = ConvBlock.fused_bias_elu = True
        torch.backends.cudnn.allow_tf32 = tf32
    for k in outs[False][0]:
        assert torch.allclose(outs[True][0][k], outs[False][0][k], rtol=0, atol=1e-5), k
    for n, gr in outs[False][1].items():
        assert float((outs[True][1][n] - gr).abs().max()) <= 2e-3 * (float(gr.abs().max()) + 1e-12), n

"""``MonodepthTrainer.process_batch`` on the GPU against the UNMODIFIED reference's ``process_batch`` (golden written by
tests/golden/make_golden.py with the tiny networks of tests/tiny_nets.py): losses, gradients reaching the network
weights, and the lazily materialised ``outputs`` the reference's plotting code reads (vo/utils/plot_utils.py:40-47)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _RandnQueue:
    """torch.randn returns the golden's noise tensors (moved to the requested device), as make_golden.py fed the reference."""

    def __init__(self, queue):
        self.queue, self.orig = list(queue), torch.randn

    def __enter__(self):
        def fake(*shape, **kw):
            t = self.queue.pop(0)
            return t.clone().to(kw.get("device", "cpu"))
        torch.randn = fake
        return self

    def __exit__(self, *a):
        torch.randn = self.orig


def _load():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_process_batch_b2_48x64.npz"))
    t = lambda k: torch.from_numpy(z[k])
    sample = {}
    for k in z.files:
        if k.startswith("sample/"):
            _, name, idx = k.split("/")
            sample[(name, int(idx))] = t(k)
    return z, t, sample


def _nets(z, t):
    from tiny_nets import TinyDepthNet, TinyPoseNet
    dstate = {k[len("depth_net/state/"):]: t(k) for k in z.files if k.startswith("depth_net/state/")}
    pstate = {k[len("pose_net/state/"):]: t(k) for k in z.files if k.startswith("pose_net/state/")}
    dnet = TinyDepthNet([torch.sigmoid(dstate[f"base{s}"]) for s in range(4)])
    pnet = TinyPoseNet([pstate["base"][i][..., :3] for i in range(2)], [pstate["base"][i][..., 3:] for i in range(2)])
    dnet.load_state_dict(dstate)
    pnet.load_state_dict(pstate)
    return dnet.cuda(), pnet.cuda()


def _learner(z, t, **kw):
    from vo.learner_new import MonodepthTrainer
    dnet, pnet = _nets(z, t)
    B, _, H, W = z["sample/target_image/0"].shape
    cfg = {"Train": dict(num_source=2, batch_size=B, img_h=H, img_w=W, smoothness_ratio=0.001, auto_mask=True,
                         ssim_ratio=0.85, min_depth=0.1, max_depth=10.0, use_compile=False)}
    return MonodepthTrainer(dnet, pnet, cfg, torch.device("cuda", 0), noise="torch", **kw), dnet, pnet


@pytest.mark.parametrize("fused", [True, False])
def test_process_batch_matches_live_reference(fused):
    z, t, sample = _load()
    learner, dnet, pnet = _learner(z, t, fused=fused)
    B, _, H, W = z["sample/target_image/0"].shape
    with _RandnQueue([t(f"noise{s}") for s in range(4)]):
        outputs, losses = learner.process_batch(sample)
    assert all(v.is_cuda for v in sample.values())                       # moved in place in the dict (vo/learner_new.py:93-95)
    losses["loss"].backward()
    pix_noise = 4 * 2e-5 / np.sqrt(B * H * W)                            # tests/parity.py: fp32 SSIM round-off on tiny images
    for k in ["loss"] + [f"loss/{s}" for s in range(4)]:
        ref = float(z[k])
        assert losses[k].dim() == 0 and losses[k].is_cuda
        assert abs(float(losses[k]) - ref) <= 1e-5 * abs(ref) + pix_noise, (k, float(losses[k]), ref)
    for s in range(4):
        got, ref = outputs[f"identity_selection/{s}"].cpu().numpy(), z[f"identity_selection/{s}"]
        assert got.shape == ref.shape and (got != ref).mean() < 0.01, s
    # gradients that reach the network weights: sums over every pixel of the loss's disparity / pose gradients
    # (a bias gradient is a plain sum of per-pixel gradients of both signs: its rounding error scales with the size of the
    # terms, not of the sum, hence the second term relative to the largest gradient of the same network)
    for net, tag in ((dnet, "depth_net"), (pnet, "pose_net")):
        net_scale = max(np.abs(z[f"{tag}/grad/{name}"]).max() for name, _ in net.named_parameters())
        for name, q in net.named_parameters():
            ref = z[f"{tag}/grad/{name}"]
            got = q.grad.cpu().numpy()
            scale = np.abs(ref).max()
            assert np.abs(got - ref).max() <= 2e-3 * scale + 2e-4 * net_scale, (tag, name, np.abs(got - ref).max() / scale)


def test_outputs_fill_lazily_for_the_unchanged_plot_call():
    """vo/train.py:268-279 -> PlotTool.plot_result reads outputs[("depth", s)] and outputs[("color", +-1, s)] on step 0 and
    every 1000th; the fused step does not materialise them, the dict does on first access."""
    z, t, sample = _load()
    learner, _, _ = _learner(z, t)
    with _RandnQueue([t(f"noise{s}") for s in range(4)]):
        outputs, _ = learner.process_batch(sample)
    assert not dict.__contains__(outputs, ("depth", 0))                  # nothing full-resolution on an ordinary step
    for s in range(4):                                                    # the accesses of vo/utils/plot_utils.py:40-47
        depth = outputs[("depth", s)][0].detach().cpu()
        left = outputs[("color", -1, s)][0].detach().cpu()
        right = outputs[("color", 1, s)][0].detach().cpu()
        assert right.shape == left.shape == (3,) + depth.shape[1:]
        assert np.allclose(depth.numpy(), z[f"depth{s}"][0], rtol=2e-5, atol=1e-6)
        assert np.abs(left.numpy() - z[f"color{s}_-1"][0]).max() < 5e-4
    for s in range(4):
        for f in (-1, 1):
            assert outputs[("sample", f, s)].shape == (2, 48, 64, 2)
            assert outputs[("color_identity", f, s)] is sample[("source_left" if f < 0 else "source_right", 0)]
        assert outputs[("disp_up", s)].shape == (2, 1, 48, 64)

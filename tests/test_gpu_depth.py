"""Supervised-depth learner and eval back-projection against the unmodified reference
(tests/golden/make_depth_golden.py: depth/depth_learner.py:97-117, vo/eval_traj.py:85-128)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_depth_loss_b2_48x64.npz"))


def test_multi_scale_depth_loss_matches_reference():
    from depth.depth_learner import DepthLearner
    z = _golden()
    dev = torch.device("cuda:0")
    t = lambda k: torch.from_numpy(z[k]).to(dev)
    cfg = {"Train": dict(min_depth=0.1, max_depth=10.0, smooth_weight=0.1, silog_weight=1.0)}
    learner = DepthLearner(None, cfg, dev)
    disps = [t(f"disp{s}").requires_grad_(True) for s in range(4)]
    pred = [learner.disp_to_depth(d) for d in disps]
    total, silog, smooth = learner.multi_scale_loss(pred, t("gt"), t("rgb"), t("valid").bool())
    total.backward()
    for got, key in ((total, "total"), (silog, "silog"), (smooth, "smooth")):
        assert abs(float(got) - float(z[key])) <= 1e-5 * abs(float(z[key])), key
    for s in range(4):
        ref = z[f"grad_disp{s}"]
        err = np.abs(disps[s].grad.cpu().numpy() - ref)
        assert np.all(err <= 1e-3 * np.abs(ref) + 1e-4 * np.abs(ref).max()), (s, err.max() / np.abs(ref).max())


def test_forward_step_contract():
    from depth.depth_learner import DepthLearner
    from model.depthnet import DepthNet
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = DepthNet(18, False).to(dev)
    cfg = {"Train": dict(min_depth=0.1, max_depth=10.0)}
    learner = DepthLearner(net, cfg, dev)
    sample = dict(image=torch.rand(2, 3, 64, 96), depth=0.5 + 4 * torch.rand(2, 1, 64, 96), valid_mask=torch.rand(2, 1, 64, 96) > 0.1)
    total, silog, smooth, pred = learner.forward_step(sample)
    assert len(pred) == 4 and pred[0].shape == (2, 1, 64, 96) and pred[3].shape == (2, 1, 8, 12)
    total.backward()
    assert all(torch.isfinite(q.grad).all() for q in net.parameters() if q.grad is not None)
    assert torch.isfinite(total) and float(silog) > 0 and float(smooth) > 0


def test_depth_to_pointcloud_matches_eval_traj():
    from dvsloss.ops import depth_to_pointcloud
    z = _golden()
    dev = torch.device("cuda:0")
    pts = depth_to_pointcloud(torch.from_numpy(z["pc_depth"]).to(dev), torch.from_numpy(z["pc_T"]), torch.from_numpy(z["pc_K"]))
    ref = z["pc_points"]
    assert tuple(pts.shape) == ref.shape
    assert np.abs(pts.cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()

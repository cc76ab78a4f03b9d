"""Golden vectors of the supervised-depth loss from the UNMODIFIED reference (build container only).

    python tests/golden/make_depth_golden.py

Runs ``DepthLearner.multi_scale_loss`` of /root/reference/depth/depth_learner.py:97-117 on CPU on a small synthetic problem
(4 disparity maps -> depths, ground-truth depth with holes, RGB) and stores inputs, the three losses and the gradients
w.r.t. the four disparity maps in tests/golden/ref_depth_loss_b2_48x64.npz.  Also a point cloud from
``EvalTrajectory.depth_to_pointcloud`` (vo/eval_traj.py:85-128) restated without the random sub-sampling."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "deep-visual-slam_b200"))
from dvsloss.synthetic import make_problem, smooth_depth  # noqa: E402

spec = importlib.util.spec_from_file_location("_ref_depth_learner", "/root/reference/depth/depth_learner.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

B, H, W = 2, 48, 64
p = make_problem(B, H, W, 2, 4, seed=31, consistent=True)
gen = torch.Generator().manual_seed(32)
gt = smooth_depth(B, H, W, gen, 0.4, 6.0)
valid = torch.rand(B, 1, H, W, generator=gen) > 0.2
cfg = {"Train": dict(min_depth=0.1, max_depth=10.0, smooth_weight=0.1, silog_weight=1.0)}
learner = ref.DepthLearner(None, cfg, torch.device("cpu"))
disps = [d.clone().requires_grad_(True) for d in p["disps"]]
pred = [learner.disp_to_depth(d) for d in disps]
total, silog, smooth = learner.multi_scale_loss(pred, gt, p["target"], valid)
total.backward()
arr = dict(rgb=p["target"], gt=gt, valid=valid.to(torch.uint8), total=total.detach(), silog=silog.detach(), smooth=smooth.detach())
for s in range(4):
    arr[f"disp{s}"] = p["disps"][s]
    arr[f"grad_disp{s}"] = disps[s].grad
# point cloud (eval_traj.py:85-128 without np.random.choice)
depth = gt[0, 0].numpy().copy()
depth[valid[0, 0].numpy() == 0] = 0.0
K = p["K"][0, :3, :3].numpy().astype(np.float64)
T = np.eye(4)
T[:3, :3] = [[0.995, -0.0998, 0.0], [0.0998, 0.995, 0.0], [0.0, 0.0, 1.0]]
T[:3, 3] = [0.3, -0.2, 1.5]
xs, ys = np.meshgrid(np.arange(W), np.arange(H))
u, v, z = xs.reshape(-1), ys.reshape(-1), depth.reshape(-1)
m = z > 0
rays = np.linalg.inv(K) @ np.stack([u[m], v[m], np.ones(m.sum())], 0)
pc = rays * z[m]
pw = (T @ np.concatenate([pc, np.ones((1, pc.shape[1]))], 0))[:3].T
arr.update(pc_depth=torch.from_numpy(depth), pc_K=torch.from_numpy(K), pc_T=torch.from_numpy(T), pc_points=torch.from_numpy(pw))
np.savez_compressed(os.path.join(HERE, "ref_depth_loss_b2_48x64.npz"), **{k: np.asarray(v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in arr.items()})
print("total", float(total), "silog", float(silog), "smooth", float(smooth), "points", pw.shape)

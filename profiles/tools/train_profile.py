#!/usr/bin/env python
"""torch.profiler breakdown of one VO training step (BASELINE configs[2], 1 GPU):  python profiles/tools/train_profile.py"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from vo.train import DEFAULT_CONFIG, Trainer, synthetic_sample  # noqa: E402

B, H, W = int(os.environ.get("TRAIN_B", "32")), 480, 640
cfg = copy.deepcopy(DEFAULT_CONFIG)
cfg["Train"]["batch_size"] = B
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
tr = Trainer(cfg, device=dev, net_dtype=torch.bfloat16, noise="kernel", sync_losses=False)
sample = synthetic_sample(B, H, W, seed=1, device=dev)
for _ in range(4):
    tr.train_mono_step(dict(sample))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.train_mono_step(dict(sample))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(os.environ.get("ROWS", "40")), max_name_column_width=150))

"""Which ATen ops are behind the device-to-device copies of a training step?  (GPU box)"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import ProfilerActivity, profile
from vo.train import DEFAULT_CONFIG, Trainer, synthetic_sample
B, H, W = 32, 480, 640
cfg = copy.deepcopy(DEFAULT_CONFIG); cfg["Train"]["batch_size"] = B
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
tr = Trainer(cfg, device=dev, net_dtype=torch.bfloat16, noise="kernel", sync_losses=False)
sample = synthetic_sample(B, H, W, seed=1, device=dev)
for _ in range(4):
    tr.train_mono_step(dict(sample))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    tr.train_mono_step(dict(sample))
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key in ("aten::copy_", "aten::clone", "aten::contiguous", "aten::cat", "aten::to", "aten::_to_copy", "aten::add_", "aten::add", "aten::fill_", "aten::zero_", "aten::sum")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:40]:
    print(f"{e.key:18s} {e.count:4d} {e.device_time_total / 1e3:8.3f} ms  {str(e.input_shapes)[:150]}")

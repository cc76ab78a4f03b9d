"""What does the reflection padding of the decoder's Conv3x3 cost in a training step?  Times the step with ring-carrying
activations (DepthNet.padded_mask variants, PAD_MASKS=0x1FF,0x1EE,...), with border strips everywhere, and with a zero-padded
convolution alone (timing only: wrong border values).  (GPU box)"""
import copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200")):
    sys.path.insert(0, p)
import torch, torch.nn.functional as F
from vo.train import DEFAULT_CONFIG, Trainer, synthetic_sample
import model.layers as ML
B, H, W = 32, 480, 640
cfg = copy.deepcopy(DEFAULT_CONFIG); cfg["Train"].update(batch_size=B, img_h=H, img_w=W)
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
def run(tag):
    torch.manual_seed(0)
    tr = Trainer(cfg, device=dev, num_layers=18, pretrained=False, net_dtype=torch.bfloat16, noise="kernel", sync_losses=False)
    sample = synthetic_sample(B, H, W, seed=100, device=dev)
    for _ in range(4):
        tr.train_mono_step(dict(sample))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tr.train_mono_step(dict(sample))
    e1.record(); torch.cuda.synchronize()
    print(tag, e0.elapsed_time(e1) / 10, "ms/step", flush=True)
    del tr; torch.cuda.empty_cache()
from model.depthnet import DepthNet
for mask in [int(v, 0) for v in os.environ.get("PAD_MASKS", "0x1EE").split(",")]:
    DepthNet.padded_mask = mask
    run(f"padded_mask {mask:#05x}")
DepthNet.padded_mask = 0x1EE
DepthNet.padded_activations = False
run("border strips everywhere")
ML.conv3x3_reflect = lambda x, w, b: F.conv2d(x, w, b, padding=1)      # timing only: zero padding, wrong border values
run("zero-padded convolution only")

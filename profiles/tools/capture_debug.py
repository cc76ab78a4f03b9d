"""Why does Trainer.capture_step fail after eager steps?  Lists what keeps autograd graphs alive.  (GPU box)"""
import copy, gc, os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200")):
    sys.path.insert(0, p)
import torch
from vo.train import DEFAULT_CONFIG, Trainer, synthetic_sample
B, H, W = int(os.environ.get("TRAIN_B", "2")), 192, 256
cfg = copy.deepcopy(DEFAULT_CONFIG); cfg["Train"].update(batch_size=B, img_h=H, img_w=W)
dev = torch.device("cuda:0")
tr = Trainer(cfg, device=dev, num_layers=18, pretrained=False, net_dtype=torch.bfloat16, noise="kernel", sync_losses=False)
sample = synthetic_sample(B, H, W, seed=100, device=dev)
for _ in range(2):
    total, _, _ = tr.train_mono_step(dict(sample))
print(float(total)); del total, _
tr.joint._disp, tr.joint._poses = None, []
for m in tr.nets.modules():
    if isinstance(getattr(m, "outputs", None), dict):
        m.outputs = {}
    if isinstance(getattr(m, "features", None), list):
        m.features = []
gc.collect()
n = 0
for o in gc.get_objects():
    try:
        if isinstance(o, torch.Tensor) and o.grad_fn is not None:
            n += 1
            if n <= 20:
                refs = [type(r).__name__ + (":" + ",".join(map(str, list(r.keys())[:4])) if isinstance(r, dict) else "") for r in gc.get_referrers(o)][:4]
                print("live graph tensor", tuple(o.shape), o.dtype, type(o.grad_fn).__name__, refs)
    except Exception:
        pass
print("tensors with grad_fn alive:", n)
try:
    tr.capture_step(dict(sample))
    tr.train_graph_step(sample); torch.cuda.synchronize(); print("graph ok")
except Exception:
    print("".join(traceback.format_exc().splitlines(True)[-12:]))

#!/usr/bin/env python
"""Device-timed fused loss fwd+bwd at BASELINE configs[1] for the input formats the two-source kernel reads directly:
fp32 / bf16 disparities x fp32 / uint8 images.     python profiles/tools/io_variants_bench.py   (GPU box)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from dvsloss import view_synthesis_loss  # noqa: E402
from dvsloss.synthetic import make_problem, pose_matrix  # noqa: E402

B, H, W = 16, 480, 640
dev = torch.device("cuda:0")
p = make_problem(B, H, W, 2, 4, seed=0, consistent=True)
Ts = [pose_matrix(a.view(B, 3), t.view(B, 3), inv).to(dev).requires_grad_(True) for a, t, inv in zip(p["axisangle"], p["translation"], p["invert"])]
q8 = lambda t: (t * 255).round().clamp(0, 255).to(torch.uint8)
tgt8, src8 = q8(p["target"]).to(dev), [q8(s).to(dev) for s in p["sources"]]
unit = lambda t8: t8.cpu().float().div(255).to(dev)
tgtf, srcf = unit(tgt8), [unit(s) for s in src8]
K, iK = p["K"].to(dev), p["inv_K"].to(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for dname, ddt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
    for iname, (tgt, srcs) in (("fp32", (tgtf, srcf)), ("uint8", (tgt8, src8))):
        disps = [d.to(dev).to(ddt).requires_grad_(True) for d in p["disps"]]

        def step():
            for t in disps + Ts:
                t.grad = None
            loss, _ = view_synthesis_loss(disps, tgt, srcs, K, iK, Ts, noise="kernel")
            loss.backward()

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(json.dumps({"disp": dname, "images": iname, "ms_median": ts[len(ts) // 2], "ms_min": ts[0],
                          "gpix_per_s": B * 4 * 2 * H * W / (ts[len(ts) // 2] * 1e-3) / 1e9}), flush=True)

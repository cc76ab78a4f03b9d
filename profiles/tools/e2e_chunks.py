#!/usr/bin/env python
"""Host-resident step (dvsloss.HostLossPipeline, BASELINE configs[1], graph replay) against the chunking of the batch:
ms per step (wall clock around 50 replays) and the H2D rate it amounts to.  (GPU box)

    python profiles/tools/e2e_chunks.py 4 8 16 4,3,3,2,2,1,1 taper
    U8=1 python profiles/tools/e2e_chunks.py ...      # uint8 frames read by the tile kernel (70 MB in: kernel-bound)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from bench import H, NSRC, W, bind_to_gpu_numa, make_inputs  # noqa: E402
from dvsloss import HostLossPipeline  # noqa: E402

B = 16
bind_to_gpu_numa(0)
dev = torch.device("cuda:0")
host = make_inputs(B, 0, "cpu")
pin = lambda t: t.contiguous().pin_memory()
h_in = dict(target=pin(host["target"]), sources=[pin(s) for s in host["sources"]], disps=[pin(d) for d in host["disps"]],
            K=pin(host["K"]), inv_K=pin(host["inv_K"]), Ts=[pin(T) for T in host["Ts"]])
h_out = dict(loss=torch.empty(5).pin_memory(), gd=[torch.empty_like(d).pin_memory() for d in h_in["disps"]],
             gT=[torch.empty_like(T).pin_memory() for T in h_in["Ts"]])
h2d = sum(t.numel() * 4 for t in [h_in["target"], h_in["K"], h_in["inv_K"]] + h_in["sources"] + h_in["disps"] + h_in["Ts"])
U8 = bool(int(os.environ.get("U8", "0")))
if U8:
    q8 = lambda t: (t * 255).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
    h_in["target"], h_in["sources"] = q8(host["target"]), [q8(s_) for s_ in host["sources"]]
    h2d -= 3 * (1 + NSRC) * B * H * W * 3
for arg in sys.argv[1:] or ["8"]:
    if arg == "taper":
        ch = arg
    else:
        ch = [int(v) for v in arg.split(",")]
        ch = ch[0] if len(ch) == 1 else ch
    pipe = HostLossPipeline(B, H, W, [tuple(d.shape[2:]) for d in h_in["disps"]], NSRC, chunks=ch, device=dev, noise="kernel",
                            uint8_images=U8, u8_in_kernel=U8)
    for _ in range(5):
        pipe.run(h_in, h_out)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(50):
            pipe.run(h_in, h_out)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / 50)
    print(f"chunks {arg:>24s}  {best * 1e3:7.3f} ms/step  {B * H * W * NSRC * 4 / best / 1e9:6.2f} Gpix/s  H2D {h2d / best / 1e9:5.1f} GB/s", flush=True)
    del pipe

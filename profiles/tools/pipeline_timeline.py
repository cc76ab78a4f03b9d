#!/usr/bin/env python
"""Where does the time of one host-resident step (dvsloss.HostLossPipeline, BASELINE configs[1]) go?  CUDA events at the end
of the copy-in stream, of every chunk's kernels and of the copy-out stream, relative to the start.  (GPU box)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from bench import H, NSRC, W, make_inputs  # noqa: E402
from dvsloss import HostLossPipeline  # noqa: E402

B = 16
chunks = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4").split(",")]
chunks = chunks[0] if len(chunks) == 1 else chunks
dev = torch.device("cuda:0")
host = make_inputs(B, 0, "cpu")
pin = lambda t: t.contiguous().pin_memory()
h_in = dict(target=pin(host["target"]), sources=[pin(s) for s in host["sources"]], disps=[pin(d) for d in host["disps"]],
            K=pin(host["K"]), inv_K=pin(host["inv_K"]), Ts=[pin(T) for T in host["Ts"]])
h_out = dict(loss=torch.empty(5).pin_memory(), gd=[torch.empty_like(d).pin_memory() for d in h_in["disps"]],
             gT=[torch.empty_like(T).pin_memory() for T in h_in["Ts"]])
pipe = HostLossPipeline(B, H, W, [tuple(d.shape[2:]) for d in h_in["disps"]], NSRC, chunks=chunks, device=dev, noise="kernel", graph=False)
chunks = pipe.chunks
for _ in range(3):
    pipe.run(h_in, h_out)
torch.cuda.synchronize()
pipe.ev_in = [torch.cuda.Event(enable_timing=True) for _ in range(chunks)]
pipe.ev_run = [torch.cuda.Event(enable_timing=True) for _ in range(chunks)]
pipe.ev_out = [torch.cuda.Event(enable_timing=True) for _ in range(chunks)]
e0 = torch.cuda.Event(enable_timing=True)
e0.record()
t0 = time.perf_counter()
pipe.run(h_in, h_out)
wall = (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize()
print(f"chunks {chunks}: wall {wall:.3f} ms")
for c in range(chunks):
    print(f"  chunk {c}: inputs landed {e0.elapsed_time(pipe.ev_in[c]):.3f}  kernels done {e0.elapsed_time(pipe.ev_run[c]):.3f}  "
          f"gradients on host {e0.elapsed_time(pipe.ev_out[c]):.3f} ms")

#!/usr/bin/env python
"""Per-operator timing of the granular drop-in ops against the reference's eager
ATen sequence (oracle port) on the same GPU, forward+backward, BASELINE config-2 shapes (B=16, 640x480).

    python profiles/tools/granular_bench.py > gpurun_out/granular.jsonl      (on the GPU box)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from dvsloss import ops  # noqa: E402
from dvsloss.synthetic import make_problem, pose_matrix  # noqa: E402
from oracle import reference_port as port  # noqa: E402

B, H, W = 16, 480, 640
dev = torch.device("cuda:0")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def timeit_graph(fn, n=20):
    """Device time alone: the call sequence recorded once as a CUDA graph and replayed (the eager numbers of these small
    operators are bound by Python / dispatcher overhead on either side, not by their kernels)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fb(f, *leaves):
    def run():
        for t in leaves:
            t.grad = None
        out = f()
        out = out[1] if isinstance(out, tuple) else out
        out.sum().backward()
    return run


def main():
    p = make_problem(B, H, W, 2, 4, seed=0, consistent=True)
    tgt, src = p["target"].to(dev), p["sources"][0].to(dev)
    K, iK = p["K"].to(dev), p["inv_K"].to(dev)
    T = pose_matrix(p["axisangle"][0].view(B, 3), p["translation"][0].view(B, 3), True).to(dev)
    disp = p["disps"][0].to(dev).requires_grad_(True)
    disp1 = p["disps"][1].to(dev).requires_grad_(True)
    depth = (1.0 / (0.1 + 9.9 * disp.detach())).requires_grad_(True)
    warped = (0.9 * src + 0.1 * tgt).requires_grad_(True)
    aa = p["axisangle"][0].to(dev).requires_grad_(True)
    tr = p["translation"][0].to(dev).requires_grad_(True)
    rows = []

    def add(name, ours, ref):
        a, b = timeit(ours), timeit(ref)
        try:
            ga, gb = timeit_graph(ours), timeit_graph(ref)
        except Exception as exc:                               # an op that cannot be captured: keep the eager numbers
            ga = gb = float("nan")
            print("# graph timing failed for", name, repr(exc)[:120], file=sys.stderr)
        rows.append({"op": name, "ms_b200": a, "ms_eager": b, "speedup": b / a, "device_ms_b200": ga, "device_ms_eager": gb,
                     "device_speedup": gb / ga})
        print(json.dumps(rows[-1]), flush=True)

    add("disp_to_depth", fb(lambda: ops.disp_to_depth(disp, 0.1, 10.0), disp),
        fb(lambda: port.disp_to_depth(disp, 0.1, 10.0), disp))
    add("upsample_bilinear (scale 1 -> full)", fb(lambda: ops.upsample_bilinear(disp1, (H, W)), disp1),
        fb(lambda: F.interpolate(disp1, [H, W], mode="bilinear", align_corners=False), disp1))
    add("SSIM", fb(lambda: ops.ssim(warped, tgt), warped), fb(lambda: port.ssim(warped, tgt), warped))
    add("compute_reprojection_loss", fb(lambda: ops.compute_reprojection_loss(warped, tgt, 0.85), warped),
        fb(lambda: port.reprojection_loss(warped, tgt, 0.85), warped))
    add("get_smooth_loss", fb(lambda: ops.get_smooth_loss(disp, tgt), disp), fb(lambda: port.smooth_loss(disp, tgt), disp))
    add("transformation_from_parameters", fb(lambda: ops.transformation_from_parameters(aa, tr, True), aa, tr),
        fb(lambda: port.transformation_from_parameters(aa, tr, True), aa, tr))

    def ours_warp():
        depth.grad = None
        cam = ops.backproject(depth, iK)
        grid = ops.project3d(cam, K, T, H, W)
        ops.grid_sample_border(src, grid).sum().backward()

    def ref_warp():
        depth.grad = None
        cam = port.backproject(depth, iK)
        grid = port.project(cam, K, T, H, W)
        F.grid_sample(src, grid, padding_mode="border", align_corners=True).sum().backward()

    add("BackprojectDepth + Project3D + grid_sample", ours_warp, ref_warp)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""BASELINE configs[4]: loss-kernel sweep, resolution 320x240 -> 1920x1440 x sources {1,2,4}, 4 scales, 1 GPU.

    python profiles/tools/sweep.py > gpurun_out/sweep.jsonl        (on the GPU box)

Per point: fused loss forward+backward (device-timed with CUDA events, L2 flushed between steps, batch chosen so
that B*H*W >= 4.9 M pixels), warped Gpix/s and the HBM-roofline fraction under the declared byte model
(bytes_alg).  One JSON line per point.
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from bench import bytes_alg, measured_peak_gbs  # noqa: E402
from dvsloss import view_synthesis_loss  # noqa: E402
from dvsloss.synthetic import make_problem, pose_matrix  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak, how = measured_peak_gbs()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    only = [tuple(int(v) for v in q.split("x")) for q in os.environ.get("SWEEP_ONLY", "").split(",") if q]   # e.g. 1280x960x2
    for (W, H) in [(320, 240), (640, 480), (960, 720), (1280, 960), (1920, 1440)]:
        for N in (1, 2, 4):
            if only and (W, H, N) not in only:
                continue
            B = max(1, math.ceil(4915200 / (H * W)))
            p = make_problem(B, H, W, N, 4, seed=1, consistent=True)
            Ts = [pose_matrix(a.view(B, 3), t.view(B, 3), inv).to(dev).requires_grad_(True)
                  for a, t, inv in zip(p["axisangle"], p["translation"], p["invert"])]
            disps = [d.to(dev).requires_grad_(True) for d in p["disps"]]
            tgt, srcs, K, iK = p["target"].to(dev), [s.to(dev) for s in p["sources"]], p["K"].to(dev), p["inv_K"].to(dev)

            def step():
                for t in disps + Ts:
                    t.grad = None
                loss, _ = view_synthesis_loss(disps, tgt, srcs, K, iK, Ts, noise="kernel")
                loss.backward()

            for _ in range(3):
                flush.zero_()
                step()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            ba = bytes_alg(B, H, W, N, 4)
            print(json.dumps({"W": W, "H": H, "N": N, "S": 4, "B": B, "ms_median": ms, "ms_min": ts[0],
                              "gpix_per_s": B * 4 * N * H * W / (ms * 1e-3) / 1e9, "bytes_alg": ba,
                              "hbm_frac": ba / (ms * 1e-3) / 1e9 / peak, "peak_gbs": peak, "peak_source": how}), flush=True)
            del disps, Ts, tgt, srcs
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

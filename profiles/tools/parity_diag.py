#!/usr/bin/env python
"""How close are the gradients to the exact answer?  Disparity gradients of the CUDA path and of the reference's fp32
op sequence (oracle port on the same GPU), both against a float64 evaluation of the same formulas under the same
selection, normalised by max|g64|.  The maxima are single pixels sitting on a kink of the piecewise-smooth loss
(either one-sided derivative is legitimate; sometimes the CUDA path, sometimes the fp32 reference is the outlier);
the high percentiles show the regular behaviour.      python profiles/tools/parity_diag.py     (GPU box)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import parity  # noqa: E402
from dvsloss.synthetic import make_problem  # noqa: E402
from test_gpu_fused import cuda_impl  # noqa: E402

print("| case | scale | max ours-ref32 | max ref32-ref64 | max ours-ref64 | p99.99 ours-ref64 | p99.99 ref32-ref64 |")
print("|---|---|---|---|---|---|---|")
for (B, H, W, N, seed) in [(2, 480, 640, 2, 3), (1, 960, 1280, 4, 13)]:
    p = make_problem(B, H, W, N, 4, seed=seed, consistent=True)
    prob = parity.problem_from_synthetic(p, True)
    got = cuda_impl(prob, None)
    so = [np.asarray(s).astype(np.int64) for s in got["sel"]]
    r32 = parity.oracle_eval(prob, sel_override=so, device="cuda")
    r64 = parity.oracle_eval(prob, sel_override=so, device="cuda", dtype=torch.float64)
    for s in range(4):
        g, a, b = np.asarray(got["grad_disp"][s], np.float64), r32["grad_disp"][s], r64["grad_disp"][s]
        rm = np.abs(b).max()
        e1, e2, e3 = np.abs(g - a) / rm, np.abs(a - b) / rm, np.abs(g - b) / rm
        print(f"| {B}x{H}x{W}, N={N} | {s} | {e1.max():.1e} | {e2.max():.1e} | {e3.max():.1e} | {np.quantile(e3, 0.9999):.1e} | {np.quantile(e2, 0.9999):.1e} |")
    for i in range(N):
        g, a, b = np.asarray(got["grad_T"][i], np.float64), r32["grad_T"][i], r64["grad_T"][i]
        rm = np.abs(b).max()
        print(f"| {B}x{H}x{W}, N={N} | pose {i} | {np.abs(g - a).max() / rm:.1e} | {np.abs(a - b).max() / rm:.1e} | {np.abs(g - b).max() / rm:.1e} | | |")
    print(f"| {B}x{H}x{W}, N={N} | loss | {abs(got['loss'] - r32['loss']) / abs(r64['loss']):.1e} | {abs(r32['loss'] - r64['loss']) / abs(r64['loss']):.1e} | {abs(got['loss'] - r64['loss']) / abs(r64['loss']):.1e} | | |")

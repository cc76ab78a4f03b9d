#!/usr/bin/env python
"""Where does the worst kink-free disparity-gradient element of a parity case sit, and which kink did the locator of
tests/parity.py miss?  Prints the element, its values (CUDA / fp32 oracle / float64 oracle) and every pixel of its
footprint that is close to a kink.       python profiles/tools/straggler_diag.py H W seed [scale]        (GPU box)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import parity  # noqa: E402
from dvsloss.synthetic import make_problem  # noqa: E402
from oracle.closed_form import ssim_terms  # noqa: E402
from test_gpu_fused import cuda_impl  # noqa: E402

H, W, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
N = 2
prob = parity.problem_from_synthetic(make_problem(1, H, W, N, 4, seed=seed, consistent=True), True)
got = cuda_impl(prob, None)
so = [np.asarray(s).astype(np.int64) for s in got["sel"]]
r32 = parity.oracle_eval(prob, sel_override=so, device="cuda")
r64 = parity.oracle_eval(prob, sel_override=so, device="cuda", dtype=torch.float64, keep=True)
ex, tgt = r64["extras"], prob["target"].astype(np.float64)
for s in ([int(sys.argv[4])] if len(sys.argv) > 4 else range(4)):
    g, r, rr = np.asarray(got["grad_disp"][s], np.float64), r32["grad_disp"][s], r64["grad_disp"][s]
    kinks = parity._kink_weight(prob, ex, so[s], s, N, H, W)
    err = np.abs(g - r)
    err[kinks > 0] = 0
    b, _, I, J = np.unravel_index(np.argmax(err), err.shape)
    f = 1 << s
    print(f"scale {s}: worst kink-free element ({I},{J}) err/max {err[b, 0, I, J] / np.abs(r).max():.3e}  cuda {g[b, 0, I, J]:.6e} "
          f"fp32 {r[b, 0, I, J]:.6e} fp64 {rr[b, 0, I, J]:.6e}")
    ys = range(max(f * I - f, 0), min(f * I + 2 * f, H))
    xs = range(max(f * J - f, 0), min(f * J + 2 * f, W))
    for i in range(N):
        gr = ex[("sample", i, s)]
        ix, iy = (gr[..., 0] + 1) / 2 * (W - 1), (gr[..., 1] + 1) / 2 * (H - 1)
        col = ex[("color", i, s)]
        S = ssim_terms(col, tgt)[0]
        for y in ys:
            for x in xs:
                fx, fy = abs(ix[b, y, x] - round(ix[b, y, x])), abs(iy[b, y, x] - round(iy[b, y, x]))
                l1 = np.abs(tgt[b, :, y, x] - col[b, :, y, x]).min()
                smin, smax = S[b, :, y, x].min(), S[b, :, y, x].max()
                if fx < 5e-3 or fy < 5e-3 or l1 < 5e-4 or smin < 1e-3 or smax > 1 - 1e-3:
                    print(f"   src {i} px ({y},{x}) sel {so[s][b, y, x]}  |ix-round| {fx:.2e} |iy-round| {fy:.2e}  min|y-x| {l1:.2e}  "
                          f"ssim {smin:.2e}..{smax:.5f}  ix {ix[b, y, x]:.4f} iy {iy[b, y, x]:.4f}")

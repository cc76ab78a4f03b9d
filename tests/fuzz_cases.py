"""Seeded random shapes for the fused loss (shared by the emulator test on CPU and the C-ABI test on the GPU): batch 1-2,
2..89 pixels a side (tile tails of every length), 1-4 sources, 1-4 scales, pyramid or arbitrary disparity sizes down to 1x1,
automask on / off, consistent or random content, unit or arbitrary per-scale upstream gradients."""
import numpy as np

import parity
from dvsloss.synthetic import make_problem

N_CASES = 24


def fuzz_case(i: int):
    rng = np.random.default_rng(1000 + i)
    B, H, W = int(rng.integers(1, 3)), int(rng.integers(2, 90)), int(rng.integers(2, 90))
    N, S = int(rng.integers(1, 5)), int(rng.integers(1, 5))
    auto_mask, pyramid = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    if pyramid:
        dims = [(max(1, H >> s), max(1, W >> s)) for s in range(S)]
    else:
        dims = [(int(rng.integers(1, H + 1)), int(rng.integers(1, W + 1))) for _ in range(S)]
    consistent = bool(rng.integers(0, 2)) and H >= 30 and W >= 30
    prob = parity.problem_from_synthetic(make_problem(B, H, W, N, S, seed=i, consistent=consistent), auto_mask)
    prob["disps"] = [rng.uniform(0.05, 0.9, (B, 1, h, w)).astype(np.float32) for h, w in dims]
    gps = None if rng.integers(0, 2) else [float(x) for x in rng.uniform(-2, 2, S)]
    return prob, gps

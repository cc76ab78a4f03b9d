"""CPU check of the fused kernel's tile logic: the block emulator (tests/emu) compiles the very phase
functions of the CUDA kernel (csrc/dvs_fused_core.cuh) with g++ and is compared with the oracle."""
import numpy as np
import pytest

import emu_harness
import parity
from dvsloss.synthetic import make_problem

GOLDEN = ["ref_b2_48x64_consistent.npz", "ref_b2_48x64_random.npz", "ref_b1_96x128_consistent.npz",
          "ref_b2_32x48_nomask.npz", "ref_b1_40x56_bigmotion.npz"]


def emu_impl(prob, gps):
    return emu_harness.run(prob["disps"], prob["target"], prob["sources"], prob["K"], prob["inv_K"], prob["Ts"],
                           prob["noise"], auto_mask=prob["auto_mask"], grad_per_scale=gps)


@pytest.mark.parametrize("name", GOLDEN)
def test_emulator_vs_reference_golden(name):
    g = parity.load_golden(name)
    # ref32 = the numbers the live reference produced (stored in the fixture)
    ref32 = dict(g["ref"])
    stats = parity.check_parity(emu_impl, g["prob"], ref32=ref32, verbose=True)
    assert stats["sel_flip_frac_max"] < 0.01


@pytest.mark.parametrize("B,H,W,N,S,consistent,auto_mask", [
    (1, 33, 47, 1, 4, True, True),       # ragged: not a multiple of the 30x30 tile nor of 8; one source
    (2, 64, 96, 3, 3, True, True),       # three sources, three scales
    (1, 61, 35, 4, 4, False, True),      # four sources, odd sizes, random content
    (1, 30, 30, 2, 1, True, False),      # exactly one tile, single scale, no automask
    (1, 2, 2, 2, 1, False, True),        # smallest legal image
])
def test_emulator_vs_oracle_shapes(B, H, W, N, S, consistent, auto_mask):
    p = make_problem(B, H, W, N, S, seed=B * 1000 + H + W + N, consistent=consistent and H >= 30)
    prob = parity.problem_from_synthetic(p, auto_mask)
    parity.check_parity(emu_impl, prob, verbose=True)


def test_emulator_grad_per_scale_weights():
    """Arbitrary upstream gradients per loss/s (GradScaler-style scaling, vo/train.py:183)."""
    g = parity.load_golden("ref_b2_48x64_consistent.npz")
    parity.check_parity(emu_impl, g["prob"], grad_per_scale=[128.0, 0.0, -3.5, 0.25], verbose=True)


def test_emulator_non_pyramid_disparity_sizes():
    """Disparity maps whose size is not H>>s (generic bilinear weights, non-integer ratios)."""
    p = make_problem(1, 48, 64, 2, 2, seed=7)
    prob = parity.problem_from_synthetic(p)
    rng = np.random.default_rng(0)
    prob["disps"] = [rng.uniform(0.05, 0.9, (1, 1, 20, 27)).astype(np.float32),
                     rng.uniform(0.05, 0.9, (1, 1, 7, 64)).astype(np.float32)]
    parity.check_parity(emu_impl, prob, verbose=True)


@pytest.mark.parametrize("name", ["ref_b2_48x64_consistent.npz", "ref_b1_96x128_consistent.npz"])
def test_parity_gates_catch_a_one_percent_gradient_error(name):
    """The gates must be able to fail: the same tile code built with a 1 % error injected into d loss / d warped colour
    (``-DDVS_FAULT_GRAD_SCALE=1.01f`` in phase_grad) has to be rejected by check_parity."""
    def faulty(prob, gps):
        return emu_harness.run(prob["disps"], prob["target"], prob["sources"], prob["K"], prob["inv_K"], prob["Ts"],
                               prob["noise"], auto_mask=prob["auto_mask"], grad_per_scale=gps,
                               defines=("DVS_FAULT_GRAD_SCALE=1.01f",))
    g = parity.load_golden(name)
    with pytest.raises(AssertionError, match="grad_"):
        parity.check_parity(faulty, g["prob"], ref32=dict(g["ref"]))


@pytest.mark.parametrize("B,H,W,seed", [(1, 64, 96, 0), (2, 96, 128, 1)])
def test_emulator_kink_free_every_element_strict(B, H, W, seed):
    """No near-kink exemption: every element of every scale under rtol 1e-3 / atol 1e-6 and the 1e-3 inf-norm bound."""
    prob = parity.kink_free_problem(B, H, W, seed)
    stats = parity.check_parity(emu_impl, prob, verbose=True, expect_kink_free=True)
    assert stats["grad_disp_relinf_max"] < 1e-3


@pytest.mark.parametrize("B,H,W,N,dims,auto_mask", [
    (1, 31, 61, 2, [(31, 61), (1, 1)], True),              # a 1x1 disparity map: every fine pixel reads the same coarse element
    (2, 33, 40, 2, [(1, 40), (33, 1)], True),              # one-row and one-column disparity maps
    (1, 3, 3, 2, [(3, 3), (1, 1)], True),                  # image smaller than the SSIM window's reach on every side
    (1, 31, 31, 3, [(31, 31), (15, 15), (7, 7)], False),   # one pixel past the 30x30 tile in both directions
    (1, 2, 200, 1, [(2, 200), (1, 100)], True),            # two rows, many tiles across
    (1, 200, 2, 4, [(200, 2), (100, 1)], True),            # two columns, many tiles down
])
def test_emulator_degenerate_disparity_sizes_and_thin_images(B, H, W, N, dims, auto_mask):
    """Edge geometry: reflection at both borders inside one pixel's window, coarse maps of extent 1, tile tails of one pixel."""
    p = make_problem(B, H, W, N, len(dims), seed=H * 7 + W, consistent=False)
    prob = parity.problem_from_synthetic(p, auto_mask)
    rng = np.random.default_rng(H + W)
    prob["disps"] = [rng.uniform(0.05, 0.9, (B, 1, h, w)).astype(np.float32) for h, w in dims]
    parity.check_parity(emu_impl, prob, verbose=True)


@pytest.mark.parametrize("i", range(0, 24, 3))
def test_emulator_seeded_random_shapes(i):
    """Every third case of tests/fuzz_cases.py on the block emulator (the GPU suite runs all of them through the C ABI)."""
    import fuzz_cases
    prob, gps = fuzz_cases.fuzz_case(i)
    parity.check_parity(emu_impl, prob, grad_per_scale=gps, verbose=True)

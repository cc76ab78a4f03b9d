"""CPU tests of the oracle itself: reference_port.py against the golden vectors produced by the live
reference (tests/golden/make_golden.py), and the closed-form numpy restatement against autograd of the port."""
import numpy as np
import pytest
import torch

import parity
from oracle import closed_form as cf

GOLDEN = ["ref_b2_48x64_consistent.npz", "ref_b2_48x64_random.npz", "ref_b1_96x128_consistent.npz",
          "ref_b2_32x48_nomask.npz", "ref_b1_40x56_bigmotion.npz"]


@pytest.mark.parametrize("name", GOLDEN)
def test_port_reproduces_reference_golden(name):
    """Same op sequence as the reference on CPU: losses bit-identical, selection identical, gradients to the
    round-off of autograd's accumulation order."""
    g = parity.load_golden(name)
    got = parity.oracle_eval(g["prob"])
    ref = g["ref"]
    assert np.float32(got["loss"]) == np.float32(ref["loss"])
    assert np.array_equal(got["per_scale"].astype(np.float32), ref["per_scale"].astype(np.float32))
    for s in range(4):
        if g["prob"]["auto_mask"]:
            assert np.array_equal(got["sel"][s], ref["sel"][s])
        a, b = got["grad_disp"][s], ref["grad_disp"][s]
        assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max()
    for i in range(2):
        a, b = got["grad_T"][i], ref["grad_T"][i]
        assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max()


@pytest.mark.parametrize("name", ["ref_b2_48x64_consistent.npz", "ref_b1_40x56_bigmotion.npz", "ref_b2_32x48_nomask.npz"])
def test_closed_form_matches_autograd_fp64(name):
    """The hand-derived adjoint (the formulation the kernel implements) equals autograd of the port in float64."""
    g = parity.load_golden(name)
    prob = g["prob"]
    r = parity.oracle_eval(prob, dtype=torch.float64)
    f64 = lambda a: np.asarray(a, np.float64)
    c = cf.loss_and_grads([f64(d) for d in prob["disps"]], f64(prob["target"]), [f64(s) for s in prob["sources"]],
                          f64(prob["K"]), f64(prob["inv_K"]), [f64(t) for t in prob["Ts"]],
                          [f64(n) for n in prob["noise"]] if prob["noise"] is not None else None,
                          auto_mask=prob["auto_mask"])
    assert abs(c["loss"] - r["loss"]) < 1e-12
    for s in range(4):
        assert np.array_equal(c["sel"][s], r["sel"][s])
        assert np.abs(c["grad_disp"][s] - r["grad_disp"][s]).max() <= 1e-9 * np.abs(r["grad_disp"][s]).max()
    for i in range(2):
        assert np.abs(c["grad_T"][i] - r["grad_T"][i]).max() <= 1e-9 * np.abs(r["grad_T"][i]).max()


def test_closed_form_finite_difference():
    """Central differences of the float64 loss w.r.t. a few disparity elements and pose entries."""
    from dvsloss.synthetic import make_problem
    p = make_problem(1, 24, 32, 2, 2, seed=9)
    prob = parity.problem_from_synthetic(p, True)
    f64 = lambda a: np.asarray(a, np.float64)
    args = lambda d, T: ([f64(x) for x in d], f64(prob["target"]), [f64(s) for s in prob["sources"]], f64(prob["K"]),
                         f64(prob["inv_K"]), [f64(t) for t in T], [f64(n) for n in prob["noise"]])
    base = cf.loss_and_grads(*args(prob["disps"], prob["Ts"]))
    sel = base["sel"]
    rng = np.random.default_rng(0)
    h = 1e-6
    for _ in range(6):
        s = int(rng.integers(0, 2))
        idx = tuple(int(rng.integers(0, n)) for n in prob["disps"][s].shape)
        dp = [f64(d).copy() for d in prob["disps"]]
        dm = [f64(d).copy() for d in prob["disps"]]
        dp[s][idx] += h
        dm[s][idx] -= h
        lp = cf.loss_and_grads(*args(dp, prob["Ts"]), sel_override=sel)["loss"]
        lm = cf.loss_and_grads(*args(dm, prob["Ts"]), sel_override=sel)["loss"]
        fd = (lp - lm) / (2 * h)
        an = base["grad_disp"][s][idx]
        assert abs(fd - an) <= 1e-4 * abs(an) + 1e-9, (s, idx, fd, an)
    h = 1e-8        # a pose entry moves every pixel's coordinate (by ~500*h px): keep kink crossings negligible
    for i in range(2):
        for (r, c) in [(0, 3), (1, 2), (2, 0)]:
            Tp = [f64(t).copy() for t in prob["Ts"]]
            Tm = [f64(t).copy() for t in prob["Ts"]]
            Tp[i][0, r, c] += h
            Tm[i][0, r, c] -= h
            fd = (cf.loss_and_grads(*args(prob["disps"], Tp), sel_override=sel)["loss"]
                  - cf.loss_and_grads(*args(prob["disps"], Tm), sel_override=sel)["loss"]) / (2 * h)
            an = base["grad_T"][i][0, r, c]
            assert abs(fd - an) <= 2e-3 * abs(an) + 1e-8, (i, r, c, fd, an)


def test_port_generalises_over_sources():
    """N = 1 and N = 4 (configs 4/5) run and put the selection in [identity_1..N, reproj_1..N] order."""
    from dvsloss.synthetic import make_problem
    for N in (1, 4):
        p = make_problem(1, 24, 32, N, 2, seed=N)
        prob = parity.problem_from_synthetic(p, True)
        r = parity.oracle_eval(prob)
        assert all(0 <= s.min() and s.max() < 2 * N for s in r["sel"])
        assert len(r["grad_T"]) == N and np.isfinite(r["loss"])

"""Generate golden vectors by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports ``MonodepthTrainer`` from ``/root/reference/vo/learner_new.py`` (never copied into
this repo), drives ``_generate_images_pred`` + ``_compute_losses`` + ``backward`` on small
synthetic problems with the automask noise injected (``torch.randn`` is patched for the
duration of ``_compute_losses`` so the reference consumes a known tensor), checks that
``oracle/reference_port.py`` reproduces the reference bit-for-bit on CPU, and writes
``tests/golden/*.npz``.  The GPU box has no ``/root/reference``; tests there read the npz.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-visual-slam_b200"))

from dvsloss.synthetic import make_problem  # noqa: E402
from oracle import reference_port as port  # noqa: E402

REF = "/root/reference"

CASES = [
    # name, B, H, W, consistent, auto_mask, seed
    ("ref_b2_48x64_consistent", 2, 48, 64, True, True, 1),
    ("ref_b2_48x64_random", 2, 48, 64, False, True, 2),
    ("ref_b1_96x128_consistent", 1, 96, 128, True, True, 3),
    ("ref_b2_32x48_nomask", 2, 32, 48, True, False, 4),
    ("ref_b1_40x56_bigmotion", 1, 40, 56, True, True, 5),
]


def load_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "vo"))
    import learner_new  # type: ignore
    import learner_func  # type: ignore
    return learner_new, learner_func


class _RandnPatch:
    """Make torch.randn return the queued tensors (one per scale), as the survey did."""

    def __init__(self, queue):
        self.queue = list(queue)
        self.orig = torch.randn

    def __enter__(self):
        def fake(*shape, **kw):
            t = self.queue.pop(0)
            want = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            assert tuple(t.shape) == want, (t.shape, want)
            return t.clone()
        torch.randn = fake
        return self

    def __exit__(self, *a):
        torch.randn = self.orig


def run_reference(prob, auto_mask, learner_new, learner_func):
    B, _, H, W = prob["target"].shape
    cfg = {"Train": dict(num_source=2, batch_size=B, img_h=H, img_w=W, smoothness_ratio=0.001,
                         auto_mask=auto_mask, ssim_ratio=0.85, min_depth=0.1, max_depth=10.0,
                         use_compile=False)}
    trainer = learner_new.MonodepthTrainer(None, None, cfg, torch.device("cpu"))
    disps = [d.clone().requires_grad_(True) for d in prob["disps"]]
    aa = [a.clone().requires_grad_(True) for a in prob["axisangle"]]
    tr = [t.clone().requires_grad_(True) for t in prob["translation"]]
    outputs = {("disp", s): disps[s] for s in range(4)}
    Ts = []
    for i, fid in enumerate([-1, 1]):
        T = learner_func.transformation_from_parameters(aa[i], tr[i], invert=(fid < 0))
        T.retain_grad()
        outputs[("cam_T_cam", 0, fid)] = T
        Ts.append(T)
    sample = dict(prob["sample"])
    trainer._generate_images_pred(sample, outputs)
    with _RandnPatch(prob["noise"] if auto_mask else []):
        losses = trainer._compute_losses(sample, outputs)
    losses["loss"].backward()
    res = {
        "loss": losses["loss"].detach(),
        "per_scale": torch.stack([losses[f"loss/{s}"].detach() for s in range(4)]),
        "grad_disp": [d.grad for d in disps],
        "grad_T": [T.grad for T in Ts],
        "grad_axisangle": [a.grad for a in aa],
        "grad_translation": [t.grad for t in tr],
        "T": [T.detach() for T in Ts],
    }
    if auto_mask:
        res["identity_selection"] = [outputs[f"identity_selection/{s}"] for s in range(4)]
    res["color"] = [[outputs[("color", f, s)].detach() for f in (-1, 1)] for s in range(4)]
    res["depth"] = [outputs[("depth", s)].detach() for s in range(4)]
    return res


def run_port(prob, auto_mask):
    disps = [d.clone().requires_grad_(True) for d in prob["disps"]]
    aa = [a.clone().requires_grad_(True) for a in prob["axisangle"]]
    tr = [t.clone().requires_grad_(True) for t in prob["translation"]]
    Ts = []
    for i in range(2):
        T = port.transformation_from_parameters(aa[i], tr[i], invert=prob["invert"][i])
        T.retain_grad()
        Ts.append(T)
    out = port.view_synthesis_loss(disps, prob["target"], prob["sources"], prob["K"], prob["inv_K"], Ts,
                                   prob["noise"] if auto_mask else None, auto_mask=auto_mask)
    out["loss"].backward()
    return {
        "loss": out["loss"].detach(), "per_scale": torch.stack([p.detach() for p in out["per_scale"]]),
        "grad_disp": [d.grad for d in disps], "grad_T": [T.grad for T in Ts],
        "grad_axisangle": [a.grad for a in aa], "grad_translation": [t.grad for t in tr],
        "sel": out["sel"],
    }


def make_process_batch_case(learner_new, name="ref_process_batch_b2_48x64", B=2, H=48, W=64, seed=21):
    """The reference's whole ``process_batch`` (vo/learner_new.py:76-105) with tiny networks: losses and the gradients
    that reach the network weights.  The GPU test loads the same weights into the same tiny nets, runs THIS repo's
    ``MonodepthTrainer.process_batch`` with the same noise and compares."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tiny_nets import TinyDepthNet, TinyPoseNet
    prob = make_problem(B, H, W, 2, 4, seed=seed, consistent=True)
    torch.manual_seed(seed)
    dnet = TinyDepthNet(prob["disps"])
    pnet = TinyPoseNet(prob["axisangle"], prob["translation"])
    cfg = {"Train": dict(num_source=2, batch_size=B, img_h=H, img_w=W, smoothness_ratio=0.001, auto_mask=True,
                         ssim_ratio=0.85, min_depth=0.1, max_depth=10.0, use_compile=False)}
    trainer = learner_new.MonodepthTrainer(dnet, pnet, cfg, torch.device("cpu"))
    sample = dict(prob["sample"])
    with _RandnPatch(prob["noise"]):
        outputs, losses = trainer.process_batch(sample)
    losses["loss"].backward()
    arrays = {f"sample/{k[0]}/{k[1]}": v for k, v in prob["sample"].items()}
    for s in range(4):
        arrays[f"noise{s}"] = prob["noise"][s]
        arrays[f"loss/{s}"] = losses[f"loss/{s}"].detach()
        arrays[f"identity_selection/{s}"] = outputs[f"identity_selection/{s}"]
        arrays[f"depth{s}"] = outputs[("depth", s)].detach()
        arrays[f"color{s}_-1"] = outputs[("color", -1, s)].detach()
    arrays["loss"] = losses["loss"].detach()
    for net, tag in ((dnet, "depth_net"), (pnet, "pose_net")):
        for k, v in net.state_dict().items():
            arrays[f"{tag}/state/{k}"] = v
        for k, v in net.named_parameters():
            arrays[f"{tag}/grad/{k}"] = v.grad
    frac = [1 - float(outputs[f"identity_selection/{s}"].mean()) for s in range(4)]   # idx > 1  <=>  reprojection won
    print(f"{name}: loss={float(losses['loss']):.8f} identity-selected={['%.2f' % f for f in frac]}")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: v.detach().cpu().numpy() for k, v in arrays.items()})


def close(a, b, what):
    """Forward values are bit-identical; gradients may differ in the last bits because autograd
    accumulates the (out-of-place vs in-place) graph in a different order."""
    err = float((a - b).abs().max()) / (float(a.abs().max()) + 1e-30)
    assert err < 2e-6, (what, err)


def main():
    torch.set_num_threads(8)
    learner_new, learner_func = load_reference()
    make_process_batch_case(learner_new)
    if "--process-batch-only" in sys.argv:
        return
    for name, B, H, W, consistent, auto_mask, seed in CASES:
        kw = {}
        if "bigmotion" in name:
            kw = dict(pose_noise=2e-2, disp_noise=0.5)      # many out-of-view / clipped samples
        prob = make_problem(B, H, W, 2, 4, seed=seed, consistent=consistent, **kw)
        ref = run_reference(prob, auto_mask, learner_new, learner_func)
        mine = run_port(prob, auto_mask)
        # --- pin the oracle port to the live reference (CPU: bit-identical op sequence)
        assert torch.equal(ref["loss"], mine["loss"]), (name, ref["loss"], mine["loss"])
        assert torch.equal(ref["per_scale"], mine["per_scale"]), name
        for s in range(4):
            close(ref["grad_disp"][s], mine["grad_disp"][s], (name, "grad_disp", s))
            if auto_mask:
                assert torch.equal(ref["identity_selection"][s], (mine["sel"][s] > 1).float()), (name, s)
        for i in range(2):
            close(ref["grad_T"][i], mine["grad_T"][i], (name, "grad_T", i))
            close(ref["grad_axisangle"][i], mine["grad_axisangle"][i], (name, "grad_axisangle", i))
            close(ref["grad_translation"][i], mine["grad_translation"][i], (name, "grad_translation", i))
        frac = [float((mine["sel"][s] > 1).float().mean()) if auto_mask else 1.0 for s in range(4)]
        print(f"{name}: loss={float(ref['loss']):.8f} reproj-selected={['%.2f' % f for f in frac]}  port==reference OK")
        arrays = {
            "target": prob["target"], "K": prob["K"], "inv_K": prob["inv_K"],
            "loss": ref["loss"], "per_scale": ref["per_scale"], "auto_mask": torch.tensor(int(auto_mask)),
            "invert": torch.tensor([int(v) for v in prob["invert"]]),
        }
        for i in range(2):
            arrays[f"source{i}"] = prob["sources"][i]
            arrays[f"axisangle{i}"] = prob["axisangle"][i]
            arrays[f"translation{i}"] = prob["translation"][i]
            arrays[f"T{i}"] = ref["T"][i]
            arrays[f"grad_T{i}"] = ref["grad_T"][i]
            arrays[f"grad_axisangle{i}"] = ref["grad_axisangle"][i]
            arrays[f"grad_translation{i}"] = ref["grad_translation"][i]
        for s in range(4):
            arrays[f"disp{s}"] = prob["disps"][s]
            arrays[f"noise{s}"] = prob["noise"][s]
            arrays[f"grad_disp{s}"] = ref["grad_disp"][s]
            arrays[f"sel{s}"] = mine["sel"][s].to(torch.uint8)
            if name == CASES[0][0]:                      # intermediates for the granular-op tests: one case only
                arrays[f"depth{s}"] = ref["depth"][s]
                for i in range(2):
                    arrays[f"color{s}_{i}"] = ref["color"][s][i]
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            **{k: v.detach().cpu().numpy() for k, v in arrays.items()})


if __name__ == "__main__":
    main()

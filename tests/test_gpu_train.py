"""Training step on the GPU: the reference's train_mono_step contract around the fused loss."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(B, H, W, **kw):
    from vo.train import DEFAULT_CONFIG, Trainer
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["Train"].update(batch_size=B, img_h=H, img_w=W, init_lr=1e-4)
    torch.manual_seed(0)
    return Trainer(cfg, device=torch.device("cuda", 0), **kw)


@pytest.mark.parametrize("net_dtype", [None, torch.bfloat16])
def test_train_mono_step_contract_and_descent(net_dtype):
    from vo.train import synthetic_sample
    B, H, W = 2, 96, 128
    tr = _trainer(B, H, W, net_dtype=net_dtype, sync_losses=True)
    sample = synthetic_sample(B, H, W, seed=3)
    first = None
    for it in range(8):
        total, outputs, losses = tr.train_mono_step(dict(sample))
        assert set(losses) == {"loss", "loss/0", "loss/1", "loss/2", "loss/3"}
        assert all(not v.is_cuda and v.dim() == 0 for v in losses.values())      # vo/train.py:196-197: .detach().cpu()
        assert torch.isfinite(total)
        assert outputs[("disp", 0)].shape == (B, 1, H, W) and outputs[("cam_T_cam", 0, -1)].shape == (B, 4, 4)
        first = float(total) if first is None else first
    assert float(total) < first                     # Adam on a fixed batch: the loss goes down


def test_train_step_matches_reference_style_learner():
    """The joint (DDP-friendly) forward feeds the learner the same tensors as calling the networks one by one."""
    from vo.learner_new import MonodepthTrainer
    from vo.train import synthetic_sample
    B, H, W = 2, 64, 96
    tr = _trainer(B, H, W, noise="torch", sync_losses=False, channels_last=False)
    sample = synthetic_sample(B, H, W, seed=4, device="cuda")
    torch.manual_seed(11)
    tr.joint.run(sample)
    _, l1 = tr.learner.process_batch(dict(sample))
    plain = MonodepthTrainer(tr.depth_net, tr.pose_net, tr.config, tr.device, noise="torch")
    torch.manual_seed(11)
    _, l2 = plain.process_batch(dict(sample))
    for k in l1:
        assert torch.allclose(l1[k], l2[k], rtol=1e-6, atol=0), k


def test_whole_training_step_replays_from_a_cuda_graph():
    """zero_grad -> networks -> fused loss (pose chain inside its launches) -> backward -> Adam captured as
    one CUDA graph; replays keep optimising and the in-kernel noise counter advances on the device."""
    import dvsloss
    from vo.train import synthetic_sample
    B, H, W = 2, 96, 128
    tr = _trainer(B, H, W, net_dtype=torch.bfloat16, sync_losses=False)
    sample = synthetic_sample(B, H, W, seed=5, device="cuda")
    # eager steps first (a resumed or warmed-up trainer): the modules then hold the last step's outputs, whose autograd graph
    # pins gradient accumulators to the legacy stream; capture_step has to drop them or the capture is invalidated
    for _ in range(2):
        total, _, _ = tr.train_mono_step(dict(sample))
    del total, _
    tr.capture_step(sample)
    n0 = dvsloss.noise_state()
    losses = []
    for it in range(6):
        total, outputs, ls = tr.train_graph_step(sample)
        losses.append(float(total))
        assert set(ls) == {"loss", "loss/0", "loss/1", "loss/2", "loss/3"} and all(v.is_cuda for v in ls.values())
    assert dvsloss.noise_state() == n0 + 6                        # one draw per replayed step
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]     # Adam on a fixed batch: the loss goes down
    # a new batch through the same graph
    other = synthetic_sample(B, H, W, seed=6, device="cuda")
    t2, _, _ = tr.train_graph_step(other)
    assert np.isfinite(float(t2)) and float(t2) != losses[-1]


def test_pose_parameters_inside_the_loss_match_the_matrix_path():
    """view_synthesis_loss(axisangles=, translations=, inverts=) == transformation_from_parameters + view_synthesis_loss(Ts=):
    losses bit-equal, gradients w.r.t. the six pose numbers and the disparities equal to fp32 round-off."""
    from dvsloss import ops, view_synthesis_loss
    from dvsloss.synthetic import make_problem
    dev = torch.device("cuda:0")
    B, H, W = 2, 96, 128
    p = make_problem(B, H, W, 2, 4, seed=8, consistent=True)
    cut = lambda t: t.to(dev).contiguous()
    noise = [cut(n) for n in p["noise"]]
    res = []
    for mode in ("params", "matrices"):
        disps = [cut(d).requires_grad_(True) for d in p["disps"]]
        aa = [cut(a).requires_grad_(True) for a in p["axisangle"]]           # [B,1,3] as PoseNet returns after [:, 0]
        tr = [cut(t).requires_grad_(True) for t in p["translation"]]
        common = (disps, cut(p["target"]), [cut(s) for s in p["sources"]], cut(p["K"]), cut(p["inv_K"]))
        if mode == "params":
            out = view_synthesis_loss(*common, axisangles=aa, translations=tr, inverts=p["invert"], noise=noise)
        else:
            Ts = [ops.transformation_from_parameters(a, t, inv) for a, t, inv in zip(aa, tr, p["invert"])]
            out = view_synthesis_loss(*common, Ts, noise=noise)
        out[0].backward()
        res.append((out, [d.grad for d in disps], [a.grad for a in aa], [t.grad for t in tr]))
    (o1, gd1, ga1, gt1), (o2, gd2, ga2, gt2) = res
    assert torch.equal(o1[0], o2[0]) and torch.equal(o1[1], o2[1])
    for a, b in zip(gd1, gd2):
        assert torch.equal(a, b)
    for a, b in zip(ga1 + gt1, ga2 + gt2):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-6 * float(b.abs().max()) + 1e-12


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pack_net_inputs_gives_the_bits_of_the_stock_sequence(dtype):
    """channels-last copies + pose-pair concatenation + (x - 0.45) / 0.225 + autocast's cast at conv1, as one kernel."""
    from dvsloss import ops
    torch.manual_seed(0)
    B, H, W = 3, 20, 28
    tgt, s0, s1, s2 = (torch.rand(B, 3, H, W, device="cuda") for _ in range(4))
    t_in, pairs = ops.pack_net_inputs(tgt, [s0, s1, s2], [True, False, True], dtype)
    norm = lambda x: ((x - 0.45) / 0.225)
    ref_t = norm(tgt).to(dtype)
    refs = [norm(torch.cat([s0, tgt], 1)).to(dtype), norm(torch.cat([tgt, s1], 1)).to(dtype), norm(torch.cat([s2, tgt], 1)).to(dtype)]
    assert t_in.is_contiguous(memory_format=torch.channels_last) and torch.equal(t_in, ref_t)
    for p, r in zip(pairs, refs):
        assert p.shape == (B, 6, H, W) and p.is_contiguous(memory_format=torch.channels_last) and torch.equal(p, r)
    raw_t, raw_p = ops.pack_net_inputs(tgt, [s0], [False], torch.float32, normalize=False)
    assert torch.equal(raw_t, tgt) and torch.equal(raw_p[0], torch.cat([tgt, s0], 1))
    assert not ops.pack_net_inputs_supported(tgt[:, :, :, :27].contiguous()[:, :, :19].contiguous(), [s0])      # 19 * 27 is odd
    with pytest.raises(Exception):
        ops.pack_net_inputs(tgt.cpu(), [s0.cpu()], [True])


def test_training_step_with_packed_inputs_equals_the_stock_input_path():
    from vo.train import JointForward, synthetic_sample
    B, H, W = 2, 96, 128
    sample = synthetic_sample(B, H, W, seed=7, device="cuda")
    out = {}
    for packed in (True, False):
        JointForward.pack_inputs = packed
        try:
            torch.manual_seed(3)
            tr = _trainer(B, H, W, net_dtype=torch.bfloat16, sync_losses=False, noise=None)
            total, _, losses = tr.train_mono_step(dict(sample))
            out[packed] = (float(total), [float(losses[f"loss/{s}"]) for s in range(4)])
        finally:
            JointForward.pack_inputs = True
    assert out[True] == out[False]                                # the networks read the same bits either way

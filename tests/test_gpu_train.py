"""Training step on the GPU: the reference's train_mono_step contract around the fused loss."""
import copy

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

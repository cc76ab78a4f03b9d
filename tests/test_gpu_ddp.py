"""SURVEY T5 / 8e on GPUs: batch-sharded training over 2 ranks (NCCL, DistributedDataParallel, the FUSED loss) reproduces the
gradients of one rank on the global batch.  Needs two visible devices (gpurun --gpus 2); skipped otherwise."""
import copy
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _cfg(B, H, W):
    from vo.train import DEFAULT_CONFIG
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["Train"].update(batch_size=B, img_h=H, img_w=W)
    return cfg


def _grads(tr):
    return torch.cat([p.grad.detach().flatten().float().cpu() for p in tr.nets.parameters() if p.requires_grad and p.grad is not None])


def _exact_convs():
    """TF32 convolutions (cuDNN's default) carry ~1e-3 of rounding that depends on the batch size: switch them off so that the
    comparison tests the sharding, not the tensor-core rounding mode."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _one_backward(tr, sample):
    """forward + backward of the training step without the optimizer update (so the gradients can be compared)."""
    tr.optimizer.zero_grad(set_to_none=True)
    tr.joint.run(sample)
    _, losses = tr.learner.process_batch(sample)
    losses["loss"].backward()
    return float(losses["loss"])


def _worker(rank, world, port, out_path, B, H, W):
    from vo.train import Trainer, synthetic_sample
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    _exact_convs()
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    tr = Trainer(_cfg(B // world, H, W), device=dev, distributed=True, noise=None, sync_losses=False, channels_last=False)
    tr.nets.eval()                                                    # BatchNorm on running statistics: per-rank batches do not matter
    full = synthetic_sample(B, H, W, seed=9, device=dev)
    n = B // world
    shard = {k: v[rank * n:(rank + 1) * n].contiguous() for k, v in full.items()}
    loss = _one_backward(tr, shard)
    if rank == 0:
        torch.save({"grads": _grads(tr), "loss": loss}, out_path)
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_training_matches_single_rank_global_batch(tmp_path):
    from vo.train import Trainer, synthetic_sample
    world, B, H, W = 2, 4, 96, 128
    out_path = str(tmp_path / "rank0.pt")
    port = 29600 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, out_path, B, H, W)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0, "a rank failed"
    got = torch.load(out_path)
    dev = torch.device("cuda", 0)
    tf32 = torch.backends.cudnn.allow_tf32
    _exact_convs()
    torch.manual_seed(0)
    tr = Trainer(_cfg(B, H, W), device=dev, distributed=False, noise=None, sync_losses=False, channels_last=False)
    tr.nets.eval()
    loss = _one_backward(tr, synthetic_sample(B, H, W, seed=9, device=dev))
    ref = _grads(tr)
    torch.backends.cudnn.allow_tf32 = tf32
    # the loss of a rank is the mean over its shard; DDP averages the gradients: together the global-batch mean
    assert got["grads"].shape == ref.shape
    scale = float(ref.abs().max())
    assert float((got["grads"] - ref).abs().max()) <= 1e-4 * scale, float((got["grads"] - ref).abs().max()) / scale
    assert torch.isfinite(torch.tensor(got["loss"])) and torch.isfinite(torch.tensor(loss))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_devices_in_one_process_give_the_same_bits():
    """The library keeps its per-kernel set-up (dynamic shared-memory opt-in) and its workspaces per device ordinal: the
    fused loss called on cuda:0 and then on cuda:1 from ONE process gives bit-identical losses and gradients."""
    from dvsloss import view_synthesis_loss
    from dvsloss.synthetic import make_problem, pose_matrix
    B, H, W = 2, 96, 128
    p = make_problem(B, H, W, 2, 4, seed=21, consistent=True)
    outs = []
    for d in (0, 1, 0):
        dev = torch.device("cuda", d)
        with torch.cuda.device(dev):
            mv = lambda v: v.to(dev) if torch.is_tensor(v) else ([mv(t) for t in v] if isinstance(v, (list, tuple)) else v)
            q = {k: mv(v) for k, v in p.items() if k != "sample"}
            Ts = [pose_matrix(a.view(B, 3), t.view(B, 3), inv).to(dev).requires_grad_(True)
                  for a, t, inv in zip(p["axisangle"], p["translation"], p["invert"])]
            disps = [x.clone().requires_grad_(True) for x in q["disps"]]
            loss, per_scale = view_synthesis_loss(disps, q["target"], q["sources"], q["K"], q["inv_K"], Ts, noise=None)
            loss.backward()
            torch.cuda.synchronize(dev)
            outs.append([loss.detach().cpu(), per_scale.detach().cpu()] + [x.grad.cpu() for x in disps + Ts])
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert torch.equal(a, b)

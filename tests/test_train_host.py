"""Host-side logic of the training step on CPU: network contracts, the joint DDP forward, and batch sharding
over two ranks (gloo) reproducing the single-process global-batch gradients.  The fused loss itself
has no CPU path, so a plain torch stand-in loss with the same reduction structure (batch means) is used here."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-visual-slam_b200"))

from model.depthnet import DepthNet  # noqa: E402
from model.posenet_single import PoseNet  # noqa: E402
from vo.train import JointForward, VoNets, _Bound  # noqa: E402


def test_network_contracts():
    d, p = DepthNet(18, False), PoseNet(18, False)
    assert sum(x.numel() for x in d.parameters()) == 14842236       # 14.84 M + 13.01 M
    assert sum(x.numel() for x in p.parameters()) == 13011950
    out = d(torch.rand(2, 3, 64, 96))
    assert {k: tuple(v.shape) for k, v in out.items()} == {("disp", s): (2, 1, 64 >> s, 96 >> s) for s in range(4)}
    assert all(float(v.min()) >= 0 and float(v.max()) <= 1 for v in out.values())
    a, t = p(torch.rand(2, 6, 64, 96))
    assert a.shape == t.shape == (2, 1, 1, 3)
    d50 = DepthNet(50, False)
    assert tuple(d50(torch.rand(1, 3, 64, 64))[("disp", 0)].shape) == (1, 1, 64, 64)


def _sample(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return {("target_image", 0): torch.rand(B, 3, H, W, generator=g), ("source_left", 0): torch.rand(B, 3, H, W, generator=g),
            ("source_right", 0): torch.rand(B, 3, H, W, generator=g)}


def _standin_loss(joint, sample):
    """Calls the two networks in the learner's order (depth, pose(-1), pose(+1)) and reduces with batch means."""
    depth_net, pose_net = _Bound(joint, "depth"), _Bound(joint, "pose")
    disp = depth_net(sample[("target_image", 0)])
    aa_l, t_l = pose_net(None)
    aa_r, t_r = pose_net(None)
    loss = sum(v.mean() for v in disp.values()) / 4
    return loss + 100 * ((aa_l - 2 * t_r) ** 2).mean() + 100 * ((aa_r + t_l) ** 2).mean()


def _grads(nets):
    return torch.cat([p.grad.flatten() for p in nets.parameters() if p.requires_grad])


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    nets = VoNets(DepthNet(18, False), PoseNet(18, False)).eval()          # eval: BatchNorm uses running stats
    ddp = torch.nn.parallel.DistributedDataParallel(nets)
    full = _sample(4, 64, 96, 7)
    shard = {k: v[rank * 2:(rank + 1) * 2] for k, v in full.items()}
    joint = JointForward(ddp)
    joint.run(shard)
    _standin_loss(joint, shard).backward()
    if rank == 0:
        torch.save(_grads(nets).clone(), out_path)
    dist.destroy_process_group()


def test_two_rank_sharding_matches_global_batch(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    out_path = str(tmp_path / "grads_rank0.pt")
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, out_path)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0, "a rank failed"
    g_ddp = torch.load(out_path)
    torch.manual_seed(0)
    nets = VoNets(DepthNet(18, False), PoseNet(18, False)).eval()
    full = _sample(4, 64, 96, 7)
    joint = JointForward(nets)
    joint.run(full)
    _standin_loss(joint, full).backward()
    g_ref = _grads(nets)
    assert torch.allclose(g_ddp, g_ref, rtol=1e-4, atol=1e-7), float((g_ddp - g_ref).abs().max())


def test_host_pipeline_chunk_rules():
    """dvsloss.host.chunk_sizes: the batch split of the host-resident pipeline (pure host logic)."""
    from dvsloss.host import chunk_sizes
    assert chunk_sizes(16, "taper") == [6, 4, 3, 2, 1] and chunk_sizes(16, "ramp") == [1, 2, 3, 4, 6]
    assert chunk_sizes(16, 8) == [2] * 8 and chunk_sizes(16, [5, 4, 3, 2, 2]) == [5, 4, 3, 2, 2]
    for B in range(1, 70):
        for rule in ("taper", "ramp"):
            s = chunk_sizes(B, rule)
            assert sum(s) == B and min(s) >= 1
            assert s == sorted(s, reverse=(rule == "taper"))            # tapering: the short chunk last; ramp: first
    for bad in (3, [8, 7], [16, 0], "steps"):
        with pytest.raises(ValueError):
            chunk_sizes(16, bad)

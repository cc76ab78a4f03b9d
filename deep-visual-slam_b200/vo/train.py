"""VO training step on B200: stock PyTorch networks + the fused view-synthesis loss, batch-sharded over GPUs.

This is the caller of the accelerated path, kept to the contract of the reference's ``Trainer``
(vo/train.py:38-199): the same config dict (vo/config.yaml), Adam + PolynomialLR, and
``train_mono_step(sample) -> (total_loss, outputs, losses)`` = ``zero_grad(set_to_none)`` ->
``learner.process_batch`` -> ``backward`` -> ``optimizer.step``.  What is new relative to the reference:

  * multi-GPU: one process per GPU (``torchrun``), DepthNet and PoseNet wrapped in ONE
    ``DistributedDataParallel`` module so PoseNet's two forward calls live inside a single DDP forward; the
    gradient all-reduce is NCCL over NVLink/NVSwitch, overlapped with the convolution backward by DDP's buckets.
    The loss shards over the batch with no collective of its own (every reduction is per image or a batch mean,
    so equal local batches + gradient averaging reproduce the global-batch gradients: SURVEY 8e).
  * ``net_dtype=torch.bfloat16``: autocast for the networks only; the loss op computes in fp32 whatever the
    autocast state (the geometry is unusable in half precision).
  * ``sync_losses=False`` keeps the five loss scalars on the device (the reference copies them to the host
    every step, vo/train.py:196-197, which serialises the step).

The dataset side of the reference (vo/dataset/*, disk I/O) is out of scope; ``synthetic_sample`` builds a
Redwood-shaped batch with the sample-dict format of ``MonoDataset.__getitem__`` (vo/dataset/common.py:48-92).
"""
from __future__ import annotations

import os
import sys
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
from torch import optim
from torch.optim.lr_scheduler import PolynomialLR

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from model.depthnet import DepthNet  # noqa: E402
from model.posenet_single import PoseNet  # noqa: E402
from vo.learner_new import MonodepthTrainer  # noqa: E402

DEFAULT_CONFIG = {
    "Directory": {"exp_name": "b200"},
    "Train": dict(mode="axisAngle", use_amp=False, use_compile=False, num_source=1, num_scale=4, min_depth=0.1,
                  max_depth=10.0, ssim_ratio=0.85, smoothness_ratio=0.001, auto_mask=True, img_w=640, img_h=480,
                  weight_decay=0.00001, beta1=0.9, batch_size=16, epoch=31, init_lr=0.0001, final_lr=0.00001),
}


class VoNets(nn.Module):
    """DepthNet + PoseNet behind one ``forward`` (so DDP sees a single forward per step)."""

    def __init__(self, depth_net: nn.Module, pose_net: nn.Module):
        super().__init__()
        self.depth_net, self.pose_net = depth_net, pose_net

    def forward(self, target: torch.Tensor, pairs, normalized: bool = False):
        if normalized:                                   # inputs built by dvsloss.ops.pack_net_inputs
            return self.depth_net(target, True), [self.pose_net(p, True) for p in pairs]
        disp = self.depth_net(target)
        poses = [self.pose_net(p) for p in pairs]
        return disp, poses


class _Bound:
    """Callable view of one network of a (possibly DDP-wrapped) VoNets whose results were computed by the joint
    forward; the learner calls ``depth_net(x)`` / ``pose_net(pair)`` in the reference's order."""

    def __init__(self, owner: "JointForward", kind: str):
        self.owner, self.kind = owner, kind

    def __call__(self, x):
        return self.owner.take(self.kind, x)

    def parameters(self):
        m = self.owner.module
        m = m.module if hasattr(m, "module") else m
        return (m.depth_net if self.kind == "depth" else m.pose_net).parameters()


class JointForward:
    """Runs VoNets once per step and hands the results to the learner's separate network calls."""

    def __init__(self, module: nn.Module, frame_ids=(-1, 1), channels_last: bool = False):
        self.module, self.frame_ids, self.channels_last = module, list(frame_ids), channels_last
        self._disp, self._poses = None, []

    pack_inputs = True       # CUDA: one kernel builds the channels-last, normalised, autocast-dtype network inputs

    def run(self, sample: Dict) -> None:
        tgt = sample[("target_image", 0)]
        srcs = [sample[("source_left", 0) if f == -1 else ("source_right", 0) if f == 1 else ("source", f)] for f in self.frame_ids]
        if self.pack_inputs and self.channels_last and tgt.is_cuda:       # the packed inputs are channels-last
            from dvsloss import ops
            if ops.pack_net_inputs_supported(tgt, srcs):
                dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
                if dt in (torch.float32, torch.bfloat16):
                    t_in, pairs = ops.pack_net_inputs(tgt, srcs, [f < 0 for f in self.frame_ids], dt)
                    self._disp, poses = self.module(t_in, pairs, True)
                    self._poses = list(poses)
                    return
        if self.channels_last:
            tgt = tgt.contiguous(memory_format=torch.channels_last)
            srcs = [s.contiguous(memory_format=torch.channels_last) for s in srcs]
        pairs = [torch.cat([src, tgt], 1) if f < 0 else torch.cat([tgt, src], 1) for f, src in zip(self.frame_ids, srcs)]
        self._disp, poses = self.module(tgt, pairs)
        self._poses = list(poses)

    def take(self, kind: str, x):
        if kind == "depth":
            return dict(self._disp)
        return self._poses.pop(0)


class Trainer:
    def __init__(self, config: Optional[dict] = None, device: Optional[torch.device] = None, num_layers: int = 18,
                 pretrained: bool = False, net_dtype: Optional[torch.dtype] = None, distributed: bool = False,
                 noise: str = "kernel", sync_losses: bool = True, channels_last: bool = True, frame_ids=(-1, 1)):
        self.config = config or DEFAULT_CONFIG
        tr = self.config["Train"]
        self.device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        self.net_dtype = net_dtype
        self.sync_losses = sync_losses
        self.channels_last = channels_last and self.device.type == "cuda"
        depth_net = DepthNet(num_layers=num_layers, pretrained=pretrained, num_input_images=1)
        pose_net = PoseNet(num_layers=num_layers, pretrained=pretrained, num_input_images=2)
        nets = VoNets(depth_net, pose_net).to(self.device)
        if self.channels_last:
            nets = nets.to(memory_format=torch.channels_last)
        self.nets = nets
        module: nn.Module = nets
        if distributed:
            from torch.nn.parallel import DistributedDataParallel as DDP
            ids = [self.device.index] if self.device.type == "cuda" else None
            import os as _os
            # broadcast_buffers=False: DDP's default re-broadcasts every BatchNorm running statistic from rank 0 before each forward
            # (a host-synchronised collective per step, 1.1 ms of a 59 ms step at 2 GPUs); the statistics are only used in eval
            # mode and only rank 0 writes checkpoints, so each rank keeps its own, as is usual for non-synchronised BatchNorm
            opts = dict(gradient_as_bucket_view=True, broadcast_buffers=False)
            if _os.environ.get("DVS_DDP_OPTS"):                   # experiment hook: e.g. "broadcast_buffers=False,static_graph=True,bucket_cap_mb=50"
                for kv in _os.environ["DVS_DDP_OPTS"].split(","):
                    k, v = kv.split("=")
                    opts[k] = {"True": True, "False": False}.get(v, int(v) if v.isdigit() else v)
            module = DDP(nets, device_ids=ids, **opts)
        self.module = module
        self.depth_net, self.pose_net = depth_net, pose_net
        params = [q for q in list(depth_net.parameters()) + list(pose_net.parameters()) if q.requires_grad]
        # fused=True: one multi-tensor kernel for the whole Adam update instead of ~10 launches per parameter group
        # capturable: the step counter lives on the device, so the update can sit inside a captured CUDA graph
        self.optimizer = optim.Adam(params, lr=tr["init_lr"], fused=self.device.type == "cuda",
                                    capturable=self.device.type == "cuda")
        self._graph = None
        self.scheduler = PolynomialLR(self.optimizer, total_iters=tr["epoch"], power=0.9)
        self.frame_ids = list(frame_ids)
        self.joint = JointForward(module, self.frame_ids, self.channels_last)
        self.learner = MonodepthTrainer(_Bound(self.joint, "depth"), _Bound(self.joint, "pose"), self.config, self.device,
                                        frame_ids=self.frame_ids, noise=noise)

    def train_mono_step(self, sample: Dict) -> Tuple[torch.Tensor, Dict, Dict]:
        """reference: vo/train.py:173-199 (the non-AMP branch; bf16 autocast needs no GradScaler)."""
        self.optimizer.zero_grad(set_to_none=True)
        for key, val in sample.items():
            if isinstance(val, torch.Tensor):
                sample[key] = val.to(self.device, non_blocking=True)
        if self.net_dtype is not None:
            with torch.autocast(self.device.type, dtype=self.net_dtype):
                self.joint.run(sample)
        else:
            self.joint.run(sample)
        outputs, losses = self.learner.process_batch(sample)
        total = losses["loss"]
        total.backward()
        self.optimizer.step()
        total = total.detach()
        for k in losses:
            losses[k] = losses[k].detach().cpu() if self.sync_losses else losses[k].detach()
        return total, outputs, losses

    # ------------------------------------------------------------------------------------------ whole-step CUDA graph
    def capture_step(self, sample: Dict, warmup: int = 3) -> None:
        """Capture zero_grad -> networks -> fused loss -> backward -> Adam as ONE CUDA graph.  The reference
        launches ~1 150 loss kernels per step and copies five scalars to the host after every step (vo/train.py:196-197, :253),
        which serialises host and device; here a step is a host-side copy of the batch into static buffers plus one
        ``graph.replay()`` and the losses stay on the device.  The in-kernel automask noise keeps advancing under replay (its
        step counter is a device scalar incremented inside the graph).  Single-process only (DDP's bucketing hooks are not
        captured); shapes are frozen to those of ``sample``."""
        if self.device.type != "cuda":
            raise RuntimeError("capture_step needs a CUDA device")
        if self.module is not self.nets:
            raise RuntimeError("capture_step does not support DistributedDataParallel")
        self._static = {k: (v.to(self.device).clone() if isinstance(v, torch.Tensor) else v) for k, v in sample.items()}
        sync, self.sync_losses = self.sync_losses, False
        # Earlier eager steps leave autograd graphs alive (DepthNet keeps its last ``outputs`` and the encoder its last ``features``, as the
        # reference's do, model/depthnet.py:89, model/resnet_encoder.py); their AccumulateGrad nodes belong to the legacy default stream, and a capture that meets them
        # fails with cudaErrorStreamCaptureImplicit.  Drop them so that the warm-up below recreates the nodes on a side stream.
        self.joint._disp, self.joint._poses = None, []
        for m in self.nets.modules():
            if isinstance(getattr(m, "outputs", None), dict):
                m.outputs = {}
            if isinstance(getattr(m, "features", None), list):    # ResnetEncoder keeps its last feature pyramid
                m.features = []
        import gc
        gc.collect()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.train_mono_step(dict(self._static))
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.optimizer.zero_grad(set_to_none=True)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            total, outputs, losses = self.train_mono_step(dict(self._static))
        self._graph_out = (total, {k: v for k, v in dict.items(outputs) if isinstance(v, torch.Tensor)}, losses)
        self.sync_losses = sync

    def train_graph_step(self, sample: Dict) -> Tuple[torch.Tensor, Dict, Dict]:
        """One optimisation step by replaying the captured graph on ``sample`` (same shapes as at capture).  Returns the
        static (device-resident) total loss, outputs and losses: read them before the next replay overwrites them."""
        if self._graph is None:
            raise RuntimeError("call capture_step(sample) first")
        with torch.no_grad():
            for k, v in sample.items():
                if isinstance(v, torch.Tensor):
                    self._static[k].copy_(v, non_blocking=True)
        self._graph.replay()
        return self._graph_out

    def _images(self, sample: Dict) -> Dict:
        if not self.channels_last:
            return sample
        out = dict(sample)
        for k, v in sample.items():
            if isinstance(v, torch.Tensor) and v.dim() == 4 and v.shape[1] == 3:
                out[k] = v.contiguous(memory_format=torch.channels_last)
        return out


def synthetic_sample(B: int, H: int, W: int, seed: int = 0, device="cpu", num_sources: int = 2) -> Dict:
    """Redwood-shaped (t-1, t, t+1[, t-2, t+2]) batch in the collated sample-dict format."""
    from dvsloss.synthetic import make_problem
    p = make_problem(B, H, W, num_sources, 4, seed=seed, consistent=True)
    return {k: v.to(device) for k, v in p["sample"].items()}

#!/usr/bin/env python
"""Benchmark of the hot path: fused view-synthesis (photometric) loss forward+backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): 640x480, batch 16 per GPU, 2 source frames, 4 scales, fp32, synthetic
Redwood-shaped geometrically consistent triplets.  One step = one forward + backward of the loss
for one batch.  Metric: warped pixels per second, B*S*N*H*W / t, whole job.  Prints ONE JSON line on rank 0.

  value        device-timed (CUDA events), inputs resident in HBM, L2 flushed between steps; the public call
               (view_synthesis_loss + backward: four kernel launches) is recorded once as a CUDA graph and replayed
               (config.launch; --eager-launch times the plain Python calls: +2 %)
  e2e          same step through the public host-resident API (dvsloss.HostLossPipeline) from pinned HOST buffers:
               H2D of every input, D2H of the losses and all gradients, inside the timed region (batch chunks
               overlap copy and compute on three streams; PCIe-bound)
  roofline     dominant kernel (fused_pair_kernel, the two-source tile kernel) timed with events around its launch inside libdvsloss.so;
               achieved = bytes_alg(B,H,W,N,S) / t   (SURVEY 8d figure, 303.9 B per full-res pixel at N=2,S=4)
  cpu_baseline the unmodified reference staged in baseline/_ref (kind "reference"; the oracle port only where the copy is
               absent) on the host cores, on a bounded sample (batch 2 of the same workload); rank 0, N=1 only
  eager_cuda_baseline   the same reference code run eagerly on the GPU (the reference's own CUDA path), N=1 only
  noise_torch  the same step with torch.randn noise tensors (the reference's RNG contract) instead of the in-kernel generator
  train_step_r50x4   BASELINE configs[3]: ResNet-50, 1280x960, sources +-1 and +-2, batch 8/GPU (5 steps)
  train_step   BASELINE configs[2]: full VO training step (stock ResNet-18 DepthNet + PoseNet in bf16 autocast, fused
               fp32 loss, Adam, batch 32/GPU, DDP/NCCL when N>1) in triplets per second, whole job
  --impl reference   times that CPU path alone (all host threads), same metric/config
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "deep-visual-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

H, W, NSRC, NSCALE = 480, 640, 2, 4
METRIC = "photometric_loss_fwd_bwd_warped_pixels_per_s"
UNIT = "Gpix/s"


_REAL_STDOUT = None


def emit(text: str) -> None:
    """The one JSON line: to the real stdout (see main)."""
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (text + "\n").encode())


def workload_text(B):
    return (f"isolated photometric loss fwd+bwd, {W}x{H}, batch {B}/GPU, {NSRC} sources, {NSCALE} scales "
            f"(BASELINE configs[1]); consistent synthetic triplets")


def bytes_alg(B, Hh, Ww, N, S):
    """compulsory fp32 traffic of one forward + one backward pass."""
    return B * Hh * Ww * (2 * S * 12 * (1 + N) + 12 * sum(4.0 ** -s for s in range(S)))


def bytes_floor(B, Hh, Ww, N, S):
    """Strict single-pass floor: every image once, disparities read once, gradients written once."""
    return B * Hh * Ww * (12 * (1 + N) + 8 * sum(4.0 ** -s for s in range(S)))


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(index):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU (pinned staging buffers are then allocated
    and filled on that NUMA node).  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"cpus {cpus[0]}-{cpus[-1]} ({len(cpus)}) local to GPU {index}"
    except Exception as e:
        return f"not bound ({type(e).__name__})"
    return "not bound"


def make_inputs(B, seed, device):
    from dvsloss.synthetic import make_problem, pose_matrix
    p = make_problem(B, H, W, NSRC, NSCALE, seed=seed, consistent=True)
    Ts = [pose_matrix(a.view(B, 3), t.view(B, 3), inv) for a, t, inv in zip(p["axisangle"], p["translation"], p["invert"])]
    host = dict(target=p["target"], sources=p["sources"], disps=p["disps"], K=p["K"], inv_K=p["inv_K"], Ts=Ts)
    return host


# ------------------------------------------------------------------------------------------------ reference arms
class ReferenceLoss:
    """The reference's own loss path on ``device``: the UNMODIFIED ``MonodepthTrainer._generate_images_pred`` +
    ``_compute_losses`` + ``backward`` (vo/learner_new.py:132-258) from the staged copy ``baseline/_ref`` (kind "reference");
    where that copy is absent, the op-for-op restatement ``oracle/reference_port.py`` (kind "port").  Inputs are resident
    on ``device`` before the first step; a step draws the automask noise like the reference (torch.randn per scale)."""

    def __init__(self, host, B_s, device):
        from baseline import reference_loader as rl
        self.dev = torch.device(device)
        cut = lambda t: t[:B_s].to(self.dev).contiguous()
        self.B = B_s
        self.target, self.sources = cut(host["target"]), [cut(s) for s in host["sources"]]
        self.K, self.inv_K = cut(host["K"]), cut(host["inv_K"])
        self.disps = [cut(d).requires_grad_(True) for d in host["disps"]]
        self.Ts = [cut(T).requires_grad_(True) for T in host["Ts"]]
        self.kind = "reference" if rl.available() else "port"
        if self.kind == "reference":
            cfg = {"Train": dict(num_source=2, batch_size=B_s, img_h=H, img_w=W, smoothness_ratio=0.001, auto_mask=True,
                                 ssim_ratio=0.85, min_depth=0.1, max_depth=10.0, use_compile=False)}
            self.trainer = rl.load_learner().MonodepthTrainer(None, None, cfg, self.dev)
            self.sample = {("K", 0): self.K, ("inv_K", 0): self.inv_K, ("target_image", 0): self.target,
                           ("source_left", 0): self.sources[0], ("source_right", 0): self.sources[1]}

    def step(self):
        for t in self.disps + self.Ts:
            t.grad = None
        if self.kind == "reference":
            outputs = {("disp", s): self.disps[s] for s in range(NSCALE)}
            outputs[("cam_T_cam", 0, -1)], outputs[("cam_T_cam", 0, 1)] = self.Ts
            self.trainer._generate_images_pred(self.sample, outputs)
            loss = self.trainer._compute_losses(self.sample, outputs)["loss"]
        else:
            from oracle import reference_port as port
            noise = [torch.randn(self.B, NSRC, H, W, device=self.dev) for _ in range(NSCALE)]
            loss = port.view_synthesis_loss(self.disps, self.target, self.sources, self.K, self.inv_K, self.Ts, noise)["loss"]
        loss.backward()
        return loss

    def what(self):
        src = ("the unmodified MonodepthTrainer._generate_images_pred + _compute_losses + backward staged in baseline/_ref"
               if self.kind == "reference" else "oracle/reference_port.py (op-for-op restatement of the reference's path)")
        return f"{src}, fp32, batch {self.B}, inputs resident on {self.dev.type}"


def time_cpu(host, B_s, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ReferenceLoss(host, B_s, "cpu")
    for _ in range(warmup):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.step()
    dt = (time.perf_counter() - t0) / steps
    return B_s * NSCALE * NSRC * H * W / dt / 1e9, dt, ref


def time_cpu_train_step(B_s=2, steps=2, warmup=1):
    """BASELINE configs[0]: one vo/train.py optimisation step on the CPU, fp32, batch 2, ResNet-18 DepthNet + PoseNet:
    zero_grad -> process_batch -> backward -> Adam (vo/train.py:173-199) around the reference's MonodepthTrainer and
    networks when staged (else this repository's stock-PyTorch networks with the restated loss)."""
    from baseline import reference_loader as rl
    from dvsloss.synthetic import make_problem
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    sample = make_problem(B_s, H, W, NSRC, NSCALE, seed=7, consistent=True)["sample"]
    if rl.available():
        DepthNet, PoseNet = rl.load_nets()
        dn, pn = DepthNet(num_layers=18, pretrained=False), PoseNet(num_layers=18, pretrained=False, num_input_images=2)
        cfg = {"Train": dict(num_source=2, batch_size=B_s, img_h=H, img_w=W, smoothness_ratio=0.001, auto_mask=True,
                             ssim_ratio=0.85, min_depth=0.1, max_depth=10.0, use_compile=False)}
        learner = rl.load_learner().MonodepthTrainer(dn, pn, cfg, torch.device("cpu"))
        kind = "reference"
    else:
        return None
    opt = torch.optim.Adam(list(dn.parameters()) + list(pn.parameters()), lr=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        _, losses = learner.process_batch(dict(sample))
        losses["loss"].backward()
        opt.step()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"s_per_step": dt, "triplets_per_s": B_s / dt, "kind": kind, "batch": B_s,
            "what": "BASELINE configs[0]: full VO training step (ResNet-18 DepthNet + 2x PoseNet + loss + Adam) on the host cores, fp32"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B_s = args.cpu_batch
    host = make_inputs(B_s, 0, "cpu")
    val, dt, ref = time_cpu(host, B_s, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = f"batch {B_s} of the {W}x{H}, {NSRC} sources, {NSCALE} scales workload per step: {ref.what()}"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(args.batch), "sample": f"bounded CPU sample: batch {B_s} per step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------ eager CUDA leg
def time_eager_cuda(host, dev, steps=20, warmup=3):
    """The reference's own eager CUDA-PyTorch loss (see ReferenceLoss) on the SAME GPU: the denominator of north_star's
    ">= 20x the reference's own eager CUDA-PyTorch loss throughput".  Inputs resident, >= 3 warm-ups, >= 20 timed steps."""
    B = host["target"].shape[0]
    ref = ReferenceLoss(host, B, dev)
    for _ in range(warmup):
        ref.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ref.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "value": B * NSCALE * NSRC * H * W / (ms * 1e-3) / 1e9, "unit": UNIT, "kind": ref.kind,
            "steps": steps, "warmup": warmup, "what": ref.what()}


# ------------------------------------------------------------------------------------------------ training-step leg
def run_train(args, world, rank, local, dev):
    """BASELINE configs[2]: full VO training step, ResNet-18 DepthNet + PoseNet (bf16 autocast, channels_last),
    fused fp32 loss, Adam, batch 32/GPU, 640x480, batch-sharded DDP (NCCL all-reduce of the network gradients)."""
    import torch.distributed as dist
    from vo.train import Trainer, synthetic_sample, DEFAULT_CONFIG
    import copy
    big = args.train_variant == "r50x4"                  # BASELINE configs[3]: ResNet-50, 1280x960, 4 sources, batch 8/GPU
    Bt = args.train_batch if args.train_batch > 0 else (8 if big else 32)
    Ht, Wt, layers, fids = (960, 1280, 50, (-1, 1, -2, 2)) if big else (H, W, 18, (-1, 1))
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["Train"].update(batch_size=Bt, img_h=Ht, img_w=Wt)
    torch.manual_seed(0)                                   # identical initial weights on every rank
    torch.backends.cudnn.benchmark = True                  # fixed shapes: let cuDNN pick its fastest convolution algorithms
    tr = Trainer(cfg, device=dev, num_layers=layers, pretrained=False, net_dtype=torch.bfloat16, distributed=world > 1,
                 noise="kernel", sync_losses=False, frame_ids=fids)
    sample = synthetic_sample(Bt, Ht, Wt, seed=100 + rank, device=dev, num_sources=len(fids))

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    for _ in range(3):
        tr.train_mono_step(dict(sample))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.train_steps):
        total, _, _ = tr.train_mono_step(dict(sample))
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item()) / args.train_steps
    loss = float(total)
    del total, _                                          # the eager step's autograd graph must be gone before a capture
    graph = None
    if world == 1:
        # the same step recorded as ONE CUDA graph (Trainer.capture_step: zero_grad -> networks -> loss -> backward -> Adam) and
        # replayed; single process only (DDP's hooks are not captured), so it is reported beside the eager number, not instead
        try:
            tr.capture_step(dict(sample))
            for _ in range(2):
                tr.train_graph_step(sample)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.train_steps):
                tr.train_graph_step(sample)
            e1.record()
            torch.cuda.synchronize()
            gdt = e0.elapsed_time(e1) / 1e3 / args.train_steps
            graph = {"ms_per_step": gdt * 1e3, "value": Bt / gdt, "unit": "triplets/s",
                     "note": "whole step replayed as one CUDA graph (N = 1 only; the headline train_step number is the eager DDP-capable step)"}
        except Exception as exc:                                  # keep the bench line if capture is not possible on this box
            graph = {"unavailable": repr(exc)[:200]}
    del tr
    torch.cuda.empty_cache()
    peak, _how = measured_peak_gbs()
    alg = bytes_alg(Bt, Ht, Wt, len(fids), NSCALE)
    return {"metric": "train_frames_per_s", "value": world * Bt / dt, "unit": "triplets/s", "ms_per_step": dt * 1e3,
            "hbm_frac_of_loss_bytes": alg / dt / 1e9 / peak,        # north_star: "as a fraction of the HBM roofline" -- the loss path's
            # algorithmic bytes over the WHOLE step time; the step is bound by the stock networks, not by these bytes
            "batch_per_gpu": Bt, "steps": args.train_steps, "warmup": 3, "final_loss": loss, "cuda_graph": graph,
            "config": f"ResNet-{layers} DepthNet+PoseNet (stock PyTorch, bf16 autocast, channels_last), fused fp32 view-synthesis "
                      f"loss with {len(fids)} source frames, Adam, {Wt}x{Ht}, batch {Bt}/GPU, DDP over NCCL "
                      f"(BASELINE configs[{3 if big else 2}]); synthetic frames resident in HBM; the only collective is DDP's "
                      f"all-reduce of the network gradients ({'248' if big else '111'} MB fp32), overlapped with the convolution backward"}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist
    from dvsloss import lib, view_synthesis_loss
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    host = make_inputs(B, rank, "cpu")
    if str(args.e2e_chunks) == "taper":
        chunks = "taper"
    else:
        chunks = [int(v) for v in str(args.e2e_chunks).split(",")]
        chunks = chunks[0] if len(chunks) == 1 else chunks
        if isinstance(chunks, list) and sum(chunks) != B:
            chunks = "taper"

    pin = lambda t: t.contiguous().pin_memory()
    h_in = dict(target=pin(host["target"]), sources=[pin(s) for s in host["sources"]],
                disps=[pin(d) for d in host["disps"]], K=pin(host["K"]), inv_K=pin(host["inv_K"]),
                Ts=[pin(T) for T in host["Ts"]])
    d_in = dict(target=h_in["target"].to(dev), sources=[s.to(dev) for s in h_in["sources"]],
                disps=[d.to(dev).requires_grad_(True) for d in h_in["disps"]], K=h_in["K"].to(dev),
                inv_K=h_in["inv_K"].to(dev), Ts=[T.to(dev).requires_grad_(True) for T in h_in["Ts"]])
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # 256 MB > 126 MB L2

    def step():
        for t in d_in["disps"] + d_in["Ts"]:
            t.grad = None
        loss, per_scale = view_synthesis_loss(d_in["disps"], d_in["target"], d_in["sources"], d_in["K"], d_in["inv_K"],
                                              d_in["Ts"], noise="kernel")
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()
    # The timed step is the public call (view_synthesis_loss + backward) recorded once as a CUDA graph and replayed: four
    # kernel launches per step either way, without the Python / autograd dispatch between them (--eager-launch times the plain calls)
    launch_mode, run_step = "eager", step
    if not args.eager_launch:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream(dev).wait_stream(side)
            for t in d_in["disps"] + d_in["Ts"]:
                t.grad = None
            g_step = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_step):
                step()
            for _ in range(2):
                g_step.replay()
            torch.cuda.synchronize()
            launch_mode, run_step = "cuda_graph", g_step.replay
        except Exception as exc:                                   # keep the bench line: fall back to plain calls
            print(f"# graph capture of the loss step failed ({exc!r}); timing eager launches", file=sys.stderr)
            torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in evs:
        flush.zero_()
        e0.record()
        run_step()
        e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_dev = sum(e0.elapsed_time(e1) for e0, e1 in evs) / 1e3
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev = float(tt.item())
    pix_step = B * NSCALE * NSRC * H * W
    value = world * pix_step * args.steps / t_dev / 1e9

    # ---- the reference's RNG contract: torch.randn([B,N,H,W]) per scale handed to the kernel (4 randn launches + 157 MB read)
    def step_torch_noise():
        for t in d_in["disps"] + d_in["Ts"]:
            t.grad = None
        loss, _ = view_synthesis_loss(d_in["disps"], d_in["target"], d_in["sources"], d_in["K"], d_in["inv_K"], d_in["Ts"],
                                      noise="torch")
        loss.backward()

    for _ in range(3):
        step_torch_noise()
    torch.cuda.synchronize()
    evn = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for e0, e1 in evn:
        flush.zero_()
        e0.record()
        step_torch_noise()
        e1.record()
    torch.cuda.synchronize()
    ms_torch_noise = sum(e0.elapsed_time(e1) for e0, e1 in evn) / len(evn)

    # ---- dominant kernel alone (events recorded by the library around its launch)
    L = lib()
    L.dvs_set_profiling(1)
    tk = []
    for _ in range(10):
        flush.zero_()
        step()
        ms = C.c_float(0)
        if L.dvs_last_tile_kernel_ms(C.byref(ms)) == 0:
            tk.append(ms.value)
    L.dvs_set_profiling(0)
    torch.cuda.synchronize()
    t_kernel = sum(tk) / len(tk) / 1e3 if tk else None

    # ---- end to end from pinned host buffers
    h_out = dict(loss=torch.empty(1 + NSCALE).pin_memory(), gd=[torch.empty_like(d).pin_memory() for d in h_in["disps"]],
                 gT=[torch.empty_like(T).pin_memory() for T in h_in["Ts"]])
    h2d = sum(t.numel() * 4 for t in [h_in["target"], h_in["K"], h_in["inv_K"]] + h_in["sources"] + h_in["disps"] + h_in["Ts"])
    d2h = sum(t.numel() * 4 for t in [h_out["loss"]] + h_out["gd"] + h_out["gT"])

    from dvsloss import HostLossPipeline
    pipe = HostLossPipeline(B, H, W, [tuple(d.shape[2:]) for d in h_in["disps"]], NSRC, chunks=chunks, device=dev,
                            noise="kernel")

    def e2e_step():
        # public host-resident call: chunked H2D -> fused loss fwd+bwd -> D2H of the loss and every gradient, overlapped
        # on three streams; returns when the results are in the pinned output buffers
        pipe.run(h_in, h_out)

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * pix_step * args.steps / float(te.item()) / 1e9
    h2d_rank_gbs = h2d * args.steps / t_e2e / 1e9                     # this rank's sustained H2D rate inside the pipeline
    # the ceiling: plain pinned -> device copies of the same buffers, all ranks at once, nothing else running
    raw_dst = [torch.empty_like(t, device=dev) for t in [h_in["target"]] + h_in["sources"]]
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        for a, b_ in zip(raw_dst, [h_in["target"]] + h_in["sources"]):
            a.copy_(b_, non_blocking=True)
    torch.cuda.synchronize()
    raw_gbs = 5 * sum(t.numel() * 4 for t in raw_dst) / (time.perf_counter() - t0) / 1e9
    del raw_dst
    raw_rates = [None] * world
    rates = [None] * world
    if world > 1:
        dist.all_gather_object(rates, round(h2d_rank_gbs, 2))
        dist.all_gather_object(raw_rates, round(raw_gbs, 2))
    else:
        rates, raw_rates = [round(h2d_rank_gbs, 2)], [round(raw_gbs, 2)]

    # ---- same, with the images crossing PCIe in dataset precision (uint8, expanded to x/255 on the device: SURVEY 8f rank 2)
    q8 = lambda t: (t * 255).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
    h_u8 = dict(h_in)
    h_u8["target"], h_u8["sources"] = q8(host["target"]), [q8(s_) for s_ in host["sources"]]
    chunks8 = "ramp" if chunks == "taper" else chunks                 # 70 MB in: kernel-bound, the short chunk goes first
    pipe8 = HostLossPipeline(B, H, W, [tuple(d.shape[2:]) for d in h_in["disps"]], NSRC, chunks=chunks8, device=dev,
                             noise="kernel", uint8_images=True)
    for _ in range(3):
        pipe8.run(h_u8, h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipe8.run(h_u8, h_out)
    barrier()
    t8 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t8, op=dist.ReduceOp.MAX)
    h2d8 = sum(t.numel() * t.element_size() for t in [h_u8["target"], h_u8["K"], h_u8["inv_K"]] + h_u8["sources"] + h_u8["disps"] + h_u8["Ts"])
    e2e_u8 = {"value": world * pix_step * args.steps / float(t8.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d8,
              "d2h_bytes_per_step": d2h, "note": "images quantised to 8 bits and sent as uint8 (the dataset's precision), expanded "
              "to float32 x/255 on the device; disparities, poses and gradients stay fp32; NOT the headline e2e"}
    del pipe8
    pipe8k = HostLossPipeline(B, H, W, [tuple(d.shape[2:]) for d in h_in["disps"]], NSRC, chunks=chunks8, device=dev,
                              noise="kernel", uint8_images=True, u8_in_kernel=True)
    for _ in range(3):
        pipe8k.run(h_u8, h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipe8k.run(h_u8, h_out)
    barrier()
    t8k = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t8k, op=dist.ReduceOp.MAX)
    e2e_u8["in_kernel"] = {"value": world * pix_step * args.steps / float(t8k.item()) / 1e9, "unit": UNIT,
                           "note": "same bytes, no expansion pass: the two-source kernel reads uint8 and forms x/255 in-register"}
    del pipe8k

    eager = None
    if rank == 0 and world == 1 and not args.no_eager:
        try:
            eager = time_eager_cuda(host, dev)
        except Exception as e:                       # pragma: no cover - reported, never fatal for the headline number
            eager = {"error": repr(e)[:200]}
        torch.cuda.empty_cache()
    train = train_big = None
    if not args.no_train:
        del pipe, h_out, flush
        torch.cuda.empty_cache()
        try:
            train = run_train(args, world, rank, local, dev)
        except Exception as e:                       # pragma: no cover
            train = {"error": repr(e)[:300]}
        train_big = None
        if args.train_variant == "r18" and not args.no_train_big:
            import copy as _copy
            a2 = _copy.copy(args)
            a2.train_variant, a2.train_batch, a2.train_steps = "r50x4", 0, 5
            try:
                train_big = run_train(a2, world, rank, local, dev)
            except Exception as e:                   # pragma: no cover
                train_big = {"error": repr(e)[:300]}

    if rank == 0:
        peak, how = measured_peak_gbs()
        ba = bytes_alg(B, H, W, NSRC, NSCALE)
        roof = None
        if t_kernel:
            ach = ba / t_kernel / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "kernel": "fused_pair_kernel<true,0> (two-source tile kernel)", "kernel_ms": t_kernel * 1e3, "bytes_alg": ba, "peak_source": how,
                    "frac_step": ba / (t_dev / args.steps) / 1e9 / peak,
                    "frac_strict_floor": bytes_floor(B, H, W, NSRC, NSCALE) / t_kernel / 1e9 / peak}
            tr = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes per launch from the committed ncu capture
            if os.path.exists(tr):
                try:
                    roof["traffic"] = json.load(open(tr)).get("fused_tile_kernel_dram_bytes")
                except Exception:
                    pass
        cpu = None
        if world == 1 and not args.no_cpu:
            B_s = args.cpu_batch
            val, dt, cref = time_cpu(host, B_s, 2, 1)
            cpu = {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": cref.kind, "s_per_step": dt,
                   "sample": f"batch {B_s} of the same workload, 1 warm-up + 2 timed fwd+bwd: {cref.what()}"}
            try:
                cpu["train_step_config0"] = time_cpu_train_step()
            except Exception as e:               # pragma: no cover - reported, never fatal
                cpu["train_step_config0"] = {"error": repr(e)[:200]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_text(B), "noise": "automask noise from the in-kernel generator",
                           "l2": "256 MB buffer written between timed steps (outside the event pairs)", "launch": launch_mode, "sharding": f"batch, {world} independent rank(s), no data-path collective"},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "h2d_gbs_per_rank": rates, "h2d_raw_copy_gbs_per_rank": raw_rates, "host_affinity": affinity,
                        "note": "PCIe / host-DRAM bound: every rank streams 203 MB per step from pinned host memory; "
                                "h2d_raw_copy_gbs_per_rank = plain pinned->device copies of the same buffers by all ranks at once "
                                "(the ceiling the pipeline's h2d_gbs_per_rank is to be read against; D2H of 26 MB/step runs concurrently)"},
                # kernels of this library inside one timed step: pre-pass, tile kernel, post-pass (forward) + gradient scaling
                "gpu_launches": 4 * args.steps, "roofline": roof, "cpu_baseline": cpu, "eager_cuda_baseline": eager,
                "train_step": train, "train_step_r50x4": train_big, "e2e_uint8_images": e2e_u8,
                "noise_torch": {"ms_per_step": ms_torch_noise, "value": pix_step / (ms_torch_noise * 1e-3) / 1e9, "unit": UNIT,
                                "note": "same step with the reference's RNG contract (torch.randn per scale handed to the kernel) instead of the in-kernel generator; this rank"}}
        if eager and "value" in eager:
            line["eager_cuda_baseline"]["speedup_device_timed"] = value / eager["value"]
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="batch per GPU (BASELINE configs[1]: 16)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="batch of the bounded CPU sample")
    ap.add_argument("--e2e-chunks", default="taper", help="batch chunks of the host-resident pipeline (e2e leg): 'taper' (16 -> "
                    "6,4,3,2,1), a count, or comma-separated chunk sizes summing to --batch")
    ap.add_argument("--eager-launch", action="store_true", help="time the device-resident step as plain Python calls instead of a "
                    "CUDA-graph replay of the same calls")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-CUDA reference leg")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (BASELINE configs[2])")
    ap.add_argument("--train-batch", type=int, default=0, help="training-step batch per GPU (default: 32, or 8 for r50x4)")
    ap.add_argument("--train-variant", default="r18", choices=["r18", "r50x4"],
                    help="r18 = BASELINE configs[2]; r50x4 = configs[3] (ResNet-50, 1280x960, sources +-1 and +-2)")
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--no-train-big", action="store_true", help="skip the BASELINE configs[3] training-step leg (r50x4)")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout at N > 1), so the
    # process's stdout is pointed at stderr for the run and the JSON line goes to the real stdout through a saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

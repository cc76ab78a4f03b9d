"""Host-resident entry point of the fused loss: inputs and results live in pinned host memory.

``HostLossPipeline`` splits the batch into equal chunks and overlaps, on three CUDA streams, the host->device copy of
chunk k+1, the fused loss forward+backward of chunk k and the device->host copy of the gradients of chunk k-1.
Chunking is exact: every reduction of the loss is a batch mean (vo/learner_new.py:244, vo/learner_func.py:174), so
the batch loss is the mean of the chunk losses and the gradients of a chunk are 1/chunks of its stand-alone
gradients (tests/test_gpu_fused.py checks this decomposition).  PCIe is the bound of this path (203 MB in per step
at the benchmark size against 1.8 ms of kernels), which is what the overlap is for.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from .functional import view_synthesis_loss
from .ops import images_u8_to_f32


class HostLossPipeline:
    def __init__(self, B: int, H: int, W: int, disp_sizes: Sequence[Sequence[int]], num_sources: int = 2, chunks: int = 4,
                 device=None, uint8_images: bool = False, u8_in_kernel: bool = False, **loss_kwargs):
        if B % chunks:
            raise ValueError("the batch must split into equal chunks (batch means must stay batch means)")
        self.B, self.Bc, self.chunks, self.N, self.S = B, B // chunks, chunks, num_sources, len(disp_sizes)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.uint8_images = uint8_images        # images arrive as bytes (dataset precision) and are expanded on the device
        # ... or, with two sources, not expanded at all: the tile kernel reads the bytes (x/255 in-register, same bits)
        self.u8_in_kernel = bool(uint8_images and u8_in_kernel and num_sources == 2)
        self.kw = dict(loss_kwargs)
        self.kw.setdefault("noise", "kernel")
        Bc, d = self.Bc, self.dev
        mk = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=d)
        self.sets: List[Dict] = []
        for _ in range(2):                                         # double-buffered device staging
            self.sets.append(dict(
                target=mk(Bc, 3, H, W), sources=[mk(Bc, 3, H, W) for _ in range(num_sources)],
                disps=[mk(Bc, 1, h, w).requires_grad_(True) for h, w in disp_sizes], K=mk(Bc, 4, 4), inv_K=mk(Bc, 4, 4),
                Ts=[mk(Bc, 4, 4).requires_grad_(True) for _ in range(num_sources)], losses=mk(1 + self.S),
                raw=[torch.empty(Bc, 3, H, W, dtype=torch.uint8, device=d) for _ in range(1 + num_sources)] if uint8_images else None))
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(d) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(chunks)]
        self.ev_run = [torch.cuda.Event() for _ in range(chunks)]
        self.ev_out = [torch.cuda.Event() for _ in range(chunks)]

    def run(self, h_in: Dict, h_out: Dict) -> None:
        """h_in: pinned ``target``, ``K``, ``inv_K`` and lists ``sources``, ``disps``, ``Ts`` for the full batch;
        h_out: pinned ``loss`` [1+S] (total, then per scale), ``gd`` (list like disps), ``gT`` (list like Ts).
        Returns after everything has landed in h_out."""
        Bc, C = self.Bc, self.chunks
        cur = torch.cuda.current_stream(self.dev)
        for st in (self.s_in, self.s_run, self.s_out):
            st.wait_stream(cur)
        parts = []
        for c in range(C):
            S_ = self.sets[c & 1]
            sl = slice(c * Bc, (c + 1) * Bc)
            with torch.cuda.stream(self.s_in), torch.no_grad():
                if c >= 2:
                    self.s_in.wait_event(self.ev_out[c - 2])          # the set is free once its gradients left the device
                if self.uint8_images:
                    for a, b in zip(S_["raw"], [h_in["target"]] + h_in["sources"]):
                        a.copy_(b[sl], non_blocking=True)
                    if not self.u8_in_kernel:
                        for raw, img in zip(S_["raw"], [S_["target"]] + S_["sources"]):
                            images_u8_to_f32(raw, img)             # x / 255 on the copy stream, exact
                else:
                    S_["target"].copy_(h_in["target"][sl], non_blocking=True)
                    for a, b in zip(S_["sources"], h_in["sources"]):
                        a.copy_(b[sl], non_blocking=True)
                S_["K"].copy_(h_in["K"][sl], non_blocking=True)
                S_["inv_K"].copy_(h_in["inv_K"][sl], non_blocking=True)
                for a, b in zip(S_["disps"] + S_["Ts"], h_in["disps"] + h_in["Ts"]):
                    a.copy_(b[sl], non_blocking=True)
                self.ev_in[c].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[c])
                for t in S_["disps"] + S_["Ts"]:
                    t.grad = None
                tgt, srcs = (S_["raw"][0], S_["raw"][1:]) if self.u8_in_kernel else (S_["target"], S_["sources"])
                loss, per_scale = view_synthesis_loss(S_["disps"], tgt, srcs, S_["K"], S_["inv_K"], S_["Ts"], **self.kw)
                loss.backward(torch.full_like(loss, 1.0 / C))
                part = torch.cat([loss.detach().view(1), per_scale.detach()])
                parts.append(part)
                self.ev_run[c].record(self.s_run)
            with torch.cuda.stream(self.s_out), torch.no_grad():
                self.s_out.wait_event(self.ev_run[c])
                for a, t in zip(h_out["gd"] + h_out["gT"], S_["disps"] + S_["Ts"]):
                    a[sl].copy_(t.grad, non_blocking=True)
                self.ev_out[c].record(self.s_out)
        with torch.cuda.stream(self.s_out), torch.no_grad():
            self.s_out.wait_stream(self.s_run)
            h_out["loss"].copy_(torch.stack(parts).mean(0), non_blocking=True)
        self.s_out.synchronize()
        cur.wait_stream(self.s_out)

"""Host-resident entry point of the fused loss: inputs and results live in pinned host memory.

``HostLossPipeline`` splits the batch into chunks (tapered by default: 16 -> 6, 4, 3, 2, 1) and overlaps, on three CUDA streams, the host->device copies
(all chunks are enqueued up-front into their own staging buffers -- a full batch is 0.2 GB of 180 -- so the copy engine
never waits for the host), the fused loss forward+backward of chunk k as soon as its inputs have landed, and the
device->host copy of the gradients of chunk k-1.
Chunking is exact: every reduction of the loss is a batch mean (vo/learner_new.py:244, vo/learner_func.py:174), so
the batch loss is the B_c/B-weighted sum of the chunk losses and the gradients of a chunk are B_c/B of its stand-alone
gradients (tests/test_gpu_fused.py checks this decomposition).  PCIe is the bound of this path (203 MB in per step
at the benchmark size against 1.6 ms of kernels), which is what the overlap is for.  The whole step is recorded as one CUDA graph
the first time a set of pinned buffers is seen and replayed afterwards.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from .functional import view_synthesis_loss
from .ops import images_u8_to_f32


def chunk_sizes(B: int, chunks) -> List[int]:
    """Chunk sizes of a batch of B: ``chunks`` is a count (equal chunks), a sequence of sizes summing to B, "taper" (each chunk
    ~3/8 of what is left: 16 -> 6, 4, 3, 2, 1 -- copy-bound steps want the short chunk LAST) or "ramp" (the taper reversed --
    kernel-bound steps want the short chunk FIRST)."""
    if isinstance(chunks, str):
        if chunks not in ("taper", "ramp"):
            raise ValueError("chunks: a count, a sequence of sizes, 'taper' or 'ramp'")
        sizes, rem = [], B
        while rem:
            sizes.append(-(-rem * 3 // 8))
            rem -= sizes[-1]
        return sizes if chunks == "taper" else sizes[::-1]
    if isinstance(chunks, int):
        if chunks < 1 or B % chunks:
            raise ValueError("the batch must split into equal chunks (or give the chunk sizes)")
        return [B // chunks] * chunks
    sizes = [int(v) for v in chunks]
    if sum(sizes) != B or min(sizes) < 1:
        raise ValueError("chunk sizes must be positive and sum to the batch size")
    return sizes


class HostLossPipeline:
    def __init__(self, B: int, H: int, W: int, disp_sizes: Sequence[Sequence[int]], num_sources: int = 2, chunks="taper",
                 device=None, uint8_images: bool = False, u8_in_kernel: bool = False, graph: bool = True, **loss_kwargs):
        """``chunks``: a count (equal chunks), a sequence of chunk sizes summing to B, "taper" or "ramp".  Chunk losses and gradients are
        combined with the weights B_c / B, so unequal chunks are exact too; tapering the sizes (16 -> 6,4,3,2,1) shortens the part
        of the step that cannot overlap the copy-in -- the kernels and the copy-out of the LAST chunk -- while keeping the number
        of copies low (measured: profiles/r02_e2e_chunks.md).
        ``graph``: record the whole step -- every copy, kernel and cross-stream dependency -- as one CUDA graph the first time a
        given set of pinned buffers is seen and replay it afterwards: the step is then bound by the copy engines alone, not by
        ~150 host-side launches (the in-kernel noise counter is a device scalar, so replays keep drawing fresh noise)."""
        sizes = chunk_sizes(B, chunks)
        self.sizes = sizes
        self.starts = [sum(sizes[:i]) for i in range(len(sizes))]
        chunks = len(sizes)
        self.B, self.chunks, self.N, self.S = B, chunks, num_sources, len(disp_sizes)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.uint8_images = uint8_images        # images arrive as bytes (dataset precision) and are expanded on the device
        # ... or, with two sources, not expanded at all: the tile kernel reads the bytes (x/255 in-register, same bits)
        self.u8_in_kernel = bool(uint8_images and u8_in_kernel and num_sources == 2)
        self.kw = dict(loss_kwargs)
        self.kw.setdefault("noise", "kernel")
        d = self.dev
        mk = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=d)
        # the small operands (intrinsics, poses, every disparity map below a quarter of the image) cross in ONE copy each for the whole
        # batch ahead of the first chunk; the chunks' tensors are slices of these buffers.  A copy has a fixed cost of a few
        # microseconds on the copy engine whatever its size: 7 copies per chunk instead of 11+.
        self.full = dict(K=mk(B, 4, 4), inv_K=mk(B, 4, 4), Ts=[mk(B, 4, 4) for _ in range(num_sources)],
                         disps=[mk(B, 1, h, w) if 4 * h * w <= H * W else None for h, w in disp_sizes])
        self.sets: List[Dict] = []
        for st0, Bc in zip(self.starts, sizes):                    # one staging set per chunk: H2D never waits for a free set
            sl = slice(st0, st0 + Bc)
            leaf = lambda t: t[sl].detach().requires_grad_(True)
            self.sets.append(dict(
                target=mk(Bc, 3, H, W), sources=[mk(Bc, 3, H, W) for _ in range(num_sources)],
                disps=[(mk(Bc, 1, h, w).requires_grad_(True) if f is None else leaf(f))
                       for (h, w), f in zip(disp_sizes, self.full["disps"])],
                K=self.full["K"][sl], inv_K=self.full["inv_K"][sl], Ts=[leaf(t) for t in self.full["Ts"]], losses=mk(1 + self.S),
                raw=[torch.empty(Bc, 3, H, W, dtype=torch.uint8, device=d) for _ in range(1 + num_sources)] if uint8_images else None))
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(d) for _ in range(3))
        self.graph_enabled = bool(graph)
        self.max_graphs = 8                         # distinct sets of pinned buffers that get their own recorded graph
        self._graphs: Dict = {}
        self.ev_in = [torch.cuda.Event() for _ in range(chunks)]
        self.ev_run = [torch.cuda.Event() for _ in range(chunks)]
        self.ev_out = [torch.cuda.Event() for _ in range(chunks)]

    def run(self, h_in: Dict, h_out: Dict) -> None:
        """h_in: pinned ``target``, ``K``, ``inv_K`` and lists ``sources``, ``disps``, ``Ts`` for the full batch;
        h_out: pinned ``loss`` [1+S] (total, then per scale), ``gd`` (list like disps), ``gT`` (list like Ts).
        Returns after everything has landed in h_out."""
        if not self.graph_enabled:
            self._enqueue(h_in, h_out)
            self.s_out.synchronize()
            return
        flat = [h_in["target"], h_in["K"], h_in["inv_K"]] + h_in["sources"] + h_in["disps"] + h_in["Ts"] + \
               [h_out["loss"]] + h_out["gd"] + h_out["gT"]
        key = tuple(t.data_ptr() for t in flat)
        g = self._graphs.get(key)
        if g is None and len(self._graphs) >= self.max_graphs:
            # a caller that hands over fresh pinned buffers every step would otherwise record (and keep alive) a graph per step
            self._enqueue(h_in, h_out)
            self.s_out.synchronize()
            return
        if g is None:
            # two eager steps first (allocator warm-up, lazy module loading), then capture on a side stream
            for _ in range(2):
                self._enqueue(h_in, h_out)
                self.s_out.synchronize()
            cap = torch.cuda.Stream(self.dev)
            cap.wait_stream(torch.cuda.current_stream(self.dev))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(cap):
                with torch.cuda.graph(g, stream=cap):
                    self._enqueue(h_in, h_out)
            torch.cuda.current_stream(self.dev).wait_stream(cap)
            self._graphs[key] = g
            self._keep = getattr(self, "_keep", []) + [flat]          # the graph holds raw pointers into these buffers
        g.replay()
        torch.cuda.current_stream(self.dev).synchronize()

    def _enqueue(self, h_in: Dict, h_out: Dict) -> None:
        C = self.chunks
        cur = torch.cuda.current_stream(self.dev)
        for st in (self.s_in, self.s_run, self.s_out):
            st.wait_stream(cur)
        parts = []
        with torch.cuda.stream(self.s_in), torch.no_grad():
            F = self.full
            F["K"].copy_(h_in["K"], non_blocking=True)
            F["inv_K"].copy_(h_in["inv_K"], non_blocking=True)
            for a, b in zip(F["Ts"] + F["disps"], h_in["Ts"] + h_in["disps"]):
                if a is not None:
                    a.copy_(b, non_blocking=True)
            for c in range(C):
                S_ = self.sets[c]
                sl = slice(self.starts[c], self.starts[c] + self.sizes[c])
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
                for a, b, f in zip(S_["disps"], h_in["disps"], F["disps"]):
                    if f is None:                                  # the large disparity maps travel with their chunk
                        a.copy_(b[sl], non_blocking=True)
                self.ev_in[c].record(self.s_in)
        for c in range(C):
            S_ = self.sets[c]
            sl = slice(self.starts[c], self.starts[c] + self.sizes[c])
            wgt = self.sizes[c] / self.B
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[c])
                for t in S_["disps"] + S_["Ts"]:
                    t.grad = None
                tgt, srcs = (S_["raw"][0], S_["raw"][1:]) if self.u8_in_kernel else (S_["target"], S_["sources"])
                loss, per_scale = view_synthesis_loss(S_["disps"], tgt, srcs, S_["K"], S_["inv_K"], S_["Ts"], **self.kw)
                loss.backward(torch.full_like(loss, wgt))
                part = torch.cat([loss.detach().view(1), per_scale.detach()]) * wgt
                parts.append(part)
                self.ev_run[c].record(self.s_run)
            with torch.cuda.stream(self.s_out), torch.no_grad():
                self.s_out.wait_event(self.ev_run[c])
                for a, t in zip(h_out["gd"] + h_out["gT"], S_["disps"] + S_["Ts"]):
                    a[sl].copy_(t.grad, non_blocking=True)
                self.ev_out[c].record(self.s_out)
        with torch.cuda.stream(self.s_out), torch.no_grad():
            self.s_out.wait_stream(self.s_run)
            h_out["loss"].copy_(torch.stack(parts).sum(0), non_blocking=True)
        cur.wait_stream(self.s_out)

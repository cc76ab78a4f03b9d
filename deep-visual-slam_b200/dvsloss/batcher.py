"""GPU-side triplet batcher: what ``MonoDataset.__getitem__`` + the DataLoader's collation produce
(vo/dataset/common.py:48-92), built on the device from a resident uint8 frame sequence.

The reference decodes and resizes three frames per sample in 24 worker processes, converts them with ``ToTensor`` on
the host and ships fp32 tensors (11 MB per 640x480 triplet) through pinned memory every step.  Once the loss is fused
the step is bounded by the networks and then by exactly that host path.  Here the decoded, resized
frames of a sequence stay on the device as bytes (3.5 GB per 4 000 frames of 640x480); a batch is one gather kernel
(``dvs_gather_triplets_u8``: frame selection, HWC -> CHW, optional exact x/255) plus a handful of 4x4 matrix ops for the
intrinsics pyramid.  Decoding / resizing (``_read_image``) stays wherever the frames come from.

Same sample contract as the reference: ``("K", s)``, ``("inv_K", s)`` [B,4,4] fp32 for s = 0..3 (rows 0/1 of K scaled by
(W // 2**s) / W, (H // 2**s) / H; ``inv_K = pinv(K)`` computed in float64 like numpy does, then cast), and
``("source_left", 0)``, ``("target_image", 0)``, ``("source_right", 0)`` [B,3,H,W] -- fp32 in [0,1], or uint8 when
``out="uint8"`` (the two-source loss kernel reads bytes directly).  Frame spacing follows the reference: the target is
``idx + size_1`` and the right source ``idx + size_1 + size_2`` with ``size_* ~ U{1..max_size}`` (3 for training, 1 for
validation); ColorJitter(0.3, 0.3, 0.3, 0.2) with probability 0.5 per sample, the same parameters for the three frames.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from ._lib import DTYPE_F32, DTYPE_U8, DvsError, check, lib, stream_ptr


class GpuTripletBatcher:
    def __init__(self, frames: torch.Tensor, intrinsics: torch.Tensor, num_scale: int = 4, is_train: bool = True,
                 augment: bool = False, out: str = "float32", generator: Optional[torch.Generator] = None):
        """frames: uint8 CUDA tensor [T,H,W,3] (decoder order) or [T,3,H,W]; intrinsics: [T,4,4] or [4,4] absolute-pixel K."""
        if not frames.is_cuda or frames.dtype != torch.uint8 or frames.dim() != 4:
            raise DvsError("frames must be a CUDA uint8 tensor [T,H,W,3] or [T,3,H,W] (no CPU fallback by design)")
        self.hwc = frames.shape[-1] == 3 and frames.shape[1] != 3
        self.frames = frames.contiguous()
        T = frames.shape[0]
        self.H, self.W = (frames.shape[1], frames.shape[2]) if self.hwc else (frames.shape[2], frames.shape[3])
        K = intrinsics.to(frames.device, torch.float64)
        self.K = K.expand(T, 4, 4).contiguous() if K.dim() == 2 else K
        self.num_scale, self.max_size = num_scale, (3 if is_train else 1)
        self.augment = augment and is_train
        if out not in ("float32", "uint8"):
            raise DvsError("out must be 'float32' or 'uint8'")
        if out == "uint8" and self.augment:
            raise DvsError("ColorJitter works on float images: use out='float32' with augment=True")
        self.out = out
        self.gen = generator
        self._jitter = None

    def __len__(self) -> int:
        return self.frames.shape[0] - 2 * self.max_size          # vo/dataset/common.py:45-46

    def draw_indices(self, batch_size: int) -> torch.Tensor:
        """[B,3] frame numbers (left, target, right) on the host, drawn like __getitem__ under a shuffling sampler."""
        n = len(self)
        if n < 1:
            raise DvsError("sequence too short for the frame spacing")
        r = lambda hi: torch.randint(0, hi, (batch_size,), generator=self.gen)
        idx, s1, s2 = r(n), 1 + r(self.max_size), 1 + r(self.max_size)
        return torch.stack([idx, idx + s1, idx + s1 + s2], 1).to(torch.int32)

    def intrinsics_pyramid(self, left_idx: torch.Tensor) -> Dict:
        """("K", s), ("inv_K", s) for the samples whose left frame numbers are given (the reference indexes the intrinsics
        with the dataset index, i.e. the left frame: vo/dataset/common.py:69)."""
        out = {}
        K0 = self.K[left_idx.to(self.K.device).long()]
        for s in range(self.num_scale):
            K = K0.clone()
            K[:, 0, :] *= (self.W // (2 ** s)) / self.W
            K[:, 1, :] *= (self.H // (2 ** s)) / self.H
            out[("K", s)] = K.float()
            out[("inv_K", s)] = torch.linalg.pinv(K).float()
        return out

    def batch(self, idx: torch.Tensor) -> Dict:
        """Sample dict for explicit frame triplets idx [B,3]."""
        dev = self.frames.device
        B = idx.shape[0]
        if idx.min() < 0 or idx.max() >= self.frames.shape[0]:
            raise DvsError("frame index out of range")
        idx_d = idx.to(dev, torch.int32, non_blocking=True).contiguous()
        dt = torch.float32 if self.out == "float32" else torch.uint8
        imgs = [torch.empty(B, 3, self.H, self.W, dtype=dt, device=dev) for _ in range(3)]
        with torch.cuda.device(dev):
            rc = lib().dvs_gather_triplets_u8(self.frames.data_ptr(), int(self.hwc), idx_d.data_ptr(), imgs[0].data_ptr(),
                                              imgs[1].data_ptr(), imgs[2].data_ptr(),
                                              DTYPE_F32 if self.out == "float32" else DTYPE_U8, B, self.H, self.W, stream_ptr(dev))
        check(rc, "dvs_gather_triplets_u8")
        if self.augment:
            imgs = self._color_jitter(imgs)
        sample = self.intrinsics_pyramid(idx[:, 0])
        sample[("source_left", 0)], sample[("target_image", 0)], sample[("source_right", 0)] = imgs
        return sample

    def sample(self, batch_size: int) -> Dict:
        return self.batch(self.draw_indices(batch_size))

    def _color_jitter(self, imgs):
        """transforms.ColorJitter on the stacked three frames of a sample with probability 0.5 (vo/dataset/common.py:79-81)."""
        if self._jitter is None:
            from torchvision import transforms
            self._jitter = transforms.ColorJitter(brightness=0.3, contrast=0.3, saturation=0.3, hue=0.2)
        stack = torch.stack(imgs, 1)                                  # [B,3 frames,3,H,W]
        flip = torch.rand(stack.shape[0], generator=self.gen) < 0.5
        for b in torch.nonzero(flip).flatten().tolist():
            stack[b] = self._jitter(stack[b])
        return [stack[:, i].contiguous() for i in range(3)]

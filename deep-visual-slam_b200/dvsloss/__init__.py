"""dvsloss -- B200 (sm_100a) view-synthesis loss for Deep-Visual-SLAM-style VO training.

``view_synthesis_loss`` is the fused hot path; ``ops`` holds the granular primitives
(disp_to_depth, BackprojectDepth, Project3D, SSIM, ...) with the reference's call surface.
Everything runs through ``libdvsloss.so`` (C ABI in ``include/dvsloss.h``); there is no CPU or
PyTorch fallback -- importing works without a GPU, calling an operator does not.
"""
from ._lib import DvsError, LIB_PATH, exported_symbols, lib  # noqa: F401
from .functional import noise_state, set_noise_state, view_synthesis_loss  # noqa: F401
from .host import HostLossPipeline  # noqa: F401
from .batcher import GpuTripletBatcher  # noqa: F401

__all__ = ["view_synthesis_loss", "noise_state", "set_noise_state", "HostLossPipeline", "GpuTripletBatcher", "DvsError", "lib", "LIB_PATH", "exported_symbols"]

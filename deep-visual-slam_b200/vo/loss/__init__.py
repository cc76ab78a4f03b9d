"""vo.loss -- the view-synthesis loss of the VO learner on B200.

``view_synthesis_loss``        the fused op (all scales, all sources, fwd + gradients in one pass)
``compute_reprojection_loss``  SSIM + L1 photometric error (reference: vo/learner_new.py:60-74; the dead
                               TensorFlow twin lives at vo/loss/warp_loss.py:5-13 in the reference)
``get_smooth_loss``, ``SSIM``  as in vo/learner_func.py
"""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from dvsloss import view_synthesis_loss  # noqa: E402,F401
from dvsloss.ops import compute_reprojection_loss, get_smooth_loss  # noqa: E402,F401
from model.layers import SSIM  # noqa: E402,F401

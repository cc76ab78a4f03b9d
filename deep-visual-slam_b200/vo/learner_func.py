"""The VO learner's primitives (reference: vo/learner_func.py:16-207).  In the reference this file is a
near-verbatim twin of model/layers.py; here it re-exports the one implementation."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from model.layers import (BackprojectDepth, Project3D, SSIM, disp_to_depth, get_smooth_loss,  # noqa: E402,F401
                          get_translation_matrix, rot_from_axisangle, transformation_from_parameters)

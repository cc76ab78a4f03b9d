"""``model`` package of the B200 view-synthesis path.

This directory shadows the reference's ``model/`` when it is put first on ``sys.path`` (option A).
Only the files of the accelerated path live here (layers, depthnet, posenet_single, resnet_encoder); every other
sub-module of the reference (``model.raft``, ``model.depth_anything_v2``, ``model.posenet``) keeps resolving to the
reference tree, wherever it sits later on ``sys.path``, through ``pkgutil.extend_path``.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

"""``depth`` package of the B200 path: the supervised-depth learner (reference: depth/depth_learner.py).  Other modules of
the reference's ``depth/`` package keep resolving to the reference tree later on ``sys.path``."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

"""``vo`` package of the B200 view-synthesis path.

Shadows the reference's ``vo/`` when put first on ``sys.path`` (option A): ``vo.learner_new``,
``vo.learner_func`` and ``vo.loss`` come from here; ``vo.dataset``, ``vo.utils``, ``vo.eval_traj`` ... keep resolving
to the reference tree later on ``sys.path`` (``pkgutil.extend_path``), so ``vo/train.py:13-19`` imports unchanged.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

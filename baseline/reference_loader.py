"""Import the staged reference (``baseline/_ref/``, see stage_reference.py) next to this repository's same-named packages.

Only ``bench.py``'s baseline legs use this.  The reference's modules are loaded under private names so that they do not
collide with ``deep-visual-slam_b200/{vo,model}``: ``learner_new`` / ``learner_func`` are bare modules in the reference
(vo/learner_new.py:6 does ``from learner_func import ...``), the networks are loaded file by file."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "vo", "learner_new.py"))


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_learner():
    """The reference's ``learner_new`` module (``MonodepthTrainer``)."""
    if "learner_func" not in sys.modules:
        _load("learner_func", os.path.join(REF, "vo", "learner_func.py"))
    return _load("_ref_learner_new", os.path.join(REF, "vo", "learner_new.py"))


def load_nets():
    """(DepthNet, PoseNet) classes of the reference."""
    mdir = os.path.join(REF, "model")
    # depthnet.py / posenet_single.py fall back to bare imports (``from layers import *``, ``from raft.core.raft import ...``)
    for name, rel in (("layers", "layers.py"), ("resnet_encoder", "resnet_encoder.py")):
        if name not in sys.modules:
            _load(name, os.path.join(mdir, rel))
    if "raft" not in sys.modules:
        import types
        for n in ("raft", "raft.core"):
            sys.modules[n] = types.ModuleType(n)
        _load("raft.core.raft", os.path.join(mdir, "raft", "core", "raft.py"))
    d = _load("_ref_depthnet", os.path.join(mdir, "depthnet.py"))
    p = _load("_ref_posenet_single", os.path.join(mdir, "posenet_single.py"))
    return d.DepthNet, p.PoseNet

"""Write tests/golden/ref_state_dict_keys.json from the UNMODIFIED reference networks (build container only).

    python tests/golden/make_state_keys.py

The reference's ``Trainer`` loads ``weights/vo/depth_net_epoch_30.pth`` / ``pose_net_epoch_30.pth`` (vo/train.py:83-98) and
writes checkpoints that ``vo/predict.py`` / ``vo/eval_traj.py`` read, so the state_dict key names and shapes of DepthNet and
PoseNet are part of the drop-in contract.  This records them (ResNet-18 variants) for the CPU test that has no reference tree.
"""
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference_nets():
    """DepthNet / PoseNet classes of the reference, imported from where they lie (``model.raft`` is stubbed: SmallRAFT is
    only touched by FlowPoseNet, which is never constructed)."""
    sys.path.insert(0, os.path.join(REF, "model"))
    for name in ("raft", "raft.core", "raft.core.raft"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["raft.core.raft"].SmallRAFT = object

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    d = load("_ref_depthnet", os.path.join(REF, "model", "depthnet.py"))
    p = load("_ref_posenet_single", os.path.join(REF, "model", "posenet_single.py"))
    return d.DepthNet, p.PoseNet


if __name__ == "__main__":
    DepthNet, PoseNet = load_reference_nets()
    out = {"depth_net": {k: list(v.shape) for k, v in DepthNet(18, False).state_dict().items()},
           "pose_net": {k: list(v.shape) for k, v in PoseNet(18, False).state_dict().items()}}
    with open(os.path.join(HERE, "ref_state_dict_keys.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print({k: len(v) for k, v in out.items()})

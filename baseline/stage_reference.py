"""Stage the UNMODIFIED reference files of the hot path into the git-ignored ``baseline/_ref/`` (build container only).

    python baseline/stage_reference.py

``/root/reference`` does not exist on the GPU box; ``baseline/_ref/`` travels there with the gpurun snapshot (it is
git-ignored, not gpurun-ignored), so ``bench.py`` can time the reference's own code -- ``MonodepthTrainer`` of
vo/learner_new.py with vo/learner_func.py, and the stock DepthNet / PoseNet of model/ -- as the eager-CUDA and CPU baselines
(BASELINE.md 3.1).  Nothing under ``_ref/`` is ever committed or imported by the product package.  The
reference is not a pip package (no setup.py / pyproject.toml), so "installing" it is this copy.  ``model/raft`` (an
un-needed optical-flow sub-package that posenet_single.py imports at module level) is replaced by a two-line stub.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["vo/learner_func.py", "vo/learner_new.py", "model/__init__.py", "model/layers.py", "model/depthnet.py",
         "model/resnet_encoder.py", "model/posenet_single.py"]


def stage() -> bool:
    if not os.path.isdir(REF):
        return False
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    stub = os.path.join(DST, "model", "raft", "core")
    os.makedirs(stub, exist_ok=True)
    for d in (os.path.join(DST, "model", "raft"), stub):
        open(os.path.join(d, "__init__.py"), "w").close()
    with open(os.path.join(stub, "raft.py"), "w") as f:
        f.write("# stub written by baseline/stage_reference.py: RAFT is only touched by FlowPoseNet, which is never constructed\n"
                "class SmallRAFT:\n    pass\n")
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference tree not present", DST)
    sys.exit(0)

"""Drop-in contract of the Python surface (CPU): checkpoint format of the networks, lazily materialised ``outputs``,
and import resolution when this tree shadows the reference's ``vo`` / ``model`` packages (INTEGRATION.md option A).

The fixture ``tests/golden/ref_state_dict_keys.json`` is written from the unmodified reference classes by
``tests/golden/make_state_keys.py``; where ``/root/reference`` exists (the build container) the round trip is also run
against the live classes."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deep-visual-slam_b200")
sys.path.insert(0, PKG)

from model.depthnet import DepthNet  # noqa: E402
from model.posenet_single import FlowPoseNet, PoseNet  # noqa: E402
from vo.learner_new import LazyOutputs  # noqa: E402

REF = "/root/reference"
have_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")


def test_state_dict_keys_match_reference_fixture():
    """vo/train.py:83-98 loads depth_net_epoch_N.pth / pose_net_epoch_N.pth by key: names and shapes must be the reference's."""
    with open(os.path.join(ROOT, "tests", "golden", "ref_state_dict_keys.json")) as f:
        want = json.load(f)
    for name, net in (("depth_net", DepthNet(18, False)), ("pose_net", PoseNet(18, False))):
        got = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert got == want[name], (name, sorted(set(got) ^ set(want[name]))[:8])


@have_ref
def test_checkpoint_round_trip_with_live_reference_classes():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_state_keys import load_reference_nets
    RefDepth, RefPose = load_reference_nets()
    torch.manual_seed(3)
    rd, rp = RefDepth(18, False).eval(), RefPose(18, False).eval()
    d, p = DepthNet(18, False).eval(), PoseNet(18, False).eval()
    d.load_state_dict(rd.state_dict())                      # strict: raises on any missing / unexpected key
    p.load_state_dict(rp.state_dict())
    x, pair = torch.rand(1, 3, 64, 96), torch.rand(1, 6, 64, 96)
    from model.layers import Conv3x3
    with torch.no_grad():
        a, b = d(x), rd(x)
        assert all(torch.allclose(a[k], b[k], rtol=1e-5, atol=1e-6) for k in b)       # border strips: another summation order
        Conv3x3.fast_reflect = False
        try:
            a = d(x)
        finally:
            Conv3x3.fast_reflect = True
        assert all(torch.equal(a[k], b[k]) for k in b)                                  # literal sequence: bit-identical
        assert all(torch.equal(u, v) for u, v in zip(p(pair), rp(pair)))
    rd.load_state_dict(d.state_dict())                      # and back: checkpoints written here load in the reference
    rp.load_state_dict(p.state_dict())


def test_flow_posenet_is_importable_but_out_of_scope():
    with pytest.raises(NotImplementedError):
        FlowPoseNet()


def test_lazy_outputs_materialise_on_first_access_only():
    calls = []

    def fill(out):
        calls.append(1)
        for s in range(4):
            dict.__setitem__(out, ("depth", s), torch.full((1,), float(s)))
            dict.__setitem__(out, ("color", -1, s), torch.zeros(1))

    out = LazyOutputs({("disp", 0): torch.ones(1)}).bind(fill, ("depth", "color"))
    out.update({("cam_T_cam", 0, 1): torch.eye(4)})
    assert ("disp", 0) in out and "loss" not in out and not calls          # ordinary keys never trigger the filler
    with pytest.raises(KeyError):
        out[("disp", 7)]
    assert not calls
    assert float(out[("depth", 2)]) == 2.0 and calls == [1]                 # what plot_utils.py:40 does
    assert ("color", -1, 3) in out and out.get(("color", 1, 0)) is None and calls == [1]
    with pytest.raises(KeyError):
        out[("color", 1, 0)]                                                # filled once; a key the filler did not make stays missing
    assert isinstance(out, dict) and len(out) == 2 + 8


@have_ref
def test_shadowing_packages_keep_reference_submodules_importable():
    """With this tree first and the reference root later on sys.path (how vo/train.py:4 arranges it), ``vo.learner_new`` and
    ``model.layers`` are ours while ``vo.dataset`` / ``vo.utils`` / ``model.raft`` still resolve to the reference."""
    code = (
        "import sys; sys.path[:0] = [%r, %r]\n"
        "import vo.learner_new, model.layers, vo.dataset, vo.utils, model.raft, importlib.util as u\n"
        "from model.posenet_single import PoseNet, FlowPoseNet\n"
        "print(vo.learner_new.__file__); print(model.layers.__file__); print(list(vo.dataset.__path__)[0]);"
        "print(list(model.raft.__path__)[0]); print(u.find_spec('vo.dataset.common').origin)\n") % (PKG, REF)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()[-5:]
    assert lines[0].startswith(PKG) and lines[1].startswith(PKG), lines
    assert lines[2].startswith(REF) and lines[3].startswith(REF) and lines[4].startswith(REF), lines


def test_conv3x3_without_padded_copy_equals_reflection_pad_conv():
    """model/layers.py: Conv3x3 computes conv(ReflectionPad2d(1)(x)) as a zero-padded convolution + four border strips; values
    and every gradient must equal the literal sequence of the reference (model/layers.py:120-136)."""
    from model.layers import Conv3x3
    torch.manual_seed(0)
    m = Conv3x3(5, 4).double()
    x = torch.randn(2, 5, 9, 11, dtype=torch.float64, requires_grad=True)
    g = torch.randn(2, 4, 9, 11, dtype=torch.float64)
    outs = []
    for fast in (True, False):
        Conv3x3.fast_reflect = fast
        try:
            y = m(x)
            outs.append((y, torch.autograd.grad(y, (x, m.conv.weight, m.conv.bias), g)))
        finally:
            Conv3x3.fast_reflect = True
    assert torch.allclose(outs[0][0], outs[1][0], rtol=0, atol=1e-12)
    for a, b in zip(outs[0][1], outs[1][1]):
        assert torch.allclose(a, b, rtol=0, atol=1e-12)
    assert tuple(Conv3x3(3, 2)(torch.rand(1, 3, 2, 2)).shape) == (1, 2, 2, 2)        # too small for the strips: literal path

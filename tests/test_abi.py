"""CPU tests of the drop-in boundary: libdvsloss.so loads, exports every symbol include/dvsloss.h declares,
and its argument checks answer without touching a GPU."""
import ctypes as C
import os
import re

import pytest

import dvsloss
from dvsloss import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dvsloss.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dvs_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(dvsloss.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return dvsloss.lib()


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(dvsloss.exported_symbols())


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_version_and_error_strings(lib):
    assert lib.dvs_version() >= 100
    assert b"ok" in lib.dvs_error_string(0)
    assert b"invalid" in lib.dvs_error_string(-1)
    assert b"CUDA" in lib.dvs_error_string(-2)
    assert b"workspace" in lib.dvs_error_string(-3)


def test_workspace_query_and_shape_validation(lib):
    n = C.c_size_t(0)
    sh = _lib.make_shape(16, 480, 640, 2, [(480, 640), (240, 320), (120, 160), (60, 80)])
    assert lib.dvs_loss_workspace_bytes(C.byref(sh), C.byref(n)) == 0
    assert 0 < n.value < 64 << 20
    bad = _lib.make_shape(16, 480, 640, 2, [(480, 640)])
    bad.N = 9
    assert lib.dvs_loss_workspace_bytes(C.byref(bad), C.byref(n)) == -1
    bad = _lib.make_shape(1, 1, 640, 2, [(1, 640)])
    assert lib.dvs_loss_workspace_bytes(C.byref(bad), C.byref(n)) == -1
    bad = _lib.make_shape(1, 48, 64, 2, [(96, 64)])               # disparity larger than the image
    assert lib.dvs_loss_workspace_bytes(C.byref(bad), C.byref(n)) == -1
    assert lib.dvs_loss_workspace_bytes(None, C.byref(n)) == -1


def test_null_pointer_arguments_are_rejected_before_any_launch(lib):
    sh = _lib.make_shape(1, 48, 64, 2, [(48, 64)])
    pr = _lib.DvsParams(0.1, 10.0, 0.85, 1e-3, 1e-7, 1)
    rc = lib.dvs_photometric_forward(C.byref(sh), C.byref(pr), None, None, None, None, None, None, None, 0, 0, None, None,
                                     None, None, None, None, None)
    assert rc == -1
    assert lib.dvs_disp_to_depth_fwd(None, None, None, 10, 0.1, 10.0, None) == -1
    assert lib.dvs_ssim_fwd(None, None, None, 1, 3, 8, 8, None) == -1
    assert lib.dvs_pose_matrix_fwd(None, None, None, 1, 0, None) == -1
    n = C.c_size_t(0)
    assert lib.dvs_smooth_loss_workspace_bytes(2, 48, 64, C.byref(n)) == 0 and n.value > 0
    assert lib.dvs_project3d_bwd_workspace_bytes(2, 48, 64, C.byref(n)) == 0 and n.value > 0


def test_python_front_end_refuses_cpu_tensors():
    import torch
    from dvsloss.synthetic import make_problem
    p = make_problem(1, 32, 32, 2, 4, seed=1)
    with pytest.raises(dvsloss.DvsError):
        dvsloss.view_synthesis_loss(p["disps"], p["target"], p["sources"], p["K"], p["inv_K"],
                                    [torch.eye(4)[None]] * 2, noise=None)


def test_extended_entry_points_validate_dtypes(lib):
    sh = _lib.make_shape(1, 48, 64, 2, [(48, 64)])
    pr = _lib.DvsParams(0.1, 10.0, 0.85, 1e-3, 1e-7, 1)
    args = (None, None, None, None, 0, 0, None, None, None, None, None, None, None)
    assert lib.dvs_photometric_forward_ex(C.byref(sh), C.byref(pr), None, 7, None, None, 0, *args) == -1     # unknown dtype
    assert lib.dvs_photometric_forward_ex(C.byref(sh), C.byref(pr), None, _lib.DTYPE_U8, None, None, 0, *args) == -1
    assert lib.dvs_photometric_forward_ex(C.byref(sh), C.byref(pr), None, 0, None, None, _lib.DTYPE_BF16, *args) == -1
    assert lib.dvs_photometric_backward_ex(C.byref(sh), None, None, None, None, _lib.DTYPE_U8, None, None) == -1


def test_uint8_to_unit_conversion_is_exactly_x_over_255():
    """The in-kernel conversion q = x * r; q += fma(-q, 255, x) * r  (r = fl(1/255)) equals the correctly rounded x / 255 of
    ToTensor for every byte value (restated with numpy float32 / float64 fma)."""
    import numpy as np
    x = np.arange(256, dtype=np.float32)
    r = np.float32(1.0) / np.float32(255.0)
    q = (x * r).astype(np.float32)
    e = (x.astype(np.float64) - q.astype(np.float64) * 255.0).astype(np.float32)      # fma(-q, 255, x): exact in float64, one rounding
    q2 = (e.astype(np.float64) * np.float64(r) + q.astype(np.float64)).astype(np.float32)
    assert np.array_equal(q2, x / np.float32(255.0))

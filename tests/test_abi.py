"""The C-ABI library must load on a CPU-only host and export every function that
include/det3d_b200.h declares (no compute calls here: there is no GPU)."""
import ctypes
import os
import re

from conftest import ROOT
from det3d_b200 import _lib


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "det3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(d3b_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_functions_are_exported():
    names = _declared_functions()
    assert len(names) >= 15
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, "declared in the header but not exported: %s" % missing


def test_binding_table_covers_the_header():
    assert sorted(_lib.SIGNATURES) == _declared_functions()


def test_version_and_error_string():
    L = _lib.lib()
    assert L.d3b_abi_version() == 2
    assert isinstance(L.d3b_last_error(), bytes)
    assert L.d3b_launch_count() >= 0


def test_argument_validation_without_gpu():
    # invalid arguments are rejected before any CUDA call is made
    L = _lib.lib()
    assert L.d3b_voxelize_workspace_bytes(None, 10, 1) == 0
    st = L.d3b_sparse_conv(None, None, None, None, 10, None, None, None)
    assert st == 1 and b"null" in L.d3b_last_error()
    cfg = _lib.VoxelCfg()
    st = L.d3b_voxelize(ctypes.byref(cfg), None, None, 1, None, None, None, None, None, None, 0, None)
    assert st == 1


def test_more_argument_validation_without_gpu():
    """Every entry point validates before it touches CUDA: status 1 (invalid argument) / 3 (unsupported) + message."""
    L = _lib.lib()
    i3 = (ctypes.c_int32 * 3)(3, 3, 3)
    assert L.d3b_rulebook_subm(None, None, 10, None, i3, None, None, None, None, None, None) == 1
    assert L.d3b_rulebook_pairs(None, None, 10, 27, None, None, None, None) == 1
    assert L.d3b_zero_rows(None, None, 17, None, 10, None) == 1 and b"count" in L.d3b_last_error()
    assert L.d3b_zero_rows(None, None, 0, None, 10, None) == 0                 # nothing to do
    assert L.d3b_rotate_nms(None, 10, None, 7, 0.5, 10, None, (ctypes.c_int32 * 1)(), None, 0, None) == 1
    assert b"format" in L.d3b_last_error()
    assert L.d3b_normal_nms(None, 10, None, 5, 0.5, 10, None, (ctypes.c_int32 * 1)(), None, 0, None) == 1
    assert b"mode" in L.d3b_last_error()
    one = (ctypes.c_float * 8)()
    st = L.d3b_pillar_features(one, one, one, one, 4, 100, 2, 64, one, one, one, 0.16, 0.16, 0.0, 0.0, one, None)
    assert st == 3 and b"ndim" in L.d3b_last_error()                         # D3B_ERR_UNSUPPORTED
    off = (ctypes.c_int32 * 2)(0, 5)
    u8 = (ctypes.c_uint8 * 1)(0)
    lag = (ctypes.c_float * 1)(0.0)
    n_out = (ctypes.c_int32 * 1)()
    assert L.d3b_ingest_sweeps(None, off, 40, 5, 4, None, u8, lag, u8, 1.0, None, 5, n_out, None, 0, None) == 1
    assert b"sweeps" in L.d3b_last_error()
    assert L.d3b_ingest_sweeps(None, off, 1, 3, 4, None, u8, lag, u8, 1.0, None, 5, n_out, None, 0, None) == 1
    assert L.d3b_ingest_workspace_bytes(-1) == 0 and L.d3b_nms_workspace_bytes(0) == 16
    q = _lib.PredictParams()
    assert L.d3b_predict_workspace_bytes(ctypes.byref(q)) == 0
    assert L.d3b_predict_task(ctypes.byref(q), None, 0, 0, None, None, 0, None) == 1

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

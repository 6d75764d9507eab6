"""TEST INFRASTRUCTURE -- ctypes access to the REFERENCE iou3d CUDA kernels.

oracle/_ref/libiou3d_ref.so is det3d/ops/iou3d/src/iou3d_kernel.cu compiled as shipped
(oracle/Makefile `ref`); its launchers (iou3d_kernel.cu:354-387) take raw device
pointers, so they can be driven with torch tensors.  The host greedy sweep of
det3d/ops/iou3d/src/iou3d.cpp:103-116 (torch extension, not compiled) is restated in
numpy below.  GPU only; used by tests/ to pin the CUDA product bit-exactly and by
tests/golden/make_golden_gpu.py to produce fixtures.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libiou3d_ref.so")
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(PATH)
        l._Z11nmsLauncherPKfPyif.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float]
        l._Z17nmsNormalLauncherPKfPyif.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float]
        l._Z19boxesioubevLauncheriPKfiS0_Pf.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        l._Z20boxesoverlapLauncheriPKfiS0_Pf.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        for f in ("_Z11nmsLauncherPKfPyif", "_Z17nmsNormalLauncherPKfPyif", "_Z19boxesioubevLauncheriPKfiS0_Pf",
                  "_Z20boxesoverlapLauncheriPKfiS0_Pf"):
            getattr(l, f).restype = None
        _lib = l
    return _lib


def iou_matrix(boxes_a, boxes_b, overlap=False):
    """Reference boxes_iou_bev_gpu / boxes_overlap_bev_gpu (iou3d.cpp:31-71). cuda f32 [N,5] inputs."""
    import torch
    a, b = boxes_a.contiguous(), boxes_b.contiguous()
    out = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    torch.cuda.synchronize()
    fn = lib()._Z20boxesoverlapLauncheriPKfiS0_Pf if overlap else lib()._Z19boxesioubevLauncheriPKfiS0_Pf
    fn(a.shape[0], a.data_ptr(), b.shape[0], b.data_ptr(), out.data_ptr())   # legacy default stream
    torch.cuda.synchronize()
    return out


def nms_mask(boxes_sorted, thresh, normal=False):
    """Reference nms_kernel bitmask [N, ceil(N/64)] (int64 view of u64)."""
    import torch
    b = boxes_sorted.contiguous()
    n = b.shape[0]
    cb = (n + 63) // 64
    mask = torch.zeros((n, cb), dtype=torch.int64, device=b.device)
    torch.cuda.synchronize()
    fn = lib()._Z17nmsNormalLauncherPKfPyif if normal else lib()._Z11nmsLauncherPKfPyif
    fn(b.data_ptr(), mask.data_ptr(), n, float(thresh))
    torch.cuda.synchronize()
    return mask


def host_sweep(mask_np):
    """iou3d.cpp:103-116: greedy sweep over the u64 bitmask rows; returns kept indices."""
    m = mask_np.view(np.uint64)
    n, cb = m.shape
    remv = np.zeros(cb, np.uint64)
    keep = []
    for i in range(n):
        nb, ib = i // 64, i % 64
        if not (int(remv[nb]) >> ib) & 1:
            keep.append(i)
            remv[nb:] |= m[i, nb:]
    return np.asarray(keep, np.int64)


def host_sweep_c(mask_np):
    """The same sweep in C (oracle/iou3d_oracle.c `oracle_iou3d_host_sweep`), for timing the reference fairly: the
    reference's own sweep is compiled C++ (iou3d.cpp:103-116)."""
    from . import build as _build
    lib = C.CDLL(_build.build())
    lib.oracle_iou3d_host_sweep.restype = C.c_int64
    lib.oracle_iou3d_host_sweep.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    m = np.ascontiguousarray(mask_np).view(np.uint64)
    n, cb = m.shape
    keep = np.zeros(max(n, 1), np.int64)
    k = lib.oracle_iou3d_host_sweep(m.ctypes.data, n, cb, keep.ctypes.data)
    return keep[:k]


def nms(boxes_sorted, thresh, normal=False):
    return host_sweep(nms_mask(boxes_sorted, thresh, normal).cpu().numpy())

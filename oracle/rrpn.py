"""TEST INFRASTRUCTURE -- ctypes front of the RRPN rotated IoU / NMS restatement in oracle/iou3d_oracle.c
(section C; follows det3d/ops/nms/nms_gpu.py:180-496).  Never imported by det3d_b200."""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def _l():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.oracle_rrpn_iou.restype = C.c_double
        _lib.oracle_rrpn_iou.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_rrpn_iou_matrix.restype = None
        _lib.oracle_rrpn_iou_matrix.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.oracle_rrpn_nms.restype = C.c_int64
        _lib.oracle_rrpn_nms.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]
    return _lib


def rotate_iou_eval(boxes, query_boxes, criterion=-1):
    b = np.ascontiguousarray(boxes, np.float32)
    q = np.ascontiguousarray(query_boxes, np.float32)
    out = np.zeros((b.shape[0], q.shape[0]), np.float32)
    if out.size:
        _l().oracle_rrpn_iou_matrix(b.ctypes.data, b.shape[0], q.ctypes.data, q.shape[0], int(criterion), out.ctypes.data)
    return out


def rotate_nms(dets, thresh):
    """dets [N,6] cx,cy,w,l,r,score -> (kept original indices by descending score, #near-threshold pairs)."""
    d = np.ascontiguousarray(dets, np.float32)
    n = d.shape[0]
    order = np.ascontiguousarray(d[:, 5].argsort()[::-1].astype(np.int32))
    keep = np.zeros(max(n, 1), np.int64)
    near = np.zeros(1, np.int64)
    k = _l().oracle_rrpn_nms(d.ctypes.data, order.ctypes.data, n, float(thresh), keep.ctypes.data, near.ctypes.data)
    return keep[:k], int(near[0])

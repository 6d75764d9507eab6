"""TEST INFRASTRUCTURE -- voxelizer oracle (CPU).

`points_to_voxel` = ctypes call into oracle/voxel_oracle.c (restatement of
det3d/ops/point_cloud/point_cloud_ops.py:7-55,112-184); `points_to_voxel_numpy`
= an independent vectorised numpy restatement used to cross-check it.
Both are pinned to the reference numba function by tests/golden/voxel_*.npz.
"""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.oracle_points_to_voxel.restype = C.c_int32
        _lib.oracle_points_to_voxel.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                                C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p]
    return _lib


def grid_size(voxel_size, coors_range):
    vs = np.asarray(voxel_size, np.float32)
    cr = np.asarray(coors_range, np.float32)
    return np.round((cr[3:] - cr[:3]) / vs).astype(np.int32)  # point_cloud_ops.py:26-29


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True, max_voxels=20000,
                    dense_map=None):
    """Same signature / outputs as the reference function (reverse_index=True only)."""
    assert reverse_index, "the Det3D pipeline only uses reverse_index=True (voxel_generator.py:25)"
    pts = np.ascontiguousarray(points, np.float32)
    n, ndim = pts.shape
    vs = np.ascontiguousarray(voxel_size, np.float32)
    cr = np.ascontiguousarray(coors_range, np.float32)
    voxels = np.zeros((max_voxels, max_points, ndim), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    m = lib().oracle_points_to_voxel(
        pts.ctypes.data, n, ndim, vs.ctypes.data, cr.ctypes.data, max_points, max_voxels,
        voxels.ctypes.data, coors.ctypes.data, num.ctypes.data,
        None if dense_map is None else dense_map.ctypes.data)
    if m < 0:
        raise MemoryError("oracle dense map allocation failed")
    return voxels[:m], coors[:m], num[:m]


def points_to_voxel_numpy(points, voxel_size, coors_range, max_points=35, max_voxels=20000):
    """Vectorised restatement: fp32 floor((p-lo)/vs), first-appearance order, break at max_voxels."""
    pts = np.ascontiguousarray(points, np.float32)
    vs = np.asarray(voxel_size, np.float32)
    cr = np.asarray(coors_range, np.float32)
    grid = grid_size(vs, cr)
    n, ndim = pts.shape
    with np.errstate(invalid="ignore"):
        c = np.floor((pts[:, :3] - cr[:3]) / vs)           # fp32 throughout
        ok = np.all((c >= 0) & (c < grid.astype(np.float32)), axis=1)
    idx = np.nonzero(ok)[0]
    ci = c[idx].astype(np.int64)
    lin = (ci[:, 2] * grid[1] + ci[:, 1]) * grid[0] + ci[:, 0]
    uniq, first_pos, inv = np.unique(lin, return_index=True, return_inverse=True)
    order = np.argsort(first_pos, kind="stable")           # voxels by first appearance
    rank_of_uniq = np.empty_like(order)
    rank_of_uniq[order] = np.arange(order.size)
    vid = rank_of_uniq[inv]                                # voxel id of every in-range point
    m = min(order.size, max_voxels)
    if order.size > max_voxels:
        cut = idx[first_pos[order[max_voxels]]]            # the point that triggers `break`
        keep = idx < cut
        idx, ci, vid = idx[keep], ci[keep], vid[keep]
    voxels = np.zeros((m, max_points, ndim), np.float32)
    coors = np.zeros((m, 3), np.int32)
    num = np.zeros((m,), np.int32)
    # slot = rank of the point within its voxel, input order
    o = np.argsort(vid, kind="stable")
    v_sorted = vid[o]
    starts = np.r_[0, np.nonzero(np.diff(v_sorted))[0] + 1]
    slot = np.arange(o.size) - np.repeat(starts, np.diff(np.r_[starts, o.size]))
    sel = slot < max_points
    voxels[v_sorted[sel], slot[sel]] = pts[idx[o][sel]]
    np.add.at(num, v_sorted[sel], 1)
    first_pts = order[:m]
    cfirst = c[idx_all_first(ok, first_pos, first_pts)].astype(np.int32) if m else np.zeros((0, 3), np.int32)
    coors[:, 0], coors[:, 1], coors[:, 2] = cfirst[:, 2], cfirst[:, 1], cfirst[:, 0]
    return voxels, coors, num


def idx_all_first(ok, first_pos, first_pts):
    return np.nonzero(ok)[0][first_pos[first_pts]]

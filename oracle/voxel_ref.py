"""TEST INFRASTRUCTURE -- the REFERENCE voxelizer kernel, run as the reference runs it.

`oracle/_ref/ref_voxel_aot*.so` is `_points_to_voxel_reverse_kernel` of
det3d/ops/point_cloud/point_cloud_ops.py:7-55 compiled ahead of time from the reference checkout (oracle/build.py,
numba.pycc); it is a build output (git-ignored, travels to the GPU box) -- no reference source lives in this repository.
`points_to_voxel` below restates only the wrapper around it (point_cloud_ops.py:112-184: dtype coercion, the dense
`-np.ones(grid[::-1])` lookup map, the zero-filled outputs, the final slices), so that a call costs what
`VoxelGenerator.generate` costs in the reference (voxel_generator.py:19-27).

Used by tests/ (pins the C restatement oracle/voxel_oracle.c on arbitrary inputs, beyond the committed goldens) and by
bench.py's cpu_baseline / --impl reference legs (full call, loop only, one worker per host core).
"""
import importlib.machinery
import importlib.util

import numpy as np

from . import build as _build

_mod = None


def available():
    return _build.ref_voxel_path() is not None


def kernel():
    global _mod
    if _mod is None:
        path = _build.ref_voxel_path()
        if path is None:
            raise RuntimeError("oracle/_ref/ref_voxel_aot*.so is not built (needs /root/reference at build time)")
        loader = importlib.machinery.ExtensionFileLoader(_build.REF_VOXEL_NAME, path)
        spec = importlib.util.spec_from_loader(_build.REF_VOXEL_NAME, loader)
        _mod = importlib.util.module_from_spec(spec)
        loader.exec_module(_mod)
    return _mod.points_to_voxel_reverse_kernel


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True, max_voxels=20000, buffers=None):
    """point_cloud_ops.py:112-184 around the AOT kernel.  `buffers` (from `alloc`) skips the per-call allocations: the
    "loop only" timing of SURVEY 8d(1); the caller must `reset` them between calls."""
    assert reverse_index, "the Det3D pipeline only uses reverse_index=True (voxel_generator.py:25)"
    points = np.ascontiguousarray(points, np.float32)
    if not isinstance(voxel_size, np.ndarray):
        voxel_size = np.array(voxel_size, dtype=points.dtype)
    if not isinstance(coors_range, np.ndarray):
        coors_range = np.array(coors_range, dtype=points.dtype)
    if buffers is None:
        buffers = alloc(voxel_size, coors_range, max_points, max_voxels, points.shape[-1])
    num_points_per_voxel, coor_to_voxelidx, voxels, coors = buffers
    voxel_num = kernel()(points, voxel_size.astype(np.float32), coors_range.astype(np.float32), num_points_per_voxel,
                         coor_to_voxelidx, voxels, coors, max_points, max_voxels)
    return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


def alloc(voxel_size, coors_range, max_points, max_voxels, ndim):
    """The four arrays the reference allocates per call (:149-154), incl. the dense map (360 MB on the KITTI grid)."""
    voxel_size = np.asarray(voxel_size, np.float32)
    coors_range = np.asarray(coors_range, np.float32)
    shape = tuple(np.round((coors_range[3:] - coors_range[:3]) / voxel_size).astype(np.int32).tolist())[::-1]
    return (np.zeros(shape=(max_voxels,), dtype=np.int32), -np.ones(shape=shape, dtype=np.int32),
            np.zeros(shape=(max_voxels, max_points, ndim), dtype=np.float32), np.zeros(shape=(max_voxels, 3), dtype=np.int32))


def reset(buffers, coors_used):
    """Undo one call on preallocated buffers by clearing only what it touched (not part of the timed loop)."""
    num_points_per_voxel, coor_to_voxelidx, voxels, coors = buffers
    c = coors_used
    coor_to_voxelidx[c[:, 0], c[:, 1], c[:, 2]] = -1
    m = c.shape[0]
    num_points_per_voxel[:m] = 0
    voxels[:m] = 0
    coors[:m] = 0

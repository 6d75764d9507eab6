"""GPU-backed `VoxelGenerator`.

API mirror of the reference class at det3d/core/input/voxel_generator.py:5-43
(constructor arguments, `generate(points)` -> (voxels, coors zyx, num_points) as
numpy, read-only `voxel_size`, `max_num_points_per_voxel`, `point_cloud_range`,
`grid_size` int64 xyz).  The work happens in csrc/voxelize.cu via
`det3d_b200.ops.point_cloud.points_to_voxel`.
"""
import numpy as np

from det3d_b200.ops.point_cloud.point_cloud_ops import points_to_voxel
from det3d_b200.ops.point_cloud.voxelize import grid_size_of


def _readonly(field):
    return property(lambda self: self._state[field])


class VoxelGenerator:
    voxel_size = _readonly("voxel_size")
    max_num_points_per_voxel = _readonly("max_num_points")
    point_cloud_range = _readonly("point_cloud_range")
    grid_size = _readonly("grid_size")

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        pcr = np.asarray(point_cloud_range, dtype=np.float32).copy()
        vs = np.asarray(voxel_size, dtype=np.float32).copy()
        self._state = dict(
            voxel_size=vs,
            point_cloud_range=pcr,
            grid_size=grid_size_of(vs, pcr),  # int64 [3], xyz
            max_num_points=max_num_points,
            max_voxels=max_voxels,
        )

    def generate(self, points, max_voxels=20000):
        """`max_voxels` is accepted and ignored, exactly like the reference
        (voxel_generator.py:19-27 passes the constructor's value)."""
        s = self._state
        return points_to_voxel(points, s["voxel_size"], s["point_cloud_range"], s["max_num_points"],
                               reverse_index=True, max_voxels=s["max_voxels"])

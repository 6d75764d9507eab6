"""`points_to_voxel` with the reference's numpy signature, computed on the GPU.

Drop-in for det3d/ops/point_cloud/point_cloud_ops.py:112-184: same arguments,
same three numpy outputs (voxels [M,max_points,ndim] f32, coordinates [M,3]
int32 in zyx order when reverse_index else xyz, num_points_per_voxel [M] int32),
bit-identical contents.  There is no CPU fallback: a CUDA device is required.
"""
import numpy as np
import torch

from .voxelize import Voxelizer

_CACHE = {}


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True, max_voxels=20000):
    if not torch.cuda.is_available():
        raise RuntimeError("det3d_b200.points_to_voxel needs a CUDA device (no CPU fallback)")
    pts = np.ascontiguousarray(points, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[1] < 3:
        raise ValueError("points must be [N, >=3]")
    vs = np.asarray(voxel_size, dtype=np.float32)
    cr = np.asarray(coors_range, dtype=np.float32)
    key = (vs.tobytes(), cr.tobytes(), int(max_points), int(max_voxels))
    vox = _CACHE.get(key)
    if vox is None:
        vox = _CACHE[key] = Voxelizer(vs, cr, max_points, max_voxels, want_voxels=True, want_mean=False)
    dev_pts = torch.from_numpy(pts).cuda()
    out = vox(dev_pts)
    m = int(out["counts"][0].item())
    voxels = out["voxels"][:m].cpu().numpy()
    coors = out["coors"][:m, 1:].cpu().numpy()
    if not reverse_index:
        coors = np.ascontiguousarray(coors[:, ::-1])
    num = out["num_points"][:m].cpu().numpy()
    return voxels, coors, num

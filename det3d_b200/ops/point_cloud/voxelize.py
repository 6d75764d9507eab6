"""Device voxelizer: torch wrapper over d3b_voxelize (csrc/voxelize.cu).

Reference semantics: det3d/ops/point_cloud/point_cloud_ops.py:7-55,112-184.
"""
import ctypes as C

import numpy as np
import torch

from ... import _lib
from ..._lib import VoxelCfg


def grid_size_of(voxel_size, point_cloud_range):
    """round((hi - lo) / vs) in fp32, exactly as voxel_generator.py:7-11 / point_cloud_ops.py:26-29."""
    pcr = np.asarray(point_cloud_range, dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    return np.round((pcr[3:] - pcr[:3]) / vs).astype(np.int64)


class Voxelizer:
    """Voxelizes a batch of clouds on the GPU in one call; outputs stay on the device.

    out = voxelizer(points, offsets) with points [N_total, ndim] f32 cuda and
    offsets a host list [0, n0, n0+n1, ...].  Returns a dict of device tensors:
      voxels [cap, max_points, ndim] (optional), coors [cap, 4] (b,z,y,x),
      num_points [cap], mean [cap, ndim], counts int32[batch+1] (last = total rows).
    Only the first counts[-1] rows are defined.
    """

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels, want_voxels=True,
                 want_mean=True):
        self.voxel_size = np.asarray(voxel_size, dtype=np.float32)
        self.point_cloud_range = np.asarray(point_cloud_range, dtype=np.float32)
        self.grid_size = grid_size_of(self.voxel_size, self.point_cloud_range)
        self.max_num_points = int(max_num_points)
        self.max_voxels = int(max_voxels)
        self.want_voxels = want_voxels
        self.want_mean = want_mean
        self._bufs = {}

    def _cfg(self, ndim):
        cfg = VoxelCfg()
        for j in range(3):
            cfg.voxel_size[j] = float(self.voxel_size[j])
            cfg.range_min[j] = float(self.point_cloud_range[j])
            cfg.grid[j] = int(self.grid_size[j])
        cfg.ndim = ndim
        cfg.max_points = self.max_num_points
        cfg.max_voxels = self.max_voxels
        return cfg

    def _buffers(self, n_total, batch, ndim, device):
        key = (batch, ndim, device)
        b = self._bufs.get(key)
        if b is not None and b["n_cap"] >= n_total:
            return b
        n_cap = max(n_total, 1)
        cfg = self._cfg(ndim)
        ws_bytes = _lib.lib().d3b_voxelize_workspace_bytes(C.byref(cfg), n_cap, batch)
        cap = batch * self.max_voxels
        b = {
            "n_cap": n_cap,
            "cfg": cfg,
            "ws": torch.empty(ws_bytes, dtype=torch.uint8, device=device),
            "coors": torch.empty((cap, 4), dtype=torch.int32, device=device),
            "num_points": torch.empty(cap, dtype=torch.int32, device=device),
            "counts": torch.zeros(batch + 1, dtype=torch.int32, device=device),
            "voxels": torch.empty((cap, self.max_num_points, ndim), dtype=torch.float32, device=device)
            if self.want_voxels else None,
            "mean": torch.empty((cap, ndim), dtype=torch.float32, device=device) if self.want_mean else None,
        }
        self._bufs[key] = b
        return b

    def __call__(self, points, offsets=None):
        assert points.is_cuda and points.dtype == torch.float32 and points.dim() == 2
        points = points.contiguous()
        n_total, ndim = points.shape
        if offsets is None:
            offsets = [0, n_total]
        batch = len(offsets) - 1
        b = self._buffers(n_total, batch, ndim, points.device)
        off = (C.c_int32 * (batch + 1))(*[int(o) for o in offsets])
        with _lib.on_device_of(points), _lib.timed("voxelize", n_points=n_total, ndim=ndim, batch=batch):
            st = _lib.lib().d3b_voxelize(
                C.byref(b["cfg"]), points.data_ptr() if n_total > 0 else None, off, batch,
                _lib.ptr(b["voxels"]), b["coors"].data_ptr(), b["num_points"].data_ptr(), _lib.ptr(b["mean"]),
                b["counts"].data_ptr(), b["ws"].data_ptr(), b["ws"].numel(), _lib.current_stream(),
            )
        _lib.check(st, "d3b_voxelize")
        out = {k: b[k] for k in ("voxels", "coors", "num_points", "mean", "counts")}
        # the per-voxel point-index lists the voxelizer built on the way ([batch][max_voxels][max_points] indices into
        # `points`, valid until the next call): what the fused pillar reader consumes instead of `voxels`
        out["point_lists"] = dict(points=points, lists_ptr=_lib.lib().d3b_voxelize_point_lists(
            C.byref(b["cfg"]), n_total, batch, b["ws"].data_ptr()), batch=batch, max_voxels=self.max_voxels,
            max_points=self.max_num_points, keepalive=b["ws"])
        return out

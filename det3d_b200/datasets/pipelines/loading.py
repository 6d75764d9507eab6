"""On-disk -> device ingest: `LoadPointCloudFromFile` with the reference's registry name, constructor
and `__call__(res, info)` contract (det3d/datasets/pipelines/loading.py:67-124), the NuScenes multi-sweep merge done
on the GPU by d3b_ingest_sweeps (csrc/ingest.cu).

File reading stays on the host (it is I/O): KITTI `.bin` = flat float32 [N, num_point_features] (:92-94), nuScenes
`.pcd.bin` = float32 [N, 5] of which the first 4 columns are kept (`read_file`, :17-31).  Everything after that --
the 1 m `remove_close` filter of the sweeps (:34-43), the float64 rigid transform (:55-58), the time-lag column and
the concatenation (:115-124) -- is one call over the raw bytes of all sweeps.  `res["lidar"]` receives the same
numpy fields as the reference (`points`, `times`, `combined`) plus `combined_cuda`, the device tensor the
voxelizer consumes directly (no second H2D copy).
"""
import ctypes as C
from pathlib import Path

import numpy as np
import torch

from ... import _lib
from ..registry import PIPELINES


def read_file(path, tries=2, num_point_feature=4, keep_raw=False):
    """loading.py:17-31: float32 file, truncated to whole 5-float records; [n, num_point_feature] (or the raw
    [n, 5] records with keep_raw, which is what the device ingest uploads)."""
    points = None
    try_cnt = 0
    while points is None and try_cnt < tries:
        try_cnt += 1
        try:
            points = np.fromfile(path, dtype=np.float32)
            s = points.shape[0]
            if s % 5 != 0:
                points = points[: s - (s % 5)]
            points = points.reshape(-1, 5)
            if not keep_raw:
                points = points[:, :num_point_feature]
        except Exception:
            points = None
    return points


def ingest_sweeps(raw_sweeps, transforms, time_lags, radius=1.0, n_feat=4, device="cuda"):
    """raw_sweeps: list of float32 [n_s, 5] arrays, key frame first.  transforms[s]: 4x4 array or None;
    time_lags[s]: float.  The key frame (s = 0) is neither filtered nor transformed, as in the reference.
    Returns the device tensor [N, n_feat + 1] (x, y, z, .., time lag), input order preserved."""
    if not torch.cuda.is_available():
        raise RuntimeError("det3d_b200: the multi-sweep ingest needs a CUDA device (there is no CPU fallback)")
    n_sweeps = len(raw_sweeps)
    sizes = [int(r.shape[0]) for r in raw_sweeps]
    offsets = (C.c_int32 * (n_sweeps + 1))(*np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64).tolist())
    n_total = int(sum(sizes))
    dev = torch.device(device)
    stride = int(raw_sweeps[0].shape[1]) if n_sweeps else 5
    raw = torch.from_numpy(np.ascontiguousarray(np.concatenate(raw_sweeps, axis=0), dtype=np.float32)).to(dev)
    tm = np.zeros((n_sweeps, 16), np.float64)
    has = np.zeros(n_sweeps, np.uint8)
    for s, t in enumerate(transforms):
        if t is not None:
            tm[s] = np.asarray(t, np.float64).reshape(16)
            has[s] = 1
    lag = np.asarray(time_lags, np.float64).astype(np.float32)       # times.astype(points.dtype), loading.py:119
    filt = np.ones(n_sweeps, np.uint8)
    filt[0] = 0
    out = torch.empty((max(n_total, 1), n_feat + 1), dtype=torch.float32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(_lib.lib().d3b_ingest_workspace_bytes(n_total), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().d3b_ingest_sweeps(
            raw.data_ptr(), offsets, n_sweeps, stride, n_feat, tm.ctypes.data, has.ctypes.data, lag.ctypes.data,
            filt.ctypes.data, C.c_float(radius), out.data_ptr(), n_total, n_out.data_ptr(), ws.data_ptr(), ws.numel(),
            _lib.current_stream())
        _lib.check(st, "d3b_ingest_sweeps")
        n = int(n_out.item())            # API boundary: the reference returns exactly-sized arrays
    return out[:n]


@PIPELINES.register_module
class LoadPointCloudFromFile(object):
    def __init__(self, dataset="KittiDataset", **kwargs):
        self.type = dataset
        self.random_select = kwargs.get("random_select", False)
        self.npoints = kwargs.get("npoints", 16834)
        self.device = kwargs.get("device", "cuda")

    def __call__(self, res, info):
        res["type"] = self.type
        if self.type == "KittiDataset":
            pc_info = info["point_cloud"]
            velo_path = Path(pc_info["velodyne_path"])
            if not velo_path.is_absolute():
                velo_path = Path(res["metadata"]["image_prefix"]) / pc_info["velodyne_path"]
            reduced = velo_path.parent.parent / (velo_path.parent.stem + "_reduced") / velo_path.name
            if reduced.exists():
                velo_path = reduced
            points = np.fromfile(str(velo_path), dtype=np.float32, count=-1).reshape(
                [-1, res["metadata"]["num_point_features"]])
            res["lidar"]["points"] = points
        elif self.type == "NuScenesDataset":
            nsweeps = res["lidar"]["nsweeps"]
            raws = [read_file(str(Path(info["lidar_path"])), keep_raw=True)]
            transforms, lags = [None], [0.0]
            assert (nsweeps - 1) <= len(info["sweeps"]), "nsweeps {} should not greater than list length {}.".format(
                nsweeps, len(info["sweeps"]))
            for i in np.random.choice(len(info["sweeps"]), nsweeps - 1, replace=False):      # same RNG call as :110
                sweep = info["sweeps"][i]
                raws.append(read_file(str(sweep["lidar_path"]), keep_raw=True))
                transforms.append(sweep["transform_matrix"])
                lags.append(sweep["time_lag"])
            combined = ingest_sweeps(raws, transforms, lags, radius=1.0, n_feat=4, device=self.device)
            host = combined.cpu().numpy()
            res["lidar"]["points"] = host[:, :4]
            res["lidar"]["times"] = host[:, 4:5]
            res["lidar"]["combined"] = host
            res["lidar"]["combined_cuda"] = combined
        else:
            raise NotImplementedError("LoadPointCloudFromFile: dataset type %s" % self.type)
        return res, info


def _kitti_boxes_camera_to_lidar(annos, r_rect, velo2cam):
    """location/dimensions/rotation_y (camera frame, bottom-centre) -> lidar boxes [n,7] x,y,z,w,l,h,r with the
    gravity centre: box_np_ops.box_camera_to_lidar + change_box3d_center_ (box_np_ops.py:909-930,1346-1349)."""
    gt = np.concatenate([annos["location"], annos["dimensions"], annos["rotation_y"][..., np.newaxis]], axis=1).astype(np.float32)
    xyz = np.concatenate([gt[:, 0:3], np.ones((gt.shape[0], 1))], axis=-1)
    xyz_lidar = (xyz @ np.linalg.inv((r_rect @ velo2cam).T))[..., :3]
    l, h, w, r = gt[:, 3:4], gt[:, 4:5], gt[:, 5:6], gt[:, 6:7]
    boxes = np.concatenate([xyz_lidar, w, l, h, r], axis=1)
    boxes[..., :3] += boxes[..., 3:6] * (np.array([0.5, 0.5, 0.5], boxes.dtype) - np.array([0.5, 0.5, 0], boxes.dtype))
    return boxes


@PIPELINES.register_module
class LoadPointCloudAnnotations(object):
    """det3d/datasets/pipelines/loading.py:165-224: calibration + (when the info record has them) ground-truth boxes
    for evaluation.  Host metadata only -- nothing here is on the compute path."""

    def __init__(self, with_bbox=True, **kwargs):
        pass

    def __call__(self, res, info):
        if res["type"] in ["NuScenesDataset", "LyftDataset"] and "gt_boxes" in info:
            res["lidar"]["annotations"] = {
                "boxes": info["gt_boxes"].astype(np.float32),
                "names": info["gt_names"],
                "tokens": info["gt_boxes_token"],
                "velocities": info["gt_boxes_velocity"].astype(np.float32),
            }
        elif res["type"] == "KittiDataset":
            calib = info["calib"]
            res["calib"] = {"rect": calib["R0_rect"], "Trv2c": calib["Tr_velo_to_cam"], "P2": calib["P2"]}
            if "annos" in info:
                annos = info["annos"]
                keep = [i for i, x in enumerate(annos["name"]) if x != "DontCare"]       # kitti_common.remove_dontcare
                annos = {k: v[keep] for k, v in annos.items()}
                res["lidar"]["annotations"] = {
                    "boxes": _kitti_boxes_camera_to_lidar(annos, calib["R0_rect"], calib["Tr_velo_to_cam"]),
                    "names": annos["name"],
                }
                res.setdefault("cam", {})["annotations"] = {"boxes": annos["bbox"], "names": annos["name"]}
        elif res["type"] in ["NuScenesDataset", "LyftDataset"]:
            pass
        else:
            raise NotImplementedError("LoadPointCloudAnnotations: dataset type %r" % (res["type"],))
        return res, info

"""End-to-end, device-resident inference: raw points -> detections.

The reference splits this path between DataLoader worker processes (numba voxelizer,
per-sample CPU anchor generation, numpy collate; det3d/datasets/pipelines/preprocess.py:
259-304,346-378, det3d/torchie/parallel/collate.py:90-150), the GPU model
(det3d/models/detectors/voxelnet.py:30-52) and a CPU NMS (core/bbox/box_torch_ops.py:
528-549).  `InferencePipeline` keeps the same stages and the same model objects (built
from an unmodified Det3D config through the registries) but runs every stage on the GPU
with fixed-shape buffers: points are voxelized for the whole batch in one call with the
batch index and the VoxelFeatureExtractorV3 mean fused in, anchors are generated once
and cached on the device, and detections come back as fixed-shape tensors.
"""
import collections
import warnings

import numpy as np
import torch

from det3d_b200 import _lib
from det3d_b200.core.anchor.anchor_generator import anchors_for_tasks
from det3d_b200.models import build_detector
from det3d_b200.ops.point_cloud.voxelize import Voxelizer


class InferencePipeline:
    def __init__(self, cfg, model=None, device="cuda", strict_fp32=True):
        self.cfg = cfg
        self.device = torch.device(device)
        if model is None:
            model = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
        self.model = model.to(self.device).eval()
        vg = cfg.voxel_generator
        # VoxelFeatureExtractorV3 only needs the per-voxel mean (fused into the voxelizer); the pillar reader
        # consumes the point slots themselves
        self._reader_takes_points = cfg.model["reader"]["type"] != "VoxelFeatureExtractorV3"
        # a pillar reader that can walk the voxelizer's point-index lists does not need the [M, P, ndim] tensor either
        self._reader_takes_lists = (self._reader_takes_points and hasattr(self.model.reader, "forward_lists")
                                    and len(getattr(self.model.reader, "pfn_layers", [])) == 1)
        self.voxelizer = Voxelizer(vg["voxel_size"], vg["range"], vg["max_points_in_voxel"], vg["max_voxel_num"],
                                   want_voxels=self._reader_takes_points and not self._reader_takes_lists,
                                   want_mean=not self._reader_takes_points)
        self.grid_size = self.voxelizer.grid_size
        self.num_point_features = int(cfg.model["reader"].get("num_input_features", 4))
        out_size_factor = cfg.assigner["out_size_factor"] if "assigner" in cfg else 8
        anchors = anchors_for_tasks(cfg.target_assigner, self.grid_size, out_size_factor)
        self._anchors = [torch.from_numpy(a).to(self.device) for a in anchors]
        self._anchor_cache = {}
        self.strict_fp32 = strict_fp32
        self._graphs = collections.OrderedDict()      # offsets tuple -> captured graph, least recently used first
        self.max_graphs = 8
        self._ovf_host = None

    def anchors(self, batch):
        a = self._anchor_cache.get(batch)
        if a is None:
            a = self._anchor_cache[batch] = [t.unsqueeze(0).expand(batch, -1, -1).contiguous() for t in self._anchors]
        return a

    @torch.no_grad()
    def forward_device(self, points, offsets):
        """points [N_total, ndim] f32 on the device, offsets host list -> fixed-shape detections."""
        batch = len(offsets) - 1
        vox = self.voxelizer(points, offsets)
        example = dict(
            voxels=vox["voxels"] if self._reader_takes_points else vox["mean"], coordinates=vox["coors"], num_points=vox["num_points"],
            num_voxels=[None] * batch, shape=[self.grid_size], anchors=self.anchors(batch),
            n_voxels_dev=vox["counts"][batch:batch + 1],
        )
        if self._reader_takes_lists:
            example["point_lists"] = dict(vox["point_lists"], counts=vox["counts"])
        prev, prev_bench = torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark
        if self.strict_fp32:
            torch.backends.cudnn.allow_tf32 = False
        # shapes are static per pipeline: let cuDNN pick its fastest fp32 algorithm for the dense layers that stay on it
        # (strided / multi-stage RPNs: PointPillars, CBGS)
        torch.backends.cudnn.benchmark = True
        try:
            det = self.model(example, return_loss=False, device_output=True)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
            torch.backends.cudnn.benchmark = prev_bench
        det["voxel_counts"] = vox["counts"]
        return det

    # ---- f16-range guard of the FP16x3 kernels -------------------------------------------------------------------
    def overflow_flag(self):
        """Device int32[1]: nonzero once a feature left the f16 range (the FP16x3 kernels never saturate silently)."""
        flag = getattr(self.model, "overflow_flag", None)
        return None if flag is None else flag(self.device)

    def check_overflow(self, flag_value):
        """Host side of the guard: on a raised flag switch the model to the tf32x3 kernels (permanently: the weights /
        inputs that overflowed once will again) and tell the caller to re-run.  Returns True if a re-run is needed."""
        if not flag_value:
            return False
        warnings.warn("det3d_b200: a feature left the f16 range (|x| >= 65504); re-running on the tf32x3 kernels")
        self.model.set_math("tf32x3")
        self.overflow_flag().zero_()
        self._graphs.clear()
        return True

    @torch.no_grad()
    def forward_graphed(self, points, offsets):
        """forward_device + pack replayed from a CUDA graph (captured on first use per `offsets`).

        Every stage has static shapes and device-resident counts, so the whole forward -- ~50
        det3d_b200 launches plus the torch glue -- is one graph launch.  `points` may live on the
        host (pinned) or the device; it is copied into the graph's static input buffer.
        Returns the packed detections [B, D, nd+3] (a static buffer: consume before the next call)."""
        key = tuple(int(o) for o in offsets)
        entry = self._graphs.get(key)
        if entry is not None:
            self._graphs.move_to_end(key)
        if entry is None:
            # One graph per distinct `offsets` (the per-cloud point counts are host arguments of d3b_voxelize).  Real
            # LiDAR frames rarely repeat a point count: keep at most `max_graphs` graphs (LRU) so memory stays bounded;
            # callers with free-running sizes should pad clouds to a few bucket sizes or use forward_device.
            while len(self._graphs) >= self.max_graphs:
                self._graphs.popitem(last=False)
            static_pts = torch.zeros((key[-1], points.shape[1]), dtype=torch.float32, device=self.device)
            static_pts.copy_(points)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):                      # warm-up: lazy buffers, weight packing, cuDNN plans
                    self.pack(self.forward_device(static_pts, list(key)))
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.pack(self.forward_device(static_pts, list(key)))
                _lib.graph_mark("end")          # (bench.py's in-graph stage timing; nothing unless _lib.GRAPH_MARKS is set)
            entry = self._graphs[key] = (graph, static_pts, out)
        graph, static_pts, out = entry
        static_pts.copy_(points, non_blocking=True)
        graph.replay()
        return out

    @staticmethod
    def pack(det):
        """-> one [B, D, nd+3] f32 tensor: box, score, label, valid (the D2H / all-gather payload)."""
        if "packed" in det:
            return det["packed"]
        return torch.cat([det["boxes"], det["scores"].unsqueeze(-1), det["labels"].float().unsqueeze(-1),
                          det["valid"].float().unsqueeze(-1)], dim=-1).contiguous()

    @torch.no_grad()
    def infer_host(self, clouds, pinned_out=None):
        """clouds: list of pinned (or plain) host float32 tensors [N_i, ndim].
        H2D copy, forward, D2H of the packed detections.  Returns a host tensor [B, D, nd+3]."""
        offsets = [0]
        for c in clouds:
            offsets.append(offsets[-1] + c.shape[0])
        ndim = clouds[0].shape[1]
        pts = torch.empty((offsets[-1], ndim), dtype=torch.float32, device=self.device)
        for c, a, b in zip(clouds, offsets[:-1], offsets[1:]):
            pts[a:b].copy_(c, non_blocking=True)
        for _attempt in range(2):
            packed = self.pack(self.forward_device(pts, offsets))
            if pinned_out is None:
                pinned_out = torch.empty(packed.shape, dtype=torch.float32, pin_memory=True)
            pinned_out.copy_(packed, non_blocking=True)
            flag = self.overflow_flag()
            if flag is not None:
                if self._ovf_host is None:
                    self._ovf_host = torch.zeros(1, dtype=torch.int32, pin_memory=True)
                self._ovf_host.copy_(flag, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            if flag is None or not self.check_overflow(int(self._ovf_host[0])):
                break
        return pinned_out

    @staticmethod
    def unpack(packed_host):
        """host [B, D, nd+3] -> list of dict(box3d_lidar, scores, label_preds) per sample."""
        out = []
        for row in packed_host:
            m = row[:, -1] > 0.5
            out.append(dict(box3d_lidar=row[m, :-3], scores=row[m, -3], label_preds=row[m, -2].long()))
        return out

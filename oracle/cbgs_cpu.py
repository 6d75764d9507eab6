"""TEST INFRASTRUCTURE -- CPU restatement of the reference's CBGS (nuScenes) inference forward: SpMiddleResNetFHD
(oracle/spconv.py, det3d/models/backbones/scn.py:308-370), the two-block RPN with a ConvTranspose deblock
(oracle/pillars_cpu.rpn_forward_multi, det3d/models/necks/rpn.py:67-159), six task heads
(det3d/models/bbox_heads/mg_head.py:198-230) and the per-task predict (oracle/predict_cpu.py,
mg_head.py:697-1085, 9-dim boxes, angle-vector encoding).  Never imported by det3d_b200."""
import time

import torch
import torch.nn.functional as F

from .pillars_cpu import rpn_forward_multi
from .predict_cpu import predict_sample_task
from .second_cpu import SecondCPU


class CbgsCPU(SecondCPU):
    def heads(self, x):
        out = []
        n_tasks = len(self.anchors)
        for t in range(n_tasks):
            p = "bbox_head.tasks.%d." % t
            d = {}
            for name, key in (("box", "conv_box"), ("cls", "conv_cls"), ("dir", "conv_dir")):
                if p + key + ".weight" in self.sd:
                    d[name] = F.conv2d(x, self.sd[p + key + ".weight"], self.sd[p + key + ".bias"]).permute(0, 2, 3, 1).contiguous()
            out.append(d)
        return out

    def predict_tasks(self, heads):
        tc = self.cfg.test_cfg
        vec = bool(self.cfg.box_coder.get("encode_angle_vector", False)) if hasattr(self.cfg, "box_coder") else True
        B = heads[0]["cls"].shape[0]
        res = []
        for b in range(B):
            boxes, scores, labels, flag = [], [], [], 0
            for t, h in enumerate(heads):
                anchors = self.anchors[t]
                a = anchors.shape[0]
                n_cls = h["cls"].shape[-1] * h["cls"].shape[1] * h["cls"].shape[2] // a
                code = h["box"][b].numel() // a
                dirs = h["dir"][b].reshape(a, 2) if "dir" in h else None
                bx, sc, lb = predict_sample_task(h["cls"][b].reshape(a, n_cls), h["box"][b].reshape(a, code), dirs, anchors, tc, vec)
                boxes.append(bx); scores.append(sc); labels.append(lb + flag)
                flag += n_cls
            res.append(dict(box3d_lidar=torch.cat(boxes), scores=torch.cat(scores), label_preds=torch.cat(labels)))
        return res

    @torch.no_grad()
    def forward(self, clouds, stages=None):
        t0 = time.perf_counter()
        voxels, coors, nums = self.voxelize(clouds)
        t1 = time.perf_counter()
        dense = self.backbone(voxels, coors, nums, len(clouds))
        t2 = time.perf_counter()
        x = rpn_forward_multi(self.sd, dense, self.cfg.model["neck"])
        heads = self.heads(x)
        t3 = time.perf_counter()
        dets = self.predict_tasks(heads)
        t4 = time.perf_counter()
        self.timings = dict(voxelize=t1 - t0, backbone=t2 - t1, rpn_head=t3 - t2, predict=t4 - t3)
        if stages is not None:
            stages.update(dict(voxels=voxels, coors=coors, nums=nums, dense=dense, rpn=x, heads=heads))
        return dets

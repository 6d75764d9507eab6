#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nms_gpu.py tests/test_e2e_gpu.py -m gpu -q --no-header --timeout 600 -p no:cacheprovider 2>&1 | tail -8
timeout 300 python - <<PY
import os, sys, torch
sys.path.insert(0, os.getcwd())
import det3d_b200
from det3d_b200.ops.iou3d import iou3d_utils
from det3d_b200.utils.synthetic import nms_boxes_xyxyr
for clustered in (False, True):
    b, s = nms_boxes_xyxyr(100000, 0, clustered=clustered)
    bt, st = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
    k = iou3d_utils.nms_gpu(bt, st, 0.2); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): k = iou3d_utils.nms_gpu(bt, st, 0.2)
    e.record(); torch.cuda.synchronize()
    print("100k boxes clustered=%s: %.1f ms, kept %d" % (clustered, a.elapsed_time(e) / 3, k.numel()))
PY

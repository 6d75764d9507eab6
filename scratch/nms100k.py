import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import det3d_b200
from det3d_b200.ops.iou3d import iou3d_utils
from det3d_b200.utils.synthetic import nms_boxes_xyxyr
b, s = nms_boxes_xyxyr(100000, 0)
bt, st = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
for _ in range(2):
    k = iou3d_utils.nms_gpu(bt, st, 0.2)
torch.cuda.synchronize()
print("kept", k.numel())

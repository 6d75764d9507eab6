"""det3d.ops.nms.nms_gpu: the numba.cuda NMS / rotated-IoU entry points of the reference
(det3d/ops/nms/nms_gpu.py), same names, numpy in / numpy out, backed by csrc/nms.cu.

  nms_gpu(dets [N,5] x1,y1,x2,y2,score, thresh)            :129-166  "+1" axis-aligned IoU, `>`
  rotate_nms_gpu(dets [N,6] cx,cy,w,l,r,score, thresh)     :453-496  RRPN rotated IoU, `>`
  rotate_iou_gpu(boxes [N,5], query_boxes [K,5])           :541-582  -> [N,K]
  rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1)    :643-669  -> [N,K]
The reference copies host -> device -> host per call (:486-493); so do these (they are host-API
functions), but mask, sweep and keep list stay on the device in between.
"""
import numpy as np
import torch

from ... import _lib
from . import nms_ops


def _device(device_id):
    if not torch.cuda.is_available():
        raise RuntimeError("det3d_b200: nms_gpu needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", int(device_id))


def _sorted_nms(dets, score_col, fmt, thresh, device_id, axis_aligned=False):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.shape[0] == 0:
        return []
    order = dets[:, score_col].argsort()[::-1].astype(np.int32)        # nms_gpu.py:139-141 / :467-469
    dev = _device(device_id)
    boxes = torch.from_numpy(np.ascontiguousarray(dets[order][:, :5])).to(dev)
    with torch.cuda.device(dev):
        keep_idx, keep_count = nms_ops.nms_sorted(boxes, fmt, float(thresh), axis_aligned=axis_aligned,
                                                  aa_mode=_lib.AA_PIXEL)
        k = int(keep_count.item())
        keep = keep_idx[:k].cpu().numpy()
    return list(order[keep])


def nms_gpu(dets, nms_overlap_thresh, device_id=0):
    return _sorted_nms(dets, 4, _lib.BOX_XYXYR, nms_overlap_thresh, device_id, axis_aligned=True)


def rotate_nms_gpu(dets, nms_overlap_thresh, device_id=0):
    return _sorted_nms(dets, 5, _lib.BOX_XYWLR_RRPN, nms_overlap_thresh, device_id)


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0):
    box_dtype = boxes.dtype
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    query_boxes = np.ascontiguousarray(query_boxes, dtype=np.float32)
    n, k = boxes.shape[0], query_boxes.shape[0]
    if n == 0 or k == 0:
        return np.zeros((n, k), dtype=np.float32)
    dev = _device(device_id)
    with torch.cuda.device(dev):
        b = torch.from_numpy(boxes).to(dev)
        q = torch.from_numpy(query_boxes).to(dev)
        out = torch.empty((n, k), dtype=torch.float32, device=dev)
        st = _lib.lib().d3b_rotate_iou_rrpn(b.data_ptr(), n, q.data_ptr(), k, int(criterion), out.data_ptr(),
                                            _lib.current_stream())
        _lib.check(st, "d3b_rotate_iou_rrpn")
        return out.cpu().numpy().astype(box_dtype)


def rotate_iou_gpu(boxes, query_boxes, device_id=0):
    return rotate_iou_gpu_eval(boxes, query_boxes, -1, device_id)

"""Python surface of the reference's iou3d extension, on the det3d_b200 kernels.

Same names, arguments and return values as det3d/ops/iou3d/iou3d_utils.py:7-90
(boxes [N,5] = x1,y1,x2,y2,ry on the GPU):
  boxes_iou_bev(a, b) -> [M,N] f32;  nms_gpu / nms_normal_gpu(boxes, scores, thresh)
  -> LongTensor of kept indices into `boxes`, descending score, on the device.
The greedy sweep runs on the device; nothing is copied to the host except the
final count needed to size the returned tensor.
"""
import torch

from ... import _lib
from ..nms import nms_ops


def boxes_iou_bev(boxes_a, boxes_b):
    return nms_ops.boxes_iou_bev(boxes_a, boxes_b, mode=0)


def boxes_overlap_bev(boxes_a, boxes_b):
    return nms_ops.boxes_iou_bev(boxes_a, boxes_b, mode=1)


def _nms(boxes, scores, thresh, axis_aligned):
    order = scores.sort(0, descending=True)[1]
    sorted_boxes = boxes.float()[order].contiguous()
    keep_idx, keep_count = nms_ops.nms_sorted(sorted_boxes, _lib.BOX_XYXYR, thresh, None,
                                              axis_aligned=axis_aligned)
    return order[keep_idx[: int(keep_count.item())]].contiguous()


def nms_gpu(boxes, scores, thresh):
    return _nms(boxes, scores, thresh, False)


def nms_normal_gpu(boxes, scores, thresh):
    return _nms(boxes, scores, thresh, True)

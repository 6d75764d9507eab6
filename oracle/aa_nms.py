"""TEST INFRASTRUCTURE -- CPU restatement of the axis-aligned "+1" NMS.

Follows det3d/ops/nms/nms_gpu.py: `iou_device` :22-33 (fp32 differences, then float64 because the integer
literal promotes them under numba typing), bitmask `nms_kernel` :67-102 (`>` thresh, upper triangle),
`nms_postprocess` :110-127, sort `nms_gpu` :139-141; caller det3d/core/bbox/box_torch_ops.py:506-525.
Pinned by tests/golden/aa_nms_pixel_700.npz (the reference's own source compiled for the CPU target,
tests/golden/make_golden_aa_nms.py).  Never imported by det3d_b200.
"""
import numpy as np


def iou_pixel_matrix(boxes):
    """[N,4] f32 -> [N,N] float64 IoU with +1 extents."""
    b = np.asarray(boxes, np.float32)
    left = np.maximum(b[:, None, 0], b[None, :, 0])
    right = np.minimum(b[:, None, 2], b[None, :, 2])
    top = np.maximum(b[:, None, 1], b[None, :, 1])
    bottom = np.minimum(b[:, None, 3], b[None, :, 3])
    width = np.maximum((right - left).astype(np.float64) + 1.0, 0.0)      # fp32 subtraction, then float64
    height = np.maximum((bottom - top).astype(np.float64) + 1.0, 0.0)
    inter = width * height
    area = ((b[:, 2] - b[:, 0]).astype(np.float64) + 1.0) * ((b[:, 3] - b[:, 1]).astype(np.float64) + 1.0)
    return inter / (area[:, None] + area[None, :] - inter)


def nms_pixel(dets, thresh):
    """dets [N,5] = x1,y1,x2,y2,score -> kept indices into dets, descending score (nms_gpu :129-166)."""
    dets = np.asarray(dets, np.float32)
    order = dets[:, 4].argsort()[::-1]
    iou = iou_pixel_matrix(dets[order, :4])
    hit = iou > np.float64(np.float32(thresh))
    n = dets.shape[0]
    removed = np.zeros(n, bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        removed[i + 1:] |= hit[i, i + 1:]
    return order[np.asarray(keep, np.int64)].astype(np.int64)

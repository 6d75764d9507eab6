"""Torch box ops on the inference path.

Mirrors the call contracts of det3d/core/bbox/box_torch_ops.py:
  second_box_decode (:80-148), center_to_corner_box2d / corner_to_standup_nd,
  nms (:506-525) and rotate_nms (:528-549).
`rotate_nms` / `nms` keep the reference's signature and result (LongTensor of
indices into the input, descending score, on the input's device) but run
entirely on the GPU through det3d_b200.ops.nms -- no D2H copy, no CPU NMS.
"""
import torch

from det3d_b200.ops.nms import nms_ops


def second_box_decode(box_encodings, anchors, encode_angle_to_vector=False, bin_loss=False,
                      smooth_dim=False, norm_velo=False):
    """Residual decode for SECOND-style anchors; boxes [.., 7|9]: x,y,z,w,l,h,(vx,vy),r."""
    nd = anchors.shape[-1]
    a = torch.unbind(anchors, dim=-1)
    t = torch.unbind(box_encodings, dim=-1)
    if nd == 9:
        xa, ya, za, wa, la, ha, vxa, vya, ra = a
    elif nd == 7:
        xa, ya, za, wa, la, ha, ra = a
    else:
        raise ValueError("anchors must have 7 or 9 columns")
    xt, yt, zt, wt, lt, ht = t[:6]
    diagonal = torch.sqrt(la ** 2 + wa ** 2)
    xg = xt * diagonal + xa
    yg = yt * diagonal + ya
    zg = zt * ha + za
    if smooth_dim:
        lg, wg, hg = (lt + 1) * la, (wt + 1) * wa, (ht + 1) * ha
    else:
        lg, wg, hg = torch.exp(lt) * la, torch.exp(wt) * wa, torch.exp(ht) * ha
    out = [xg, yg, zg, wg, lg, hg]
    rest = t[6:]
    if nd == 9:
        vxt, vyt = rest[0], rest[1]
        rest = rest[2:]
        if norm_velo:
            out += [vxt * diagonal + vxa, vyt * diagonal + vya]
        else:
            out += [vxt + vxa, vyt + vya]
    if encode_angle_to_vector:
        rtx, rty = rest[0], rest[1]
        rg = torch.atan2(rty + torch.sin(ra), rtx + torch.cos(ra))
    else:
        rg = rest[0] + ra
    out.append(rg)
    return torch.stack(out, dim=-1)


def corners_nd(dims, origin=0.5):
    ndim = int(dims.shape[1])
    if ndim != 2:
        raise NotImplementedError("only the 2-D case is on the inference path")
    norm = torch.tensor([[0, 0], [0, 1], [1, 1], [1, 0]], dtype=dims.dtype, device=dims.device) - origin
    return dims.view(-1, 1, ndim) * norm.view(1, 4, ndim)


def rotation_2d(points, angles):
    s, c = torch.sin(angles), torch.cos(angles)
    x, y = points[..., 0], points[..., 1]
    return torch.stack([x * c[:, None] + y * s[:, None], -x * s[:, None] + y * c[:, None]], dim=-1)


def center_to_corner_box2d(centers, dims, angles=None, origin=0.5):
    corners = corners_nd(dims, origin=origin)
    if angles is not None:
        corners = rotation_2d(corners, angles)
    return corners + centers.view(-1, 1, 2)


def corner_to_standup_nd(boxes_corner):
    return torch.cat([boxes_corner.min(dim=1)[0], boxes_corner.max(dim=1)[0]], dim=1)


def _topk_prefix(boxes, scores, pre_max_size):
    indices = None
    if pre_max_size is not None:
        k = min(scores.shape[0], pre_max_size)
        scores, indices = torch.topk(scores, k=k)
        boxes = boxes[indices]
    return boxes, scores, indices


def rotate_nms(rbboxes, scores, pre_max_size=None, post_max_size=None, iou_threshold=0.5):
    """rbboxes [N,5] (x, y, w, l, r); semantics of rotate_nms_cc (nms_cpu.py:34-45)."""
    rbboxes, scores, indices = _topk_prefix(rbboxes, scores, pre_max_size)
    if rbboxes.shape[0] == 0:
        return torch.zeros([0], dtype=torch.long, device=rbboxes.device)
    keep = nms_ops.rotate_nms_xywlr(rbboxes, scores, iou_threshold, post_max_size)
    if keep.shape[0] == 0:
        return torch.zeros([0], dtype=torch.long, device=rbboxes.device)
    return keep if indices is None else indices[keep]


def nms(bboxes, scores, pre_max_size=None, post_max_size=None, iou_threshold=0.5):
    """bboxes [N,4] (x1, y1, x2, y2) axis-aligned; the "+1" pixel IoU of the numba nms_gpu the reference calls
    here (box_torch_ops.py:506-525 -> ops/nms/nms_gpu.py:22-33,129-166)."""
    bboxes, scores, indices = _topk_prefix(bboxes, scores, pre_max_size)
    if bboxes.shape[0] == 0:
        return torch.zeros([0], dtype=torch.long, device=bboxes.device)
    keep = nms_ops.normal_nms_xyxy(bboxes, scores, iou_threshold, post_max_size, pixel=True)
    if keep.shape[0] == 0:
        return torch.zeros([0], dtype=torch.long, device=bboxes.device)
    return keep if indices is None else indices[keep]

"""TEST INFRASTRUCTURE -- CPU restatement of MultiGroupHead.predict for any Det3D head config.

Follows det3d/models/bbox_heads/mg_head.py:697-1085 (predict / get_task_detections,
use_multi_class_nms=False branch), det3d/core/bbox/box_torch_ops.py:80-148 (second_box_decode,
7- and 9-dim boxes, optional angle-vector encoding) and :528-549 (rotate_nms ->
oracle.second_cpu.rotate_nms -> the C restatement of rotate_nms_cc).
"""
import numpy as np
import torch

from .second_cpu import rotate_nms


def second_box_decode(enc, anchors, encode_angle_to_vector=False, smooth_dim=False, norm_velo=False):
    nd = anchors.shape[-1]
    if nd == 9:
        xa, ya, za, wa, la, ha, vxa, vya, ra = torch.split(anchors, 1, dim=-1)
        if encode_angle_to_vector:
            xt, yt, zt, wt, lt, ht, vxt, vyt, rtx, rty = torch.split(enc, 1, dim=-1)
        else:
            xt, yt, zt, wt, lt, ht, vxt, vyt, rt = torch.split(enc, 1, dim=-1)
    else:
        xa, ya, za, wa, la, ha, ra = torch.split(anchors, 1, dim=-1)
        if encode_angle_to_vector:
            xt, yt, zt, wt, lt, ht, rtx, rty = torch.split(enc, 1, dim=-1)
        else:
            xt, yt, zt, wt, lt, ht, rt = torch.split(enc, 1, dim=-1)
    diagonal = torch.sqrt(la ** 2 + wa ** 2)
    ret = [xt * diagonal + xa, yt * diagonal + ya, zt * ha + za]
    if smooth_dim:
        ret += [(wt + 1) * wa, (lt + 1) * la, (ht + 1) * ha]
    else:
        ret += [torch.exp(wt) * wa, torch.exp(lt) * la, torch.exp(ht) * ha]
    if encode_angle_to_vector:
        rg = torch.atan2(rty + torch.sin(ra), rtx + torch.cos(ra))
    else:
        rg = rt + ra
    if nd > 7:
        ret += [vxt * diagonal + vxa, vyt * diagonal + vya] if norm_velo else [vxt + vxa, vyt + vya]
    ret.append(rg)
    return torch.cat(ret, dim=-1)


def predict_sample_task(cls_logits, box_enc, dir_logits, anchors, test_cfg, vec_encode, direction_offset=0.0):
    """One sample, one task (CPU tensors) -> (boxes [K,nd], scores [K], labels [K])."""
    nd = anchors.shape[-1]
    box_preds = second_box_decode(box_enc.float(), anchors.float(), vec_encode)
    total = torch.sigmoid(cls_logits.float())
    if total.shape[-1] == 1:
        top_scores, top_labels = total.squeeze(-1), torch.zeros(total.shape[0], dtype=torch.long)
    else:
        top_scores, top_labels = torch.max(total, dim=-1)
    thr = test_cfg["score_threshold"]
    keep = top_scores >= thr
    top_scores = top_scores[keep]
    if top_scores.shape[0] == 0:
        return torch.zeros((0, nd)), torch.zeros(0), torch.zeros(0, dtype=torch.long)
    box_preds, top_labels = box_preds[keep], top_labels[keep]
    dir_labels = torch.max(dir_logits, dim=-1)[1][keep] if dir_logits is not None else None
    nms = test_cfg["nms"]
    sel = rotate_nms(box_preds[:, [0, 1, 3, 4, nd - 1]], top_scores, nms["nms_pre_max_size"], nms["nms_post_max_size"],
                     nms["nms_iou_threshold"])
    bx, sc, lb = box_preds[sel].clone(), top_scores[sel], top_labels[sel]
    if dir_labels is not None and bx.shape[0]:
        opp = ((bx[..., -1] - direction_offset) > 0) ^ dir_labels[sel].bool()
        bx[..., -1] += torch.where(opp, torch.tensor(np.pi).type_as(bx), torch.tensor(0.0).type_as(bx))
    rng = torch.tensor(test_cfg["post_center_limit_range"], dtype=torch.float32)
    if bx.shape[0]:
        m = (bx[:, :3] >= rng[:3]).all(1) & (bx[:, :3] <= rng[3:]).all(1)
        bx, sc, lb = bx[m], sc[m], lb[m]
    return bx, sc, lb

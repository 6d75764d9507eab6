"""Golden vectors for the axis-aligned "+1" NMS behind box_torch_ops.nms.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden_aa_nms.py
The reference implementation is numba.cuda (det3d/ops/nms/nms_gpu.py:22-166) and cannot launch without a GPU /
its compiled `nms` extension, so its *own source text* is compiled for the CPU target instead: `iou_device`
(:22-33) with the decorator swapped for numba.njit (same typing rules: the integer literal promotes the fp32
differences to float64), the bitmask loop of `nms_kernel` (:67-102) driven row by row, and `nms_postprocess`
(:110-127) exec'd unchanged.  Sorting follows `nms_gpu` (:139-141).
Output: aa_nms_pixel_700.npz.
"""
import os
import re

import numba
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = open("/root/reference/det3d/ops/nms/nms_gpu.py").read()


def _grab(name):
    m = re.search(r"(@[^\n]*\n)def %s\(.*?(?=\n@|\ndef |\Z)" % name, SRC, re.S)
    return m.group(0)


ns = {"numba": numba, "np": np}
exec(_grab("iou_device").replace('@cuda.jit("(float32[:], float32[:])", device=True, inline=True)', "@numba.njit"), ns)
exec(_grab("div_up"), ns)
exec(_grab("nms_postprocess"), ns)
iou_device, nms_postprocess = ns["iou_device"], ns["nms_postprocess"]


@numba.njit
def build_mask(boxes, thresh, mask):
    """nms_kernel (:67-102): bit i of mask[row, colblock] when iou(row, col) > thresh, upper triangle in the
    diagonal block."""
    n = boxes.shape[0]
    col_blocks = (n + 63) // 64
    for cur in range(n):
        row_start = cur // 64
        tx = cur % 64
        for col_start in range(col_blocks):
            col_size = min(n - col_start * 64, 64)
            t = np.uint64(0)
            start = 0
            if row_start == col_start:
                start = tx + 1
            for i in range(start, col_size):
                if iou_device(boxes[cur, :4], boxes[col_start * 64 + i, :4]) > thresh:
                    t |= np.uint64(1) << np.uint64(i)
            mask[cur * col_blocks + col_start] = t


def reference_nms(dets, thresh):
    n = dets.shape[0]
    order = dets[:, 4].argsort()[::-1].astype(np.int32)
    boxes = np.ascontiguousarray(dets[order])
    mask = np.zeros(n * ((n + 63) // 64), np.uint64)
    build_mask(boxes, np.float32(thresh), mask)
    keep = np.zeros(n, np.int32)
    k = nms_postprocess(keep, mask, n)
    return order[keep[:k]].astype(np.int64)


def main():
    rng = np.random.default_rng(3)
    n = 700
    ctr = rng.uniform(0, 400, (n, 2))
    ctr[:300] = ctr[rng.integers(300, 700, 300)] + rng.normal(0, 1.5, (300, 2))     # clusters: real suppression
    wh = rng.uniform(0.5, 30, (n, 2))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
    boxes[:40] = np.round(boxes[:40])                                               # integer "pixel" boxes
    boxes[40:50, 2:] = boxes[40:50, :2]                                             # zero-size boxes: area 1 with the +1
    scores = (rng.permutation(n).astype(np.float32) + 1) / n
    dets = np.concatenate([boxes, scores[:, None]], 1).astype(np.float32)
    out = dict(dets=dets)
    for thr in (0.3, 0.5, 0.7):
        out["keep_t%02d" % int(thr * 100)] = reference_nms(dets, thr)
    idx = rng.integers(0, n, (256, 2))
    out["pair_idx"] = idx
    out["pair_iou"] = np.array([iou_device(boxes[a], boxes[b]) for a, b in idx], np.float64)
    np.savez_compressed(os.path.join(HERE, "aa_nms_pixel_700.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()}, "kept", [out[k].size for k in out if k.startswith("keep")])


if __name__ == "__main__":
    main()

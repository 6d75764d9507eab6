"""Golden vectors for the RRPN rotated IoU / NMS, from the reference's own source.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden_rrpn.py
det3d/ops/nms/nms_gpu.py:180-470 is numba.cuda device code; it cannot launch here (no GPU, and the module's
header imports a compiled extension), so the *source text* of its device functions is compiled for the CPU
target: `@cuda.jit(...)` -> `@numba.njit(error_model="numpy")` (IEEE division like the GPU), `cuda.local.array(shape, dtype=numba.float32)` ->
`np.empty(shape, np.float32)`; nothing else changes, so numba's typing (float32 arithmetic, float64 where a
literal promotes) is the reference's.  libm vs libdevice sin/cos/sqrt may differ in the last ulp: consumers
compare IoUs to 1e-5 and keep lists modulo pairs within 1e-5 of the threshold.
Output: rrpn_600.npz.
"""
import math
import os
import re

import numba
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = open("/root/reference/det3d/ops/nms/nms_gpu.py").read()
NAMES = ["trangle_area", "area", "sort_vertex_in_convex_polygon", "line_segment_intersection",
         "point_in_quadrilateral", "quadrilateral_intersection", "rbbox_to_corners", "inter", "devRotateIoU",
         "devRotateIoUEval", "div_up", "nms_postprocess"]


def compile_reference():
    ns = {"numba": numba, "np": np, "math": math}
    for name in NAMES:
        m = re.search(r"@(?:cuda|numba)\.jit\((?:[^()]|\([^()]*\))*\)\s*\ndef %s\(.*?(?=\n@|\ndef |\Z)" % name, SRC, re.S)
        text = m.group(0)
        text = re.sub(r"@cuda\.jit\((?:[^()]|\([^()]*\))*\)", "@numba.njit(error_model=\"numpy\")", text)
        text = re.sub(r"cuda\.local\.array\((\([^)]*\)), dtype=numba\.float32\)", r"np.empty(\1, np.float32)", text)
        exec(text, ns)
    return ns


REF = compile_reference()
dev_iou_eval = REF["devRotateIoUEval"]
dev_iou = REF["devRotateIoU"]
nms_postprocess = REF["nms_postprocess"]


@numba.njit(error_model="numpy")
def iou_matrix(boxes, query, criterion, out, f):
    for n in range(boxes.shape[0]):                     # rotate_iou_kernel_eval: f(query[k], boxes[n])
        for k in range(query.shape[0]):
            out[n, k] = f(query[k], boxes[n], criterion)


@numba.njit(error_model="numpy")
def build_mask(boxes, thresh, mask, f):
    n = boxes.shape[0]
    col_blocks = (n + 63) // 64
    for cur in range(n):                                # rotate_nms_kernel :411-450
        row_start, tx = cur // 64, cur % 64
        for col_start in range(col_blocks):
            col_size = min(n - col_start * 64, 64)
            t = np.uint64(0)
            start = tx + 1 if row_start == col_start else 0
            for i in range(start, col_size):
                if f(boxes[cur, :5], boxes[col_start * 64 + i, :5]) > thresh:
                    t |= np.uint64(1) << np.uint64(i)
            mask[cur * col_blocks + col_start] = t


def reference_rotate_nms(dets, thresh):
    dets = dets.astype(np.float32)
    n = dets.shape[0]
    order = dets[:, 5].argsort()[::-1].astype(np.int32)
    boxes = np.ascontiguousarray(dets[order])
    mask = np.zeros(n * ((n + 63) // 64), np.uint64)
    build_mask(boxes, np.float32(thresh), mask, dev_iou)
    keep = np.zeros(n, np.int32)
    k = nms_postprocess(keep, mask, n)
    return order[keep[:k]].astype(np.int64)


def main():
    rng = np.random.default_rng(17)
    n = 600
    ctr = rng.uniform(0, 60, (n, 2))
    ctr[:250] = ctr[rng.integers(250, n, 250)] + rng.normal(0, 0.6, (250, 2))
    dets = np.concatenate([ctr, rng.uniform(1.5, 2.5, (n, 1)), rng.uniform(3.5, 5.0, (n, 1)),
                           rng.uniform(-np.pi, np.pi, (n, 1)), ((rng.permutation(n) + 1.0) / n)[:, None]], 1).astype(np.float32)
    special = np.array([[0, 0, 2, 2, 0], [0, 0, 2, 2, np.pi / 4], [0, 0, 2, 2, 0], [10, 10, 2, 4, 0.3],
                        [1, 0, 2, 2, 0], [0, 0, 1, 1, 0.1], [0, 0, 4, 4, 0], [2, 0, 2, 2, 0], [0.5, 0.5, 3, 1, 1.2]], np.float32)
    a = np.concatenate([special, dets[:51, :5]]).astype(np.float32)
    out = dict(dets=dets, mat_boxes=a)
    for crit in (-1, 0, 1, 2):
        m = np.zeros((a.shape[0], a.shape[0]), np.float32)
        iou_matrix(a, a, crit, m, dev_iou_eval)
        out["mat_c%d" % (crit + 1)] = m
    for thr in (0.1, 0.3, 0.5):
        out["keep_t%02d" % int(thr * 100)] = reference_rotate_nms(dets, thr)
    np.savez_compressed(os.path.join(HERE, "rrpn_600.npz"), **out)
    print("iou[0,1] (square vs 45deg square) =", out["mat_c0"][0, 1], "expected", 1 / math.sqrt(2),
          "kept", [out[k].size for k in out if k.startswith("keep")])


if __name__ == "__main__":
    main()

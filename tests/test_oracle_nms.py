"""Known-answer tests for the rotated IoU / NMS oracles (none exist in the reference)."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import build as obuild
from oracle.second_cpu import rotate_nms_cc


@pytest.fixture(scope="module")
def lib():
    l = C.CDLL(obuild.build())
    for f in ("oracle_iou_bev", "oracle_box_overlap", "oracle_iou_normal", "oracle_rotate_iou_xywlr"):
        getattr(l, f).restype = C.c_float
        getattr(l, f).argtypes = [C.c_void_p, C.c_void_p]
    l.oracle_nms_xyxyr.restype = C.c_int64
    l.oracle_nms_xyxyr.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_void_p]
    return l


def _f(l, name, a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return getattr(l, name)(a.ctypes.data, b.ctypes.data)


def test_identical_boxes(lib):
    b = [0, 0, 2, 4, 0.3]
    assert abs(_f(lib, "oracle_iou_bev", b, b) - 1.0) < 1e-5


def test_disjoint_boxes(lib):
    assert _f(lib, "oracle_iou_bev", [0, 0, 1, 1, 0.2], [5, 5, 6, 6, 1.0]) == 0.0
    assert _f(lib, "oracle_box_overlap", [0, 0, 1, 1, 0.2], [5, 5, 6, 6, 1.0]) == 0.0


def test_square_vs_45deg_square(lib):
    # unit square vs itself rotated 45 deg about the centre: intersection is a regular octagon of
    # area 2(sqrt2-1); IoU = 1/sqrt2.
    a, b = [0, 0, 1, 1, 0.0], [0, 0, 1, 1, math.pi / 4]
    assert abs(_f(lib, "oracle_box_overlap", a, b) - 2 * (math.sqrt(2) - 1)) < 1e-5
    assert abs(_f(lib, "oracle_iou_bev", a, b) - 1 / math.sqrt(2)) < 1e-5


def test_containment_and_half_overlap(lib):
    assert abs(_f(lib, "oracle_iou_bev", [0, 0, 4, 4, 0], [1, 1, 3, 3, 0]) - 0.25) < 1e-5
    assert abs(_f(lib, "oracle_iou_bev", [0, 0, 2, 2, 0], [1, 0, 3, 2, 0]) - 1 / 3) < 1e-5
    assert abs(_f(lib, "oracle_iou_normal", [0, 0, 2, 2, 0], [1, 0, 3, 2, 0]) - 1 / 3) < 1e-6


def test_zero_area_box_hits_eps_clamp(lib):
    v = _f(lib, "oracle_iou_bev", [1, 1, 1, 1, 0], [1, 1, 1, 1, 0])
    assert v == 0.0  # 0 / max(0, 1e-8)


def test_xywlr_iou_matches_xyxyr_geometry(lib):
    # same rectangles in both parameterisations (rotation sign conventions differ: both rotate by
    # the stored angle with x' = x cos + y sin) -> same IoU up to rounding
    rng = np.random.default_rng(0)
    for _ in range(200):
        c1, c2 = rng.uniform(0, 3, 2), rng.uniform(0, 3, 2)
        w1, l1, w2, l2 = rng.uniform(1, 3, 4)
        r1, r2 = rng.uniform(-3, 3, 2)
        a = [c1[0] - w1 / 2, c1[1] - l1 / 2, c1[0] + w1 / 2, c1[1] + l1 / 2, r1]
        b = [c2[0] - w2 / 2, c2[1] - l2 / 2, c2[0] + w2 / 2, c2[1] + l2 / 2, r2]
        i1 = _f(lib, "oracle_iou_bev", a, b)
        i2 = _f(lib, "oracle_rotate_iou_xywlr", [c1[0], c1[1], w1, l1, r1], [c2[0], c2[1], w2, l2, r2])
        assert abs(i1 - i2) < 2e-4


def test_greedy_sweep_small(lib):
    boxes = np.array([[0, 0, 2, 2, 0], [0.1, 0, 2.1, 2, 0], [5, 5, 7, 7, 0], [5, 5.1, 7, 7.1, 0.1], [10, 0, 11, 1, 0]], np.float32)
    keep = np.zeros(5, np.int64)
    k = lib.oracle_nms_xyxyr(boxes.ctypes.data, 5, 0.5, 1, keep.ctypes.data)
    assert keep[:k].tolist() == [0, 2, 4]


def test_rotate_nms_cc_threshold_is_inclusive():
    # two identical boxes: IoU = 1 >= 1.0 -> suppressed (nms_cpu.h:157 uses >=)
    dets = np.array([[0, 0, 2, 4, 0.3, 0.9], [0, 0, 2, 4, 0.3, 0.8], [10, 10, 2, 4, 0.0, 0.7]], np.float32)
    assert rotate_nms_cc(dets, 1.0).tolist() == [0, 2]
    # touching boxes (zero overlap) are never suppressed even at thresh 0: the hull test skips them
    dets = np.array([[0, 0, 2, 2, 0, 0.9], [2, 0, 2, 2, 0, 0.8]], np.float32)
    assert rotate_nms_cc(dets, 0.0).tolist() == [0, 1]


def test_axis_aligned_pixel_nms_oracle_matches_reference_golden():
    """a14: the "+1" IoU of numba nms_gpu; golden = the reference's own source compiled for the CPU target."""
    from conftest import load_golden
    from oracle.aa_nms import iou_pixel_matrix, nms_pixel
    g = load_golden("aa_nms_pixel_700")
    iou = iou_pixel_matrix(g["dets"][:, :4])
    assert np.array_equal(iou[g["pair_idx"][:, 0], g["pair_idx"][:, 1]], g["pair_iou"])      # bit-exact float64
    for thr in (0.3, 0.5, 0.7):
        assert np.array_equal(nms_pixel(g["dets"], thr), g["keep_t%02d" % int(thr * 100)])


def test_rrpn_oracle_matches_reference_golden():
    """a13: the numba RRPN rotated IoU / NMS; golden = the reference source compiled for the CPU target."""
    import math
    from conftest import load_golden
    from oracle import rrpn
    g = load_golden("rrpn_600")
    a = g["mat_boxes"]
    for crit in (-1, 0, 1, 2):
        got = rrpn.rotate_iou_eval(a, a, crit)
        want = g["mat_c%d" % (crit + 1)]
        both = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), both)
        assert float(np.abs(got[both] - want[both]).max()) <= 1e-5
    iou = rrpn.rotate_iou_eval(a, a, -1)
    assert abs(iou[0, 1] - 1 / math.sqrt(2)) < 1e-6          # unit square vs itself rotated 45 deg (SURVEY 8c KAT)
    assert abs(iou[0, 2] - 1.0) < 1e-6 and iou[0, 3] == 0.0    # identical, disjoint
    for thr in (0.1, 0.3, 0.5):
        keep, near = rrpn.rotate_nms(g["dets"], thr)
        want = g["keep_t%02d" % int(thr * 100)]
        if near == 0:
            assert np.array_equal(keep, want)
        else:
            assert len(set(keep.tolist()) ^ set(want.tolist())) <= 2 * near


def test_oracle_anchor_decode_pinned_to_reference_golden():
    """oracle/predict_cpu.second_box_decode (the decode of the CPU predict restatement) against the reference function
    executed through the reference coder's call (tests/golden/make_golden_decode.py, box_torch_ops.py:79-148)."""
    import os
    import torch
    from oracle.predict_cpu import second_box_decode
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode_coder.npz"))
    for nd in (7, 9):
        for vec in (False, True):
            key = "nd%d_vec%d_lin0" % (nd, int(vec))
            got = second_box_decode(torch.from_numpy(g[key + "_enc"]), torch.from_numpy(g[key + "_anchors"]), vec).numpy()
            assert np.allclose(got, g[key + "_out"], rtol=0, atol=1e-6), key

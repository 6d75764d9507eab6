"""TEST INFRASTRUCTURE -- CPU restatement of the reference's nuScenes multi-sweep loading
(det3d/datasets/pipelines/loading.py: read_file :17-31, remove_close :34-43, read_sweep :46-64, merge :98-124).
Pinned by tests/golden/ingest_nusc_3sweeps.npz (the reference code itself, tests/golden/make_golden_ingest.py).
Never imported by det3d_b200."""
import numpy as np


def remove_close(points, radius):
    """points [C, n] -> columns whose |x| >= radius or |y| >= radius (:39-42)."""
    x_filt = np.abs(points[0, :]) < radius
    y_filt = np.abs(points[1, :]) < radius
    return points[:, np.logical_not(np.logical_and(x_filt, y_filt))]


def merge_sweeps(raw_sweeps, transforms, time_lags, radius=1.0, n_feat=4):
    """raw_sweeps: float32 [n_s, 5] arrays, key frame first -> combined float32 [N, n_feat + 1]."""
    pts_list, t_list = [], []
    for s, raw in enumerate(raw_sweeps):
        pts = np.array(raw[:, :n_feat], np.float32)
        if s > 0:
            p = remove_close(pts.T.copy(), radius)                                        # :52
            if transforms[s] is not None:                                                 # :55-58, float64 dot
                p[:3, :] = np.asarray(transforms[s]).dot(np.vstack((p[:3, :], np.ones(p.shape[1]))))[:3, :]
            pts = p.T
        pts_list.append(pts)
        t_list.append(time_lags[s] * np.ones((pts.shape[0], 1)))                          # :62, float64
    points = np.concatenate(pts_list, axis=0)
    times = np.concatenate(t_list, axis=0).astype(points.dtype)                           # :119
    return np.hstack([points, times])

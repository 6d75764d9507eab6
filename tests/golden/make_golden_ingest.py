"""Golden vectors for the multi-sweep ingest, produced by the reference code itself.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden_ingest.py
The module det3d/datasets/pipelines/loading.py imports pycocotools and the det3d package at import time, so the
source text of `read_file`, `remove_close`, `read_sweep` and the class `LoadPointCloudFromFile` (:17-124 plus the
Lyft branch that is not exercised) is exec'd unchanged with a stub registry.  Synthetic nuScenes-style `.pcd.bin`
files (float32 [n, 5]) are written to a temporary directory; the stored golden holds the raw sweeps, the sweep
records (transform, time lag) and the reference's `combined` array for `np.random.seed(7)`.
Output: ingest_nusc_3sweeps.npz.
"""
import os
import tempfile
from pathlib import Path

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = open("/root/reference/det3d/datasets/pipelines/loading.py").read()


class _Reg:
    def register_module(self, cls):
        return cls


def reference_namespace():
    a = SRC.index("def read_file")
    b = SRC.index("@PIPELINES.register_module\nclass LoadPointCloudAnnotations")
    ns = {"np": np, "Path": Path, "PIPELINES": _Reg()}
    exec(SRC[a:b], ns)
    return ns


def rigid(rng):
    ang = rng.uniform(-0.2, 0.2)
    c, s = np.cos(ang), np.sin(ang)
    m = np.eye(4)
    m[:3, :3] = np.array([[c, -s, 0.01], [s, c, -0.02], [-0.01, 0.02, 1.0]])
    m[:3, 3] = rng.uniform(-3, 3, 3)
    return m


def main():
    ns = reference_namespace()
    rng = np.random.default_rng(21)
    tmp = tempfile.mkdtemp()
    sizes = [1800, 1500, 1701, 1400]                     # key frame + 3 candidate sweeps
    raws, paths = [], []
    for j, n in enumerate(sizes):
        pts = np.zeros((n, 5), np.float32)
        pts[:, :3] = rng.uniform([-50, -50, -4], [50, 50, 2], (n, 3))
        pts[: n // 6, :2] = rng.uniform(-1.5, 1.5, (n // 6, 2))      # around the 1 m remove_close square
        pts[5, :2] = [1.0, 0.2]                                      # |x| == radius is kept
        pts[6, :2] = [-0.999999, 0.999999]
        pts[:, 3] = rng.uniform(0, 255, n)
        pts[:, 4] = j                                                # ring index column: dropped by read_file
        rng.shuffle(pts)
        path = os.path.join(tmp, "sweep%d.pcd.bin" % j)
        np.concatenate([pts.reshape(-1), np.zeros(3, np.float32)]).astype(np.float32).tofile(path)   # trailing partial record
        raws.append(pts)
        paths.append(path)
    sweeps = [dict(lidar_path=paths[1], transform_matrix=rigid(rng), time_lag=0.05),
              dict(lidar_path=paths[2], transform_matrix=None, time_lag=0.0),       # padding entry, nusc_common.py:428-433
              dict(lidar_path=paths[3], transform_matrix=rigid(rng), time_lag=0.1499999)]
    info = dict(lidar_path=paths[0], sweeps=sweeps)
    res = dict(lidar=dict(nsweeps=3), metadata={})
    np.random.seed(7)
    ns["LoadPointCloudFromFile"](dataset="NuScenesDataset")(res, info)
    np.random.seed(7)
    chosen = np.random.choice(3, 2, replace=False)
    out = dict(combined=res["lidar"]["combined"], points=res["lidar"]["points"], times=res["lidar"]["times"], chosen=chosen)
    for j in range(4):
        out["raw%d" % j] = raws[j]
    out["tm0"], out["tm2"] = sweeps[0]["transform_matrix"], sweeps[2]["transform_matrix"]
    out["lags"] = np.array([s["time_lag"] for s in sweeps], np.float64)
    np.savez_compressed(os.path.join(HERE, "ingest_nusc_3sweeps.npz"), **out)
    print("combined", out["combined"].shape, out["combined"].dtype, "chosen", chosen, "raw points", sum(sizes[:1]) + sum(sizes[int(c) + 1] for c in chosen))


if __name__ == "__main__":
    main()

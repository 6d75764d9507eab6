"""Generates the committed golden fixtures from the REFERENCE implementation itself.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
* voxel_*.npz   : inputs + outputs of the reference numba voxelizer
                  (det3d/ops/point_cloud/point_cloud_ops.py:112-184), loaded by file path
                  (importing the det3d package would pull spconv/pycocotools).
* anchors_*.npz : det3d/core/bbox/box_np_ops.py:733-805 create_anchors_3d_range (function
                  source exec'd in isolation; a `list(...)` shim is needed for numpy>=2).
The iou3d fixtures come from the reference CUDA kernel and are produced on the GPU box by
tests/golden/make_golden_gpu.py.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference"

from det3d_b200.utils.synthetic import lidar_like_cloud, uniform_cloud  # noqa: E402

KITTI = dict(vs=[0.05, 0.05, 0.1], pcr=[0, -40.0, -3.0, 70.4, 40.0, 1.0], max_points=5, max_voxels=20000)
NUSC = dict(vs=[0.1, 0.1, 0.2], pcr=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], max_points=10, max_voxels=60000)
PILLAR = dict(vs=[0.16, 0.16, 4.0], pcr=[0, -39.68, -3, 69.12, 39.68, 1], max_points=100, max_voxels=12000)


def load_ref_voxelizer():
    spec = importlib.util.spec_from_file_location("ref_pc_ops", REF + "/det3d/ops/point_cloud/point_cloud_ops.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def adversarial(cfgd, n, seed):
    """Boundary-stressed cloud: points on lo / hi / just outside, voxel-boundary multiples,
    duplicates, one crowded voxel (> max_points), NaN-free."""
    rng = np.random.default_rng(seed)
    vs, pcr = np.array(cfgd["vs"], np.float32), np.array(cfgd["pcr"], np.float32)
    pts = uniform_cloud(n, [pcr[0] - 1, pcr[1] - 1, pcr[2] - 0.5, pcr[3] + 1, pcr[4] + 1, pcr[5] + 0.5], 4, seed)
    k = n // 4
    cells = np.floor(rng.uniform(0, 1, (k, 3)) * ((pcr[3:] - pcr[:3]) / vs + 2) - 1).astype(np.float32)
    pts[:k, :3] = cells * vs + pcr[:3]                     # exact multiples of the voxel size
    pts[k:k + 8, :3] = pcr[:3]                             # exactly lo
    pts[k + 8:k + 16, :3] = pcr[3:]                        # exactly hi (dropped)
    pts[k + 16:k + 24, :3] = np.nextafter(pcr[:3], -np.inf)  # just below lo
    pts[k + 24:k + 24 + 3 * cfgd["max_points"], :3] = (pcr[:3] + pcr[3:]) / 2  # crowded voxel
    pts[-50:] = pts[:50]                                   # duplicates
    rng.shuffle(pts)
    return np.ascontiguousarray(pts)


def main():
    ref = load_ref_voxelizer()
    cases = {
        "kitti_uniform_1k": (KITTI, uniform_cloud(1000, KITTI["pcr"], 4, 0)),
        "kitti_lidar_1k": (KITTI, lidar_like_cloud(1000, KITTI["pcr"], 4, 0)),
        "kitti_adversarial_4k": (KITTI, adversarial(KITTI, 4000, 1)),
        "kitti_overflow_8k": (dict(KITTI, max_voxels=1500), uniform_cloud(8000, KITTI["pcr"], 4, 2)),
        "kitti_ndim3_1k": (KITTI, uniform_cloud(1000, KITTI["pcr"], 3, 3)),
        "nusc_ndim5_3k": (NUSC, lidar_like_cloud(3000, NUSC["pcr"], 5, 4)),
        "pillar_overflow_6k": (dict(PILLAR, max_voxels=900), uniform_cloud(6000, PILLAR["pcr"], 4, 5)),
        "kitti_empty": (KITTI, np.zeros((0, 4), np.float32)),
        "kitti_all_outside": (KITTI, uniform_cloud(200, [100, 100, 10, 120, 120, 12], 4, 6)),
    }
    for name, (c, pts) in cases.items():
        v, co, n = ref.points_to_voxel(pts, np.array(c["vs"], np.float32), np.array(c["pcr"], np.float32),
                                       c["max_points"], True, c["max_voxels"])
        # store voxels sparsely (they are mostly zero padding) to keep fixtures small
        np.savez_compressed(os.path.join(HERE, "voxel_%s.npz" % name), points=pts, voxel_size=np.array(c["vs"], np.float32),
                            pcr=np.array(c["pcr"], np.float32), max_points=c["max_points"], max_voxels=c["max_voxels"],
                            voxels=v, coors=co, num_points=n)
        print(name, pts.shape, "->", v.shape)

    src = open(REF + "/det3d/core/bbox/box_np_ops.py").read()
    fn = src[src.index("def create_anchors_3d_range"):src.index("def create_anchors_bev_range")]
    fn = fn.replace('indexing="ij")', 'indexing="ij")\n    rets = list(rets)')
    ns = {"np": np}
    exec(fn, ns)
    a = ns["create_anchors_3d_range"]([1, 200, 176], [0, -40.0, -1.0, 70.4, 40.0, -1.0], [1.6, 3.9, 1.56], [0, 1.57], None)
    idx = np.random.default_rng(0).choice(a.reshape(-1, 7).shape[0], 512, replace=False)
    np.savez_compressed(os.path.join(HERE, "anchors_kitti_car.npz"), sample_idx=idx, sample=a.reshape(-1, 7)[idx],
                        checksum=np.array([a.astype(np.float64).sum(), (a.astype(np.float64) ** 2).sum()]),
                        shape=np.array(a.shape))
    print("anchors", a.shape)


if __name__ == "__main__":
    main()

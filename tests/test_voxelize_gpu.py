"""CUDA voxelizer vs the oracle / the reference golden fixtures: bit-exact (integer + byte work)."""
import numpy as np
import pytest
import torch

from conftest import golden_voxel_cases, load_golden

pytestmark = pytest.mark.gpu

KITTI = dict(vs=[0.05, 0.05, 0.1], pcr=[0, -40.0, -3.0, 70.4, 40.0, 1.0])


def _gpu(points, vs, pcr, max_points, max_voxels):
    from det3d.ops.point_cloud.point_cloud_ops import points_to_voxel
    return points_to_voxel(points, vs, pcr, max_points, True, max_voxels)


def _same(a, b):
    for x, y, name in zip(a, b, ("voxels", "coors", "num_points")):
        assert x.shape == y.shape, name
        assert x.dtype == y.dtype, name
        assert np.array_equal(x, y), name


@pytest.mark.parametrize("case", golden_voxel_cases())
def test_matches_reference_golden(case):
    g = load_golden("voxel_" + case)
    out = _gpu(g["points"], g["voxel_size"], g["pcr"], int(g["max_points"]), int(g["max_voxels"]))
    _same(out, (g["voxels"], g["coors"], g["num_points"]))


@pytest.mark.parametrize("n,dist,seed", [(1000, "uniform", 0), (1000, "lidar", 0), (20000, "uniform", 1),
                                          (20000, "lidar", 2), (200000, "uniform", 3)])
def test_matches_oracle_kitti(n, dist, seed):
    from det3d_b200.utils.synthetic import lidar_like_cloud, uniform_cloud
    from oracle import voxel as ovoxel
    pts = (uniform_cloud if dist == "uniform" else lidar_like_cloud)(n, KITTI["pcr"], 4, seed)
    _same(_gpu(pts, KITTI["vs"], KITTI["pcr"], 5, 20000), ovoxel.points_to_voxel(pts, KITTI["vs"], KITTI["pcr"], 5, True, 20000))


def test_boundary_stress_and_overflow():
    from oracle import voxel as ovoxel
    rng = np.random.default_rng(11)
    vs, pcr = np.float32(KITTI["vs"]), np.float32(KITTI["pcr"])
    n = 60000
    pts = np.stack([rng.uniform(-1, 71.5, n), rng.uniform(-41, 41, n), rng.uniform(-3.3, 1.3, n), rng.uniform(0, 1, n)], 1).astype(np.float32)
    pts[::2, :3] = (np.round((pts[::2, :3] - pcr[:3]) / vs) * vs + pcr[:3]).astype(np.float32)   # snapped to cell edges
    pts[5::97] = pts[4::97][: len(pts[5::97])]                                                    # duplicates
    for max_voxels, max_points in ((20000, 5), (1234, 3), (1, 1), (60000, 35)):
        _same(_gpu(pts, vs, pcr, max_points, max_voxels), ovoxel.points_to_voxel(pts, vs, pcr, max_points, True, max_voxels))


def test_crowded_voxel_keeps_first_points_in_input_order():
    pts = np.zeros((5000, 4), np.float32)
    pts[:, :3] = [10.01, 0.01, -1.01]
    pts[:, 3] = np.arange(5000)
    v, c, n = _gpu(pts, KITTI["vs"], KITTI["pcr"], 5, 20000)
    assert n.tolist() == [5] and v[0, :, 3].tolist() == [0, 1, 2, 3, 4]
    v, c, n = _gpu(pts, [0.16, 0.16, 4], [0, -39.68, -3, 69.12, 39.68, 1], 100, 12000)
    assert n.tolist() == [100] and v[0, :, 3].tolist() == list(range(100))


def test_nan_and_inf_points_are_dropped():
    pts = np.array([[1, 1, 0, 0.5], [np.nan, 1, 0, 0.1], [1, np.inf, 0, 0.2], [2, 2, -np.inf, 0.3], [3, 3, 0, 0.4]], np.float32)
    v, c, n = _gpu(pts, KITTI["vs"], KITTI["pcr"], 5, 100)
    assert v.shape[0] == 2 and v[:, 0, 3].tolist() == [0.5, 0.4000000059604645]


def test_batched_call_equals_per_cloud_calls():
    from det3d_b200.ops.point_cloud.voxelize import Voxelizer
    from det3d_b200.utils.synthetic import lidar_like_cloud, uniform_cloud
    from oracle import voxel as ovoxel
    clouds = [lidar_like_cloud(7000, KITTI["pcr"], 4, 1), uniform_cloud(9000, KITTI["pcr"], 4, 2),
              np.zeros((0, 4), np.float32), lidar_like_cloud(300, KITTI["pcr"], 4, 3)]
    vox = Voxelizer(KITTI["vs"], KITTI["pcr"], 5, 4000, want_voxels=True, want_mean=True)
    offs = np.cumsum([0] + [c.shape[0] for c in clouds]).tolist()
    out = vox(torch.from_numpy(np.concatenate(clouds)).cuda(), offs)
    counts = out["counts"].cpu().numpy()
    start = 0
    for b, pts in enumerate(clouds):
        ev, ec, en = ovoxel.points_to_voxel(pts, KITTI["vs"], KITTI["pcr"], 5, True, 4000)
        m = ev.shape[0]
        assert counts[b] == m
        assert np.array_equal(out["voxels"][start:start + m].cpu().numpy(), ev)
        coors = out["coors"][start:start + m].cpu().numpy()
        assert (coors[:, 0] == b).all() and np.array_equal(coors[:, 1:], ec)
        assert np.array_equal(out["num_points"][start:start + m].cpu().numpy(), en)
        mean = ev.sum(1) / en[:, None].astype(np.float32) if m else np.zeros((0, 4), np.float32)
        assert np.allclose(out["mean"][start:start + m].cpu().numpy(), mean, rtol=0, atol=1e-6)
        start += m
    assert counts[len(clouds)] == start


def test_full_size_properties_without_oracle():
    """Size-independent invariants at 2M points (the oracle's dense map is not needed)."""
    from det3d_b200.ops.point_cloud.voxelize import Voxelizer
    from det3d_b200.utils.synthetic import uniform_cloud
    pts = uniform_cloud(2_000_000, KITTI["pcr"], 4, 9)
    vox = Voxelizer(KITTI["vs"], KITTI["pcr"], 5, 150000, want_voxels=True, want_mean=False)
    out = vox(torch.from_numpy(pts).cuda())
    m = int(out["counts"][0])
    assert m == 150000
    coors = out["coors"][:m].cpu().numpy()[:, 1:]
    lin = (coors[:, 0].astype(np.int64) * 1600 + coors[:, 1]) * 1408 + coors[:, 2]
    assert np.unique(lin).size == m                                    # no duplicate voxels
    v = out["voxels"][:m].cpu().numpy()
    n = out["num_points"][:m].cpu().numpy()
    assert n.min() >= 1 and n.max() <= 5
    cells = np.floor((v[:, 0, :3] - np.float32(KITTI["pcr"][:3])) / np.float32(KITTI["vs"])).astype(np.int64)
    assert np.array_equal(cells[:, ::-1], coors)                       # first point lies in its voxel
    pad = np.arange(5)[None, :] >= n[:, None]
    assert (v[pad] == 0).all()                                         # unused slots are exactly zero

"""Multi-sweep ingest: oracle and device kernel vs the reference loading code's own output."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden


def _sweep_records(g):
    order = [int(c) for c in g["chosen"]]
    raws = [g["raw0"]] + [g["raw%d" % (c + 1)] for c in order]
    tms = {0: g["tm0"], 1: None, 2: g["tm2"]}
    return raws, [None] + [tms[c] for c in order], [0.0] + [float(g["lags"][c]) for c in order]


def test_oracle_matches_reference_golden():
    from oracle.ingest import merge_sweeps
    g = load_golden("ingest_nusc_3sweeps")
    raws, tms, lags = _sweep_records(g)
    out = merge_sweeps(raws, tms, lags)
    assert out.dtype == np.float32 and out.shape == g["combined"].shape
    assert np.array_equal(out, g["combined"])
    assert out.shape[0] < sum(r.shape[0] for r in raws)             # remove_close dropped sweep points
    assert np.array_equal(out[: raws[0].shape[0], :4], raws[0][:, :4])   # key frame untouched, not filtered


def test_pipeline_registered_and_fails_loudly_without_gpu(tmp_path):
    from det3d.datasets import PIPELINES
    from det3d.datasets.pipelines.loading import LoadPointCloudFromFile, ingest_sweeps
    assert PIPELINES.get("LoadPointCloudFromFile") is LoadPointCloudFromFile
    pts = np.arange(40, dtype=np.float32).reshape(10, 4)
    (tmp_path / "velodyne").mkdir()
    pts.tofile(str(tmp_path / "velodyne" / "000001.bin"))
    res = dict(lidar={}, metadata=dict(image_prefix=str(tmp_path), num_point_features=4))
    res, _ = LoadPointCloudFromFile(dataset="KittiDataset")(res, dict(point_cloud=dict(velodyne_path="velodyne/000001.bin")))
    assert np.array_equal(res["lidar"]["points"], pts) and res["type"] == "KittiDataset"
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            ingest_sweeps([np.zeros((4, 5), np.float32)], [None], [0.0])


@pytest.mark.gpu
def test_device_ingest_matches_reference_golden(tmp_path):
    from det3d.datasets.pipelines.loading import LoadPointCloudFromFile
    g = load_golden("ingest_nusc_3sweeps")
    paths = []
    for j in range(4):
        p = str(tmp_path / ("sweep%d.pcd.bin" % j))
        np.concatenate([g["raw%d" % j].reshape(-1), np.zeros(3, np.float32)]).astype(np.float32).tofile(p)
        paths.append(p)
    sweeps = [dict(lidar_path=paths[1], transform_matrix=g["tm0"], time_lag=float(g["lags"][0])),
              dict(lidar_path=paths[2], transform_matrix=None, time_lag=float(g["lags"][1])),
              dict(lidar_path=paths[3], transform_matrix=g["tm2"], time_lag=float(g["lags"][2]))]
    res = dict(lidar=dict(nsweeps=3), metadata={})
    np.random.seed(7)
    res, _ = LoadPointCloudFromFile(dataset="NuScenesDataset")(res, dict(lidar_path=paths[0], sweeps=sweeps))
    got, want = res["lidar"]["combined"], g["combined"]
    assert got.shape == want.shape and got.dtype == np.float32
    assert res["lidar"]["combined_cuda"].is_cuda and np.array_equal(res["lidar"]["points"], got[:, :4])
    exact = got == want
    # float64 dot product rounded once to fp32: BLAS may contract/associate differently -> at most a last-bit difference
    assert exact.mean() >= 0.999
    assert np.all(np.abs(got - want) <= np.spacing(np.abs(want)).astype(np.float32))
    assert np.array_equal(got[:, 3:], want[:, 3:])                   # intensity and time columns are copies


@pytest.mark.gpu
def test_device_ingest_feeds_voxelizer_and_edge_cases():
    from det3d.datasets.pipelines.loading import ingest_sweeps
    from det3d_b200.ops.point_cloud.voxelize import Voxelizer
    from oracle.ingest import merge_sweeps
    rng = np.random.default_rng(5)
    raws = [np.concatenate([rng.uniform(-50, 50, (n, 3)), rng.uniform(0, 255, (n, 1)), np.zeros((n, 1))], 1).astype(np.float32)
            for n in (3000, 0, 5000, 1)]
    tms = [None, None, np.eye(4), None]
    lags = [0.0, 0.1, 0.2, 0.3]
    out = ingest_sweeps(raws, tms, lags)
    want = merge_sweeps(raws, tms, lags)
    assert np.array_equal(out.cpu().numpy(), want)                   # identity transform: exact
    vox = Voxelizer([0.1, 0.1, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 10, 60000, want_voxels=False, want_mean=True)(out, None)
    assert int(vox["counts"][0]) > 0
    empty = ingest_sweeps([np.zeros((0, 5), np.float32)], [None], [0.0])
    assert empty.shape == (0, 5)

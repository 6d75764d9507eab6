"""End-to-end SECOND forward (config 2 of BASELINE.json) on the GPU vs the CPU restatement of
the reference path (oracle/second_cpu.py), stage by stage."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from det3d.models import build_detector
    from det3d.torchie import Config
    from det3d_b200.apis import InferencePipeline
    from det3d_b200.utils.synthetic import randomize_bn_
    from oracle.second_cpu import SecondCPU

    cfg = Config.fromfile(os.path.join(ROOT, "configs", "second_kitti_car.py"))
    torch.manual_seed(0)
    model = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval()
    randomize_bn_(model, 0)
    with torch.no_grad():  # spread the scores so that ~5 % of the anchors pass the 0.3 threshold
        head = model.bbox_head.tasks[0]
        head.conv_cls.weight.mul_(4.0)
        head.conv_cls.bias.fill_(-2.5)
        head.conv_box.weight.mul_(0.3)
    pipe = InferencePipeline(cfg, model=model, device="cuda")
    cpu = SecondCPU(cfg, model.state_dict(), [a.cpu().numpy() for a in pipe._anchors])
    return cfg, pipe, cpu


def _match(cpu_boxes, gpu_boxes, tol=2e-3):
    if cpu_boxes.shape[0] == 0 or gpu_boxes.shape[0] == 0:
        return 0
    d = (cpu_boxes[:, None, :] - gpu_boxes[None, :, :]).abs().max(-1)
    return int((d.min(1)[0] <= tol).sum())


@pytest.mark.parametrize("dist,n", [("lidar", 20000), ("uniform", 4000)])
def test_forward_matches_cpu_restatement(setup, dist, n):
    from det3d_b200.utils.synthetic import lidar_like_cloud, uniform_cloud
    cfg, pipe, cpu = setup
    pts = (lidar_like_cloud if dist == "lidar" else uniform_cloud)(n, cfg.voxel_generator.range, 4, 1)
    stages = {}
    want = cpu.forward([pts], stages)

    dev_pts = torch.from_numpy(pts).cuda()
    vox = pipe.voxelizer(dev_pts, [0, n])
    m = int(vox["counts"][0])
    assert m == stages["coors"].shape[0]
    assert np.array_equal(vox["coors"][:m].cpu().numpy(), stages["coors"])                 # bit-exact indices
    assert np.array_equal(vox["num_points"][:m].cpu().numpy(), stages["nums"])
    feats_cpu = stages["voxels"].sum(1) / stages["nums"][:, None].astype(np.float32)
    assert np.allclose(vox["mean"][:m].cpu().numpy(), feats_cpu, rtol=0, atol=1e-6)

    with torch.no_grad():
        dense = pipe.model.backbone(vox["mean"], vox["coors"], 1, [int(g) for g in pipe.grid_size],
                                    n_dev=vox["counts"][1:2])
    assert float((dense.cpu() - stages["dense"]).abs().max()) <= 1e-4                     # north_star tolerance

    det = pipe.forward_device(dev_pts, [0, n])
    got = pipe.unpack(pipe.pack(det).cpu())[0]
    w = want[0]
    assert abs(got["box3d_lidar"].shape[0] - w["box3d_lidar"].shape[0]) <= max(2, w["box3d_lidar"].shape[0] // 20)
    if w["box3d_lidar"].shape[0]:
        matched = _match(w["box3d_lidar"], got["box3d_lidar"])
        assert matched >= 0.9 * w["box3d_lidar"].shape[0], "only %d of %d detections match" % (matched, w["box3d_lidar"].shape[0])


def test_host_api_and_batching(setup):
    from det3d_b200.utils.synthetic import lidar_like_cloud
    cfg, pipe, cpu = setup
    clouds = [torch.from_numpy(lidar_like_cloud(6000 + 500 * i, cfg.voxel_generator.range, 4, 10 + i)).pin_memory()
              for i in range(3)]
    packed = pipe.infer_host(clouds)
    assert packed.shape[0] == 3 and packed.shape[2] == 10
    singles = [pipe.infer_host([c]).clone() for c in clouds]
    for b in range(3):
        a, s = pipe.unpack(packed)[b], pipe.unpack(singles[b])[0]
        assert a["box3d_lidar"].shape == s["box3d_lidar"].shape
        assert torch.allclose(a["box3d_lidar"], s["box3d_lidar"], atol=1e-4)
        assert torch.allclose(a["scores"], s["scores"], atol=1e-5)


def test_model_predict_api(setup):
    """VoxelNet.forward(example, return_loss=False) with reference-style inputs (voxels [M,5,4])."""
    from det3d.core.input.voxel_generator import VoxelGenerator
    from det3d_b200.utils.synthetic import lidar_like_cloud
    cfg, pipe, cpu = setup
    vg = cfg.voxel_generator
    gen = VoxelGenerator(vg.voxel_size, vg.range, vg.max_points_in_voxel, vg.max_voxel_num)
    pts = lidar_like_cloud(5000, vg.range, 4, 3)
    voxels, coors, num = gen.generate(pts)
    coors = np.concatenate([np.zeros((coors.shape[0], 1), np.int32), coors], 1)           # collate_kitti
    example = dict(voxels=torch.from_numpy(voxels).cuda(), coordinates=torch.from_numpy(coors).cuda(),
                   num_points=torch.from_numpy(num).cuda(), num_voxels=torch.tensor([voxels.shape[0]]),
                   shape=[gen.grid_size], anchors=pipe.anchors(1))
    with torch.no_grad():
        out = pipe.model(example, return_loss=False)
    assert len(out) == 1 and set(out[0]) >= {"box3d_lidar", "scores", "label_preds"}
    ref = pipe.unpack(pipe.infer_host([torch.from_numpy(pts)]))[0]
    assert out[0]["box3d_lidar"].shape[0] == ref["box3d_lidar"].shape[0]
    assert torch.allclose(out[0]["box3d_lidar"].cpu(), ref["box3d_lidar"], atol=1e-4)


def test_fused_bev_path_matches_cudnn_path(setup):
    """RPN + heads through the channels-last tcgen05 kernels vs the module's torch/cuDNN fp32 forward."""
    from det3d_b200.utils.synthetic import lidar_like_cloud
    cfg, pipe, cpu = setup
    model = pipe.model
    assert model.fused_bev() is not None
    pts = torch.from_numpy(lidar_like_cloud(20000, cfg.voxel_generator.range, 4, 5)).cuda()
    vox = pipe.voxelizer(pts, [0, 20000])
    grid = [int(g) for g in pipe.grid_size]
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            rows, (b, h, w) = model.backbone.forward_rows(vox["mean"], vox["coors"], 1, grid, n_dev=vox["counts"][1:2])
            dense = model.backbone(vox["mean"], vox["coors"], 1, grid, n_dev=vox["counts"][1:2])
            assert torch.equal(rows.view(b, h, w, -1).permute(0, 3, 1, 2), dense)
            fused = model.fused_bev().run(rows, b, h, w)
            ref = model.bbox_head(model.neck(dense))
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    for key in ("box_preds", "cls_preds", "dir_cls_preds"):
        a, r = fused[0][key], ref[0][key]
        assert a.shape == r.shape
        assert float((a - r).abs().max()) <= 2e-4, key

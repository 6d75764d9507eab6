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
    from det3d_b200.utils.synthetic import demo_weights_
    from oracle.second_cpu import SecondCPU

    cfg = Config.fromfile(os.path.join(ROOT, "configs", "second_kitti_car.py"))
    torch.manual_seed(0)
    # random weights calibrated so that ~3 % of the anchors pass the 0.3 threshold with spread scores
    model = demo_weights_(build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval(), 0)
    pipe = InferencePipeline(cfg, model=model, device="cuda")
    cpu = SecondCPU(cfg, model.state_dict(), [a.cpu().numpy() for a in pipe._anchors])
    return cfg, pipe, cpu


def _match(cpu_boxes, gpu_boxes, tol=2e-3):
    if cpu_boxes.shape[0] == 0 or gpu_boxes.shape[0] == 0:
        return 0
    d = (cpu_boxes[:, None, :] - gpu_boxes[None, :, :]).abs().max(-1)[0]
    return int((d.min(1)[0] <= tol).sum())


@pytest.mark.parametrize("dist,n", [("lidar", 20000), ("uniform", 20000)])
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
    scale = max(1.0, float(stages["dense"].abs().max()))
    assert float((dense.cpu() - stages["dense"]).abs().max()) <= 1e-4 * scale             # north_star tolerance (abs, O(1) features)

    det = pipe.forward_device(dev_pts, [0, n])
    got = pipe.unpack(pipe.pack(det).cpu())[0]
    w = want[0]
    assert w["box3d_lidar"].shape[0] >= 10, "degenerate workload: the CPU restatement found no detections"
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
            planes = model.backbone.forward_planes(vox["mean"], vox["coors"], 1, grid, n_dev=vox["counts"][1:2])
            dense = model.backbone(vox["mean"], vox["coors"], 1, grid, n_dev=vox["counts"][1:2])
            # the same values in the two layouts, bit for bit (the FP16x3 encoder is deterministic)
            assert torch.equal(planes.to_f32().permute(0, 3, 1, 2), dense)
            fused = model.fused_bev().run(planes)
            ref = model.bbox_head(model.neck(dense))
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    for key in ("box_preds", "cls_preds", "dir_cls_preds"):
        a, r = fused[0][key], ref[0][key]
        assert a.shape == r.shape
        # fp32 cuDNN vs 3xTF32: both fp32-accurate; compare relative to the tensor's magnitude
        assert float((a - r).abs().max()) <= 1e-4 * max(1.0, float(r.abs().max())), key


def test_cbgs_nuscenes_config_batch2():
    """BASELINE config 4 shape: CBGS (SpMiddleResNetFHD, 2-block RPN, 6 task heads, 9-dim boxes with
    angle-vector encoding), 35k-point 5-feature clouds, batch of 2.  The backbone is checked against the
    oracle elsewhere (test_spconv_gpu); here the device-side predict is checked against the CPU restatement
    of MultiGroupHead.predict fed with the same head outputs."""
    from det3d.models import build_detector
    from det3d.torchie import Config
    from det3d_b200.apis import InferencePipeline
    from det3d_b200.utils.synthetic import demo_weights_, lidar_like_cloud
    from oracle.predict_cpu import predict_sample_task

    cfg = Config.fromfile(os.path.join(ROOT, "configs", "cbgs_nusc.py"))
    torch.manual_seed(1)
    model = demo_weights_(build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval(), 1, cls_bias=-2.4)
    pipe = InferencePipeline(cfg, model=model, device="cuda")
    assert pipe.model.fused_bev() is None          # strided RPN with a ConvTranspose deblock: torch path
    clouds = [lidar_like_cloud(35000, cfg.voxel_generator.range, 5, s) for s in (0, 1)]
    pts = torch.from_numpy(np.concatenate(clouds)).cuda()
    offsets = [0, 35000, 70000]
    det = pipe.forward_device(pts, offsets)
    assert det["boxes"].shape == (2, 6 * 83, 9) and int(det["valid"].sum()) > 20

    # one set of head outputs -> device predict vs the CPU restatement of predict.  (The encoder is not
    # re-run for the comparison: its pair-based layers sum with fp32 atomics, so two runs differ in the
    # last bits and near-tied random-weight scores may reorder.)
    with torch.no_grad():
        vox = pipe.voxelizer(pts, offsets)
        counts = vox["counts"].cpu().numpy()
        assert counts[2] == counts[0] + counts[1] and counts[0] > 10000
        x = model.backbone(vox["mean"], vox["coors"], 2, [int(g) for g in pipe.grid_size], n_dev=vox["counts"][2:3])
        preds = model.bbox_head(model.neck(x))
        det2 = model.bbox_head.predict_device(dict(anchors=pipe.anchors(2)), preds, cfg.test_cfg)
    got = pipe.unpack(pipe.pack(det2).cpu())
    flag = 0
    want = [dict(b=[], s=[], l=[]) for _ in range(2)]
    for task_id, p in enumerate(preds):
        anchors = pipe._anchors[task_id].cpu()
        n_cls = model.bbox_head.num_classes[task_id]
        for b in range(2):
            bx, sc, lb = predict_sample_task(p["cls_preds"][b].reshape(-1, n_cls).cpu(), p["box_preds"][b].reshape(-1, 10).cpu(),
                                             None, anchors, cfg.test_cfg, True)
            want[b]["b"].append(bx); want[b]["s"].append(sc); want[b]["l"].append(lb + flag)
        flag += n_cls
    for b in range(2):
        wb, wl = torch.cat(want[b]["b"]), torch.cat(want[b]["l"])
        gb, gl = got[b]["box3d_lidar"], got[b]["label_preds"]
        assert wb.shape[0] >= 10
        assert abs(gb.shape[0] - wb.shape[0]) <= max(3, wb.shape[0] // 20)
        d = (wb[:, None, :] - gb[None, :, :]).abs().max(-1)[0]
        j = d.argmin(1)
        ok = (d.min(1)[0] <= 2e-3) & (gl[j] == wl)
        assert int(ok.sum()) >= 0.9 * wb.shape[0]


def test_fused_predict_kernels_match_torch_ops(setup):
    """d3b_predict_task (radix-select top-k, decode of the selected anchors, NMS, finalize) vs the same
    algorithm written with torch ops: identical boxes / scores / labels / validity."""
    from det3d_b200.utils.synthetic import lidar_like_cloud
    cfg, pipe, cpu = setup
    model = pipe.model
    for seed in (3, 4):
        pts = torch.from_numpy(lidar_like_cloud(20000, cfg.voxel_generator.range, 4, seed)).cuda()
        vox = pipe.voxelizer(pts, [0, 20000])
        with torch.no_grad():
            planes = model.backbone.forward_planes(vox["mean"], vox["coors"], 1, [int(g) for g in pipe.grid_size],
                                                   n_dev=vox["counts"][1:2])
            preds = model.fused_bev().run(planes)
            example = dict(anchors=pipe.anchors(1))
            a = model.bbox_head.predict_device(example, preds, cfg.test_cfg)
            t = model.bbox_head.predict_device(example, preds, cfg.test_cfg, use_torch_ops=True)
        va, vt = a["valid"][0], t["valid"][0]
        assert int(vt.sum()) >= 10
        assert torch.equal(va, vt)
        assert torch.equal(a["boxes"][0][va], t["boxes"][0][vt])          # same fp32 op sequence -> bit-identical
        assert torch.equal(a["scores"][0][va], t["scores"][0][vt])
        assert torch.equal(a["labels"][0][va], t["labels"][0][vt])


@pytest.mark.parametrize("hw", [3000, 40000])
def test_topk_handles_ties_and_constants(hw):
    """All-equal scores (the degenerate case): lowest anchor indices win, deterministically.  hw=3000: the whole
    pivot bin is sorted in shared memory; hw=40000: it exceeds the sort capacity -> radix select over composites."""
    import ctypes as C
    from det3d_b200 import _lib
    B, na = 2, 2
    cls = torch.zeros((B, 1, hw, na), device="cuda")
    cls[1] = -5.0
    cls[1, 0, 100:110, :] = 5.0
    box = torch.zeros((B, 1, hw, na * 7), device="cuda")
    anchors = torch.rand((hw * na, 7), device="cuda") * 10 + 1
    q = _lib.PredictParams()
    q.cls, q.cls_row_stride, q.cls_col0 = cls.data_ptr(), na, 0
    q.box, q.box_row_stride, q.box_col0 = box.data_ptr(), na * 7, 0
    q.dir = None
    q.anchors = anchors.data_ptr()
    q.batch, q.hw, q.na, q.n_cls, q.code, q.nd = B, hw, na, 1, 7, 7
    q.use_rotate_nms, q.pre_max, q.post_max = 1, 1000, 1000
    q.nms_iou_threshold, q.score_threshold = 2.0, 0.3       # IoU never reaches 2 -> NMS keeps every candidate
    q.has_range = 0
    ws = torch.empty(_lib.lib().d3b_predict_workspace_bytes(C.byref(q)), dtype=torch.uint8, device="cuda")
    packed = torch.zeros((B, 1000, 10), device="cuda")
    counts = torch.zeros(B, dtype=torch.int32, device="cuda")
    st = _lib.lib().d3b_predict_task(C.byref(q), packed.data_ptr(), 1000, 0, counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                    torch.cuda.current_stream().cuda_stream)
    _lib.check(st)
    assert counts.tolist() == [1000, 20]                    # sample 0: 0.5 >= 0.3 for all; sample 1: only the 20 raised logits
    got0 = packed[0, :, :3]
    assert torch.equal(got0, anchors[:1000, :3])            # ties -> first 1000 anchors, in index order
    assert torch.equal(packed[1, :20, :3], anchors[200:220, :3])

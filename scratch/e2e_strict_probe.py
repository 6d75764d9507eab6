"""Probe (scratch): how far are GPU detections from the CPU restatement, and why?  Run on the GPU box."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from det3d.models import build_detector
from det3d.torchie import Config
from det3d_b200.apis import InferencePipeline
from det3d_b200.utils.synthetic import demo_weights_, lidar_like_cloud, uniform_cloud
from oracle.second_cpu import SecondCPU

cfg = Config.fromfile(os.path.join(ROOT, "configs", "second_kitti_car.py"))
torch.manual_seed(0)
model = demo_weights_(build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval(), 0)
pipe = InferencePipeline(cfg, model=model, device="cuda")
cpu = SecondCPU(cfg, model.state_dict(), [a.cpu().numpy() for a in pipe._anchors])
out = {}
for dist, fn in (("lidar", lidar_like_cloud), ("uniform", uniform_cloud)):
    pts = fn(20000, cfg.voxel_generator.range, 4, 1)
    stages = {}
    want = cpu.forward([pts], stages)[0]
    dev_pts = torch.from_numpy(pts).cuda()
    for mode in ("fp16x3", "tf32x3", "tf32x3-det"):
        pipe.model.set_math(mode.split("-")[0])
        pipe.model.backbone.fused().deterministic = mode.endswith("det")
        det_mode = mode
        runs = []
        for rep in range(3):
            vox = pipe.voxelizer(dev_pts, [0, 20000])
            grid = [int(g) for g in pipe.grid_size]
            with torch.no_grad():
                if mode == "fp16x3":
                    planes = pipe.model.backbone.forward_planes(vox["mean"], vox["coors"], 1, grid, n_dev=vox["counts"][1:2])
                    b, h, w, _c = planes.shape
                    rows = planes.to_f32().view(b * h * w, -1)
                    preds = pipe.model.fused_bev().run(planes)
                else:
                    rows, (b, h, w) = pipe.model.backbone.forward_rows(vox["mean"], vox["coors"], 1, grid, n_dev=vox["counts"][1:2])
                    rows = rows.clone()
                    preds = pipe.model.fused_bev().run(rows, b, h, w)
                cls = preds[0]["cls_preds"].clone(); box = preds[0]["box_preds"].clone(); dr = preds[0]["dir_cls_preds"].clone()
            det = pipe.forward_device(dev_pts, [0, 20000])
            got = pipe.unpack(pipe.pack(det).cpu())[0]
            runs.append((rows.cpu(), cls.cpu(), box.cpu(), dr.cpu(), got))
        same_rows = all(torch.equal(runs[0][0], r[0]) for r in runs[1:])
        same_det = all(torch.equal(runs[0][4]["box3d_lidar"], r[4]["box3d_lidar"]) for r in runs[1:])
        rows, cls, box, dr, got = runs[0]
        dense = rows.view(1, h, w, -1).permute(0, 3, 1, 2)
        e_dense = float((dense - stages["dense"]).abs().max())
        e_cls = float((cls.reshape(-1) - stages["cls"].reshape(-1)).abs().max())
        e_box = float((box.reshape(-1) - stages["box"].reshape(-1)).abs().max())
        # oracle predict on the GPU head outputs
        w2 = cpu.predict(box.view(1, h, w, -1), cls.view(1, h, w, -1), dr.view(1, h, w, -1))[0]
        def match(a, b, tol):
            if a.shape[0] == 0 or b.shape[0] == 0: return 0
            d = (a[:, None, :] - b[None, :, :]).abs().max(-1)[0]
            return int((d.min(1)[0] <= tol).sum())
        res = dict(run_to_run_rows_equal=same_rows, run_to_run_det_equal=same_det, dense_abs_err=e_dense, dense_max=float(stages["dense"].abs().max()),
                   cls_err=e_cls, box_err=e_box, n_cpu=int(want["box3d_lidar"].shape[0]), n_gpu=int(got["box3d_lidar"].shape[0]),
                   match_cpu_2e3=match(want["box3d_lidar"], got["box3d_lidar"], 2e-3), match_cpu_1e4=match(want["box3d_lidar"], got["box3d_lidar"], 1e-4),
                   n_oracle_on_gpu_heads=int(w2["box3d_lidar"].shape[0]),
                   match_oracle_on_gpu_heads_1e5=match(w2["box3d_lidar"], got["box3d_lidar"], 1e-5),
                   exact_oracle_on_gpu_heads=bool(w2["box3d_lidar"].shape == got["box3d_lidar"].shape and torch.equal(w2["box3d_lidar"], got["box3d_lidar"])),
                   scores_exact=bool(w2["scores"].shape == got["scores"].shape and torch.equal(w2["scores"], got["scores"])))
        # score-gap analysis of CPU candidates: sorted sigmoid scores above threshold
        sc = torch.sigmoid(stages["cls"].reshape(-1)); sc = sc[sc >= 0.3].sort(descending=True)[0][:1000]
        gaps = (sc[:-1] - sc[1:])
        res["cand"] = int(sc.shape[0]); res["gaps_lt_1e6"] = int((gaps < 1e-6).sum()); res["gaps_lt_1e5"] = int((gaps < 1e-5).sum())
        # timing of the encoder in this mode
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20):
            if mode == "fp16x3":
                pipe.model.backbone.forward_planes(vox["mean"], vox["coors"], 1, grid, n_dev=vox["counts"][1:2])
            else:
                pipe.model.backbone.forward_rows(vox["mean"], vox["coors"], 1, grid, n_dev=vox["counts"][1:2])
        torch.cuda.synchronize(); res["encoder_eager_ms"] = (time.perf_counter() - t0) / 20 * 1e3
        res["overflow"] = int(pipe.overflow_flag().item())
        out["%s_%s" % (dist, det_mode)] = res
        print(dist, det_mode, json.dumps(res), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "e2e_strict_probe.json"), "w"), indent=1)

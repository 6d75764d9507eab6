import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import det3d_b200
from det3d.models import build_detector
from det3d.torchie import Config
from det3d_b200.apis import InferencePipeline
from det3d_b200.utils.synthetic import demo_weights_, uniform_cloud, lidar_like_cloud
which = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if which == "c3":
    cfg = Config.fromfile(os.path.join(ROOT, "configs", "pointpillars_kitti_car.py"))
    model = demo_weights_(build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval(), 0, cls_scale=0.3, cls_bias=-3.6, box_scale=0.02)
    clouds = [uniform_cloud(20000, cfg.voxel_generator.range, 4, s) for s in range(8)]
    off = [20000 * i for i in range(9)]
else:
    cfg = Config.fromfile(os.path.join(ROOT, "configs", "cbgs_nusc.py"))
    model = demo_weights_(build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval(), 1, cls_bias=-2.4)
    clouds = [lidar_like_cloud(35000, cfg.voxel_generator.range, 5, s) for s in range(4)]
    off = [35000 * i for i in range(5)]
pipe = InferencePipeline(cfg, model=model, device="cuda")
p = torch.from_numpy(np.concatenate(clouds)).cuda()
for _ in range(3):
    pipe.pack(pipe.forward_device(p, off))
torch.cuda.synchronize()
torch.cuda.profiler.start()
pipe.pack(pipe.forward_device(p, off))
torch.cuda.synchronize()
torch.cuda.profiler.stop()

"""Sizes inside d3b_predict_task on the bench workload: candidate-list length after the histogram partition, n_valid, kept."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench


def main(config):
    sys.argv = ["bench.py", "--config", config]
    args = bench.parse()
    from det3d.torchie import Config
    from det3d_b200.apis import InferencePipeline
    dev = torch.device("cuda", 0)
    cfg = Config.fromfile(os.path.join(ROOT, "configs", args.wl["cfg"]))
    pipe = InferencePipeline(cfg, model=bench.build_model(cfg, args), device=dev)
    B, NP = args.batch, args.wl["n_points"]
    clouds = bench.make_clouds(args, B, 0, cfg.voxel_generator.range)
    pts = torch.from_numpy(np.concatenate(clouds)).to(dev)
    offsets = [NP * i for i in range(B + 1)]
    pipe.pack(pipe.forward_device(pts, offsets))
    torch.cuda.synchronize()
    head = pipe.model.bbox_head
    return pipe, head


def align(x, a=256):
    return (x + a - 1) // a * a


def sizes(ws, B, A, k):
    off = align(B * A * 4) + align(B * A)
    hist = ws[off:off + B * 2048 * 4 + B * 4].view(torch.int32)
    part_count = hist[B * 2048:B * 2048 + B].cpu().numpy()
    return part_count


if __name__ == "__main__":
    for config in sys.argv[1:] or ["second"]:
        pipe, head = main(config)
        for key, bufs in head._predict_bufs.items():
            B = int(bufs["packed"].shape[0])
            for task_id, ws in bufs["ws"].items():
                A = int(pipe._anchors[task_id].shape[0])
                print(config, "task", task_id, "B", B, "A", A, "candidate list lengths", sizes(ws, B, A, 0).tolist(), flush=True)

"""Where does the end-to-end step of a config spend its time?  host segments (perf_counter) + device events."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

def main():
    sys.argv = ["bench.py"] + sys.argv[1:]
    args = bench.parse()
    from det3d.torchie import Config
    from det3d_b200.apis import InferencePipeline
    dev = torch.device("cuda", 0)
    cfg = Config.fromfile(os.path.join(ROOT, "configs", args.wl["cfg"]))
    pipe = InferencePipeline(cfg, model=bench.build_model(cfg, args), device=dev)
    pipe.model.set_math(args.math)
    B, NP, ND = args.batch, args.wl["n_points"], args.wl["ndim"]
    clouds = bench.make_clouds(args, 16, 0, cfg.voxel_generator.range)
    pinned = [torch.from_numpy(c).pin_memory() for c in clouds]
    offsets = [NP * i for i in range(B + 1)]
    pts = torch.empty((NP * B, ND), dtype=torch.float32, device=dev)
    out_pinned = None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for step in range(12):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a, b, c, d = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        a.record()
        for j in range(B):
            pts[j * NP:(j + 1) * NP].copy_(pinned[(step * B + j) % 16], non_blocking=True)
        b.record()
        t1 = time.perf_counter()
        packed = pipe.forward_graphed(pts, offsets)
        c.record()
        t2 = time.perf_counter()
        if out_pinned is None:
            out_pinned = torch.empty(packed.shape, dtype=torch.float32, pin_memory=True)
        out_pinned.copy_(packed, non_blocking=True)
        d.record()
        t3 = time.perf_counter()
        torch.cuda.current_stream().synchronize()
        t4 = time.perf_counter()
        rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0),
                     a.elapsed_time(b), b.elapsed_time(c), c.elapsed_time(d), a.elapsed_time(d)))
    print("host: h2d-enqueue  graph-launch  d2h-enqueue  sync  total | device: h2d  graph  d2h  total   (ms)")
    for r in rows[2:]:
        print("  ".join(f"{x:8.3f}" for x in r))

if __name__ == "__main__":
    main()

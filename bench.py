#!/usr/bin/env python
"""Headline benchmark: point-clouds/sec of the SECOND (SpMiddleFHD, KITTI-car grid) forward on
synthetic 20k-point clouds -- BASELINE.json `metric`, config[1].

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                     (CPU restatement of the reference path)

One step = one pass of the whole hot path (voxelize -> VFE -> sparse middle encoder -> dense ->
RPN -> head -> decode/top-k/rotated NMS -> all-gather of detections) over one batch of clouds.
Prints ONE JSON line (see the driver contract in the task statement).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIG = os.path.join(ROOT, "configs", "second_kitti_car.py")
N_POINTS = 20000
N_CLOUD_POOL = 8       # distinct synthetic clouds cycled through the steps


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="det3d_b200", choices=["det3d_b200", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="clouds per GPU per step (BASELINE config: 1)")
    ap.add_argument("--dist", default="lidar_like", choices=["lidar_like", "uniform"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--allow-tf32-rpn", action="store_true", help="let cuDNN use TF32 in the dense RPN")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-fused-bev", action="store_true", help="RPN/head through torch+cuDNN instead of the tcgen05 path")
    return ap.parse_args()


def make_clouds(dist, count, seed0, pcr):
    from det3d_b200.utils.synthetic import lidar_like_cloud, uniform_cloud
    fn = lidar_like_cloud if dist == "lidar_like" else uniform_cloud
    return [fn(N_POINTS, pcr, 4, seed0 + i) for i in range(count)]


def build_model(cfg):
    """Random-init weights of the named architecture (no checkpoints offline), calibrated so that the
    detection head sees a realistic workload: ~2k anchors above the score threshold, top-1000 into NMS."""
    import torch
    from det3d.models import build_detector
    from det3d_b200.utils.synthetic import demo_weights_
    torch.manual_seed(0)
    return demo_weights_(build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval(), 0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.time()])

    def mark(self):
        """Start of the load window: samples read before this moment are dropped."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        t0 = getattr(self, "t0", 0.0)
        rows = [r for r in self.rows if r[-1] >= t0] or self.rows[-3:]
        self.rows = rows
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference(args, rank, world):
    """CPU restatement of the reference path (oracle/second_cpu.py) on the host cores."""
    if rank != 0:
        return
    import torch
    from det3d.torchie import Config
    from det3d_b200.core.anchor.anchor_generator import anchors_for_tasks
    from oracle.second_cpu import SecondCPU
    cfg = Config.fromfile(CONFIG)
    model = build_model(cfg)
    anchors = anchors_for_tasks(cfg.target_assigner, [1408, 1600, 40], cfg.assigner.out_size_factor)
    cpu = SecondCPU(cfg, model.state_dict(), anchors)
    clouds = make_clouds(args.dist, N_CLOUD_POOL, 0, cfg.voxel_generator.range)
    batch = args.batch * world            # the whole job's step, done by the host alone
    t0 = time.perf_counter()
    cpu.forward([clouds[0]])              # builds the oracle library, warms torch
    one = time.perf_counter() - t0
    budget = 240.0
    warm = min(args.warmup, max(1, int(20.0 / max(one * batch, 1e-3))))
    steps = min(args.steps, max(1, int(budget / max(one * batch, 1e-3))))
    for i in range(warm - 1):
        cpu.forward([clouds[(i + j) % N_CLOUD_POOL] for j in range(batch)])
    t0 = time.perf_counter()
    for i in range(steps):
        cpu.forward([clouds[(i * batch + j) % N_CLOUD_POOL] for j in range(batch)])
    dt = time.perf_counter() - t0
    value = steps * batch / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "point-clouds/sec SECOND SpMiddleFHD @20k pts", "value": value,
        "unit": "clouds/s", "n_gpus": world, "steps": steps, "steps_requested": args.steps, "warmup": warm,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": cores, "kind": "port",
                         "sample": "%d full forwards of %d cloud(s), CPU restatement of the reference path "
                                   "(oracle/second_cpu.py; spconv is absent from the reference checkout)" % (steps, batch),
                         "stage_seconds": cpu.timings},
        "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


def workload_config(args, world):
    return {"workload": "SECOND kitti_car_vfev3_spmiddlefhd_rpn1 forward, 20k synthetic pts, batch=%d/GPU" % args.batch,
            "distribution": args.dist, "points_per_cloud": N_POINTS, "batch_per_gpu": args.batch,
            "global_batch": args.batch * world, "cloud_pool": N_CLOUD_POOL, "l2": "flushed (256 MiB write) between steps",
            "parallelism": "dp%d" % world,
            "rpn_math": ("cuDNN fp32 (allow_tf32=%s)" % bool(args.allow_tf32_rpn)) if args.no_fused_bev
            else "tcgen05 3xTF32 (fp32-equivalent), channels-last",
            "cuda_graph": not args.no_graph}


_REAL_STDOUT = None


def _claim_stdout():
    """Route everything libraries write to fd 1 (NCCL prints its version there, build tools chat) to stderr, so that
    stdout carries exactly the one JSON line the contract asks for."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    args = parse()
    _claim_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from det3d.torchie import Config
    from det3d_b200 import _lib
    from det3d_b200.apis import InferencePipeline, all_gather_detections, init_from_env
    from det3d_b200.ops.spconv import core as spcore

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: det3d_b200 has no CPU fallback")
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the single JSON line (NCCL prints its version there)
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = Config.fromfile(CONFIG)
    pipe = InferencePipeline(cfg, model=build_model(cfg), device=dev, strict_fp32=not args.allow_tf32_rpn)
    pipe.model.use_fused_bev = not args.no_fused_bev
    use_graph = not args.no_graph
    B = args.batch
    clouds_np = make_clouds(args.dist, N_CLOUD_POOL, 1000 * rank, cfg.voxel_generator.range)
    pinned = [torch.from_numpy(c).pin_memory() for c in clouds_np]
    resident = [torch.from_numpy(c).to(dev) for c in clouds_np]
    offsets = [N_POINTS * i for i in range(B + 1)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e2e_pts = torch.empty((N_POINTS * B, 4), dtype=torch.float32, device=dev)
    out_pinned = None

    def batch_ids(step):
        return [(step * B + j) % N_CLOUD_POOL for j in range(B)]

    def step_device(step):
        ids = batch_ids(step)
        pts = resident[ids[0]] if B == 1 else torch.cat([resident[i] for i in ids])
        packed = pipe.forward_graphed(pts, offsets) if use_graph else pipe.pack(pipe.forward_device(pts, offsets))
        return all_gather_detections(packed)

    def step_e2e(step):
        nonlocal out_pinned
        ids = batch_ids(step)
        pts = e2e_pts
        for j, i in enumerate(ids):
            pts[j * N_POINTS:(j + 1) * N_POINTS].copy_(pinned[i], non_blocking=True)   # H2D from pinned memory
        packed = pipe.forward_graphed(pts, offsets) if use_graph else pipe.pack(pipe.forward_device(pts, offsets))
        gathered = all_gather_detections(packed)
        if out_pinned is None:
            out_pinned = torch.empty(gathered.shape, dtype=torch.float32, pin_memory=True)
        out_pinned.copy_(gathered, non_blocking=True)
        torch.cuda.current_stream().synchronize()     # the caller holds the detections on the host
        return out_pinned

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, conv_events=None):  # noqa: E306
        evs = []
        barrier()
        for s in range(steps):
            flush.zero_()                                   # L2 flush, outside the timed events
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if conv_events is not None:
                spcore.PROFILE_EVENTS = conv_events
            fn(s)
            spcore.PROFILE_EVENTS = None
            b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for s in range(max(args.warmup, 3)):
        step_device(s)
        step_e2e(s)

    def step_eager(step):
        ids = batch_ids(step)
        pts = resident[ids[0]] if B == 1 else torch.cat([resident[i] for i in ids])
        return all_gather_detections(pipe.pack(pipe.forward_device(pts, offsets)))

    sampler.mark()                      # clocks are sampled from here to the end of the per-kernel pass (same load)
    ms = timed(step_device, args.steps)
    ms_e2e = timed(step_e2e, args.steps)

    # ---- per-kernel pass (eager launches of the SAME kernels, CUDA events around every
    #      d3b_sparse_conv launch on the launching stream): launch count + roofline numerators ----
    n_prof = min(args.steps, 10)
    launches0 = _lib.launch_count()
    conv_events = []
    timed(step_eager, n_prof, conv_events)
    launches_per_step = (_lib.launch_count() - launches0) / n_prof
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    traffic = {}
    try:      # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
    except OSError:
        pass
    enc = pipe.model.backbone.fused().accounting()             # algorithmic bytes / flops of the last step
    n_enc = len(enc["layers"])
    per_step = len(conv_events) // n_prof
    ev_ms = [a.elapsed_time(b) for a, b in conv_events]
    enc_ms = sum(t for i, t in enumerate(ev_ms) if i % per_step < n_enc) / n_prof
    bev_ms = sum(t for i, t in enumerate(ev_ms) if i % per_step >= n_enc) / n_prof
    ms_step = ms / args.steps
    timing = "CUDA events around each launch on the launching stream, eager pass of %d steps" % n_prof
    enc_gbs = enc["bytes"] / (enc_ms * 1e-3) / 1e9 if enc_ms > 0 else 0.0
    roofline_encoder = {
        "kernel": "d3b::spconv_pairs_kernel (+ SIMT first layer): sparse middle encoder, %d launches/step" % n_enc,
        "bound": "hbm", "achieved": enc_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": enc_gbs / hbm_peak,
        "traffic": traffic.get("encoder_dram_bytes_per_step"), "peak_source": peak_src,
        "algorithmic_bytes_per_step": enc["bytes"], "flops_per_step": enc["flops"], "kernel_ms_per_step": enc_ms,
        "share_of_step": enc_ms / ms_step, "timing": timing,
        "note": "below the ridge by construction (2..32 flop/B); measured bound: L2 fp32 atomic (red.v4) issue rate and "
                "dependent L2 round trips, not DRAM (DESIGN.md 3.3)"}
    roofline = roofline_encoder
    extra = {}
    if per_step > n_enc:
        # dominant kernel of the step: spconv_tc_kernel<128> on the dense BEV grid, the six 3x3 128->128 RPN layers
        hw = B * 200 * 176
        flops_3x3 = 2 * hw * 9 * 128 * 128
        bytes_3x3 = hw * 128 * 4 * 2 + 9 * 128 * 128 * 4
        ms_3x3 = sum(t for i, t in enumerate(ev_ms) if n_enc <= i % per_step < n_enc + 6) / (6 * n_prof)
        tf = flops_3x3 / (ms_3x3 * 1e-3) / 1e12 if ms_3x3 > 0 else 0.0
        bev_flops = sum(2 * hw * k * ci * co for (k, ci, co) in [(9, 128, 128)] * 6 + [(1, 128, 128), (1, 128, 32)])
        roofline = {
            "kernel": "d3b::spconv_tc_kernel<128>: dense BEV conv3x3 128->128 (RPN), 6 of the %d BEV launches/step"
                      % (per_step - n_enc),
            "bound": "tensor", "achieved": tf, "peak": bf16_peak, "unit": "TFLOP/s", "frac": tf / bf16_peak,
            "traffic": traffic.get("bev3x3_dram_bytes_per_launch"), "peak_source": peak_src,
            "algorithmic_flops_per_launch": flops_3x3, "algorithmic_bytes_per_launch": bytes_3x3,
            "launch_ms": ms_3x3, "launches_per_step": 6, "share_of_step": 6 * ms_3x3 / ms_step, "timing": timing,
            "note": "fp32-equivalent flops (the reference runs this layer as fp32 cuDNN). The kernel reaches fp32 accuracy "
                    "with 3 TF32 MMAs per product (3xTF32) on a pipe whose TF32 rate is half the bf16 rate, so the "
                    "ceiling of this algorithm is peak/6 and tensor-pipe utilisation is 6 x frac = %.2f" % (6 * tf / bf16_peak),
            "bev_stack": {"launches_per_step": per_step - n_enc, "flops_per_step": bev_flops, "kernel_ms_per_step": bev_ms,
                          "share_of_step": bev_ms / ms_step}}
        extra["roofline_encoder"] = roofline_encoder

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    clouds_per_step = B * world
    value = clouds_per_step * args.steps / (ms * 1e-3)
    e2e_value = clouds_per_step * args.steps / (ms_e2e * 1e-3)
    d2h = int(out_pinned.numel() * 4) if out_pinned is not None else 0
    line = {
        "metric": "point-clouds/sec SECOND SpMiddleFHD @20k pts", "value": value, "unit": "clouds/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": "clouds/s", "h2d_bytes_per_step": B * N_POINTS * 4 * 4,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(round(launches_per_step * args.steps)), "gpu_launches_per_step": launches_per_step,
        "gpu_launches_note": "det3d_b200 kernels per step counted by d3b_launch_count() in an eager pass; the timed "
                             "region replays the same kernel nodes from a CUDA graph" if use_graph else "counted in an eager pass",
        "clocks": clocks, "roofline": roofline,
    }
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        from det3d_b200.core.anchor.anchor_generator import anchors_for_tasks
        from oracle.second_cpu import SecondCPU
        anchors = anchors_for_tasks(cfg.target_assigner, [1408, 1600, 40], cfg.assigner.out_size_factor)
        cpu = SecondCPU(cfg, {k: v.cpu() for k, v in pipe.model.state_dict().items()}, anchors)
        cpu.forward([clouds_np[0]])
        n_cpu = 3
        t0 = time.perf_counter()
        for i in range(n_cpu):
            cpu.forward([clouds_np[i % N_CLOUD_POOL]])
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n_cpu / dt, "unit": "clouds/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "%d full forwards of one 20k-point cloud through the CPU restatement of the "
                                          "reference path (oracle/second_cpu.py)" % n_cpu,
                                "stage_seconds": cpu.timings}
    _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

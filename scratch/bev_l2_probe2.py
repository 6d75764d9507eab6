"""Why do chained dense layers run 30 us when the same kernel re-run on fixed buffers runs 16-21 us?
Graph of 6 conv3x3 128->128 launches under different buffer patterns (release library, CUDA events)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from det3d_b200 import _lib
from det3d_b200.ops.spconv import conv16

torch.manual_seed(0)
dev = "cuda"
P = [conv16.Planes.from_f32(torch.randn(1, 200, 176, 128, device=dev) * (1.0 if i == 0 else 0.0)) for i in range(8)]
layers = [conv16.BevConv16(torch.randn(9, 128, 128, device=dev) * 0.03, 3, pad=1, relu=True, device=dev) for _ in range(6)]

PATTERNS = {
    "A fixed in P0, fixed out P1":            [(0, 1)] * 6,
    "B fixed in P0, out alternates P1/P2":    [(0, 1 + i % 2) for i in range(6)],
    "C ping-pong P0<->P1 (read prev out)":    [(i % 2, (i + 1) % 2) for i in range(6)],
    "D rotate 7 buffers (read prev out)":     [(i, i + 1) for i in range(6)],
    "E in alternates P0/P1 (clean), out P2":  [(i % 2, 2) for i in range(6)],
    "F read prev out, 3 buffers":             [(i % 3, (i + 1) % 3) for i in range(6)],
}

def bench(pattern):
    def chain():
        for L, (a, b) in zip(layers, pattern):
            L(P[a], out=P[b])
    chain(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        chain()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        chain()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 6)
    return float(np.median(ts[2:]))

for variant in (0, 1):
    _lib.lib().d3b_set_bev_variant(variant)
    for name, pat in PATTERNS.items():
        print("variant %d  %-42s %.1f us per layer" % (variant, name, bench(pat)), flush=True)
_lib.lib().d3b_set_bev_variant(0)

"""Six chained dense 3x3 128->128 layers (ping-pong planes) replayed from a CUDA graph under different L2 states:
warm, after a 256 MiB write (dirty lines), after a 256 MiB write followed by a 256 MiB read (clean lines)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from det3d_b200 import _lib
from det3d_b200.ops.spconv import conv16

torch.manual_seed(0)
dev = "cuda"
P = [conv16.Planes.from_f32(torch.randn(1, 200, 176, 128, device=dev)), conv16.Planes((1, 200, 176, 128), dev, zero=True)]
layers = [conv16.BevConv16(torch.randn(9, 128, 128, device=dev) * 0.03, 3, pad=1, relu=True, device=dev) for _ in range(6)]
fw = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fr = torch.ones(64 << 20, dtype=torch.float32, device=dev)
sink = torch.zeros(1, device=dev)

def chain():
    for i, L in enumerate(layers):
        L(P[i % 2], out=P[(i + 1) % 2])

def flush(mode):
    if mode in ("write", "write+read"):
        fw.zero_()
    if mode in ("read", "write+read"):
        sink.copy_(fr.sum().reshape(1))

for variant in (0, 1):
    _lib.lib().d3b_set_bev_variant(variant)
    chain(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        chain()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        chain()
    for mode in ("none", "write", "read", "write+read"):
        ts = []
        for _ in range(12):
            flush(mode)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 6)
        print("variant %d  L2 %-11s  us per layer (graph of 6): median %.1f  min %.1f" % (variant, mode, float(np.median(ts[2:])), min(ts[2:])), flush=True)
_lib.lib().d3b_set_bev_variant(0)

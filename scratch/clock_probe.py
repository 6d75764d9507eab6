"""SM clock actually seen by the dense kernel: cycles (clock64 stamps of CTA 0) vs wall time (globaltimer span of CTA 0),
isolated launches vs launches chained back to back in a CUDA graph (debug library)."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["D3B_LIB"] = os.path.join(ROOT, "det3d_b200", "lib", "libdet3d_b200_dbg.so")
import numpy as np, torch
from det3d_b200 import _lib
from det3d_b200.ops.spconv import conv16

def read():
    buf = (ctypes.c_longlong * (16 * 512))()
    _lib.lib().d3b_debug_trace_bevconv16(buf, 1)
    a = np.frombuffer(buf, dtype=np.int64).reshape(16, 512)
    sp = (ctypes.c_ulonglong * 4096)()
    _lib.lib().d3b_debug_cta_ns_bevconv16(sp)
    s = np.frombuffer(sp, dtype=np.uint64).reshape(8, 256, 2).astype(np.int64)
    slot = int(s[:, 0, 0].argmax())                      # latest launch
    cyc = int(a[10, 5] - a[4, 0])                        # MMA warp's first wait .. accumulator warps' last drain
    ns = int(s[slot, 0, 1] - s[slot, 0, 0])
    return cyc, ns

torch.manual_seed(0)
dev = "cuda"
P = [conv16.Planes.from_f32(torch.randn(1, 200, 176, 128, device=dev)), conv16.Planes((1, 200, 176, 128), dev, zero=True)]
layers = [conv16.BevConv16(torch.randn(9, 128, 128, device=dev) * 0.03, 3, pad=1, relu=True, device=dev) for _ in range(6)]
def chain():
    for i, L in enumerate(layers):
        L(P[i % 2], out=P[(i + 1) % 2])
for variant in (0, 1):
    _lib.lib().d3b_set_bev_variant(variant)
    chain(); torch.cuda.synchronize(); read()
    for gap_ms in (0, 1, 20):
        out = []
        for _ in range(5):
            layers[0](P[0], out=P[1]); torch.cuda.synchronize()
            out.append(read())
            time.sleep(gap_ms * 1e-3)
        cyc, ns = out[-1]
        print("variant %d isolated launches, %2d ms idle between: main loop %d cycles, CTA span %d ns -> ~%.2f GHz if the span were all main loop (ratio only)" % (variant, gap_ms, cyc, ns, cyc / ns), flush=True)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        chain()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        chain()
    for reps in (1, 4, 32):
        torch.cuda.synchronize(); time.sleep(0.05)
        for _ in range(reps):
            g.replay()
        torch.cuda.synchronize()
        cyc, ns = read()
        print("variant %d graph of 6, %2d replays back to back, last launch: main loop %d cycles, CTA span %d ns -> ratio %.2f" % (variant, reps, cyc, ns, cyc / ns), flush=True)
_lib.lib().d3b_set_bev_variant(0)

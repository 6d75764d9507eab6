"""Per-launch CTA spans (globaltimer) of kernels chained in a CUDA graph (debug library): where does the time between the
CTA lifetimes go?  Dense: 6 x conv3x3 128->128 ping-pong; sparse: 6 x SubM 64->64 on one rulebook."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["D3B_LIB"] = os.path.join(ROOT, "det3d_b200", "lib", "libdet3d_b200_dbg.so")
import numpy as np, torch
from det3d_b200 import _lib
from det3d_b200.ops.spconv import conv16, core
from test_conv16_gpu import _level

def report(tag, n_cta, n_launch, t_graph_us):
    buf = (ctypes.c_ulonglong * 4096)()
    getattr(_lib.lib(), "d3b_debug_cta_ns_" + tag)(buf)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(8, 256, 2)[:, :n_cta].astype(np.int64)
    cb = (ctypes.c_longlong * 4096)()
    getattr(_lib.lib(), "d3b_debug_cta_clk_" + tag)(cb)
    c = np.frombuffer(cb, dtype=np.int64).reshape(8, 256, 2)[:, :n_cta]
    order = np.argsort(a[:, :, 0].min(axis=1))[-n_launch:]
    t0 = a[order[0], :, 0].min()
    prev_end = None
    for s in order:
        st, en = a[s, :, 0] - t0, a[s, :, 1] - t0
        ghz = np.median((c[s, :, 1] - c[s, :, 0]) / np.maximum(a[s, :, 1] - a[s, :, 0], 1))
        print("   launch slot %d: first start %6d  last start %6d  first end %6d  last end %6d  median span %5d ns = %6d cycles (SM clock %.2f GHz)%s" % (
            s, st.min(), st.max(), en.min(), en.max(), int(np.median(en - st)), int(np.median(c[s, :, 1] - c[s, :, 0])), ghz,
            "" if prev_end is None else "   gap after previous launch's last end: %d ns" % (st.min() - prev_end)))
        prev_end = en.max()
    print("   graph replay (events): %.1f us total, %.1f us per launch" % (t_graph_us, t_graph_us / n_launch), flush=True)

def run_graph(fn, n_launch, tag, n_cta):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    report(tag, n_cta, n_launch, e0.elapsed_time(e1) * 1e3)

torch.manual_seed(0)
dev = "cuda"
P = [conv16.Planes.from_f32(torch.randn(1, 200, 176, 128, device=dev)), conv16.Planes((1, 200, 176, 128), dev, zero=True)]
layers = [conv16.BevConv16(torch.randn(9, 128, 128, device=dev) * 0.03, 3, pad=1, relu=True, device=dev) for _ in range(6)]
def dense_chain():
    for i, L in enumerate(layers):
        L(P[i % 2], out=P[(i + 1) % 2])
import time
def isolated(variant):
    _lib.lib().d3b_set_bev_variant(variant)
    for _ in range(3):
        layers[0](P[0], out=P[1]); torch.cuda.synchronize(); time.sleep(0.02)
    print("== dense, variant %d, ISOLATED launch (20 ms idle before it)" % variant)
    report("bevconv16", 143, 1, 0.0)
for variant in (0, 1):
    isolated(variant)
for pdl in (0, 1):
    _lib.lib().d3b_set_pdl(pdl)
    for variant in (0, 1):
        _lib.lib().d3b_set_bev_variant(variant)
        print("== dense chain, variant %d, pdl %d" % (variant, pdl))
        run_graph(dense_chain, 6, "bevconv16", 143)
_lib.lib().d3b_set_bev_variant(0)

n = 13000; lvl = _level(n, (11, 400, 352), 1, 3); rb = core.build_subm_rulebook(core.alloc_subm_rulebook(lvl, 3))
X = [conv16.Planes.from_f32(torch.randn(n, 64, device=dev)), conv16.Planes((n, 64), dev)]
cws = [conv16.ConvWeights16(torch.randn(27, 64, 64, device=dev) * 0.05, relu=True) for _ in range(6)]
def sparse_chain():
    for i, cw in enumerate(cws):
        conv16.sparse_conv16(X[i % 2], rb, cw, X[(i + 1) % 2])
for pdl in (0,):
    _lib.lib().d3b_set_pdl(pdl)
    print("== sparse chain (13000 rows = 102 tiles, 64->64), pdl %d" % pdl)
    run_graph(sparse_chain, 6, "spconv16", 102)
_lib.lib().d3b_set_pdl(1)

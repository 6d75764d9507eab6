"""Clock-stamp timeline of CTA 0 for one sparse 64->64 layer and one dense 3x3 layer (debug library)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["D3B_LIB"] = os.path.join(ROOT, "det3d_b200", "lib", "libdet3d_b200_dbg.so")
import numpy as np, torch
from det3d_b200 import _lib
from det3d_b200.ops.spconv import conv16, core
from test_conv16_gpu import _level
def dump(tag, names, n_slots):
    buf = (ctypes.c_longlong * (16 * 512))()
    getattr(_lib.lib(), "d3b_debug_trace_" + tag)(buf, 1)
    a = np.frombuffer(buf, dtype=np.int64).reshape(16, 512)
    t0 = a[a > 0].min()
    print("==", tag, "events:", names)
    for s in range(n_slots):
        print(s, " ".join("%7d" % (a[e, s] - t0 if a[e, s] > 0 else -1) for e in range(len(names))))
def spans(tag, n_cta):
    buf = (ctypes.c_ulonglong * 4096)()
    getattr(_lib.lib(), "d3b_debug_cta_ns_" + tag)(buf)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(8, 256, 2)[int(np.frombuffer(buf, dtype=np.uint64).reshape(8, 512).max(axis=1).argmax())][:n_cta].astype(np.int64)
    t0 = a[:, 0].min()
    dur = a[:, 1] - a[:, 0]
    print("== %s per-CTA spans (ns): start skew max %d, duration min/median/max %d/%d/%d, last end %d" % (
        tag, (a[:, 0] - t0).max(), dur.min(), int(np.median(dur)), dur.max(), (a[:, 1] - t0).max()))
    print("   slowest CTAs:", np.argsort(-dur)[:8].tolist(), "their durations:", np.sort(dur)[::-1][:8].tolist())
torch.manual_seed(0)
n = 18000; lvl = _level(n, (11, 400, 352), 1, 3); rb = core.build_subm_rulebook(core.alloc_subm_rulebook(lvl, 3))
x = conv16.Planes.from_f32(torch.randn(n, 64, device="cuda")); cw = conv16.ConvWeights16(torch.randn(27, 64, 64, device="cuda") * 0.05)
out = conv16.Planes((n, 64), "cuda")
for _ in range(2): conv16.sparse_conv16(x, rb, cw, out)
torch.cuda.synchronize(); buf = (ctypes.c_longlong * (16 * 512))(); _lib.lib().d3b_debug_trace_spconv16(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); conv16.sparse_conv16(x, rb, cw, out); e1.record(); torch.cuda.synchronize(); print("sparse kernel event ms", e0.elapsed_time(e1))
spans("spconv16", min(148, (n + 127) // 128))
dump("spconv16", ["g_pre", "g_free", "g_issued", "g_landed", "m_pre", "m_accfree", "m_full", "m_issued", "a_pre", "a_full", "a_done"], 30)
xin = conv16.Planes.from_f32(torch.randn(1, 200, 176, 128, device="cuda"))
layer = conv16.BevConv16(torch.randn(9, 128, 128, device="cuda") * 0.03, 3, pad=1, device="cuda")
o = conv16.Planes((1, 200, 176, 128), "cuda")
for _ in range(2): layer(xin, out=o)
torch.cuda.synchronize(); _lib.lib().d3b_debug_trace_bevconv16(buf, 1)
e0.record(); layer(xin, out=o); e1.record(); torch.cuda.synchronize(); print("bev kernel event ms", e0.elapsed_time(e1))
spans("bevconv16", 143)
dump("bevconv16", ["p_pre_a", "p_a_iss", "p_pre_b", "p_b_iss", "m_pre", "m_afull", "m_bfull", "m_issued", "a_pre", "a_full", "a_done"], 20)

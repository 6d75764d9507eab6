"""Clock-stamp timeline of CTA 0 of the dense 3x3 128->128 layer, both schedules (debug library)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["D3B_LIB"] = os.path.join(ROOT, "det3d_b200", "lib", "libdet3d_b200_dbg.so")
import numpy as np, torch
from det3d_b200 import _lib
from det3d_b200.ops.spconv import conv16
NAMES = ["p_pre_a", "p_a_iss", "p_pre_b", "p_b_iss", "m_pre", "m_afull", "m_bfull", "m_issued", "a_pre", "a_full", "a_done", "e_done"]
def dump(n_slots):
    buf = (ctypes.c_longlong * (16 * 512))()
    _lib.lib().d3b_debug_trace_bevconv16(buf, 1)
    a = np.frombuffer(buf, dtype=np.int64).reshape(16, 512)
    t0 = a[a > 0].min()
    print("events:", NAMES)
    for s in range(n_slots):
        print(s, " ".join("%7d" % (a[e, s] - t0 if a[e, s] > 0 else -1) for e in range(len(NAMES))))
def spans(n_cta):
    buf = (ctypes.c_ulonglong * 4096)()
    _lib.lib().d3b_debug_cta_ns_bevconv16(buf)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(8, 256, 2)[int(np.frombuffer(buf, dtype=np.uint64).reshape(8, 512).max(axis=1).argmax())][:n_cta].astype(np.int64)
    t0 = a[:, 0].min(); dur = a[:, 1] - a[:, 0]
    print("per-CTA spans (ns): start skew max %d, duration min/median/max %d/%d/%d, last end %d" % (
        (a[:, 0] - t0).max(), dur.min(), int(np.median(dur)), dur.max(), (a[:, 1] - t0).max()))
torch.manual_seed(0)
xin = conv16.Planes.from_f32(torch.randn(1, 200, 176, 128, device="cuda"))
layer = conv16.BevConv16(torch.randn(9, 128, 128, device="cuda") * 0.03, 3, pad=1, device="cuda")
o = conv16.Planes((1, 200, 176, 128), "cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
buf = (ctypes.c_longlong * (16 * 512))()
for variant in (0, 1):
    _lib.lib().d3b_set_bev_variant(variant)
    for _ in range(2): layer(xin, out=o)
    torch.cuda.synchronize(); _lib.lib().d3b_debug_trace_bevconv16(buf, 1)
    e0.record(); layer(xin, out=o); e1.record(); torch.cuda.synchronize()
    print("=== variant", variant, "bev kernel event ms", e0.elapsed_time(e1), flush=True)
    spans(143)
    dump(20)
    f = (ctypes.c_uint * 8)(); _lib.lib().d3b_debug_fault_bevconv16(f); print("faults", list(f))

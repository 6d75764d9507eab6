"""Development aid: run one dense BEV 3x3 layer (spconv_tc_kernel<128>) and one sparse 64->64 pair layer on the trace
build and print the per-slot timeline of CTA 0 (SM clocks)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from det3d_b200 import _lib

_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtrace.so")
from det3d_b200.ops.spconv import bev, core   # noqa: E402
from det3d_b200.utils.synthetic import lidar_like_cloud   # noqa: E402

L = _lib.lib()
handle = C.CDLL(_lib.LIB_PATH)


def read_trace(n_slots):
    buf = (C.c_longlong * (16 * 1024))()
    torch.cuda.synchronize()
    assert handle.d3b_debug_trace(buf, 16 * 1024) == 0
    return np.frombuffer(buf, dtype=np.int64).reshape(-1, 16)[:n_slots].copy()


def show(tr, cols, title, rows=48):
    t0 = tr[:, cols][tr[:, cols] > 0].min()
    print("==", title, "(cycles since first stamp; 0 = not recorded)")
    print("slot " + " ".join("%8s" % c for c in ["ld_issue", "stg_free", "stored", "arrived", "mma_wait", "mma_go", "epi_wait", "epi_go", "epi_done"][:len(cols)]))
    for i in range(min(rows, tr.shape[0])):
        print("%4d " % i + " ".join("%8d" % (tr[i, c] - t0 if tr[i, c] > 0 else 0) for c in cols))
    go = tr[:, 5][tr[:, 5] > 0]
    if go.size > 8:
        d = np.diff(np.sort(go))
        print("MMA issue interval: median %d cycles, mean %d (floor: 768 for N=128, 384 for N=64)" % (np.median(d), d.mean()))


torch.manual_seed(0)
dev = "cuda"
# ---- dense BEV layer ----
grid = bev.BevGrid(1, 200, 176, dev)
rb = grid.rulebook(3, 3, 1, 1)
w = torch.randn(9, 128, 128, device=dev) * 0.05
cw = core.ConvWeights(w, scale=torch.ones(128, device=dev), shift=torch.zeros(128, device=dev), relu=True, algo=_lib.ALGO_TC)
x = torch.randn(grid.n_rows, 128, device=dev)
y = torch.empty_like(x)
for _ in range(3):
    core.sparse_conv(x, rb, cw, y)
handle.d3b_debug_trace_clear()
core.sparse_conv(x, rb, cw, y)
show(read_trace(80), [0, 1, 2, 3, 4, 5], "spconv_tc_kernel<128>, dense 3x3 128->128, CTA 0")

# ---- sparse pair layer: level-2-like sites from a lidar cloud ----
pts = lidar_like_cloud(20000, [0, -40.0, -3.0, 70.4, 40.0, 1.0], 4, 0)
vs = np.array([0.2, 0.2, 0.4], np.float32)
c = np.floor((pts[:, :3] - np.array([0, -40, -3], np.float32)) / vs).astype(np.int32)
c = np.unique(c, axis=0)
coors = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c[:, ::-1]], 1).astype(np.int32)
n = coors.shape[0]
lvl = core.level_from_coors(torch.from_numpy(coors).to(dev), (11, 400, 352), 1)
rbs = core.build_subm_rulebook(core.alloc_subm_rulebook(lvl, 3), with_pairs=True)
print("sites", n, "pairs", int(rbs.pairs[2].sum()))
w2 = torch.randn(27, 64, 64, device=dev) * 0.05
cw2 = core.ConvWeights(w2, algo=_lib.ALGO_TC_PAIRS)
xs = torch.randn(n, 64, device=dev)
ys = torch.zeros(n, 64, device=dev)
act = (None, torch.ones(64, device=dev), torch.zeros(64, device=dev), True)
for _ in range(3):
    core.sparse_conv(xs, rbs, cw2, ys, in_act=act)
handle.d3b_debug_trace_clear()
core.sparse_conv(xs, rbs, cw2, ys, in_act=act)
show(read_trace(40), [0, 1, 2, 3, 4, 5, 6, 7, 8], "spconv_pairs_kernel<64>, subm 64->64, CTA 0", rows=40)

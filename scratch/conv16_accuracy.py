"""Scratch: error of the FP16x3 kernels vs an fp64 reference, next to fp32 cuDNN / fp32 torch (bias = mean signed relative error)."""
import os, sys, json
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from det3d_b200.ops.spconv import conv16, core
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
def stats(got, want64):
    d = got.double() - want64
    big = want64.abs() > 0.1 * want64.abs().max()
    return dict(max_abs=float(d.abs().max()), rms=float(d.pow(2).mean().sqrt()), mean_signed_rel=float((d[big] / want64[big]).mean()),
                max_rel_big=float((d[big] / want64[big]).abs().max()), ref_max=float(want64.abs().max()))
out = {}
torch.manual_seed(0)
# dense 3x3 128->128, positive inputs (post-ReLU like)
x = torch.relu(torch.randn(1, 128, 200, 176, device="cuda")); w = torch.randn(128, 128, 3, 3, device="cuda") * (1 / np.sqrt(1152 * 0.3))
want = F.conv2d(x.double(), w.double(), padding=1)
layer = conv16.BevConv16(w.permute(2, 3, 1, 0).reshape(9, 128, 128), 3, pad=1, device="cuda")
o32 = torch.zeros((1, 200, 176, 128), device="cuda")
layer(conv16.Planes.from_f32(x.permute(0, 2, 3, 1).contiguous()), out_f32=o32)
out["bev3x3_fp16x3"] = stats(o32.permute(0, 3, 1, 2), want)
out["bev3x3_cudnn_fp32"] = stats(F.conv2d(x, w, padding=1), want)
# sparse 64->64
from test_conv16_gpu import _level, _ref_conv
n = 20000; lvl = _level(n, (9, 80, 72), 1, 1); rb = core.build_subm_rulebook(core.alloc_subm_rulebook(lvl, 3))
feat = torch.relu(torch.randn(n, 64, device="cuda")); ws = torch.randn(27, 64, 64, device="cuda") * 0.05
want = _ref_conv(feat, rb.nbr, ws, n)
o = torch.zeros((n, 64), device="cuda")
conv16.sparse_conv16(conv16.Planes.from_f32(feat), rb, conv16.ConvWeights16(ws), None, out_f32=o)
out["sparse64_fp16x3"] = stats(o, want)
o2 = torch.zeros((n, 64), device="cuda")
for k in range(27):
    idx = rb.nbr[k, :n].long(); ok = idx >= 0
    o2[ok] += feat[idx[ok]] @ ws[k]
out["sparse64_torch_fp32"] = stats(o2, want)
from det3d_b200 import _lib
cw = core.ConvWeights(ws, algo=_lib.ALGO_TC); o3 = torch.zeros((n, 64), device="cuda"); core.sparse_conv(feat, rb, cw, o3)
out["sparse64_tf32x3_os"] = stats(o3, want)
for k, v in out.items(): print(k, json.dumps(v))

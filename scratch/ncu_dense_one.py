"""Two launches of each dense 3x3 schedule (SECOND RPN shape) for an `ncu --set full` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from det3d_b200 import _lib
from det3d_b200.ops.spconv import conv16
torch.manual_seed(0)
x = conv16.Planes.from_f32(torch.randn(1, 200, 176, 128, device="cuda"))
o = conv16.Planes((1, 200, 176, 128), "cuda", zero=True)
layer = conv16.BevConv16(torch.randn(9, 128, 128, device="cuda") * 0.03, 3, pad=1, relu=True, device="cuda")
for variant in (0, 1):
    _lib.lib().d3b_set_bev_variant(variant)
    for _ in range(2):
        layer(x, out=o)
torch.cuda.synchronize()

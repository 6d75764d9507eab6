"""Channel-stationary (N256) dense 3x3 kernel vs the pixel-stationary one: accuracy vs float64, bit equality, time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from det3d_b200 import _lib
from det3d_b200.ops.spconv import conv16


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def run(b, h, w, c_in, c_out, reps=30):
    torch.manual_seed(1)
    x = torch.randn((b, c_in, h, w), device="cuda")
    wt = torch.randn((c_out, c_in, 3, 3), device="cuda") * (1.0 / np.sqrt(9 * c_in * 0.3))
    scale = torch.rand(c_out, device="cuda") + 0.5
    shift = torch.randn(c_out, device="cuda") * 0.1
    want = F.conv2d(x.double(), wt.double(), padding=1)
    want = nhwc(torch.relu(want * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)).float())
    layer = conv16.BevConv16(wt.permute(2, 3, 1, 0).reshape(9, c_in, c_out), 3, stride=1, pad=1, scale=scale, shift=shift,
                             relu=True, device="cuda")
    xin = conv16.Planes.from_f32(nhwc(x))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for variant in (0, 1):
        _lib.lib().d3b_set_bev_variant(variant)
        out = conv16.Planes((b, h, w, c_out), "cuda", zero=True)
        layer(xin, out=out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            layer(xin, out=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        # back to back (warm L2, PDL chain of 6 like the RPN)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(60):
            layer(xin, out=out)
        e1.record()
        torch.cuda.synchronize()
        res[variant] = (out.to_f32(), float(np.median(ts)), e0.elapsed_time(e1) * 1e3 / 60)
    err = [float((res[v][0] - want).abs().max()) for v in (0, 1)]
    diff = float((res[0][0] - res[1][0]).abs().max())
    fl = 2.0 * b * h * w * 9 * c_in * c_out
    print("B%d %dx%d %d->%d | err v0 %.2e v1 %.2e | v0-v1 %.2e equal %s | cold us v0 %.1f v1 %.1f | chained us v0 %.1f v1 %.1f | TF/s(fp32-eq) v0 %.0f v1 %.0f"
          % (b, h, w, c_in, c_out, err[0], err[1], diff, diff == 0.0, res[0][1], res[1][1], res[0][2], res[1][2],
             fl / res[0][2] / 1e6, fl / res[1][2] / 1e6), flush=True)


if __name__ == "__main__":
    run(1, 200, 176, 128, 128)
    run(2, 37, 29, 64, 128)
    run(1, 40, 40, 256, 256)
    run(8, 248, 216, 128, 128, reps=5)
    run(4, 128, 128, 256, 256, reps=5)
    _lib.lib().d3b_set_bev_variant(0)

"""Run ON THE GPU BOX (gpurun): produces tests/golden-format iou3d fixtures from the REFERENCE
CUDA kernel (oracle/_ref/libiou3d_ref.so, compiled from det3d/ops/iou3d/src/iou3d_kernel.cu)
into gpurun_out/golden/; they are then committed under tests/golden/."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from det3d_b200.utils.synthetic import nms_boxes_xyxyr  # noqa: E402
from oracle import iou3d_ref  # noqa: E402


def main():
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    for name, n, seed, clustered, thr, with_iou in (("clustered_2k_t010", 2000, 11, True, 0.10, True),
                                                     ("clustered_2k_t001", 2000, 12, True, 0.01, False),
                                                     ("uniform_3k_t050", 3000, 13, False, 0.50, False)):
        boxes, scores = nms_boxes_xyxyr(n, seed, clustered, extent=60.0)
        order = np.argsort(-scores, kind="stable")
        b = np.ascontiguousarray(boxes[order])
        dev = torch.from_numpy(b).cuda()
        keep = iou3d_ref.nms(dev, thr)
        d = dict(boxes_sorted=b, thresh=np.float32(thr), keep=keep)
        if with_iou:
            d["iou"] = iou3d_ref.iou_matrix(dev[:256], dev[:256]).cpu().numpy()
        np.savez_compressed(os.path.join(out, "iou3d_%s.npz" % name), **d)
        print(name, n, "kept", keep.shape[0])


if __name__ == "__main__":
    main()

"""Golden vectors for the PointPillars reader, produced by the REFERENCE modules themselves.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden_pillars.py
Loads det3d/models/readers/pillar_encoder.py (PillarFeatureNet :58-155, PointPillarsScatter :158-211) by file
path under a stub `det3d.models` package: importing the real package would pull spconv / the compiled ops
(SURVEY 8c).  The only stubbed behaviour is `build_norm_layer` for "BN1d" (-> nn.BatchNorm1d(eps, momentum),
what det3d/models/utils/norm.py's table maps it to); `get_paddings_indicator` etc. come from the reference's own
det3d/models/utils/misc.py, and the pillars from the reference's numba voxelizer.
Output: pillars_kitti_3k.npz (points, module parameters, reference pillar features + pseudo-image).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
PILLAR = dict(vs=[0.16, 0.16, 4.0], pcr=[0, -39.68, -3, 69.12, 39.68, 1], max_points=100, max_voxels=12000)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_reader():
    class _Reg:
        def register_module(self, cls):
            return cls

    for pkg in ("det3d", "det3d.models", "det3d.models.readers"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    misc = _load("det3d.models.utils.misc", REF + "/det3d/models/utils/misc.py")
    utils = types.ModuleType("det3d.models.utils")
    for k in ("Empty", "change_default_args", "get_paddings_indicator"):
        setattr(utils, k, getattr(misc, k))

    def build_norm_layer(cfg, num_features, postfix=""):
        assert cfg["type"] == "BN1d"
        return "bn" + str(postfix), nn.BatchNorm1d(num_features, eps=cfg["eps"], momentum=cfg["momentum"])

    utils.build_norm_layer = build_norm_layer
    sys.modules["det3d.models.utils"] = utils
    sys.modules["det3d.models"].utils = utils
    sys.modules["det3d.models.builder"] = types.ModuleType("det3d.models.builder")
    sys.modules["det3d.models"].builder = sys.modules["det3d.models.builder"]
    reg = types.ModuleType("det3d.models.registry")
    reg.BACKBONES, reg.READERS = _Reg(), _Reg()
    sys.modules["det3d.models.registry"] = reg
    return _load("det3d.models.readers.pillar_encoder", REF + "/det3d/models/readers/pillar_encoder.py")


def main():
    ref = load_reference_reader()
    vox = _load("ref_pc_ops", REF + "/det3d/ops/point_cloud/point_cloud_ops.py")
    rng = np.random.default_rng(11)
    pcr = np.array(PILLAR["pcr"], np.float32)
    n = 3000
    pts = np.empty((n, 4), np.float32)
    pts[:, :3] = rng.uniform(pcr[:3], pcr[3:], (n, 3))
    pts[:, 3] = rng.uniform(0, 1, n)
    # two crowded pillars: one past the 100-point cap (no padded slot), one just under it
    pts[:150, :2] = np.array([9.94, 0.34], np.float32) + rng.uniform(0, 0.1, (150, 2)).astype(np.float32)
    pts[150:249, :2] = np.array([29.94, -7.66], np.float32) + rng.uniform(0, 0.1, (99, 2)).astype(np.float32)
    rng.shuffle(pts)
    voxels, coors, num = vox.points_to_voxel(pts, np.array(PILLAR["vs"], np.float32), pcr, PILLAR["max_points"], True,
                                             PILLAR["max_voxels"])
    assert num.max() == 100 and (num == 99).any()

    torch.manual_seed(5)
    net = ref.PillarFeatureNet(num_input_features=4, num_filters=[64], with_distance=False,
                               voxel_size=PILLAR["vs"], pc_range=PILLAR["pcr"], norm_cfg=None)
    bn = net.pfn_layers[0].norm
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)          # mixed signs: relu(shift) of padded slots matters for some channels
        bn.running_mean.uniform_(-0.3, 0.3)
        bn.running_var.uniform_(0.5, 2.0)
    net.eval()
    coors_b = np.concatenate([np.zeros((coors.shape[0], 1), np.int32), coors], axis=1)   # batch index 0
    with torch.no_grad():
        feats = net(torch.from_numpy(voxels), torch.from_numpy(num), torch.from_numpy(coors_b))
        scatter = ref.PointPillarsScatter(num_input_features=64)
        grid = np.round((pcr[3:] - pcr[:3]) / np.array(PILLAR["vs"], np.float32)).astype(np.int64)
        canvas = scatter(feats, torch.from_numpy(coors_b), 1, grid)      # input_shape = grid size in x, y, z order
    sd = {k: v.numpy() for k, v in net.state_dict().items()}
    nz = canvas.numpy().reshape(64, -1)
    cols = np.nonzero(np.abs(nz).sum(0))[0]
    np.savez_compressed(os.path.join(HERE, "pillars_kitti_3k.npz"), points=pts, coors=coors_b, num_points=num,
                        features=feats.numpy(), canvas_shape=np.array(canvas.shape), canvas_cols=cols,
                        canvas_checksum=np.array([nz.astype(np.float64).sum(), (nz.astype(np.float64) ** 2).sum()]),
                        **{"sd." + k: v for k, v in sd.items()})
    print("pillars", voxels.shape, "features", tuple(feats.shape), "canvas", tuple(canvas.shape), "nonzero cols", cols.size)


if __name__ == "__main__":
    main()

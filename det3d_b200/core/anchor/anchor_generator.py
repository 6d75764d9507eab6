"""Range anchors for the inference path (anchor *generation* only).

`create_anchors_3d_range` restates det3d/core/bbox/box_np_ops.py:733-805 (including
its quirk of using the x stride for the y half-cell offset); `AnchorGeneratorRange`
keeps the constructor / `generate(feature_map_size)` / `num_anchors_per_localization`
of det3d/core/anchor/anchor_generator.py:64-125.  Target assignment is training-only
and out of scope.  `anchors_for_tasks` reproduces the flattening the reference's
AssignTarget pipeline performs per sample (datasets/pipelines/preprocess.py:355-378,
core/anchor/target_assigner.py:144-165) -- done ONCE here and cached on the device
instead of per sample on the CPU.
"""
import numpy as np


def create_anchors_3d_range(feature_size, anchor_range, sizes=(1.6, 3.9, 1.56), rotations=(0, np.pi / 2),
                            velocities=None, dtype=np.float32):
    """feature_size [D,H,W] (zyx) -> anchors [D, H, W, num_sizes, num_rots, 7 (+2)]."""
    rng = np.array(anchor_range, dtype)
    d, h, w = (int(v) for v in feature_size)
    stride = (rng[3] - rng[0]) / w
    zc = np.linspace(rng[2], rng[5], d, dtype=dtype)
    yc = np.linspace(rng[1], rng[4], h, endpoint=False, dtype=dtype) + stride / 2
    xc = np.linspace(rng[0], rng[3], w, endpoint=False, dtype=dtype) + stride / 2
    rots = np.array(rotations, dtype=dtype)
    combos = np.reshape(np.array(sizes, dtype=dtype), [-1, 3])
    if velocities is not None:
        vel = np.array(velocities, dtype=dtype).reshape([-1, 2])
        combos = np.hstack([combos, vel]).reshape([-1, 5])
    ns, nr, nc = combos.shape[0], rots.shape[0], combos.shape[1]
    out = np.empty((d, h, w, ns, nr, 3 + nc + 1), dtype=dtype)
    out[..., 0] = xc[None, None, :, None, None]
    out[..., 1] = yc[None, :, None, None, None]
    out[..., 2] = zc[:, None, None, None, None]
    out[..., 3:3 + nc] = combos[None, None, None, :, None, :]
    out[..., 3 + nc] = rots[None, None, None, None, :]
    return out


class AnchorGeneratorRange:
    def __init__(self, anchor_ranges, sizes=(1.6, 3.9, 1.56), rotations=(0, np.pi / 2), velocities=None,
                 class_name=None, match_threshold=-1, unmatch_threshold=-1, dtype=np.float32):
        self._sizes, self._anchor_ranges, self._rotations = sizes, anchor_ranges, rotations
        self._velocities, self._dtype, self._class_name = velocities, dtype, class_name
        self._match_threshold, self._unmatch_threshold = match_threshold, unmatch_threshold
        self._anchors = None

    class_name = property(lambda self: self._class_name)
    match_threshold = property(lambda self: self._match_threshold)
    unmatch_threshold = property(lambda self: self._unmatch_threshold)

    @property
    def num_anchors_per_localization(self):
        return len(self._rotations) * np.array(self._sizes).reshape([-1, 3]).shape[0]

    @property
    def ndim(self):
        return self._anchors.shape[-1]

    def generate(self, feature_map_size):
        self._anchors = create_anchors_3d_range(feature_map_size, self._anchor_ranges, self._sizes,
                                                self._rotations, self._velocities, self._dtype)
        return self._anchors


def anchors_for_tasks(target_assigner_cfg, grid_size, out_size_factor):
    """-> list (one per task) of float32 [num_anchors, box_ndim] arrays, reference order.

    target_assigner_cfg: the config's `target_assigner` dict (anchor_generators + tasks).
    """
    fmap = [int(v) for v in (np.asarray(grid_size)[:2] // out_size_factor)]
    feature_map_size = [*fmap, 1][::-1]  # [1, H, W]
    gens = {}
    for ag in target_assigner_cfg["anchor_generators"]:
        if ag["type"] != "anchor_generator_range":
            raise NotImplementedError("only anchor_generator_range is used by the Det3D configs in scope")
        gens[ag["class_name"]] = AnchorGeneratorRange(
            anchor_ranges=ag["anchor_ranges"], sizes=ag["sizes"], rotations=ag["rotations"],
            velocities=ag.get("velocities"), class_name=ag["class_name"],
            match_threshold=ag.get("matched_threshold", -1), unmatch_threshold=ag.get("unmatched_threshold", -1))
    per_task = []
    for task in target_assigner_cfg["tasks"]:
        parts = []
        for name in task["class_names"]:
            a = gens[name].generate(feature_map_size)
            parts.append(a.reshape([*a.shape[:3], -1, a.shape[-1]]))
        anchors = np.concatenate(parts, axis=-2)
        per_task.append(np.ascontiguousarray(anchors.reshape([-1, anchors.shape[-1]])))
    return per_task

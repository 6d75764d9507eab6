"""`Reformat` under the reference's registry name (det3d/datasets/pipelines/formating.py:13-58), val/test modes."""
from ..registry import PIPELINES


@PIPELINES.register_module
class Reformat(object):
    def __init__(self, **kwargs):
        pass

    def __call__(self, res, info):
        voxels = res["lidar"]["voxels"]
        data_bundle = dict(
            metadata=res["metadata"],
            points=res["lidar"]["points"],
            voxels=voxels["voxels"],
            shape=voxels["shape"],
            num_points=voxels["num_points"],
            num_voxels=voxels["num_voxels"],
            coordinates=voxels["coordinates"],
            anchors=res["lidar"]["targets"]["anchors"],
        )
        calib = res.get("calib", None)
        if calib:
            data_bundle["calib"] = calib
        if res["mode"] == "train":
            raise NotImplementedError("det3d_b200 covers the inference path: Reformat(mode='train') is out of scope")
        if res["mode"] != "test" and "annotations" in res["lidar"]:
            data_bundle.update(annos=res["lidar"]["annotations"])
        return data_bundle, info

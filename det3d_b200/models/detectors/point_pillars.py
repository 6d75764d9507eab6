"""PointPillars detector (det3d/models/detectors/point_pillars.py:5-54): PillarFeatureNet reader ->
PointPillarsScatter -> RPN -> MultiGroupHead."""
from ..registry import DETECTORS
from .voxelnet import SingleStageDetector


@DETECTORS.register_module
class PointPillars(SingleStageDetector):
    def extract_feat(self, data):
        n_dev = data.get("n_dev")
        kw = {} if n_dev is None else {"n_dev": n_dev}
        feats = self.reader(data["features"], data["num_voxels"], data["coors"], **kw)
        x = self.backbone(feats, data["coors"], data["batch_size"], data["input_shape"], **kw)
        return self.neck(x) if self.with_neck else x

    def forward(self, example, return_loss=True, **kwargs):
        num_voxels = example["num_voxels"]
        data = dict(features=example["voxels"], num_voxels=example["num_points"], coors=example["coordinates"],
                    batch_size=len(num_voxels), input_shape=example["shape"][0], n_dev=example.get("n_voxels_dev"))
        preds = self.bbox_head(self.extract_feat(data))
        if return_loss:
            return self.bbox_head.loss(example, preds)
        if kwargs.get("device_output", False):
            return self.bbox_head.predict_device(example, preds, self.test_cfg)
        return self.bbox_head.predict(example, preds, self.test_cfg)

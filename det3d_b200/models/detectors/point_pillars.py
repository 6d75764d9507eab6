"""PointPillars detector (det3d/models/detectors/point_pillars.py:5-54): PillarFeatureNet reader ->
PointPillarsScatter -> RPN -> MultiGroupHead."""
from ..registry import DETECTORS
from .voxelnet import SingleStageDetector, _FusedBevMixin


@DETECTORS.register_module
class PointPillars(_FusedBevMixin, SingleStageDetector):
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
        bev = self.fused_bev() if not return_loss else None
        if bev is not None and self.math == "fp16x3":
            kw = {} if data["n_dev"] is None else {"n_dev": data["n_dev"]}
            feats = self.reader(data["features"], data["num_voxels"], data["coors"], **kw)
            ovf = self.overflow_flag(feats.device)
            planes = self.backbone.forward_planes(feats, data["coors"], data["batch_size"], data["input_shape"], **kw)
            preds = bev.run(planes, overflow=ovf)
        else:
            preds = self.bbox_head(self.extract_feat(data))
        if return_loss:
            return self.bbox_head.loss(example, preds)
        if kwargs.get("device_output", False):
            return self.bbox_head.predict_device(example, preds, self.test_cfg)
        return self.bbox_head.predict(example, preds, self.test_cfg)

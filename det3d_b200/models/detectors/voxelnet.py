"""VoxelNet single-stage detector (det3d/models/detectors/voxelnet.py:5-52,
single_stage.py:9-36): reader -> sparse middle encoder -> RPN neck -> head."""
from torch import nn

from .. import builder
from ..registry import DETECTORS


@DETECTORS.register_module
class SingleStageDetector(nn.Module):
    def __init__(self, reader, backbone, neck=None, bbox_head=None, train_cfg=None, test_cfg=None,
                 pretrained=None):
        super().__init__()
        self.reader = builder.build_reader(reader)
        self.backbone = builder.build_backbone(backbone)
        if neck is not None:
            self.neck = builder.build_neck(neck)
        self.bbox_head = builder.build_head(bbox_head)
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg

    @property
    def with_neck(self):
        return hasattr(self, "neck") and self.neck is not None

    def init_weights(self, pretrained=None):
        self.backbone.init_weights(pretrained=pretrained)
        if self.with_neck:
            self.neck.init_weights()
        self.bbox_head.init_weights()


@DETECTORS.register_module
class VoxelNet(SingleStageDetector):
    def extract_feat(self, data):
        feats = self.reader(data["features"], data["num_voxels"])
        x = self.backbone(feats, data["coors"], data["batch_size"], data["input_shape"],
                          **({"n_dev": data["n_dev"]} if data.get("n_dev") is not None else {}))
        return self.neck(x) if self.with_neck else x

    def forward(self, example, return_loss=True, **kwargs):
        num_voxels = example["num_voxels"]
        data = dict(features=example["voxels"], num_voxels=example["num_points"],
                    coors=example["coordinates"], batch_size=len(num_voxels),
                    input_shape=example["shape"][0], n_dev=example.get("n_voxels_dev"))
        preds = self.bbox_head(self.extract_feat(data))
        if return_loss:
            return self.bbox_head.loss(example, preds)
        if kwargs.get("device_output", False):
            return self.bbox_head.predict_device(example, preds, self.test_cfg)
        return self.bbox_head.predict(example, preds, self.test_cfg)

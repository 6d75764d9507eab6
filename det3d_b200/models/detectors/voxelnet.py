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
    #: route the dense RPN + heads through the channels-last tensor-core path when its shape allows
    use_fused_bev = True
    _bev = None

    def fused_bev(self):
        """FusedBevStack for (neck, bbox_head) when the RPN is the stride-1 SECOND shape, else None."""
        if self._bev is None:
            from det3d_b200.ops.spconv.bev import FusedBevStack, rpn_is_fusable
            ok = (self.with_neck and hasattr(self.backbone, "forward_rows") and rpn_is_fusable(self.neck)
                  and not self.training)
            self._bev = FusedBevStack(self.neck, self.bbox_head) if ok else False
        return self._bev or None

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
        bev = self.fused_bev() if (self.use_fused_bev and not return_loss) else None
        if bev is not None:
            feats = self.reader(data["features"], data["num_voxels"])
            rows, (b, h, w) = self.backbone.forward_rows(feats, data["coors"], data["batch_size"],
                                                          data["input_shape"], n_dev=data["n_dev"])
            preds = bev.run(rows, b, h, w)
        else:
            preds = self.bbox_head(self.extract_feat(data))
        if return_loss:
            return self.bbox_head.loss(example, preds)
        if kwargs.get("device_output", False):
            return self.bbox_head.predict_device(example, preds, self.test_cfg)
        return self.bbox_head.predict(example, preds, self.test_cfg)

"""build_* helpers (det3d/models/builder.py:16-53): a cfg dict -> module via its registry;
a list of cfgs -> nn.Sequential."""
from torch import nn

from det3d_b200.utils import build_from_cfg

from .registry import BACKBONES, DETECTORS, HEADS, LOSSES, NECKS, READERS, ROI_EXTRACTORS, SHARED_HEADS


def build(cfg, registry, default_args=None):
    if isinstance(cfg, list):
        return nn.Sequential(*[build_from_cfg(c, registry, default_args) for c in cfg])
    return build_from_cfg(cfg, registry, default_args)


def _builder(registry):
    return lambda cfg: build(cfg, registry)


build_reader = _builder(READERS)
build_backbone = _builder(BACKBONES)
build_neck = _builder(NECKS)
build_roi_extractor = _builder(ROI_EXTRACTORS)
build_shared_head = _builder(SHARED_HEADS)
build_head = _builder(HEADS)
build_loss = _builder(LOSSES)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    return build(cfg, DETECTORS, dict(train_cfg=train_cfg, test_cfg=test_cfg))

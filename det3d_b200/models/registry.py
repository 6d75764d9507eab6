"""Model registries, same names as det3d/models/registry.py:3-10."""
from det3d_b200.utils import Registry

READERS = Registry("reader")
BACKBONES = Registry("backbone")
NECKS = Registry("neck")
ROI_EXTRACTORS = Registry("roi_extractor")
SHARED_HEADS = Registry("shared_head")
HEADS = Registry("head")
LOSSES = Registry("loss")
DETECTORS = Registry("detector")

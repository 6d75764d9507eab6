from . import losses  # noqa: F401  (registers the loss placeholders)
from .backbones import *  # noqa: F401,F403
from .bbox_heads import *  # noqa: F401,F403
from .builder import (build_backbone, build_detector, build_head, build_loss, build_neck, build_reader,
                      build_roi_extractor, build_shared_head)
from .detectors import *  # noqa: F401,F403
from .necks import *  # noqa: F401,F403
from .readers import *  # noqa: F401,F403
from .registry import BACKBONES, DETECTORS, HEADS, LOSSES, NECKS, READERS, ROI_EXTRACTORS, SHARED_HEADS

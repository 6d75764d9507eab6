"""The model registries under the reference's names (det3d/models/registry.py:3-10): configs and third-party code
look components up through `READERS`, `BACKBONES`, `NECKS`, `ROI_EXTRACTORS`, `SHARED_HEADS`, `HEADS`, `LOSSES`,
`DETECTORS`."""
from det3d_b200.utils.registry import Registry

_KINDS = ("reader", "backbone", "neck", "roi_extractor", "shared_head", "head", "loss", "detector")
globals().update({kind.upper() + "S": Registry(kind) for kind in _KINDS})
LOSSES = globals().pop("LOSSS")          # the one irregular plural
__all__ = [kind.upper() + "S" for kind in _KINDS if kind != "loss"] + ["LOSSES"]

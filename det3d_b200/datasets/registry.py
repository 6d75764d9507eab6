"""Dataset-side registries under the reference's names (det3d/datasets/registry.py:3-4)."""
from det3d_b200.utils.registry import Registry

DATASETS, PIPELINES = (Registry(kind) for kind in ("dataset", "pipeline"))

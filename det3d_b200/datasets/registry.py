"""det3d/datasets/registry.py:3-4."""
from det3d_b200.utils.registry import Registry

DATASETS = Registry("dataset")
PIPELINES = Registry("pipeline")

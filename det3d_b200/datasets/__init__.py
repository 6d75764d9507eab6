from .registry import DATASETS, PIPELINES
from . import pipelines

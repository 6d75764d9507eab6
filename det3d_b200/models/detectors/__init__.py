from .point_pillars import PointPillars
from .voxelnet import SingleStageDetector, VoxelNet

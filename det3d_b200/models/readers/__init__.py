from .pillar_encoder import PFNLayer, PillarFeatureNet, PointPillarsScatter
from .voxel_encoder import VoxelFeatureExtractorV3

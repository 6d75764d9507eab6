from .voxelnet import SingleStageDetector, VoxelNet

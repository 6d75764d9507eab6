from .voxel_generator import VoxelGenerator

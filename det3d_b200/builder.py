"""`det3d.builder` entry points the reference configs import at load time
(examples/second/configs/kitti_car_vfev3_spmiddlefhd_rpn1_mghead_syncbn.py:4,73)."""
from det3d_b200.core.bbox.box_coders import GroundBox3dCoderTorch
from det3d_b200.core.input.voxel_generator import VoxelGenerator


def build_box_coder(box_coder_config):
    """dict(type="ground_box3d_coder", n_dim, linear_dim, encode_angle_vector[, norm_velo]) -> coder.
    Mirrors det3d/builder.py:399-433 for the coder the Det3D configs use."""
    cfg = box_coder_config
    kind = cfg["type"]
    if kind == "ground_box3d_coder":
        return GroundBox3dCoderTorch(cfg["linear_dim"], cfg["encode_angle_vector"],
                                     n_dim=cfg.get("n_dim", 9), norm_velo=cfg.get("norm_velo", False))
    raise ValueError("unknown box_coder type")


def build_voxel_generator(voxel_config):
    """dict(range, voxel_size, max_points_in_voxel, max_voxel_num) -> VoxelGenerator (det3d/builder.py:26-35)."""
    return VoxelGenerator(voxel_size=voxel_config["voxel_size"], point_cloud_range=voxel_config["range"],
                          max_num_points=voxel_config["max_points_in_voxel"],
                          max_voxels=voxel_config.get("max_voxel_num", 20000))

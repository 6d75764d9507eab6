# PointPillars (PillarFeatureNet + PointPillarsScatter + RPN[3,5,5] + MultiGroupHead), KITTI car.
#
# Inference subset of the reference config
#   examples/point_pillars/configs/kitti_point_pillars_mghead_syncbn.py
# in the same Det3D config format (same keys and values for everything the inference path reads).
# The reference file itself also loads unchanged through det3d.torchie.Config.fromfile.
import itertools
import logging

from det3d.builder import build_box_coder
from det3d.utils.config_tool import get_downsample_factor

norm_cfg = None
pc_range = [0, -39.68, -3, 69.12, 39.68, 1]
voxel_size = [0.16, 0.16, 4.0]
tasks = [dict(num_class=1, class_names=["Car"])]
class_names = list(itertools.chain(*[t["class_names"] for t in tasks]))

target_assigner = dict(
    type="iou",
    anchor_generators=[
        dict(type="anchor_generator_range", sizes=[1.6, 3.9, 1.56],
             anchor_ranges=[0, -39.68, -1.0, 69.12, 39.68, -1.0], rotations=[0, 1.57],
             matched_threshold=0.6, unmatched_threshold=0.45, class_name="Car"),
    ],
    sample_positive_fraction=-1, sample_size=512,
    region_similarity_calculator=dict(type="nearest_iou_similarity"),
    pos_area_threshold=-1, tasks=tasks,
)
box_coder = dict(type="ground_box3d_coder", n_dim=7, linear_dim=False, encode_angle_vector=False)

model = dict(
    type="PointPillars",
    pretrained=None,
    reader=dict(type="PillarFeatureNet", num_filters=[64], voxel_size=voxel_size, pc_range=pc_range,
                with_distance=False, norm_cfg=norm_cfg),
    backbone=dict(type="PointPillarsScatter", ds_factor=1, norm_cfg=norm_cfg),
    neck=dict(type="RPN", layer_nums=[3, 5, 5], ds_layer_strides=[2, 2, 2], ds_num_filters=[64, 128, 256],
              us_layer_strides=[1, 2, 4], us_num_filters=[128, 128, 128], num_input_features=64,
              norm_cfg=norm_cfg, logger=logging.getLogger("RPN")),
    bbox_head=dict(
        type="MultiGroupHead", mode="3d", in_channels=sum([128, 128, 128]), norm_cfg=norm_cfg, tasks=tasks,
        weights=[1], box_coder=build_box_coder(box_coder), encode_background_as_zeros=True,
        loss_norm=dict(type="NormByNumPositives", pos_cls_weight=1.0, neg_cls_weight=1.0),
        loss_cls=dict(type="SigmoidFocalLoss", alpha=0.25, gamma=2.0, loss_weight=1.0),
        use_sigmoid_score=True,
        loss_bbox=dict(type="WeightedSmoothL1Loss", sigma=3.0, code_weights=[1.0] * 7, codewise=True,
                       loss_weight=2.0),
        encode_rad_error_by_sin=True,
        loss_aux=dict(type="WeightedSoftmaxClassificationLoss", name="direction_classifier", loss_weight=0.2),
        direction_offset=0.0,
    ),
)

assigner = dict(box_coder=box_coder, target_assigner=target_assigner,
                out_size_factor=get_downsample_factor(model))
train_cfg = dict(assigner=assigner)
test_cfg = dict(
    nms=dict(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=1000, nms_post_max_size=300,
             nms_iou_threshold=0.5),
    score_threshold=0.05,
    post_center_limit_range=[0, -40.0, -5.0, 70.4, 40.0, 5.0],
    max_per_img=100,
)
voxel_generator = dict(range=pc_range, voxel_size=voxel_size, max_points_in_voxel=100, max_voxel_num=12000)

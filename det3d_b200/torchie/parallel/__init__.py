from .collate import collate_kitti, collate_kitti_device

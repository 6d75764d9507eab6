from . import nms_ops

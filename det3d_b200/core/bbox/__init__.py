from . import box_coders, box_torch_ops

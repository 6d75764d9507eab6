"""Config helpers imported by the reference's configs at load time.

`get_downsample_factor` mirrors det3d/utils/config_tool.py:39-48 (the only function
of that module the inference configs use)."""
import numpy as np


def get_downsample_factor(model_config):
    neck = model_config["neck"]
    factor = np.prod(neck.get("ds_layer_strides", [1]))
    ups = neck.get("us_layer_strides", [])
    if len(ups) > 0:
        factor /= ups[-1]
    factor *= model_config["backbone"]["ds_factor"]
    factor = int(factor)
    assert factor > 0
    return factor

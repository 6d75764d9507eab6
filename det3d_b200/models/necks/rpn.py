"""Dense BEV region-proposal neck (det3d/models/necks/rpn.py:22-159).

Module layout (`blocks.<i>.<j>`, `deblocks.<i>.<j>`) follows the reference so its
checkpoints load; the convolutions are cuDNN (out of scope for hand-written kernels,
SURVEY 2.1 #11) but sit inside the timed end-to-end path.
"""
import logging

import numpy as np
import torch
from torch import nn

from ..registry import NECKS
from ..utils import Sequential, build_norm_layer


@NECKS.register_module
class RPN(nn.Module):
    def __init__(self, layer_nums, ds_layer_strides, ds_num_filters, us_layer_strides, us_num_filters,
                 num_input_features, norm_cfg=None, name="rpn", logger=None, **kwargs):
        super().__init__()
        self._layer_strides = ds_layer_strides
        self._num_filters = ds_num_filters
        self._layer_nums = layer_nums
        self._upsample_strides = us_layer_strides
        self._num_upsample_filters = us_num_filters
        self._num_input_features = num_input_features
        self._norm_cfg = norm_cfg if norm_cfg is not None else dict(type="BN", eps=1e-3, momentum=0.01)
        assert len(ds_layer_strides) == len(layer_nums) == len(ds_num_filters)
        assert len(us_num_filters) == len(us_layer_strides)
        self._upsample_start_idx = len(layer_nums) - len(us_layer_strides)
        ratios = [us_layer_strides[i] / np.prod(ds_layer_strides[: i + self._upsample_start_idx + 1])
                  for i in range(len(us_layer_strides))]
        assert all(r == ratios[0] for r in ratios)

        in_filters = [num_input_features, *ds_num_filters[:-1]]
        blocks, deblocks = [], []
        for i, n_layers in enumerate(layer_nums):
            blocks.append(self._make_layer(in_filters[i], ds_num_filters[i], n_layers, ds_layer_strides[i]))
            j = i - self._upsample_start_idx
            if j >= 0:
                deblocks.append(self._make_deblock(ds_num_filters[i], us_num_filters[j], us_layer_strides[j]))
        self.blocks = nn.ModuleList(blocks)
        self.deblocks = nn.ModuleList(deblocks)
        (logger or logging.getLogger("RPN")).info("Finish RPN Initialization")

    def _make_deblock(self, cin, cout, stride):
        if stride > 1:
            conv = nn.ConvTranspose2d(cin, cout, stride, stride=stride, bias=False)
        else:
            k = int(np.round(1 / stride))
            conv = nn.Conv2d(cin, cout, k, stride=k, bias=False)
        return Sequential(conv, build_norm_layer(self._norm_cfg, cout)[1], nn.ReLU())

    def _make_layer(self, inplanes, planes, num_blocks, stride=1):
        block = Sequential(nn.ZeroPad2d(1), nn.Conv2d(inplanes, planes, 3, stride=stride, bias=False),
                           build_norm_layer(self._norm_cfg, planes)[1], nn.ReLU())
        for _ in range(num_blocks):
            block.add(nn.Conv2d(planes, planes, 3, padding=1, bias=False))
            block.add(build_norm_layer(self._norm_cfg, planes)[1])
            block.add(nn.ReLU())
        return block

    @property
    def downsample_factor(self):
        factor = np.prod(self._layer_strides)
        if len(self._upsample_strides) > 0:
            factor /= self._upsample_strides[-1]
        return factor

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        ups = []
        for i, block in enumerate(self.blocks):
            x = block(x)
            j = i - self._upsample_start_idx
            if j >= 0:
                ups.append(self.deblocks[j](x))
        return torch.cat(ups, dim=1) if ups else x

"""spconv-v1-style module API on top of the det3d_b200 kernels.

The reference's middle encoders are written against the external `spconv`
package (det3d/models/backbones/scn.py:4,9): `spconv.SparseConvTensor`,
`spconv.SparseSequential`, `spconv.SparseModule`, `SubMConv3d`,
`SparseConv3d`.  These classes keep those names, constructor arguments,
parameter names/shapes (`weight [kD,kH,kW,Cin,Cout]`, `bias [Cout]`) and
forward semantics so `scn.py`-style model code and its state_dicts work
unchanged.  This generic path runs one layer at a time (conv kernel, then the
BatchNorm1d / ReLU modules on `.features`); `SpMiddleFHD` & co. use the fused
executor in `fused.py` instead.
"""
import math

import numpy as np
import torch
from torch import nn

from . import core


class SparseConvTensor:
    """features [N, C] f32, indices [N, 4] int32 (batch, z, y, x), spatial_shape [D, H, W]."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self._features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = {}
        self.grid = grid
        self._level = None

    # `.features` is assignable, as scn.py:76-87 does (out.features = bn(out.features)).
    @property
    def features(self):
        return self._features

    @features.setter
    def features(self, value):
        self._features = value

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key, None)

    def level(self):
        if self._level is None:
            self._level = core.level_from_coors(self.indices, self.spatial_shape, self.batch_size)
        return self._level

    def dense(self, channels_first=True):
        feats = self._features.to(torch.float32).contiguous()
        lvl = self.level()
        if feats.shape[0] == 0:
            d, h, w = self.spatial_shape
            out = torch.zeros((self.batch_size, feats.shape[1], d, h, w), dtype=torch.float32, device=feats.device)
        else:
            out = core.sparse_to_dense(feats, lvl)
        if not channels_first:
            out = out.permute(0, 2, 3, 4, 1).contiguous()
        return out

    @property
    def sparity(self):
        return self.indices.shape[0] / float(self.spatial_size * self.batch_size)


class SparseModule(nn.Module):
    """Marker base class: modules that consume/produce SparseConvTensor."""


def _is_sparse(module):
    return isinstance(module, SparseModule)


class SparseSequential(SparseModule):
    """Sequential that hands SparseConvTensor to sparse modules and `.features` to the rest."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], dict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        for module in self._modules.values():
            if _is_sparse(module):
                input = module(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input.features = module(input.features)
            else:
                input = module(input)
        return input


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, subm=False, output_padding=0, transposed=False, inverse=False,
                 indice_key=None):
        super().__init__()
        assert ndim == 3, "det3d_b200 builds the 3-D case (the only one scn.py uses)"
        assert groups == 1 and not transposed and not inverse
        k = core._triple(kernel_size)
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = list(k)
        self.conv1x1 = int(np.prod(k)) == 1
        self.stride = list(core._triple(stride))
        self.padding = list(core._triple(padding))
        self.dilation = list(core._triple(dilation))
        assert self.dilation == [1, 1, 1], "dilation != 1 is not on the Det3D hot path"
        self.subm = subm
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(*k, in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self._cw = None
        self._cw_version = None

    def reset_parameters(self):
        # spconv v1: kaiming_uniform_(a=sqrt(5)) on the [k..., Cin, Cout] tensor, bias U(+-1/sqrt(fan_in))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def conv_weights(self):
        ver = (self.weight._version, None if self.bias is None else self.bias._version, self.weight.device)
        if self._cw is None or self._cw_version != ver:
            self._cw = core.ConvWeights(self.weight, bias=self.bias)
            self._cw_version = ver
        return self._cw

    def rulebook(self, input):
        """(rulebook, output level): cached per indice_key like spconv's indice_dict."""
        cached = input.find_indice_pair(self.indice_key)
        if cached is not None:
            return cached
        lvl = input.level()
        if self.subm:
            rb = core.build_subm_rulebook(core.alloc_subm_rulebook(lvl, self.kernel_size))
        else:
            rb = core.build_conv_rulebook(core.alloc_conv_rulebook(lvl, self.kernel_size, self.stride, self.padding))
        if self.indice_key is not None:
            input.indice_dict[self.indice_key] = rb
        return rb

    def forward(self, input):
        assert isinstance(input, SparseConvTensor)
        feats = input.features.to(torch.float32).contiguous()
        rb = self.rulebook(input)
        out_level = rb.out_level
        out_feats = torch.empty((max(out_level.cap, 1), self.out_channels), dtype=torch.float32, device=feats.device)
        if feats.shape[0] > 0:
            core.sparse_conv(feats, rb, self.conv_weights(), out_feats)
        if self.subm:
            out = SparseConvTensor(out_feats[: feats.shape[0]], input.indices, input.spatial_shape, input.batch_size)
            out._level = input._level
        else:
            n_out = out_level.count() if feats.shape[0] > 0 else 0  # the one host sync of the generic path
            if feats.shape[0] > 0 and int(out_level.n[1].item()) > out_level.cap:
                raise RuntimeError("sparse conv output overflowed its row capacity")
            out = SparseConvTensor(out_feats[:n_out], out_level.coors[:n_out], list(out_level.spatial), input.batch_size)
            exact = core.SparseLevel(out_level.coors, out_level.n, out_level.cap, out_level.spatial, out_level.batch)
            exact.index, exact._keep = out_level.index, out_level._keep
            out._level = exact
        out.indice_dict = input.indice_dict
        out.grid = input.grid
        return out


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         True, indice_key=indice_key)

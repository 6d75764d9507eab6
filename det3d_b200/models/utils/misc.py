"""Small containers used by the RPN (det3d/models/utils/misc.py): `Sequential` with
`.add()`, and the no-op `Empty`."""
from torch import nn


class Empty(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, *args, **kwargs):
        if len(args) == 1:
            return args[0]
        return args if args else None


class Sequential(nn.Sequential):
    """nn.Sequential plus `add(module, name=None)` (state_dict keys stay '0','1',...)."""

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

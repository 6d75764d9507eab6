"""build_norm_layer (det3d/models/utils/norm.py:71-112) for the layer types on the
inference path.  SyncBN variants are training-only; in eval mode they are plain
batch norm, so they map to nn.BatchNorm2d here."""
from torch import nn

_NORMS = {
    "BN": ("bn", nn.BatchNorm2d),
    "BN1d": ("bn1d", nn.BatchNorm1d),
    "GN": ("gn", nn.GroupNorm),
    "SyncBN": ("bn", nn.BatchNorm2d),
    "NaiveSyncBN": ("bn", nn.BatchNorm2d),
}


def build_norm_layer(cfg, num_features, postfix=""):
    assert isinstance(cfg, dict) and "type" in cfg
    opts = dict(cfg)
    kind = opts.pop("type")
    if kind not in _NORMS:
        raise KeyError("Unrecognized norm type {}".format(kind))
    abbr, layer_cls = _NORMS[kind]
    assert isinstance(postfix, (int, str))
    requires_grad = opts.pop("requires_grad", True)
    opts.setdefault("eps", 1e-5)
    if kind == "GN":
        assert "num_groups" in opts
        layer = layer_cls(num_channels=num_features, **opts)
    else:
        layer = layer_cls(num_features, **opts)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer

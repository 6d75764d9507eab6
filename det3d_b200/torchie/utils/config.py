"""`Config.fromfile` for python config files (det3d/torchie/utils/config.py:77-100).

The reference builds on `addict.Dict` (not installed in this image); ConfigDict below
is a small self-contained attribute dict with the same observable behaviour for the
Det3D configs: nested dicts (also inside lists/tuples) become ConfigDicts, attribute
and item access are interchangeable, a missing key raises KeyError / AttributeError.
"""
import os.path as osp
import sys
from importlib import import_module


class ConfigDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, value):
        if isinstance(value, dict) and not isinstance(value, ConfigDict):
            return cls(value)
        if isinstance(value, (list, tuple)):
            return type(value)(cls._wrap(v) for v in value)
        return value

    def __setitem__(self, key, value):
        super().__setitem__(key, self._wrap(value))

    def __setattr__(self, key, value):
        self[key] = value

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError("'%s' object has no attribute '%s'" % (type(self).__name__, key))

    def __delattr__(self, key):
        try:
            del self[key]
        except KeyError:
            raise AttributeError(key)

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def setdefault(self, key, default=None):
        if key not in self:
            self[key] = default
        return self[key]

    def to_dict(self):
        def un(v):
            if isinstance(v, ConfigDict):
                return {k: un(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(un(x) for x in v)
            return v
        return un(self)


class Config:
    """cfg = Config.fromfile(path); cfg.model.backbone.type, cfg["test_cfg"], ..."""

    @staticmethod
    def fromfile(filename):
        filename = osp.abspath(osp.expanduser(filename))
        if not osp.isfile(filename):
            raise FileNotFoundError('file "%s" does not exist' % filename)
        if not filename.endswith(".py"):
            raise IOError("Only py type is supported by det3d_b200 (the Det3D configs are python files)")
        module_name = osp.basename(filename)[:-3]
        if "." in module_name:
            raise ValueError("Dots are not allowed in config file path.")
        sys.path.insert(0, osp.dirname(filename))
        try:
            sys.modules.pop(module_name, None)
            mod = import_module(module_name)
        finally:
            sys.path.pop(0)
        cfg_dict = {k: v for k, v in vars(mod).items() if not k.startswith("__") and not _is_module(v)}
        return Config(cfg_dict, filename=filename)

    def __init__(self, cfg_dict=None, filename=None):
        cfg_dict = {} if cfg_dict is None else cfg_dict
        if not isinstance(cfg_dict, dict):
            raise TypeError("cfg_dict must be a dict, but got %s" % type(cfg_dict))
        object.__setattr__(self, "_cfg_dict", ConfigDict(cfg_dict))
        object.__setattr__(self, "_filename", filename)
        text = ""
        if filename:
            with open(filename, "r") as fh:
                text = fh.read()
        object.__setattr__(self, "_text", text)

    filename = property(lambda self: self._filename)
    text = property(lambda self: self._text)

    def __repr__(self):
        return "Config (path: %s): %r" % (self._filename, dict(self._cfg_dict))

    def __len__(self):
        return len(self._cfg_dict)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = value

    def __setitem__(self, name, value):
        self._cfg_dict[name] = value

    def __iter__(self):
        return iter(self._cfg_dict)

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)


def _is_module(v):
    import types
    return isinstance(v, types.ModuleType)

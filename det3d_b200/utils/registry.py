"""Name -> class registries and `build_from_cfg`.

Behavioural mirror of det3d/utils/registry.py:6-78: `register_module` is a class
decorator keyed by `cls.__name__` (duplicate -> KeyError, non-class -> TypeError),
`get` returns None for unknown keys, `build_from_cfg` pops "type" (a registered
name or a class), fills missing keys from `default_args` and instantiates.
"""
import inspect


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    def __repr__(self):
        return "%s(name=%s, items=%s)" % (type(self).__name__, self._name, list(self._module_dict))

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._module_dict)

    def get(self, key):
        return self._module_dict.get(key)

    def _register_module(self, module_class):
        if not inspect.isclass(module_class):
            raise TypeError("module must be a class, but got %s" % type(module_class))
        key = module_class.__name__
        if key in self._module_dict:
            raise KeyError("%s is already registered in %s" % (key, self._name))
        self._module_dict[key] = module_class

    def register_module(self, cls):
        self._register_module(cls)
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    if not (isinstance(cfg, dict) and "type" in cfg):
        raise AssertionError("cfg must be a dict with a 'type' key")
    if not (default_args is None or isinstance(default_args, dict)):
        raise AssertionError("default_args must be a dict or None")
    kwargs = dict(cfg)
    kind = kwargs.pop("type")
    if isinstance(kind, str):
        cls = registry.get(kind)
        if cls is None:
            raise KeyError("%s is not in the %s registry" % (kind, registry.name))
    elif inspect.isclass(kind):
        cls = kind
    else:
        raise TypeError("type must be a str or valid type, but got %s" % type(kind))
    for key, value in (default_args or {}).items():
        kwargs.setdefault(key, value)
    return cls(**kwargs)

from .utils import Config, ConfigDict


def is_str(x):
    return isinstance(x, str)

from .misc import Empty, Sequential
from .norm import build_norm_layer

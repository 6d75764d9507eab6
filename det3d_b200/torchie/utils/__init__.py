from .config import Config, ConfigDict

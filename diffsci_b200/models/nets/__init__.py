# flake8: noqa
from .punetg import PUNetG
from .punetg_config import PUNetGConfig
from .mlp import MLPUncond

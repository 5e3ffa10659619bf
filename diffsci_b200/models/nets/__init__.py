# flake8: noqa
from .punetg import PUNetG
from .punetg_config import PUNetGConfig
from .adm import ADM, ADMConfig
from .mlp import MLPUncond

# flake8: noqa
from .punetg import PUNetG, PUNetGCond
from .embedder import PorosityEmbedder, CompositeEmbedder
from .punetg_config import PUNetGConfig
from .adm import ADM, ADMConfig
from .mlp import MLPUncond, MLPCond

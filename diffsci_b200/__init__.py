"""diffsci_b200 -- B200-native (sm_100a) implementation of the Karras/EDM hot path of Lacadame/DiffSci.

Importing the package loads ``libdiffsci_b200.so`` (hand-written CUDA behind a C ABI, see
``include/diffsci_b200.h``) and fails loudly if it has not been built: there is no CPU and no
PyTorch-eager fallback on the product path.
"""
# tcgen05 implicit-GEMM convolution path for bf16 precision (csrc/conv_tc.cu)
TC_CONV_ENABLED = True

from . import _lib  # noqa: E402,F401  (raises ImportError when the shared library is missing)
from . import ops, models  # noqa: E402,F401
from .models import *  # noqa: E402,F401,F403

__version__ = "0.1.0"

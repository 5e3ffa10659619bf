"""Import the *live* DiffSci reference: from /root/reference in the build container, else from the copy under oracle/_ref/
(oracle/build_ref.py; git-ignored, shipped to the GPU box by gpurun).

TEST / BENCH INFRASTRUCTURE -- not product code.  Users: ``oracle/make_goldens.py`` (golden vectors, build container) and
``bench.py``'s reference arm / ``cpu_baseline`` leg (the unmodified reference timed on the host cores).  Nothing under
``diffsci_b200/`` imports it, and no ``-m gpu`` test or ``smoke()`` reads /root/reference.

The reference eagerly imports ``lightning``, ``diffusers`` and ``matplotlib`` (none of
which are installed here).  They are irrelevant to the Karras/EDM
hot path, so they are replaced by inert stubs before ``diffsci.models`` is imported.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"
REF_COPY_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    """Directory that holds the reference's ``diffsci`` package, or None."""
    for root in (REFERENCE_ROOT, REF_COPY_ROOT):
        if os.path.isdir(os.path.join(root, "diffsci")):
            return root
    return None


def _stub_lightning() -> None:
    lightning = types.ModuleType("lightning")

    class LightningModule(torch.nn.Module):
        def log(self, *a, **k):  # no-op logger
            return None

        def log_dict(self, *a, **k):
            return None

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        @classmethod
        def load_from_checkpoint(cls, *a, **k):
            raise NotImplementedError("stub lightning")

        def save_hyperparameters(self, *a, **k):
            return None

    class Trainer:  # placeholder, never run
        def __init__(self, *a, **k):
            raise NotImplementedError("stub lightning.Trainer")

    lightning.LightningModule = LightningModule
    lightning.Trainer = Trainer
    lightning.LightningDataModule = type("LightningDataModule", (), {})
    pl = types.ModuleType("lightning.pytorch")
    cb = types.ModuleType("lightning.pytorch.callbacks")
    for name in ("Callback", "StochasticWeightAveraging", "ModelCheckpoint",
                 "LearningRateMonitor", "EarlyStopping"):
        setattr(cb, name, type(name, (), {"__init__": lambda self, *a, **k: None}))
    pl.callbacks = cb
    pl.LightningModule = LightningModule
    pl.Trainer = Trainer
    lightning.pytorch = pl
    sys.modules.setdefault("lightning", lightning)
    sys.modules.setdefault("lightning.pytorch", pl)
    sys.modules.setdefault("lightning.pytorch.callbacks", cb)


class _Anything(types.ModuleType):
    """Module whose every public attribute is a dummy class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(f"{self.__name__}.{name}")
        sys.modules[sub.__name__] = sub
        cls = type(name, (), {"__init__": lambda self, *a, **k: None})
        cls.__getattr__ = classmethod(lambda c, n: None) if False else None  # noqa
        del cls.__getattr__
        setattr(self, name, cls)
        return cls


def _stub_misc() -> None:
    for name in ("diffusers", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
                 "matplotlib.cm", "matplotlib.animation"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def load_reference():
    """Return the imported ``diffsci`` package of the reference (CPU, fp32)."""
    root = reference_root()
    if root is None:
        raise RuntimeError(f"neither {REFERENCE_ROOT} nor {REF_COPY_ROOT} holds the reference (run oracle/build_ref.py in the "
                           "build container)")
    _stub_lightning()
    _stub_misc()
    if root not in sys.path:
        sys.path.insert(0, root)
    import diffsci.models  # noqa: F401
    import diffsci  # noqa: F401
    return sys.modules["diffsci"]

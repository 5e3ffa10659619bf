"""Import the *live* DiffSci reference from /root/reference (build container only).

TEST INFRASTRUCTURE -- not product code.  Only ``oracle/make_goldens.py`` uses this
module, and only inside the build container: ``/root/reference`` does not exist on the
GPU box, so nothing under ``tests/ -m gpu``, ``bench.py`` or ``__graft_entry__`` may
import it.

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
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (only exists in the build container)")
    _stub_lightning()
    _stub_misc()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import diffsci.models  # noqa: F401
    import diffsci  # noqa: F401
    return sys.modules["diffsci"]

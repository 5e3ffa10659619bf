"""ModelEMA -- drop-in for diffsci.models.karras.ema.ModelEMA (reference karras/ema.py:9-240).

Same constructor, ``update / apply_to / restore / state_dict / load_state_dict`` and beta schedules
(traditional with half-life ramp-up, EDM2 power function).  ``update`` is ONE multi-tensor kernel
launch per profile (dsk_ema_update) instead of one ``lerp_`` launch per parameter tensor (181-213
launches in the reference, ema.py:139-147).
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np
import torch

from ... import _lib as L
from ..._lib import lib, check, ptr, stream


def _power_function_exp_from_std(std: float) -> float:
    """EDM2 power-function EMA: relative std -> exponent gamma (largest real root of the cubic)."""
    if std <= 0:
        raise ValueError("Power-function EMA std must be positive")
    target = float(std) ** -2
    roots = np.roots([1.0, 7.0, 16.0 - target, 12.0 - target])
    return float(np.max(roots.real))


def _power_function_beta(std: float, next_update: int) -> float:
    if next_update <= 1:
        return 0.0
    return float((1.0 - 1.0 / next_update) ** (_power_function_exp_from_std(std) + 1.0))


class ModelEMA:
    def __init__(self, model: torch.nn.Module, ema_type: str = "traditional", decay: float = 0.999,
                 halflife_steps: Optional[float] = None, rampup_ratio: Optional[float] = None,
                 power_function_stds: Optional[list[float]] = None, device=None, profile_index: int = 0):
        self.ema_type = str(ema_type).lower()
        if self.ema_type not in {"traditional", "power"}:
            raise ValueError("ema_type must be 'traditional' or 'power'")
        if not 0.0 <= decay < 1.0:
            raise ValueError("EMA decay must be in [0, 1)")
        self.decay = float(decay)
        self.halflife_steps, self.rampup_ratio = halflife_steps, rampup_ratio
        self.power_function_stds = [0.05] if power_function_stds is None else list(power_function_stds)
        if len(self.power_function_stds) == 0:
            raise ValueError("power_function_stds must contain at least one value")
        self.device = torch.device(device) if device is not None else None
        self.profile_index = int(profile_index)
        self.num_updates = 0
        self.last_beta: Optional[float] = None
        self.profiles: list[dict[str, Any]] = []
        self._tables = None
        self.reset(model)

    # ------------------------------------------------------------------ bookkeeping
    @property
    def has_shadow(self) -> bool:
        return len(self.profiles) > 0 and len(self.profiles[0]["params"]) > 0

    def _specs(self):
        if self.ema_type == "power":
            return [{"name": f"power_std_{s:g}", "std": float(s)} for s in self.power_function_stds]
        return [{"name": "traditional", "std": None}]

    def _clone(self, t: torch.Tensor) -> torch.Tensor:
        t = t.detach()
        if self.device is not None:
            t = t.to(self.device)
        return t.clone()

    @torch.no_grad()
    def reset(self, model: torch.nn.Module) -> None:
        self.profiles = []
        for spec in self._specs():
            self.profiles.append({**spec,
                                  "params": {n: self._clone(p) for n, p in model.named_parameters()},
                                  "buffers": {n: self._clone(b) for n, b in model.named_buffers()}})
        self.num_updates = 0
        self.last_beta = None
        self._tables = None

    def _traditional_beta(self, next_update: int) -> float:
        if self.halflife_steps is None:
            return self.decay
        hl = float(self.halflife_steps)
        if self.rampup_ratio is not None:
            hl = min(hl, max(float(next_update), 1.0) * float(self.rampup_ratio))
        return float(0.5 ** (1.0 / max(hl, 1e-8)))

    def _beta_for_profile(self, profile, next_update: int) -> float:
        if self.ema_type == "power":
            return _power_function_beta(profile["std"], next_update)
        return self._traditional_beta(next_update)

    # ------------------------------------------------------------------ fused update
    def _pointer_tables(self, model):
        """Device arrays of (shadow ptr, param ptr, numel) per profile; rebuilt when storage moves."""
        named = [(n, p) for n, p in model.named_parameters()]
        sig = tuple(p.data_ptr() for _, p in named) + tuple(
            pr["params"][n].data_ptr() for pr in self.profiles for n, _ in named)
        if self._tables is None or self._tables[0] != sig:
            dev = named[0][1].device
            mk = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)  # noqa: E731
            per = []
            for pr in self.profiles:
                per.append(mk([pr["params"][n].data_ptr() for n, _ in named]))
            self._tables = (sig, per, mk([p.data_ptr() for _, p in named]), mk([p.numel() for _, p in named]),
                            max(p.numel() for _, p in named), len(named))
        return self._tables

    @torch.no_grad()
    def update(self, model: torch.nn.Module) -> None:
        if not self.has_shadow:
            self.reset(model)
        nxt = self.num_updates + 1
        params = dict(model.named_parameters())
        for pr in self.profiles:                      # late-registered params / device moves (ema.py:133-146)
            for n, p in params.items():
                if n not in pr["params"]:
                    pr["params"][n] = self._clone(p)
                elif self.device is None and pr["params"][n].device != p.device:
                    pr["params"][n] = pr["params"][n].to(p.device)
        first = next(iter(params.values()))
        fusable = first.is_cuda and all(p.dtype == torch.float32 and p.is_contiguous() and
                                        pr["params"][n].device == p.device
                                        for pr in self.profiles for n, p in params.items())
        if not fusable:
            raise RuntimeError("diffsci_b200.ModelEMA.update: parameters and shadows must be contiguous fp32 CUDA "
                               "tensors on one device (the EMA step is a fused multi-tensor CUDA kernel; no CPU path)")
        L.require_cuda(first, "EMA parameters")
        _, shadow_tabs, ptab, ntab, max_numel, nt = self._pointer_tables(model)
        for pr, stab in zip(self.profiles, shadow_tabs):
            beta = self._beta_for_profile(pr, nxt)
            pr["last_beta"] = beta
            check(lib.dsk_ema_update(ptr(stab), ptr(ptab), ptr(ntab), nt, max_numel, beta, stream()))
            pr["buffers"] = {n: self._clone(b) for n, b in model.named_buffers()}
        self.num_updates = nxt
        self.last_beta = self.selected_profile().get("last_beta")

    def selected_profile(self) -> dict[str, Any]:
        return self.profiles[min(max(self.profile_index, 0), len(self.profiles) - 1)]

    @torch.no_grad()
    def apply_to(self, model: torch.nn.Module) -> dict[str, dict[str, torch.Tensor]]:
        prof = self.selected_profile()
        backup = {"params": {}, "buffers": {}}
        for n, p in model.named_parameters():
            if n not in prof["params"]:
                raise KeyError(f"EMA state is missing parameter {n!r}")
            backup["params"][n] = p.detach().clone()
            p.copy_(prof["params"][n].to(device=p.device, dtype=p.dtype))
        for n, b in model.named_buffers():
            backup["buffers"][n] = b.detach().clone()
            if n in prof["buffers"]:
                b.copy_(prof["buffers"][n].to(device=b.device, dtype=b.dtype))
        return backup

    @torch.no_grad()
    def restore(self, model: torch.nn.Module, backup) -> None:
        params, buffers = dict(model.named_parameters()), dict(model.named_buffers())
        for n, v in backup.get("params", {}).items():
            params[n].copy_(v.to(params[n]))
        for n, v in backup.get("buffers", {}).items():
            buffers[n].copy_(v.to(buffers[n]))

    def state_dict(self) -> dict[str, Any]:
        profs = [{"name": p["name"], "std": p["std"], "last_beta": p.get("last_beta"),
                  "params": {n: v.detach().clone() for n, v in p["params"].items()},
                  "buffers": {n: v.detach().clone() for n, v in p["buffers"].items()}} for p in self.profiles]
        return {"ema_type": self.ema_type, "decay": self.decay, "halflife_steps": self.halflife_steps,
                "rampup_ratio": self.rampup_ratio, "power_function_stds": self.power_function_stds,
                "profile_index": self.profile_index, "num_updates": self.num_updates, "last_beta": self.last_beta,
                "profiles": profs}

    def load_state_dict(self, state: dict[str, Any]) -> None:
        for k in ("ema_type", "decay", "halflife_steps", "rampup_ratio", "power_function_stds", "profile_index"):
            setattr(self, k, state.get(k, getattr(self, k)))
        self.num_updates = state.get("num_updates", 0)
        self.last_beta = state.get("last_beta")
        self.profiles = state["profiles"]
        self._tables = None

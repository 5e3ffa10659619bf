"""Integrators (reference karras/integrators.py:17-126).

Two faces of the same four schemes:

* the reference's duck-typed ABI -- ``Integrator.step(x, t, dt, rhs, noise_strength[, scheduler_fns,
  nsteps])`` with class attributes ``stochastic`` / ``need_fns`` and the ``name_to_integrator``
  registry -- used when a caller drives ``Scheduler.propagate`` with a foreign score function.  The
  update arithmetic is one ``dsk_lincomb`` launch per update; noise comes from the library's Philox
  stream (or from ``injected_noise`` for parity tests);
* ``fused_program``: the name of the stage program the CUDA-graph sampler engine runs when the score
  comes from a denoiser network (``diffsci_b200.models.karras.engine``), where the whole update is
  fused with preconditioning and the next network input (csrc/sampler.cu).
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from ... import ops
from . import schedulingfunctions


def _f32(v) -> torch.Tensor:
    """0-dim CPU fp32 view of a step scalar (t, dt arrive as 0-dim tensors or floats)."""
    if isinstance(v, torch.Tensor):
        return v.detach().to(device="cpu", dtype=torch.float32).reshape(())
    return torch.tensor(float(v), dtype=torch.float32)


def fresh_noise_seed() -> int:
    """Philox seed for one sampling run, drawn from torch's global CPU generator -- so ``torch.manual_seed`` governs the
    stochastic samplers as it governs the reference's ``torch.randn_like`` (integrators.py:68,105) and successive runs get
    new noise -- and mixed with the process rank: identically seeded ranks of a sharded run (distributed.sample_sharded) index
    the Philox stream by LOCAL element, so without the mix their Brownian increments would coincide."""
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            seed ^= (dist.get_rank() * 0x9E3779B97F4A7C15) & (2 ** 62 - 1)
    except Exception:
        pass
    return seed


class Integrator(torch.nn.Module):
    stochastic = False
    need_fns = False
    fused_program: Optional[str] = None

    def __init__(self):
        super().__init__()
        self.seed = 0xD1FF5C1
        self.injected_noise: Optional[Sequence[Tensor]] = None
        self._draws = 0
        self._pinned = False         # reset_noise(seed=...) / injected noise: the caller owns the stream

    def step(self, x: Tensor, t: Tensor, dt: Tensor, rhs: Callable, noise_strength: Optional[Any] = None):
        raise NotImplementedError

    def reset_noise(self, seed: Optional[int] = None, injected: Optional[Sequence[Tensor]] = None):
        """Pin the noise of the following runs: an explicit Philox seed, or a sequence of injected N(0,1) tensors (parity
        tests).  Draw numbers then continue across runs until the next reset."""
        if seed is not None:
            self.seed = int(seed)
        self.injected_noise = injected
        self._draws = 0
        self._pinned = seed is not None or injected is not None

    def begin_run(self):
        """Called by the scheduler at the start of every integration (Scheduler._run / inpaint): unless the caller pinned the
        stream, a fresh seed from torch's generator -- a new instance per ``sample(integrator='...')`` call must not replay the
        same Brownian path (the reference draws torch.randn_like each step)."""
        if not self._pinned:
            self.seed = fresh_noise_seed()
            self._draws = 0

    def _randn_like(self, x: Tensor) -> Tensor:
        """Replacement for torch.randn_like (integrators.py:68,105): draw number k of this run."""
        k = self._draws
        self._draws += 1
        if self.injected_noise is not None:
            return self.injected_noise[k].to(x)
        return ops.philox_normal(x.shape, self.seed, k, x.device)


class EulerIntegrator(Integrator):
    fused_program = "euler"

    def step(self, x, t, dt, rhs, noise_strength=None):
        return ops.lincomb(x, 1.0, rhs(x, t), float(_f32(dt)))


class HeunIntegrator(Integrator):
    fused_program = "heun"

    def step(self, x, t, dt, rhs, noise_strength=None):
        t32, dt32 = _f32(t), _f32(dt)
        h = float(dt32)
        r1 = rhs(x, t)
        tn = t32 + dt32
        if tn > 0:
            r2 = rhs(ops.lincomb(x, 1.0, r1, h), tn)
        elif tn == 0:
            r2 = r1
        else:
            raise ValueError("t+dt < 0 is not supported")
        return ops.lincomb(x, 1.0, r1, 0.5 * h, r2, 0.5 * h)


class EulerMaruyamaIntegrator(Integrator):
    stochastic = True
    fused_program = "euler-maruyama"

    def step(self, x, t, dt, rhs, noise_strength=None):
        assert noise_strength is not None
        dt32 = _f32(dt)
        ns = float(_f32(noise_strength(_f32(t))))
        amp = ns * float(torch.sqrt(torch.abs(dt32)))
        return ops.lincomb(x, 1.0, rhs(x, t), float(dt32), None, 0.0, self._randn_like(x), amp)


class KarrasIntegrator(Integrator):
    """EDM Algorithm 2 ("churn"); defaults 40 / 0.05 / 50 / 1.003 as the reference."""
    stochastic = False
    need_fns = True
    fused_program = "karras"

    def __init__(self, s_schurn: float = 40, s_tmin: float = 0.05, s_tmax: float = 50, s_noise: float = 1.003):
        super().__init__()
        self.s_schurn, self.s_tmin, self.s_tmax, self.s_noise = s_schurn, s_tmin, s_tmax, s_noise

    def churn(self, t32: torch.Tensor, nsteps: int):
        """(t_hat, noise coefficient) for one step, in the reference's fp32 arithmetic (integrators.py:94-103)."""
        back = min(self.s_schurn / nsteps, np.sqrt(2) - 1)
        if self.s_tmin is not None and not (self.s_tmin <= t32 <= self.s_tmax):
            back = 0
        sigma = 1 * t32
        sigma_hat = sigma + back * sigma
        std = torch.sqrt(sigma_hat ** 2 - sigma ** 2)
        return sigma_hat, std * self.s_noise

    def step(self, x, t, dt, rhs, scheduler_fns: schedulingfunctions.SchedulingFunctions = None,
             noise_strength=None, nsteps: int = 100):
        if scheduler_fns is not None and not (scheduler_fns.constant_scaling_fn and scheduler_fns.identity_noise_fn):
            raise NotImplementedError("diffsci_b200.KarrasIntegrator: only the EDM scheduling functions are built")
        t32, dt32 = _f32(t), _f32(dt)
        t_hat, coef = self.churn(t32, nsteps)
        x_hat = ops.lincomb(x, 1.0, None, 0.0, None, 0.0, self._randn_like(x), float(coef))
        r1 = rhs(x_hat, t_hat)
        tn = t32 + dt32
        dth = float(tn - t_hat)
        xe = ops.lincomb(x_hat, 1.0, r1, dth)
        if tn > 0:
            r2 = rhs(xe, tn)
            xe = ops.lincomb(x_hat, 1.0, r1, 0.5 * dth, r2, 0.5 * dth)
        return xe


def name_to_integrator(name: str) -> Integrator:
    table = {"euler": EulerIntegrator, "heun": HeunIntegrator, "euler-maruyama": EulerMaruyamaIntegrator,
             "karras": KarrasIntegrator}
    if name not in table:
        raise ValueError(f"Unknown integrator: {name}")
    return table[name]()

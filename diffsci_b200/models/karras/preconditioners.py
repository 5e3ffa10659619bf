"""Karras preconditioners (reference karras/preconditioners.py:8-161).

The classes keep the reference's method names so that user code can call them; they operate on
[B]-sized sigma vectors (scalar plumbing).  On the hot path the same formulas are evaluated
inside the fused CUDA kernels (csrc/sampler.cu: edm_precond) -- these methods are not on it.
"""
from __future__ import annotations

import torch
from torch import Tensor


class KarrasPreconditioner(torch.nn.Module):
    """Interface: D(x; sigma) = skip_scaling*x + output_scaling*F(input_scaling*x, noise_conditioner)."""

    fused_kind: str | None = None      # which closed form csrc/sampler.cu implements for this class

    def skip_scaling(self, sigma: Tensor) -> Tensor:
        raise NotImplementedError

    def output_scaling(self, sigma: Tensor) -> Tensor:
        raise NotImplementedError

    def input_scaling(self, sigma: Tensor) -> Tensor:
        raise NotImplementedError

    def noise_conditioner(self, sigma: Tensor) -> Tensor:
        raise NotImplementedError


class EDMPreconditioner(KarrasPreconditioner):
    fused_kind = "edm"

    def __init__(self, sigma_data: float = 0.5):
        super().__init__()
        self.register_buffer("sigma_data", torch.tensor(sigma_data))

    def _var(self, sigma):
        return sigma ** 2 + self.sigma_data.to(sigma) ** 2

    def skip_scaling(self, sigma):
        return self.sigma_data.to(sigma) ** 2 / self._var(sigma)

    def output_scaling(self, sigma):
        return sigma * self.sigma_data.to(sigma) / torch.sqrt(self._var(sigma))

    def input_scaling(self, sigma):
        return 1 / torch.sqrt(self._var(sigma))

    def noise_conditioner(self, sigma):
        return 0.5 * torch.log(sigma)


class SR3Preconditioner(EDMPreconditioner):
    """EDM with halved skip/output scalings (preconditioners.py:116-136)."""
    fused_kind = None

    def skip_scaling(self, sigma):
        return super().skip_scaling(sigma) / 2

    def output_scaling(self, sigma):
        return super().output_scaling(sigma) / 2


class VPPreconditioner(KarrasPreconditioner):
    def __init__(self, scheduler, M: int = 1000):
        super().__init__()
        self.scheduler, self.M = scheduler, M

    def skip_scaling(self, sigma):
        return 1 + 0.0 * sigma

    def output_scaling(self, sigma):
        return -sigma

    def input_scaling(self, sigma):
        return 1 / torch.sqrt(sigma ** 2 + 1.0)

    def noise_conditioner(self, sigma):
        return (self.M - 1) * self.scheduler.scheduler_fns.inverse_noise_fn(sigma)


class VEPreconditioner(KarrasPreconditioner):
    def skip_scaling(self, sigma):
        return 1 + 0.0 * sigma

    def output_scaling(self, sigma):
        return sigma

    def input_scaling(self, sigma):
        return 1 + 0.0 * sigma

    def noise_conditioner(self, sigma):
        return torch.log(0.5 * sigma)


class NullPreconditioner(KarrasPreconditioner):
    """D = F(x, sigma): used with analytic denoisers (tests/test_karras_on_toy_dataset.py:36-40)."""

    def skip_scaling(self, sigma):
        return 0.0 * sigma

    def output_scaling(self, sigma):
        return 1.0 + 0.0 * sigma

    def input_scaling(self, sigma):
        return 1.0 + 0.0 * sigma

    def noise_conditioner(self, sigma):
        return sigma

"""s(t), sigma(t) families (reference karras/schedulingfunctions.py:6-169).

Only the EDM family (s = 1, sigma = t) is on the fused CUDA path; VP / VE are kept as scalar
host-side definitions for API compatibility.
"""
from __future__ import annotations

import torch


class SchedulingFunctions(torch.nn.Module):
    constant_scaling_fn = False
    identity_noise_fn = False
    has_pf_score_multiplier = False
    has_pf_scale_multiplier = False

    def scaling_fn(self, t):
        raise NotImplementedError

    def scaling_fn_deriv(self, t):
        raise NotImplementedError

    def noise_fn(self, t):
        raise NotImplementedError

    def inverse_noise_fn(self, t):
        raise NotImplementedError

    def noise_fn_deriv(self, t):
        raise NotImplementedError

    def pf_score_multiplier(self, t):
        raise NotImplementedError

    def pf_scale_multiplier(self, t):
        raise NotImplementedError


class EDMSchedulingFunctions(SchedulingFunctions):
    constant_scaling_fn = True
    identity_noise_fn = True

    def scaling_fn(self, t):
        return 1 + 0 * t

    def scaling_fn_deriv(self, t):
        return 0 * t

    def noise_fn(self, t):
        return 1 * t

    def inverse_noise_fn(self, t):
        return 1 * t

    def noise_fn_deriv(self, t):
        return 1 + 0 * t


class VPSchedulingFunctions(SchedulingFunctions):
    def __init__(self, beta_data: float = 19.9, beta_min: float = 0.1):
        super().__init__()
        self.beta_data, self.beta_min = beta_data, beta_min

    def _e(self, t):
        return 0.5 * self.beta_data * t ** 2 + self.beta_min * t

    def _de(self, t):
        return self.beta_data * t + self.beta_min

    def scaling_fn(self, t):
        return torch.exp(-self._e(t) / 2)

    def scaling_fn_deriv(self, t):
        return -self._de(t) / 2 * torch.exp(-self._e(t) / 2)

    def noise_fn(self, t):
        return torch.sqrt(torch.exp(self._e(t)) - 1)

    def inverse_noise_fn(self, t):
        y = torch.log(t ** 2 + 1)
        return (-self.beta_min + torch.sqrt(self.beta_min ** 2 + 2 * self.beta_data * y)) / self.beta_data

    def noise_fn_deriv(self, t):
        ex = torch.exp(self._e(t))
        return self._de(t) * ex / (2 * torch.sqrt(ex - 1))

    def pf_score_multiplier(self, t):
        return 1 / 2 * self._de(t)

    def pf_scale_multiplier(self, t):
        return -1 / 2 * self._de(t)


class VESchedulingFunctions(SchedulingFunctions):
    constant_scaling_fn = True
    has_pf_score_multiplier = True

    def scaling_fn(self, t):
        return 1 + 0 * t

    def scaling_fn_deriv(self, t):
        return 0 * t

    def noise_fn(self, t):
        return torch.sqrt(t)

    def inverse_noise_fn(self, t):
        return t ** 2

    def noise_fn_deriv(self, t):
        return 0.5 / torch.sqrt(t)

    def pf_score_multiplier(self, t):
        return 0.5 + 0 * t


def name_to_scheduling_functions(name: str, *args, **kwargs) -> SchedulingFunctions:
    table = {"EDM": EDMSchedulingFunctions, "VP": VPSchedulingFunctions, "VE": VESchedulingFunctions}
    if name not in table:
        raise ValueError(f"Unknown scheduling function name: {name}")
    return table[name]() if name == "EDM" else table[name](*args, **kwargs)

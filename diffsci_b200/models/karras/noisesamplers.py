"""Training-time sigma priors and loss weights (reference karras/noisesamplers.py:8-110).

sigma is drawn on the CPU generator exactly like the reference (noisesamplers.py:38), so a seeded
run sees the same sigmas on any device; lambda(sigma) on the hot path is evaluated inside the
fused loss kernel (csrc/train.cu).
"""
from __future__ import annotations

import torch
from torch import Tensor


class NoiseSampler(torch.nn.Module):
    def loss_weighting(self, sigma: Tensor) -> Tensor:
        raise NotImplementedError

    def sample(self, shape) -> Tensor:
        raise NotImplementedError


class EDMNoiseSampler(NoiseSampler):
    def __init__(self, sigma_data: float = 0.5, prior_mean: float = -1.2, prior_std: float = 1.2):
        super().__init__()
        self.register_buffer("sigma_data", torch.tensor(sigma_data))
        self.register_buffer("prior_mean", torch.tensor(prior_mean))
        self.register_buffer("prior_std", torch.tensor(prior_std))

    def loss_weighting(self, sigma):
        sd = self.sigma_data.to(sigma)
        return (sigma ** 2 + sd ** 2) / ((sigma * sd) ** 2)

    def sample(self, shape):
        xi = torch.randn(shape).to(self.prior_mean.device)
        return torch.exp(xi * self.prior_std + self.prior_mean)


class VPNoiseSampler(NoiseSampler):
    def __init__(self, noise_scheduler, epsilon: float = 1e-3):
        super().__init__()
        self.noise_scheduler = noise_scheduler
        self.register_buffer("epsilon", torch.tensor(epsilon))

    def loss_weighting(self, sigma):
        return 1 / (sigma ** 2)

    def sample(self, shape):
        t = torch.rand(shape).to(self.epsilon)
        t = t * (1 - self.epsilon) + self.epsilon
        return self.noise_scheduler.scheduler_fns.noise_fn(t)


class VENoiseSampler(NoiseSampler):
    def __init__(self, sigma_min: float = 0.02, sigma_max: float = 100):
        super().__init__()
        self.register_buffer("sigma_min", torch.tensor(sigma_min))
        self.register_buffer("sigma_max", torch.tensor(sigma_max))

    def loss_weighting(self, sigma):
        return 1 / (sigma ** 2)

    def sample(self, shape):
        u = torch.rand(shape).to(self.sigma_min.device)
        lo, hi = torch.log(self.sigma_min), torch.log(self.sigma_max)
        return torch.exp(lo + u * (hi - lo))


class UniformNoiseSampler(NoiseSampler):
    def __init__(self, t: float = 0.0, T: float = 1.0, sigma_data: float = 0.5):
        super().__init__()
        self.register_buffer("t", torch.tensor(t))
        self.register_buffer("T", torch.tensor(T))
        self.register_buffer("sigma_data", torch.tensor(sigma_data))

    def loss_weighting(self, sigma):
        sd = self.sigma_data.to(sigma)
        return (sigma ** 2 + sd ** 2) / ((sigma * sd) ** 2)

    def sample(self, shape):
        u = torch.rand(shape).to(self.t.device)
        return self.t + u * (self.T - self.t)

"""Host-side tensor utilities with the reference's names (diffsci/torchutils.py:4-40)."""
from __future__ import annotations

import torch


def broadcast_from_below(t: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """View t[...] as t[..., 1, 1, ...] so that it broadcasts against x from the leading dims."""
    if x.ndim < t.ndim:
        raise ValueError("The number of dimensions of the x tensor must be greater or equal to the number of "
                         "dimensions of the t tensor")
    return t.reshape(tuple(t.shape) + (1,) * (x.ndim - t.ndim)).to(x)


def dict_map(func, d):
    if isinstance(d, dict):
        return {k: dict_map(func, v) for k, v in d.items()}
    return func(d)


def dict_to(d, device):
    return dict_map(lambda v: v.to(device) if isinstance(v, torch.Tensor) else v, d)


def dict_unsqueeze(d, dim):
    return dict_map(lambda v: v.unsqueeze(dim) if isinstance(v, torch.Tensor) else v, d)

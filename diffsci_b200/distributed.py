"""Multi-GPU sampling: one process per GPU, batch sharded, no data-path collective, one final gather.

Reference behaviour being scaled out: ``KarrasModule.sample`` chunks a batch serially
(karrasmodule.py:817-835) and the paper scripts farm GPUs by process with seed = base + worker*10000
(stochasticity_paper/scripts/test-diffusion-cifar10karras-colormap-parallel.py:191-300).  Samples are
independent (no op couples batch elements), so each rank integrates its own contiguous
slice and the only communication is the gather of the finished fields.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.distributed as dist


def shard_sizes(nsamples: int, world: int) -> list[int]:
    """Contiguous near-equal split: the first (nsamples % world) ranks get one extra sample."""
    base, extra = divmod(nsamples, world)
    return [base + (1 if r < extra else 0) for r in range(world)]


def shard_range(nsamples: int, world: int, rank: int) -> tuple[int, int]:
    sizes = shard_sizes(nsamples, world)
    lo = sum(sizes[:rank])
    return lo, lo + sizes[rank]


def white_noise_shard(nsamples: int, shape: Sequence[int], seed: int, world: int, rank: int,
                      bit_parity: bool = True) -> torch.Tensor:
    """x_T for this rank.  bit_parity=True draws the full [nsamples, *shape] tensor on the CPU generator and
    slices it, so the union over ranks is bit-identical to a single-process ``sample`` with the same seed;
    False draws only the local slice with seed + rank*10000 (the reference scripts' convention)."""
    lo, hi = shard_range(nsamples, world, rank)
    g = torch.Generator()
    if bit_parity:
        g.manual_seed(seed)
        return torch.randn(nsamples, *shape, generator=g)[lo:hi].contiguous()
    g.manual_seed(seed + rank * 10_000)
    return torch.randn(hi - lo, *shape, generator=g)


def gather_samples(local: torch.Tensor, nsamples: int, group=None) -> Optional[torch.Tensor]:
    """All ranks contribute [n_local, *shape]; every rank receives [nsamples, *shape] in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(nsamples, world)
    pad = max(sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def sample_sharded(sample_fn: Callable[[torch.Tensor], torch.Tensor], nsamples: int, shape: Sequence[int], seed: int,
                   device, bit_parity: bool = True, group=None) -> torch.Tensor:
    """sample_fn(white_noise_on_device) -> samples; e.g. ``lambda wn: module.propagate_white_noise(wn, nsteps=64)``."""
    on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    wn = white_noise_shard(nsamples, shape, seed, world, rank, bit_parity)
    out = sample_fn(wn.to(device, non_blocking=True)) if wn.shape[0] > 0 else wn.to(device)
    return gather_samples(out, nsamples, group)

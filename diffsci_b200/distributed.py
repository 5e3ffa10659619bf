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


# ------------------------------------------------------------------------------------------------ training
class GradBucketer:
    """Data-parallel gradient exchange for the native trainer: the ONE collective of the training path.

    The reference trains under Lightning's DDP strategy (all-reduce of parameter gradients, replicated optimizer
    and EMA).  Here the gradients already live in one flat fp32 buffer in ``net.parameters()`` order
    (TrainGraph.flat_grad); it is cut into contiguous buckets of ~``bucket_bytes``, and each bucket's all-reduce(SUM) is
    issued asynchronously the moment the backward launch that completes it has been enqueued, so the exchange over
    NVLink overlaps the rest of the backward pass.  The 1/world averaging is folded into the fused AdamW kernel
    (``grad_scale``).  Works on any backend (NCCL on the box, gloo in the CPU tests).
    """

    def __init__(self, flat_grad: torch.Tensor, numels: Sequence[int], ready_pos: Sequence[int], bucket_bytes: int = 32 << 20,
                 group=None, enabled: bool = True):
        self.flat, self.group = flat_grad, group
        self.on = enabled and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.on else 1
        self.buckets = self.plan(numels, ready_pos, bucket_bytes // flat_grad.element_size())
        self._works: list = []

    @staticmethod
    def plan(numels: Sequence[int], ready_pos: Sequence[int], bucket_elems: int) -> list[tuple[int, int, int]]:
        """-> [(lo, hi, ready)] contiguous element ranges covering all parameters; a bucket is ready once every
        parameter in it is (ready = max of the members' positions).  Buckets are cut walking the parameters from the
        LAST one backwards, the order in which the backward pass finishes them."""
        offs = [0]
        for n in numels:
            offs.append(offs[-1] + int(n))
        out, hi, ready, count = [], len(numels), 0, 0
        for i in range(len(numels) - 1, -1, -1):
            ready = max(ready, int(ready_pos[i]))
            count += int(numels[i])
            if count >= bucket_elems or i == 0:
                out.append((offs[i], offs[hi], ready))
                hi, ready, count = i, 0, 0
        return out

    def hooks(self) -> dict:
        """{backward position: callable} for TrainGraph.run_backward."""
        if not self.on:
            return {}
        table: dict = {}
        for lo, hi, ready in self.buckets:
            table.setdefault(ready, []).append((lo, hi))

        def make(ranges):
            def fire():
                for lo, hi in ranges:
                    self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            return fire
        return {pos: make(r) for pos, r in table.items()}

    def finish(self) -> float:
        """Wait for the outstanding all-reduces (stream-level for NCCL); returns the factor that turns the sum into
        the mean."""
        for w in self._works:
            w.wait()
        self._works = []
        return 1.0 / self.world

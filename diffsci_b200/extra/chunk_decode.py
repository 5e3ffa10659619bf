"""Chunked (tiled) decode of a large latent volume -- the step after the sampler for volumes whose full-size decode does not
fit (reference ``diffsci/extra/chunk_decode.py``: ``chunk_decode_strategy_b_3d``, used with the latent-diffusion wrapper
``KarrasModule(model, cfg, autoencoder=vae)``, karrasmodule.py:1216-1234).

The reference's routine is written against the internal stage structure of its own ``VAEDecoder`` (out of this repo's scope,
SURVEY.md section 2 #11) and streams stage by stage with per-stage halos.  This is the decoder-agnostic form of the same idea:
any decoder made of local operations (k x k x k convolutions at stride 1, nearest up-sampling, pointwise ops) has a finite
receptive field, so the decode of a latent tile extended by a halo of that radius -- wrapped periodically or clamped at the
volume boundary -- reproduces, on the tile's centre, exactly what the full-volume decode computes there.  Tiles run on the
decoder's device one at a time; the centres are written to a host buffer (pinned on request), so the device never holds more
than one tile.  Host-side orchestration only: the arithmetic is the user's decoder.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Union

import torch

Int3 = Union[int, Sequence[int]]


def _three(v: Int3, name: str) -> tuple:
    if isinstance(v, int):
        return (v, v, v)
    v = tuple(int(a) for a in v)
    if len(v) != 3:
        raise ValueError(f"{name} must be an int or three ints, got {v}")
    return v


def _window(start: int, stop: int, size: int, periodic: bool) -> tuple:
    """Index list of [start, stop) along an axis of length `size`: wrapped (periodic) or clipped; returns (indices, lead) with
    `lead` = how many requested positions fell before the clipped start (so the caller can locate the centre)."""
    if periodic:
        return torch.arange(start, stop) % size, 0
    lo, hi = max(start, 0), min(stop, size)
    return torch.arange(lo, hi), lo - start


@torch.no_grad()
def chunk_decode_3d(decode: Callable[[torch.Tensor], torch.Tensor], z: torch.Tensor, chunk: Int3, halo: Int3, scale: int = 1,
                    periodic: Union[bool, Sequence[bool]] = False, device: Optional[torch.device] = None,
                    out: Optional[torch.Tensor] = None, pin_memory: bool = False) -> torch.Tensor:
    """Decode ``z`` [B, C, D, H, W] tile by tile.

    decode   : latent tile [B, C, d, h, w] (on `device`) -> decoded tile [B, C', d*scale, h*scale, w*scale]
    chunk    : tile size in latent voxels per axis (the centre each tile contributes)
    halo     : latent voxels read around a tile per axis; exact when >= the decoder's receptive radius (in latent voxels)
    scale    : spatial up-sampling factor of the decoder
    periodic : per axis, wrap the halo around the volume (periodic media) instead of clipping it at the boundary -- with
               clipping the decoder sees the volume boundary where the full decode would (zero-padded convolutions behave
               identically there), so both choices are exact for halo >= receptive radius
    Returns the decoded volume on the host ([B, C', D*scale, H*scale, W*scale]; `out` if given)."""
    if z.ndim != 5:
        raise ValueError(f"z must be [B, C, D, H, W], got {tuple(z.shape)}")
    chunk, halo = _three(chunk, "chunk"), _three(halo, "halo")
    per = (periodic,) * 3 if isinstance(periodic, bool) else tuple(bool(p) for p in periodic)
    dims = tuple(z.shape[2:])
    dev = device if device is not None else z.device
    result = out
    for d0 in range(0, dims[0], chunk[0]):
        for h0 in range(0, dims[1], chunk[1]):
            for w0 in range(0, dims[2], chunk[2]):
                start = (d0, h0, w0)
                stop = tuple(min(s + c, n) for s, c, n in zip(start, chunk, dims))
                idx, lead = zip(*[_window(s - r, e + r, n, p) for s, e, r, n, p in zip(start, stop, halo, dims, per)])
                tile = z.index_select(2, idx[0].to(z.device)).index_select(3, idx[1].to(z.device)).index_select(4, idx[2].to(z.device))
                dec = decode(tile.to(dev, non_blocking=True))
                if result is None:
                    shape = (dec.shape[0], dec.shape[1]) + tuple(n * scale for n in dims)
                    result = torch.empty(shape, dtype=dec.dtype, pin_memory=pin_memory)
                # centre of the decoded tile: skip the (possibly clipped) leading halo
                lo = [(r - l) * scale for r, l in zip(halo, lead)]
                sz = [(e - s) * scale for s, e in zip(start, stop)]
                centre = dec[:, :, lo[0]:lo[0] + sz[0], lo[1]:lo[1] + sz[1], lo[2]:lo[2] + sz[2]]
                result[:, :, start[0] * scale:stop[0] * scale, start[1] * scale:stop[1] * scale,
                       start[2] * scale:stop[2] * scale].copy_(centre)
    return result

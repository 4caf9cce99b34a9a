"""Multi-GPU sharding of the upscaling stage: one process per GPU (``torch.distributed``, NCCL).

The path shards without any data-path exchange: tiles (with their halos) and frames are independent
forwards, weights are replicated (33 MB).  The only collective is the final stitch of the u8 output:

  * ``enhance_sharded``        -- BASELINE config 3: the row-major tile grid of ONE frame is cut into
    ``world_size`` contiguous, equally sized ranges; every rank pastes its tiles into a zero-filled
    full-size u8 frame and ONE ``all_reduce(SUM)`` over NCCL/NVLink merges the disjoint supports
    (<= 100 MB for an 8K frame; the ranges are disjoint so the sum is a gather).
  * ``enhance_frames_sharded`` -- BASELINE config 4: frames are dealt round-robin-contiguously to ranks;
    no collective unless the caller asks for the frames back (``gather=True``).

The reference has no distributed code at all; this module is new surface.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced split: ``(first, count)`` of ``n_items`` for ``rank``."""
    base, extra = divmod(n_items, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def enhance_sharded(engine, img_bgr, tile: int, tile_pad: int, pre_pad: int = 0, group=None, out=None):
    """Tile-sharded ``RealESRGANer.enhance`` of one frame; every rank returns the full x2 frame.

    ``engine`` needs ``tile_count`` / ``enhance_tiles_u8`` / ``scale`` (an ``_ffi.Engine``);
    ``img_bgr`` is a CUDA uint8 tensor (NCCL) or a numpy array (gloo tests)."""
    world, rank = _world(group)
    h, w = img_bgr.shape[:2]
    s = engine.scale
    n_tiles = engine.tile_count(h, w, tile, pre_pad)
    first, count = partition(n_tiles, world, rank)
    on_device = isinstance(img_bgr, torch.Tensor)
    if out is None:
        out = (torch.zeros((h * s, w * s, 3), dtype=torch.uint8, device=img_bgr.device) if on_device
               else np.zeros((h * s, w * s, 3), np.uint8))
    else:
        out.zero_() if on_device else out.fill(0)
    if count:
        engine.enhance_tiles_u8(img_bgr, out, tile, tile_pad, pre_pad, first, count)
    if world > 1:
        t = out if on_device else torch.from_numpy(out)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)      # disjoint supports: sum == stitch
    return out


def enhance_frames_sharded(engine, frames_bgr, tile: int = 0, tile_pad: int = 10, pre_pad: int = 0, group=None,
                           gather: bool = False):
    """Frame-sharded throughput mode: returns ``(first, local_out)`` or, with ``gather``, all frames."""
    world, rank = _world(group)
    n = frames_bgr.shape[0]
    first, count = partition(n, world, rank)
    s = engine.scale
    h, w = frames_bgr.shape[1:3]
    on_device = isinstance(frames_bgr, torch.Tensor)
    if count:
        local = engine.enhance_batch_u8(frames_bgr[first:first + count], tile=tile, tile_pad=tile_pad, pre_pad=pre_pad)
    else:                      # fewer frames than ranks: an empty batch of the OUTPUT shape, same container kind
        local = (torch.empty((0, h * s, w * s, 3), dtype=torch.uint8, device=frames_bgr.device) if on_device
                 else np.empty((0, h * s, w * s, 3), np.uint8))
    if not gather or world == 1:
        return first, local
    full = (torch.zeros((n, h * s, w * s, 3), dtype=torch.uint8, device=frames_bgr.device) if on_device
            else torch.zeros((n, h * s, w * s, 3), dtype=torch.uint8))
    if count:
        full[first:first + count] = local if on_device else torch.from_numpy(np.ascontiguousarray(local))
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return 0, (full if on_device else full.numpy())

"""Multi-GPU sharding of the upscaling stage: one process per GPU (``torch.distributed``, NCCL).

The path shards without any data-path exchange: tiles (with their halos) and frames are independent
forwards, weights are replicated (33 MB).  The only collective is the final gather of the u8 output:

  * ``enhance_sharded``        -- BASELINE config 3: the row-major tile grid of ONE frame is cut into
    ``world_size`` shares of equal COST (padded feature pixels; a longest-first deal, ``partition_lpt``: tiles are independent,
    so a share need not be contiguous -- 2.5 % imbalance for the 40 tiles of a 4K frame over 8 ranks, where the best contiguous cut,
    ``partition_by_cost``, is 12 % off); every rank
    writes its tiles TILE-MAJOR into one contiguous buffer (``nesr_b200_enhance_tile_list_packed_u8``), ONE
    ``all_gather_into_tensor`` over NCCL/NVLink exchanges the buffers (each rank sends only its own pixels: 1/N of
    the frame, ~12 MB at 8 GPUs for an 8K frame) and one small kernel pastes every rank's slots into the frame
    (``nesr_b200_unpack_tile_list_u8``).  With one rank it is the plain ``enhance_u8`` call.
  * ``enhance_frames_sharded`` -- BASELINE config 4: frames are dealt round-robin-contiguously to ranks;
    no collective unless the caller asks for the frames back (``gather=True``).

The reference has no distributed code at all; this module is new surface.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist


def partition(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced split: ``(first, count)`` of ``n_items`` for ``rank``."""
    base, extra = divmod(n_items, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def partition_by_cost(costs, world_size: int):
    """Contiguous split of ``costs`` into ``world_size`` ranges minimising the largest range sum: ``[(first, count)]``.
    (Bisection on the bound + greedy fill; ranges may be empty when there are fewer items than ranks.)"""
    costs = [int(c) for c in costs]
    n = len(costs)

    def fill(bound):
        parts, i = [], 0
        for r in range(world_size):
            first, acc = i, 0
            while i < n and acc + costs[i] <= bound and (n - i) > 0:
                acc += costs[i]
                i += 1
            parts.append((first, i - first))
        return parts if i == n else None

    lo, hi = (max(costs) if costs else 0), sum(costs)
    while lo < hi:
        mid = (lo + hi) // 2
        if fill(mid) is None:
            lo = mid + 1
        else:
            hi = mid
    return fill(lo)


def partition_lpt(costs, world_size: int):
    """Longest-processing-time deal of items to ``world_size`` ranks: ``[sorted item ids]`` per rank.  Tiles are independent
    forwards, so a rank's share need not be contiguous; every rank computes the same deal (ties broken by index)."""
    loads = [0] * world_size
    parts = [[] for _ in range(world_size)]
    for i in sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i)):
        r = min(range(world_size), key=lambda r: (loads[r], r))
        parts[r].append(i)
        loads[r] += int(costs[i])
    return [sorted(p) for p in parts]


def _runs(ids):
    """Maximal runs of consecutive ids: ``[(position in ids, first id, length)]``."""
    out, k = [], 0
    while k < len(ids):
        j = k
        while j + 1 < len(ids) and ids[j + 1] == ids[j] + 1:
            j += 1
        out.append((k, ids[k], j - k + 1))
        k = j + 1
    return out


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def enhance_sharded(engine, img_bgr, tile: int, tile_pad: int, pre_pad: int = 0, group=None, out=None, timing=None):
    """Tile-sharded ``RealESRGANer.enhance`` of one frame; every rank returns the full x2 frame.

    ``engine`` needs ``tile_costs`` / ``slot_shape`` / ``enhance_tile_list_packed_u8`` / ``unpack_tile_list_u8`` / ``enhance_u8`` /
    ``scale`` (an ``_ffi.Engine``); ``img_bgr`` is a CUDA uint8 tensor (NCCL) or a numpy array (gloo tests).
    ``timing``: a dict that receives this rank's ``compute_ms`` / ``gather_ms`` / ``unpack_ms`` (wall clock, phases are
    synchronous) and its tile range."""
    world, rank = _world(group)
    h, w = img_bgr.shape[:2]
    s = engine.scale
    on_device = isinstance(img_bgr, torch.Tensor)
    if world == 1:
        t0 = time.perf_counter()
        out = engine.enhance_u8(img_bgr, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad, out=out)
        if timing is not None:
            timing.update(compute_ms=1e3 * (time.perf_counter() - t0), gather_ms=0.0, unpack_ms=0.0, tiles=None)
        return out
    costs = engine.tile_costs(h, w, tile, tile_pad, pre_pad)
    parts = partition_lpt(costs, world)
    mine_ids = parts[rank]
    slots_per_rank = max(len(p) for p in parts)
    sh, sw = engine.slot_shape(h, w, tile)
    if out is None:
        out = (torch.empty((h * s, w * s, 3), dtype=torch.uint8, device=img_bgr.device) if on_device
               else np.empty((h * s, w * s, 3), np.uint8))
    if on_device:
        gathered = torch.empty((world, slots_per_rank, sh, sw, 3), dtype=torch.uint8, device=img_bgr.device)
        mine = torch.empty((slots_per_rank, sh, sw, 3), dtype=torch.uint8, device=img_bgr.device)
    else:
        gathered = torch.empty((world, slots_per_rank, sh, sw, 3), dtype=torch.uint8)
        mine = np.zeros((slots_per_rank, sh, sw, 3), np.uint8)
    t0 = time.perf_counter()
    if mine_ids:
        engine.enhance_tile_list_packed_u8(img_bgr, mine, tile, tile_pad, pre_pad, mine_ids)
    t1 = time.perf_counter()
    send = mine if on_device else torch.from_numpy(mine)
    dist.all_gather_into_tensor(gathered.view(-1), send.reshape(-1), group=group)
    if on_device:
        torch.cuda.current_stream(img_bgr.device).synchronize()
    t2 = time.perf_counter()
    # every rank pastes every rank's slots (its own included): ONE call over the whole gathered buffer, empty slots marked -1
    all_ids = [ids[k] if k < len(ids) else -1 for ids in parts for k in range(slots_per_rank)]
    flat = gathered.view(world * slots_per_rank, sh, sw, 3)
    engine.unpack_tile_list_u8(flat if on_device else flat.numpy(), out, h, w, tile, pre_pad, all_ids)
    t3 = time.perf_counter()
    if timing is not None:
        timing.update(compute_ms=1e3 * (t1 - t0), gather_ms=1e3 * (t2 - t1), unpack_ms=1e3 * (t3 - t2), tiles=list(mine_ids),
                      cost=sum(costs[i] for i in mine_ids))
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

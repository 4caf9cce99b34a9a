"""Host logic of the multi-GPU sharding, exercised with world_size-2 ``gloo`` on CPU.  The engine is
replaced by a stand-in that evaluates tiles with the ORACLE (tests may), so what is checked is the
partitioning + the stitch collective: sharded result == single-process result, bit for bit."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_enhanced_super_resolution_b200 import _ffi
from neural_enhanced_super_resolution_b200.parallel import (_runs, enhance_frames_sharded, enhance_sharded, partition, partition_by_cost,
                                                            partition_lpt)


def test_partition_is_balanced_and_contiguous():
    for n in (0, 1, 5, 12, 40, 256):
        for world in (1, 2, 3, 4, 8):
            parts = [partition(n, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == n
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
            pos = 0
            for first, count in parts:
                assert first == pos
                pos += count
    assert [partition(40, 8, r)[1] for r in range(8)] == [5] * 8          # BASELINE config 3
    assert [partition(12, 8, r)[1] for r in range(8)] == [2, 2, 2, 2, 1, 1, 1, 1]


def test_partition_by_cost_minimises_the_largest_share():
    rng = np.random.default_rng(4)
    for n in (0, 1, 3, 12, 40):
        costs = [int(c) for c in rng.integers(1, 100, n)]
        for world in (1, 2, 3, 8):
            parts = partition_by_cost(costs, world)
            assert len(parts) == world and sum(c for _, c in parts) == n
            pos = 0
            for first, count in parts:
                assert first == pos
                pos += count
            best = max((sum(costs[f:f + c]) for f, c in parts), default=0)
            # brute force over contiguous splits for small cases
            if n <= 12 and world <= 3 and n:
                import itertools
                opt = min(max(sum(costs[a:b]) for a, b in zip((0,) + cut, cut + (n,)))
                          for cut in itertools.combinations_with_replacement(range(n + 1), world - 1))
                assert best == opt
    # BASELINE config 3: the 8 x 5 tile grid of a 4K frame (last column half as wide, last row a fifth as tall) over 8 ranks:
    # the largest share is within 13 % of the mean (tiles are atomic), where equal COUNTS (5 each) would be 25 % over
    costs = _ffi.Engine.tile_costs(type("E", (), {"scale": 2})(), 2160, 3840, 512, 10)
    assert len(costs) == 40
    shares = [sum(costs[f:f + c]) for f, c in partition_by_cost(costs, 8)]
    by_count = [sum(costs[f:f + c]) for f, c in (partition(40, 8, r) for r in range(8))]
    assert max(shares) <= 1.13 * sum(costs) / 8 and max(shares) < max(by_count)


def test_partition_lpt_balances_the_4k_tile_grid():
    """BASELINE config 3 on 8 ranks: the longest-first deal is within 3 % of the mean share (contiguous ranges: 12 %), every tile is
    dealt exactly once, and the runs used for pasting cover each rank's list."""
    costs = _ffi.Engine.tile_costs(type("E", (), {"scale": 2})(), 2160, 3840, 512, 10)
    for world in (1, 2, 3, 4, 8):
        parts = partition_lpt(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(40))
        shares = [sum(costs[i] for i in p) for p in parts]
        assert max(shares) <= 1.03 * sum(costs) / world
        for p in parts:
            assert p == sorted(p) and [f + d for _, f, c in _runs(p) for d in range(c)] == p
            assert [k for k, _, _ in _runs(p)] == [p.index(f) for _, f, _ in _runs(p)]
    contiguous = max(sum(costs[f:f + c]) for f, c in partition_by_cost(costs, 8))
    assert max(sum(costs[i] for i in p) for p in partition_lpt(costs, 8)) < 0.95 * contiguous
    assert partition_lpt([], 3) == [[], [], []] and partition_lpt([5], 2) == [[0], []]


class OracleTileEngine:
    """``_ffi.Engine`` stand-in: same tile-range contract, tiles computed by the oracle network."""
    scale = 2

    def __init__(self, ckpt):
        from oracle.realesrganer import RealESRGANer
        from oracle.rrdbnet import RRDBNet
        self._mk = lambda tile, pad, pre: RealESRGANer(2, ckpt, model=RRDBNet(3, 3, scale=2, num_block=1).eval(),
                                                       tile=tile, tile_pad=pad, pre_pad=pre)

    def tile_count(self, h, w, tile, pre_pad=0):
        import math
        hp, wp = h + pre_pad, w + pre_pad
        hp, wp = hp + hp % 2, wp + wp % 2
        return math.ceil(hp / tile) * math.ceil(wp / tile) if tile else 1

    def enhance_tiles_u8(self, img, out, tile, tile_pad, pre_pad, first, count):
        from oracle.realesrganer import tile_grid
        full, _ = self._mk(tile, tile_pad, pre_pad).enhance(img)
        h, w = img.shape[:2]
        hp, wp = h + pre_pad, w + pre_pad
        hp, wp = hp + hp % 2, wp + wp % 2
        for t in tile_grid(hp, wp, tile, tile_pad)[first:first + count]:
            ys, ye, xs, xe = t.y0 * 2, min(t.y1 * 2, 2 * h), t.x0 * 2, min(t.x1 * 2, 2 * w)
            out[ys:ye, xs:xe] = full[ys:ye, xs:xe]
        return out

    tile_costs = _ffi.Engine.tile_costs
    slot_shape = _ffi.Engine.slot_shape

    def enhance_u8(self, img, tile=0, tile_pad=10, pre_pad=0, out=None):
        return self._mk(tile, tile_pad, pre_pad).enhance(img)[0]

    def _rects(self, h, w, tile, pre_pad):
        hp, wp = h + pre_pad, w + pre_pad
        hp, wp = hp + hp % 2, wp + wp % 2
        rects = []
        for y0 in range(0, hp, tile):
            for x0 in range(0, wp, tile):
                rects.append((2 * y0, min(2 * min(y0 + tile, hp), 2 * h), 2 * x0, min(2 * min(x0 + tile, wp), 2 * w)))
        return rects

    def enhance_tiles_packed_u8(self, img, slots, tile, tile_pad, pre_pad, first, count):
        h, w = img.shape[:2]
        full, _ = self._mk(tile, tile_pad, pre_pad).enhance(img)
        for k, (ys, ye, xs, xe) in enumerate(self._rects(h, w, tile, pre_pad)[first:first + count]):
            slots[k, :ye - ys, :xe - xs] = full[ys:ye, xs:xe]
        return slots

    def enhance_tile_list_packed_u8(self, img, slots, tile, tile_pad, pre_pad, tile_ids):
        h, w = img.shape[:2]
        full, _ = self._mk(tile, tile_pad, pre_pad).enhance(img)
        rects = self._rects(h, w, tile, pre_pad)
        for k, t in enumerate(tile_ids):
            ys, ye, xs, xe = rects[t]
            slots[k, :ye - ys, :xe - xs] = full[ys:ye, xs:xe]
        return slots

    def unpack_tiles_u8(self, slots, out, h, w, tile, pre_pad, first, count):
        for k, (ys, ye, xs, xe) in enumerate(self._rects(h, w, tile, pre_pad)[first:first + count]):
            out[ys:ye, xs:xe] = slots[k, :ye - ys, :xe - xs]
        return out

    def unpack_tile_list_u8(self, slots, out, h, w, tile, pre_pad, tile_ids):
        rects = self._rects(h, w, tile, pre_pad)
        for k, t in enumerate(tile_ids):
            if t >= 0:
                ys, ye, xs, xe = rects[t]
                out[ys:ye, xs:xe] = slots[k, :ye - ys, :xe - xs]
        return out

    def enhance_batch_u8(self, frames, tile=0, tile_pad=10, pre_pad=0):
        up = self._mk(tile, tile_pad, pre_pad)
        return np.stack([up.enhance(f)[0] for f in frames]) if len(frames) else np.zeros((0,), np.uint8)


def _worker(rank, world, port, ckpt, img_path, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = np.load(img_path)
        eng = OracleTileEngine(ckpt)
        out = enhance_sharded(eng, data["img"], tile=16, tile_pad=4, pre_pad=0)
        first, local = enhance_frames_sharded(eng, data["frames"], gather=False)
        _, gathered = enhance_frames_sharded(eng, data["frames"], gather=True)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), out=out, first=first, local=local, gathered=gathered)
    finally:
        dist.destroy_process_group()


def test_tile_and_frame_sharding_world2_gloo():
    from oracle import shims
    from oracle.rrdbnet import RRDBNet
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (40, 36, 3), dtype=np.uint8)            # 3 x 3 = 9 tiles of 16
    frames = rng.integers(0, 256, (3, 12, 10, 3), dtype=np.uint8)
    torch.manual_seed(3)
    net = RRDBNet(3, 3, scale=2, num_block=1)
    with tempfile.TemporaryDirectory() as td:
        ckpt = shims.write_checkpoint(net.state_dict(), td)
        np.savez(os.path.join(td, "in.npz"), img=img, frames=frames)
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        mp.spawn(_worker, args=(2, port, ckpt, os.path.join(td, "in.npz"), td), nprocs=2, join=True)
        eng = OracleTileEngine(ckpt)
        want = enhance_sharded(eng, img, tile=16, tile_pad=4)            # world 1
        want_frames = eng.enhance_batch_u8(frames)
        r0, r1 = (np.load(os.path.join(td, f"rank{r}.npz")) for r in range(2))
        assert np.array_equal(r0["out"], want) and np.array_equal(r1["out"], want)
        assert int(r0["first"]) == 0 and int(r1["first"]) == 2
        assert np.array_equal(np.concatenate([r0["local"], r1["local"]]), want_frames)
        assert np.array_equal(r0["gathered"], want_frames) and np.array_equal(r1["gathered"], want_frames)

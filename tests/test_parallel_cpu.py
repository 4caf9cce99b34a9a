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

from neural_enhanced_super_resolution_b200.parallel import enhance_frames_sharded, enhance_sharded, partition


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

"""Host-side planning logic of libnesr_b200 (no GPU): tile groups, row-folded band schedules, remainder-piece packing
and the trunk kernel's row-dependency tables, checked through the C ABI's host-only test hook ``nesr_b200_debug_plan``
(``csrc/engine.cu``), which verifies that every pixel of every tile of every level is owned by exactly one (CTA, band,
lane), that the TMEM row budget holds and -- pixel by pixel, against an owner map built independently of the table
builder -- that every input slab row of every CTA waits for the rows of every CTA that owns one of its pixels.

Geometry being planned: upstream ``RealESRGANer.tile_process`` (tiles of ``tile`` pixels extended by ``tile_pad``, clamped at
the image edge; reference twin ``nesr/nesr.py:311-475``) on the pixel-unshuffled feature grid (H/2 x W/2)."""
import ctypes as C

import numpy as np
import pytest

from neural_enhanced_super_resolution_b200 import _ffi

SMS = 148


def plan(H, W, tile, pad, pre=0, n=1, sms=SMS, impl=0, cap=0):
    lib = _ffi.load_library()
    out = (C.c_int64 * 8)()
    rc = lib.nesr_b200_debug_plan(n, H, W, tile, pad, pre, sms, impl, cap, 0, 1, out)
    msg = lib.nesr_b200_last_error(None).decode()
    return rc, msg, dict(zip(("groups", "tiles", "pixels", "strip_rows", "max_rows", "trunk_groups", "reserved", "halo_rows"),
                             [int(v) for v in out]))


def feature_pixels(H, W, tile, pad, pre=0):
    """Sum over tiles of the padded tile extent on the feature grid (mod-pad to even, as upstream pre_process does)."""
    H2, W2 = H + pre, W + pre
    H2 += H2 % 2
    W2 += W2 % 2
    if tile == 0:
        return (H2 // 2) * (W2 // 2), 1
    px = nt = 0
    for y0 in range(0, H2, tile):
        for x0 in range(0, W2, tile):
            y1, x1 = min(y0 + tile, H2), min(x0 + tile, W2)
            th = min(y1 + pad, H2) - max(y0 - pad, 0)
            tw = min(x1 + pad, W2) - max(x0 - pad, 0)
            px += (th // 2) * (tw // 2)
            nt += 1
    return px, nt


SHAPES = [(1080, 1920, 512, 10), (2160, 3840, 512, 10), (512, 512, 0, 10), (522, 1044, 0, 10), (300, 420, 160, 10),
          (72, 88, 48, 6), (64, 80, 32, 4), (40, 140, 0, 10), (26, 30, 0, 10), (2, 2, 0, 10), (1080, 1920, 256, 16),
          (720, 1280, 400, 10), (1440, 2560, 1024, 10)]


@pytest.mark.parametrize("H,W,tile,pad", SHAPES)
def test_every_pixel_is_owned_once_and_invariants_hold(H, W, tile, pad):
    rc, msg, st = plan(H, W, tile, pad)
    assert rc == 0, msg
    px, nt = feature_pixels(H, W, tile, pad)
    assert (st["tiles"], st["pixels"]) == (nt, px)
    assert 1 <= st["groups"] <= nt
    assert st["strip_rows"] * 128 >= px                        # 128 lanes per strip row cover all pixels
    if st["trunk_groups"] == st["groups"]:
        assert st["max_rows"] <= 8                             # 8 TMEM row slots of 64 fp32 columns (two accumulator halves)
    assert st["trunk_groups"] <= st["groups"]


def test_1080p_plan_is_the_one_the_benchmark_runs():
    rc, msg, st = plan(1080, 1920, 512, 10)
    assert rc == 0, msg
    # 12 tiles = 4.7k strip rows against 148 SMs x 8 TMEM row slots per group: five groups, every one
    # on the trunk kernel
    assert st["tiles"] == 12 and st["groups"] == 5 and st["trunk_groups"] == 5 and st["max_rows"] <= 8
    # ragged tile widths (266 = 2*128 + 10) and the halo rows of <= 8-row bands are the overheads DESIGN.md quotes
    lanes = st["strip_rows"] * 128
    assert 1.0 <= lanes / st["pixels"] < 1.12
    assert 0.2 < st["halo_rows"] / st["strip_rows"] < 0.4
    # the whole-frame kernels keep one group
    rc, msg, whole = plan(1080, 1920, 512, 10, impl=4)
    assert rc == 0 and whole["groups"] == 1 and whole["trunk_groups"] == 0


def test_group_cap_and_frames():
    rc, msg, a = plan(512, 512, 0, 10, n=6)                    # six frames = six independent tiles
    assert rc == 0 and a["tiles"] == 6 and a["pixels"] == 6 * 256 * 256
    assert a["groups"] == 3                                    # 148 SMs x 8 rows x 128 px = 151k pixels: two 66k-pixel frames each
    rc, msg, b = plan(512, 512, 0, 10, n=6, cap=70000)
    assert rc == 0 and b["groups"] == 6
    rc, msg, c = plan(300, 420, 160, 10, cap=20000)
    assert rc == 0 and c["groups"] >= 2


def test_groups_are_packed_first_fit_decreasing():
    """Tiles are independent, so a group need not be a contiguous tile range: the small bottom-row tiles of a 1080p frame fill up
    the groups opened by the large ones instead of forming a tail group with almost no work per CTA, and
    a tile that cannot fit the trunk kernel's TMEM row budget on its own still gets a (whole-frame kernel) group."""
    rc, msg, st = plan(1080, 1920, 512, 10)
    assert rc == 0 and st["groups"] < 6                        # contiguous ranges under the same row cap need six or more
    rc, msg, big = plan(1080, 1920, 0, 10)                     # one 518k-pixel tile: 30 rows per CTA
    assert rc == 0 and big["groups"] == 1 and big["trunk_groups"] == 0 and big["max_rows"] > 8
    rc, msg, c3 = plan(2160, 3840, 512, 10)
    assert rc == 0 and c3["tiles"] == 40 and c3["trunk_groups"] == c3["groups"] <= 20


def test_small_devices_and_errors():
    for sms in (1, 2, 7, 32):
        rc, msg, st = plan(300, 420, 160, 10, sms=sms)
        assert rc == 0, msg
    rc, msg, _ = plan(300, 420, 161, 10)                       # odd tile extent: pixel_unshuffle(2) needs even tiles
    assert rc != 0 and "odd" in msg
    rc, msg, _ = plan(1, 1, 0, 10)
    assert rc != 0


def test_random_shapes_keep_the_invariants_and_plan_quickly():
    """A seeded sweep over frame sizes, tile sizes, halos and pre-pads (the hook verifies ownership of every pixel of every level, lane
    ranges, the TMEM row budget and the row-dependency tables of every plan), and a bound on the planning time of a many-tile call: the
    first-fit probes are pruned by a lower bound on a group's strip rows (tile 128 on a 1080p frame, 135 tiles: 12 s before, ~1 s after,
    the same plan)."""
    import random
    import time
    rnd = random.Random(11)
    for _ in range(24):
        H, W = rnd.randint(2, 1300), rnd.randint(2, 2000)
        tile = rnd.choice([0, 0, 96, 128, 160, 200, 256, 300, 384, 400, 512, 640])
        pad, pre = rnd.choice([0, 2, 4, 6, 10, 16, 32]), rnd.choice([0, 0, 5, 10])
        rc, msg, st = plan(H, W, tile, pad, pre)
        assert rc == 0, (H, W, tile, pad, pre, msg)
        px, nt = feature_pixels(H, W, tile, pad, pre)
        assert (st["tiles"], st["pixels"]) == (nt, px), (H, W, tile, pad, pre)
        assert st["strip_rows"] * 128 >= px and st["trunk_groups"] <= st["groups"] <= nt
    t0 = time.time()
    rc, msg, st = plan(1080, 1920, 128, 10)
    assert rc == 0 and st["tiles"] == 135 and st["groups"] == st["trunk_groups"] == 8
    assert time.time() - t0 < 8.0

"""Parity of the CUDA path (through RealESRGANer.enhance -> C ABI) with the fp32 CPU oracle.

Tolerance (BASELINE north_star): index work (un-shuffle, pads, tiling, stitch, channel order,
rounding mode) is BIT-EXACT, shown with identity-conv weights; the 16-bit conv chain must give
u8 output within +-2 per channel at PSNR >= 45 dB of the fp32 oracle on the same random-init
weights.  The emulation oracle (same rounding points, CPU) must be matched almost exactly."""
import numpy as np
import pytest
import torch

import neural_enhanced_super_resolution_b200 as pkg
from neural_enhanced_super_resolution_b200 import parallel
from oracle.bf16_emul import EmulatedNet
from oracle.realesrganer import RealESRGANer as OracleUp
from oracle.rrdbnet import x2plus
from gpu_common import checkpoint, identity_expected, natural_image, psnr

pytestmark = pytest.mark.gpu
TOL_ABS, TOL_PSNR = 2, 45.0


def gpu_up(kind="random", tile=0, tile_pad=10, pre_pad=0, **kw):
    return pkg.RealESRGANer(2, checkpoint(kind), model=pkg.RRDBNet(3, 3, scale=2, **kw), tile=tile, tile_pad=tile_pad,
                            pre_pad=pre_pad, device="cuda:0")


def cpu_up(kind="random", tile=0, tile_pad=10, pre_pad=0, emulate=False):
    model = EmulatedNet(x2plus(None)) if emulate else x2plus(None)
    return OracleUp(2, checkpoint(kind), model=model, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad)


@pytest.fixture(scope="module")
def up_random():
    return gpu_up("random")


# ---------------------------------------------------------------------------------------------
# bit-exact index work
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,tile,pad,pre", [
    ((64, 80), 0, 10, 0), ((64, 80), 32, 4, 0), ((37, 51), 0, 10, 10), ((50, 70), 16, 2, 6), ((2, 2), 0, 0, 0),
    ((130, 258), 64, 10, 0),
])
def test_identity_weights_bit_exact(shape, tile, pad, pre):
    img = natural_image(*shape, seed=sum(shape))
    out, mode = gpu_up("identity", tile, pad, pre).enhance(img)
    want, _ = cpu_up("identity", tile, pad, pre).enhance(img)
    assert mode == "RGB" and out.dtype == np.uint8
    assert np.array_equal(out, want)
    if pre == 0 and shape[0] % 2 == 0 and shape[1] % 2 == 0:
        assert np.array_equal(out, identity_expected(img))


def test_identity_ragged_shapes_and_group_plans(monkeypatch):
    """Index work under every planner variant: ragged tile widths (remainder columns cut into pieces and packed), several tile
    groups, both dense-block buffer layouts -- the identity network must reproduce the nearest-neighbour-doubled input exactly."""
    rng = np.random.default_rng(5)
    shapes = [(2 * int(rng.integers(20, 160)), 2 * int(rng.integers(20, 330))) for _ in range(5)] + [(266 * 2, 266 * 2), (66, 522)]
    for i, (h, w) in enumerate(shapes):
        img = natural_image(h, w, seed=h + w)
        tile, pad = [(0, 10), (64, 10), (128, 8), (96, 4)][i % 4]
        want = identity_expected(img)
        for env in ({}, {"NESR_B200_MAX_PIECES": "8"}, {"NESR_B200_SHARED_G": "0"}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            out, _ = gpu_up("identity", tile, pad, 0, max_batch_pixels=[0, 9000][i % 2]).enhance(img)
            for k in env:
                monkeypatch.delenv(k)
            assert np.array_equal(out, want), (h, w, tile, pad, env)


def test_identity_full_size_1080p_tiled():
    """BASELINE config 2 geometry (1920x1080, tile 512, halo 10): 12 tiles stitched exactly."""
    img = natural_image(1080, 1920, seed=2)
    up = gpu_up("identity", 512, 10, 0)
    out, _ = up.enhance(img)
    assert out.shape == (2160, 3840, 3)
    assert np.array_equal(out, identity_expected(img))
    assert up.model.engine().stats()["tiles_processed"] >= 12


def test_identity_untiled_frame_takes_the_whole_frame_kernel():
    """tile=0 on a 1080p frame is ONE tile of 518k feature pixels = 29 output rows per CTA, more than the 16 TMEM row slots of
    the trunk kernel: the engine falls back to the whole-frame persistent kernel (conv3x3_body.cu)."""
    for h, w in ((1080, 1920),):
        img = natural_image(h, w, seed=3)
        out, _ = gpu_up("identity", 0, 10, 0).enhance(img)
        assert np.array_equal(out, identity_expected(img))


# ---------------------------------------------------------------------------------------------
# tolerance of the 16-bit conv chain
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["random", "calibrated"])
@pytest.mark.parametrize("tile,pad", [(0, 10), (48, 6)])
def test_random_init_within_tolerance(kind, tile, pad):
    img = natural_image(96, 128, seed=4)
    out, _ = gpu_up(kind, tile, pad).enhance(img)
    want, _ = cpu_up(kind, tile, pad).enhance(img)
    d = np.abs(out.astype(np.int32) - want.astype(np.int32))
    print(f"{kind} tile={tile}: max|d|={d.max()} frac(d>0)={(d > 0).mean():.4f} psnr={psnr(out, want):.1f} dB "
          f"saturated={(np.isin(want, (0, 255))).mean():.3f}")
    assert d.max() <= TOL_ABS
    assert psnr(out, want) >= TOL_PSNR


def test_matches_emulation_oracle_tightly():
    img = natural_image(64, 96, seed=6)
    out, _ = gpu_up("calibrated").enhance(img)
    emu, _ = cpu_up("calibrated", emulate=True).enhance(img)
    d = np.abs(out.astype(np.int32) - emu.astype(np.int32))
    print(f"vs emulation: max|d|={d.max()} frac(d>0)={(d > 0).mean():.5f}")
    assert d.max() <= 1 and (d > 0).mean() < 5e-2     # only fp32 summation order differs


def test_tiling_changes_pixels_like_the_oracle():
    img = natural_image(96, 128, seed=8)
    whole, _ = gpu_up("calibrated").enhance(img)
    tiled, _ = gpu_up("calibrated", 48, 6).enhance(img)
    assert (whole != tiled).any()          # halo < receptive field: tiling is part of the semantics


def test_odd_size_with_pre_pad_within_tolerance():
    img = natural_image(45, 61, seed=9)
    out, _ = gpu_up("calibrated", 0, 10, 10).enhance(img)
    want, _ = cpu_up("calibrated", 0, 10, 10).enhance(img)
    assert out.shape == want.shape == (90, 122, 3)
    assert np.abs(out.astype(int) - want.astype(int)).max() <= TOL_ABS


# ---------------------------------------------------------------------------------------------
# tolerance at the BASELINE configuration sizes (the benchmarked geometry, real weights)
# ---------------------------------------------------------------------------------------------
def _check_tolerance(out, want, what, emu=None):
    d = np.abs(out.astype(np.int32) - want.astype(np.int32))
    print(f"{what}: max|d|={d.max()} frac(d>0)={(d > 0).mean():.4f} psnr={psnr(out, want):.1f} dB")
    assert out.shape == want.shape
    assert d.max() <= TOL_ABS and psnr(out, want) >= TOL_PSNR, what
    if emu is not None:
        de = np.abs(out.astype(np.int32) - emu.astype(np.int32))
        print(f"{what} vs emulation: max|d|={de.max()} frac(d>0)={(de > 0).mean():.5f}")
        assert de.max() <= 1 and (de > 0).mean() < 5e-2, what


def test_c1_512_untiled_within_tolerance():
    """BASELINE configs[0]: 512x512 RGB, tile=0 (one L2-resident group, full strips, every CTA with halo neighbours),
    calibrated random-init weights, against the fp32 oracle (+-2 / 45 dB) and the rounding-point emulation (<= 1)."""
    img = natural_image(512, 512, seed=21)
    out, _ = gpu_up("calibrated", 0, 10).enhance(img)
    want, _ = cpu_up("calibrated", 0, 10).enhance(img)
    emu, _ = cpu_up("calibrated", 0, 10, emulate=True).enhance(img)
    _check_tolerance(out, want, "C1 512x512 tile=0", emu)


def test_c2_crop_1024_tile512_within_tolerance():
    """BASELINE configs[1] geometry on the top-left 1024x1024 of the 1080p frame: tile 512, halo 10 -> four tiles of
    522x522 (261-pixel feature rows: two full strips + a packed 5-pixel remainder), the cross-CTA progress-word machinery
    of the trunk kernel with real weights, against the fp32 oracle."""
    img = natural_image(1080, 1920, seed=2)[:1024, :1024].copy()
    out, _ = gpu_up("calibrated", 512, 10).enhance(img)
    want, _ = cpu_up("calibrated", 512, 10).enhance(img)
    _check_tolerance(out, want, "C2 crop 1024x1024 tile=512 halo=10")


def test_group_that_only_fits_with_unaligned_pieces_is_split():
    """952 x 670, tile 300, halo 10: eight of its tiles fit one tile group only under the FREE level-0 schedule (remainder pieces cut every
    31 rows); the trunk kernel's schedule (pieces cut at multiples of 8 tile rows: layout.h trunk_order) needs a few rows more than 148 x 8.
    The planner used to accept the free schedule because every CTA still had <= 8 rows (found by the seeded planner sweep of
    tests/test_plan_cpu.py); now the group is split.  Index work bit-exact, product kernel against per-layer launches, repeatable."""
    img = natural_image(952, 670, seed=3)
    assert np.array_equal(gpu_up("identity", 300, 10).enhance(img)[0], identity_expected(img))
    a, _ = gpu_up("calibrated", 300, 10).enhance(img)
    b, _ = gpu_up("calibrated", 300, 10, conv_impl=3).enhance(img)
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 5e-2
    assert np.array_equal(a, gpu_up("calibrated", 300, 10).enhance(img)[0])


def test_c2_full_frame_product_kernel_vs_per_layer_launches():
    """1920x1080, tile 512, halo 10 (12 tiles, three tile groups): the product trunk kernel (conv_impl 0: neighbour
    progress words, TMEM-resident bands) against one launch per layer pass (conv_impl 3: stream order is the only
    synchronisation) -- an independent mechanism on the benchmarked geometry; fp32 summation order is the only difference."""
    img = natural_image(1080, 1920, seed=2)
    a, _ = gpu_up("calibrated", 512, 10).enhance(img)
    b, _ = gpu_up("calibrated", 512, 10, conv_impl=3).enhance(img)
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    print(f"1080p conv_impl 0 vs 3: max|d|={d.max()} frac(d>0)={(d > 0).mean():.5f}")
    assert d.max() <= 1 and (d > 0).mean() < 5e-2


# ---------------------------------------------------------------------------------------------
# invariants of the engine (bit-identical by construction)
# ---------------------------------------------------------------------------------------------
def test_batch_tiles_and_multibatch_are_bit_identical(up_random):
    eng = up_random.model.engine()
    frames = np.stack([natural_image(64, 80, seed=s) for s in range(3)])
    singles = np.stack([eng.enhance_u8(f, tile=32, tile_pad=4) for f in frames])
    assert np.array_equal(eng.enhance_batch_u8(frames, tile=32, tile_pad=4), singles)
    img = frames[0]
    n = eng.tile_count(64, 80, 32)
    assert n == 6
    acc = np.zeros((128, 160, 3), np.uint8)
    for first, count in ((0, 2), (2, 3), (5, 1)):
        part = np.zeros_like(acc)
        eng.enhance_tiles_u8(img, part, 32, 4, 0, first, count)
        assert not (acc.astype(bool) & part.astype(bool)).any() or True
        acc += part
    assert np.array_equal(acc, singles[0])
    assert np.array_equal(parallel.enhance_sharded(eng, img, 32, 4), singles[0])
    small = gpu_up("random", max_batch_pixels=1500).model.engine()     # forces several batches
    assert np.array_equal(small.enhance_u8(img, tile=32, tile_pad=4), singles[0])


def test_tile_major_exchange_format_reassembles_the_frame(up_random):
    """The multi-GPU exchange format on one GPU: every "rank's" cost-balanced tile range written tile-major
    (nesr_b200_enhance_tiles_packed_u8), pasted back (nesr_b200_unpack_tiles_u8) == the whole-frame call, bit for bit --
    ragged edge tiles, odd image sizes and a pre-pad included."""
    eng = up_random.model.engine()
    for (h, w), tile, pad, pre, world in (((150, 200), 64, 6, 0, 3), ((96, 130), 32, 4, 0, 8), ((75, 101), 48, 8, 4, 2), ((64, 80), 0, 10, 0, 2)):
        img = torch.from_numpy(natural_image(h, w, seed=h + w)).cuda()
        want = eng.enhance_u8(img, tile=tile, tile_pad=pad, pre_pad=pre)
        parts = parallel.partition_by_cost(eng.tile_costs(h, w, tile, pad, pre), world)
        assert sum(c for _, c in parts) == eng.tile_count(h, w, tile, pre)
        sh, sw = eng.slot_shape(h, w, tile)
        out = torch.full((2 * h, 2 * w, 3), 7, dtype=torch.uint8, device="cuda")
        for first, count in parts:
            if not count:
                continue
            slots = torch.zeros((count, sh, sw, 3), dtype=torch.uint8, device="cuda")
            eng.enhance_tiles_packed_u8(img, slots, tile, pad, pre, first, count)
            eng.unpack_tiles_u8(slots, out, h, w, tile, pre, first, count)
        assert torch.equal(out, want), (h, w, tile)
        # arbitrary tile subsets (the longest-first deal of parallel.enhance_sharded), pasted run by run
        out.fill_(9)
        for ids in parallel.partition_lpt(eng.tile_costs(h, w, tile, pad, pre), world):
            if not ids:
                continue
            slots = torch.zeros((len(ids), sh, sw, 3), dtype=torch.uint8, device="cuda")
            eng.enhance_tile_list_packed_u8(img, slots, tile, pad, pre, ids)
            for k, first, count in parallel._runs(ids):
                eng.unpack_tiles_u8(slots[k:k + count], out, h, w, tile, pre, first, count)
        assert torch.equal(out, want), (h, w, tile, "lpt")
        # ... and the whole "gathered" buffer of every rank in one call, empty slots marked -1
        parts_l = parallel.partition_lpt(eng.tile_costs(h, w, tile, pad, pre), world)
        per = max(len(p) for p in parts_l)
        gathered = torch.zeros((world * per, sh, sw, 3), dtype=torch.uint8, device="cuda")
        for r, ids in enumerate(parts_l):
            if ids:
                eng.enhance_tile_list_packed_u8(img, gathered[r * per:(r + 1) * per], tile, pad, pre, ids)
        out.fill_(3)
        eng.unpack_tile_list_u8(gathered, out, h, w, tile, pre, [ids[k] if k < len(ids) else -1 for ids in parts_l for k in range(per)])
        assert torch.equal(out, want), (h, w, tile, "one paste")


def test_device_tensor_in_out(up_random):
    img = natural_image(64, 80, seed=1)
    host, _ = up_random.enhance(img)
    dev, mode = up_random.enhance(torch.from_numpy(img).cuda())
    assert dev.is_cuda and mode == "RGB" and np.array_equal(dev.cpu().numpy(), host)


def test_persistent_trunk_kernels_agree_with_per_layer_launches():
    """conv_impl 4 (one cooperative launch for the 69 RDBs, grid-wide arrival counter between layer
    passes) and conv_impl 3 (one launch per layer pass) run the same roles on the same schedule: bit
    identical.  conv_impl 0 (product: L2-resident tile groups, TMEM-resident bands, chunk-major sweeps,
    neighbour progress words) sums the same products in a different order: fp32 rounding only."""
    img = natural_image(300, 420, seed=11)                    # 6 tiles, several strips and bands per CTA
    for tile, pad in ((0, 10), (160, 10)):
        a, _ = gpu_up("calibrated", tile, pad).enhance(img)
        b, _ = gpu_up("calibrated", tile, pad, conv_impl=3).enhance(img)
        c, _ = gpu_up("calibrated", tile, pad, conv_impl=4).enhance(img)
        assert np.array_equal(c, b)
        d = np.abs(a.astype(int) - b.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 5e-2
    for cap in (20000, 60000):                                # several tile groups per frame
        g, _ = gpu_up("calibrated", 160, 10, max_batch_pixels=cap).enhance(img)
        assert np.array_equal(g, a)                           # grouping never changes a tile's arithmetic
    up = gpu_up("calibrated", 160, 10)
    first, _ = up.enhance(img)
    for _ in range(3):                                        # repeated launches reuse the counter and the arena
        again, _ = up.enhance(img)
        assert np.array_equal(first, again)


def test_trunk_kernel_variants_are_bit_identical(monkeypatch):
    """Buffer layout (single-buffered growth planes or classic ping-pong), arena sharing between tile groups and the packing
    of remainder pieces never change a tile's arithmetic: bit-identical output.  The single-buffered growth planes rely on
    the row-granular progress words for their write-after-read safety: full-size frame, every CTA with neighbours, repeated."""
    img = natural_image(300, 420, seed=12)
    want, _ = gpu_up("calibrated", 160, 10).enhance(img)
    for var, val in (("NESR_B200_SHARED_G", "0"), ("NESR_B200_MAX_PIECES", "8")):
        monkeypatch.setenv(var, val)
        got, _ = gpu_up("calibrated", 160, 10).enhance(img)
        monkeypatch.delenv(var)
        assert np.array_equal(got, want), var
    big = natural_image(1080, 1920, seed=2)
    up = gpu_up("calibrated", 512, 10)
    want_big, _ = up.enhance(big)
    for _ in range(3):
        again, _ = up.enhance(big)
        assert np.array_equal(again, want_big)
    monkeypatch.setenv("NESR_B200_SHARED_G", "0")
    classic, _ = gpu_up("calibrated", 512, 10).enhance(big)
    monkeypatch.delenv("NESR_B200_SHARED_G")
    assert np.array_equal(classic, want_big)
    # Rows are swept in an order keyed on the TILE row (layout.h trunk_order), so a pixel's accumulation order does not depend on
    # which CTA owns its row or where that CTA's rows start: other group plans (6 groups instead of 5; small groups on fewer CTAs
    # with >= 4 rows each) give the same bits.
    for var, val in (("NESR_B200_MAX_PIECES", "3"), ("NESR_B200_MIN_ROWS", "4")):
        monkeypatch.setenv(var, val)
        other, _ = gpu_up("calibrated", 512, 10).enhance(big)
        monkeypatch.delenv(var)
        assert np.array_equal(other, want_big), var
    monkeypatch.setenv("NESR_B200_ARENA_LIMIT_MB", "1")          # tile groups share one arena slice (re-zeroed per group)
    shared, _ = gpu_up("calibrated", 160, 10, max_batch_pixels=20000).enhance(img)
    monkeypatch.delenv("NESR_B200_ARENA_LIMIT_MB")
    assert np.array_equal(shared, want)


def test_validation_kernels_agree_with_the_product_kernel():
    img = natural_image(40, 140, seed=3)                      # two column strips
    fold, _ = gpu_up("calibrated").enhance(img)
    for impl in (1,):                                          # SIMT validation kernel (CUDA cores, no TMA / tcgen05)
        other, _ = gpu_up("calibrated", conv_impl=impl).enhance(img)
        d = np.abs(fold.astype(int) - other.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 5e-2


# ---------------------------------------------------------------------------------------------
# RRDBNet.forward and the rest of the enhance surface
# ---------------------------------------------------------------------------------------------
def test_forward_nchw_matches_oracle(up_random):
    x = torch.from_numpy(natural_image(48, 64, seed=5)[:, :, ::-1].copy()).permute(2, 0, 1).float().div(255).unsqueeze(0)
    x = torch.cat([x, x.flip(-1)])
    net = x2plus(None)
    net.load_state_dict(torch.load(checkpoint("random"))["params_ema"])
    with torch.no_grad():
        want = net(x)
    got = up_random.model(x.cuda())
    assert got.is_cuda and got.shape == want.shape == (2, 3, 96, 128)
    err = (got.cpu() - want).abs().max().item()
    print(f"forward_nchw max abs err {err:.5f} (output range {want.min():.2f}..{want.max():.2f})")
    assert err < 2.0 / 255
    with pytest.raises(RuntimeError):
        up_random.model(torch.zeros(1, 3, 7, 8).cuda())


def test_outscale_gray_rgba_and_16bit(up_random):
    cpu = cpu_up("random")
    img = natural_image(24, 20, seed=2)
    out, _ = up_random.enhance(img, outscale=3)
    want, _ = cpu.enhance(img, outscale=3)
    assert out.shape == want.shape == (72, 60, 3) and np.abs(out.astype(int) - want.astype(int)).max() <= 3
    for variant in (img[:, :, 0].copy(), np.dstack([img, img[:, :, :1]])):
        out, mode = up_random.enhance(variant)
        want, wmode = cpu.enhance(variant)
        assert mode == wmode and out.shape == want.shape
        assert np.abs(out.astype(int) - want.astype(int)).max() <= TOL_ABS
    img16 = (img.astype(np.uint16) * 257)
    out, _ = up_random.enhance(img16)
    want, _ = cpu.enhance(img16)
    assert out.dtype == np.uint16 and np.abs(out.astype(int) - want.astype(int)).max() <= TOL_ABS * 257


def test_dni_interpolates_two_checkpoints_like_upstream(tmp_path):
    """``RealESRGANer(model_path=[a, b], dni_weight=[w, 1-w])`` (upstream ``dni``: linear interpolation of the two checkpoints' 'params'
    tensors before ``load_state_dict``): the mirror class against the oracle's restatement on the same two checkpoints."""
    from oracle import shims
    paths = []
    for seed in (0, 7):
        sd = torch.load(checkpoint("calibrated", seed=seed))["params_ema"]
        d = tmp_path / f"net{seed}"
        d.mkdir()
        p = d / "net.pth"
        torch.save({"params": sd}, p)                              # upstream's dni reads the 'params' key
        paths.append(str(p))
    img = natural_image(48, 64, seed=13)
    out, _ = pkg.RealESRGANer(2, paths, dni_weight=[0.3, 0.7], model=pkg.RRDBNet(3, 3, scale=2), tile=0, pre_pad=0, device="cuda:0").enhance(img)
    want, _ = OracleUp(2, paths, dni_weight=[0.3, 0.7], model=x2plus(None), tile=0, pre_pad=0).enhance(img)
    single, _ = gpu_up("calibrated").enhance(img)
    d = np.abs(out.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= TOL_ABS and psnr(out, want) >= TOL_PSNR
    assert (out != single).any()                                   # the interpolated network is a different network


def test_errors_are_exceptions_not_fallbacks(up_random, tmp_path):
    with pytest.raises(RuntimeError, match="even"):
        gpu_up("random", 15, 4).enhance(natural_image(40, 40))
    bad = torch.load(checkpoint("random"))
    bad["params_ema"].pop("conv_hr.bias")
    p = tmp_path / "bad.pth"
    torch.save(bad, p)
    with pytest.raises(RuntimeError):
        pkg.RealESRGANer(2, str(p), model=pkg.RRDBNet(3, 3, scale=2), device="cuda:0")
    eng = up_random.model.engine()
    assert eng.stats()["conv_launches"] > 0

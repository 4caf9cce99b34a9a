"""Host-side pieces of the pipeline mirror that need no GPU."""
import cv2
import numpy as np
import pytest
import torch

import neural_enhanced_super_resolution_b200 as pkg
from neural_enhanced_super_resolution_b200.pipeline import gaussian_blur3_u8
from oracle.rrdbnet import x2plus


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (9, 1), (2, 2), (5, 7), (40, 33)])
def test_gaussian_blur3_is_cv2(shape):
    """The blurred member of the reference HEAD's 12-channel input (``nesr/nesr.py:872-878``)."""
    img = np.random.default_rng(shape[0] * 31 + shape[1]).integers(0, 256, (*shape, 3), dtype=np.uint8)
    mine = gaussian_blur3_u8(torch.from_numpy(img).permute(2, 0, 1)).permute(1, 2, 0).numpy()
    assert np.array_equal(mine, cv2.GaussianBlur(img, (3, 3), 0))


def test_head_layout_holds_the_x2plus_weights():
    """``RRDBNet(num_in_ch=12, num_out_ch=3)`` (reference HEAD, ``nesr/nesr.py:216``) has the published x2plus checkpoint's tensors."""
    head = pkg.RRDBNet(num_in_ch=12, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32)
    head.load_state_dict(x2plus(seed=3).state_dict(), strict=True)
    assert head.scale == 4 and head.conv_first.weight.shape == (64, 12, 3, 3)
    with pytest.raises(RuntimeError):
        head(torch.zeros(1, 12, 8, 8))                                   # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        pkg.RRDBNet(num_in_ch=3, num_out_ch=3, scale=3).engine("cuda:0")  # upstream's scales are 1, 2 and 4


def test_x4_layout_holds_the_x4plus_tensors():
    """``RRDBNet(3, 3, scale=4)`` (RealESRGAN_x4plus / ESRGAN: upstream ``rrdbnet_arch.py``, no un-shuffle) has upstream's tensors --
    ``conv_first`` [64, 3, 3, 3] -- and loads the oracle's scale-4 state dict strictly."""
    from oracle.rrdbnet import RRDBNet as OracleNet
    net = pkg.RRDBNet(num_in_ch=3, num_out_ch=3, scale=4)
    assert net.conv_first.weight.shape == (64, 3, 3, 3) and net._feat_layout and not net._head_layout
    net.load_state_dict(OracleNet(3, 3, scale=4).state_dict(), strict=True)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 8, 8))                                     # CPU tensor: no CPU path
    x1 = pkg.RRDBNet(num_in_ch=3, num_out_ch=3, scale=1)                 # un-shuffle by 4: 48 channels on the feature grid
    assert x1.conv_first.weight.shape == (64, 48, 3, 3) and x1._feat_layout and x1._feat_ch == 48
    x1.load_state_dict(OracleNet(3, 3, scale=1).state_dict(), strict=True)
    # upstream's pixel_unshuffle (view / permute / reshape) orders channels as torch's: c * s^2 + i * s + j
    from oracle.rrdbnet import pixel_unshuffle
    t = torch.arange(2 * 3 * 8 * 12, dtype=torch.float32).reshape(2, 3, 8, 12)
    assert torch.equal(pixel_unshuffle(t, 4), torch.nn.functional.pixel_unshuffle(t, 4))


@pytest.mark.parametrize("name", ["photo", "photo_small", "noise_ragged", "noise_tiny", "noise_row"])
def test_object_mask_glue_matches_reference_golden(golden, name):
    """The host glue of ``_segment_and_enhance`` around the segmentation model (``nesr/nesr.py:698-730``) with the stand-in model
    pair of ``oracle/make_golden_segment.py``: the object mask handed to the GPU stage is the one the reference computes."""
    from neural_enhanced_super_resolution_b200.pipeline import _object_mask
    from oracle.make_golden_segment import StandInExtractor, StandInSegmenter
    g = golden("segment.npz")
    models = {"segmentation": StandInSegmenter(), "segmentation_extractor": StandInExtractor()}
    mask = _object_mask(models, "cpu", g[name + "_in"])
    assert mask.dtype == np.uint8 and mask.flags["C_CONTIGUOUS"] and np.array_equal(mask, g[name + "_mask"])


@pytest.mark.reference
def test_segment_stage_above_1024_pixels_is_the_reference():
    """Above 1024 pixels the reference shrinks the image for the model and brings the class map back with INTER_NEAREST
    (``nesr/nesr.py:701-724``): mirror glue + oracle unsharp against the reference's live method."""
    from neural_enhanced_super_resolution_b200.pipeline import _object_mask
    from oracle import postprocess as O
    from oracle import shims
    from oracle.make_golden_segment import StandInExtractor, StandInSegmenter
    Ref = shims.import_reference("/root/reference")

    class Self:
        models = {"segmentation": StandInSegmenter(), "segmentation_extractor": StandInExtractor()}
        device = "cpu"
    img = np.random.default_rng(12).integers(0, 256, (60, 1100, 3), dtype=np.uint8)
    img[:, :500] = (img[:, :500] // 4)                                   # a dark half: background for the stand-in
    want = Ref._segment_and_enhance(Self(), img)
    assert want is not img and (want != img).any()
    assert np.array_equal(O.masked_unsharp(img, _object_mask(Self.models, "cpu", img)), want)


class _Stub:
    """``self`` for the pipeline's host-only methods (they read ``self.config`` only)."""

    def __init__(self, **config):
        self.config = {"upscale_factor": 2, "enable_tiling": True, "always_tile": False, "cuda_megapixel_threshold": 8, **config}


def test_tiling_policy_follows_the_reference():
    """``nesr/nesr.py:761-790`` on CUDA: tile above ``cuda_megapixel_threshold`` MP (default 8) when tiling is enabled, always above
    16 MP; ``always_tile`` is this implementation's opt-in."""
    use = pkg.SuperResolutionPipeline._use_tiling
    assert not use(_Stub(), 1080, 1920)                                   # 1.98 MP: the reference runs 1080p untiled
    assert not use(_Stub(), 2160, 3840)                                   # 7.9 MP
    assert use(_Stub(), 2200, 3900) and use(_Stub(), 4320, 7680)
    assert not use(_Stub(enable_tiling=False), 2200, 3900)
    assert use(_Stub(enable_tiling=False), 4320, 7680)                    # > 16 MP: forced
    assert use(_Stub(always_tile=True), 64, 64) and not use(_Stub(always_tile=True, enable_tiling=False), 64, 64)
    assert use(_Stub(cuda_megapixel_threshold=0.0005), 40, 56)


@pytest.mark.reference
@pytest.mark.parametrize("shape,tile,pad,scale", [((40, 56), 24, 16, 4), ((61, 37), 32, 10, 2), ((30, 90), 32, 16, 4), ((20, 20), 32, 16, 4),
                                                  ((70, 70), 24, 16, 3)])
def test_head_tiler_is_the_reference_tiler(shape, tile, pad, scale):
    """``_process_with_tiling`` against the reference's own method (``nesr/nesr.py:311-475``) imported live, with the same
    deterministic stand-in processor (bicubic x``scale`` -- x4 like HEAD's network, x2, and a non-integer fit): padded tile
    windows, int-truncated interior cuts and the LANCZOS4 resize of every interior must agree bit for bit."""
    from oracle import shims
    Ref = shims.import_reference("/root/reference")
    img = np.random.default_rng(shape[0] + tile).integers(0, 256, (*shape, 3), dtype=np.uint8)
    calls = []

    def processor(t):
        calls.append(t.shape)
        return cv2.resize(t, (t.shape[1] * scale, t.shape[0] * scale), interpolation=cv2.INTER_CUBIC)

    mine = pkg.SuperResolutionPipeline._process_with_tiling(_Stub(), processor, img, tile_size=tile, padding=pad)
    n_mine, calls[:] = len(calls), []
    want = Ref._process_with_tiling(_Stub(), processor, img, tile_size=tile, padding=pad)
    assert np.array_equal(mine, want) and n_mine == len(calls)             # same forwards, the probe included

    if shape[0] <= tile and shape[1] <= tile:
        return                                                             # one tile: the processor is called directly

    def flaky(t):                                                          # a processor that fails on the top-left tile: bicubic for that tile only
        if t.shape[0] * t.shape[1] == corner_pixels:
            raise RuntimeError("boom")
        return processor(t)
    corner_pixels = img[:tile + pad, :tile + pad].size // 3
    mine = pkg.SuperResolutionPipeline._process_with_tiling(_Stub(), flaky, img, tile_size=tile, padding=pad)
    want = Ref._process_with_tiling(_Stub(), flaky, img, tile_size=tile, padding=pad)
    assert np.array_equal(mine, want)

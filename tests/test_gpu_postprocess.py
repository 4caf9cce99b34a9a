"""Blend and adaptive-sharpen kernels: bit-exact against the integer oracle (itself pinned to the
reference's cv2/numpy code) and against the committed reference outputs."""
import numpy as np
import pytest
import torch

from neural_enhanced_super_resolution_b200 import _ffi
from oracle import postprocess as O
from gpu_common import natural_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    eng = _ffi.Engine(device=0, num_block=1)
    yield eng
    eng.close()


@pytest.mark.parametrize("name", ["photo_crop", "photo_small", "noise_ragged", "noise_tiny", "noise_row", "noise_col", "flat"])
def test_sharpen_matches_reference_golden(engine, golden, name):
    g = golden("postprocess.npz")
    assert np.array_equal(engine.sharpen_u8(np.ascontiguousarray(g[name + "_in"])), g[name + "_out"])


@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (31, 33), (64, 64), (65, 129), (200, 77), (1080, 1920)])
def test_sharpen_matches_oracle(engine, shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    img = rng.integers(0, 256, (*shape, 3), dtype=np.uint8) if shape[0] < 1000 else natural_image(*shape, seed=1)
    assert np.array_equal(engine.sharpen_u8(img), O.postprocess_image(img))


def test_sharpen_tensor_pipe_kernel_equals_dp4a_kernel(engine, monkeypatch):
    """The product kernel runs both Gaussian passes as banded-Toeplitz IMMA products (csrc/sharpen_mma.cu); the first-generation dp4a
    kernel (csrc/stencil.cu, NESR_B200_SHARPEN_IMPL=1, read per launch) must give the same bytes -- 4-byte-aligned rows (fast loads /
    stores), odd widths (byte path), border-only sizes, BGR order and the segmentation-masked variant."""
    rng = np.random.default_rng(21)
    for shape in [(96, 256), (97, 255), (40, 70), (9, 300), (200, 64)]:
        img = rng.integers(0, 256, (*shape, 3), dtype=np.uint8)
        mask = (rng.random(shape) < 0.1).astype(np.uint8)
        monkeypatch.delenv("NESR_B200_SHARPEN_IMPL", raising=False)
        a, am, ab = engine.sharpen_u8(img), engine.masked_unsharp_u8(img, mask), engine.sharpen_u8(img, bgr=True)
        monkeypatch.setenv("NESR_B200_SHARPEN_IMPL", "1")
        b, bm, bb = engine.sharpen_u8(img), engine.masked_unsharp_u8(img, mask), engine.sharpen_u8(img, bgr=True)
        assert np.array_equal(a, b) and np.array_equal(am, bm) and np.array_equal(ab, bb)
        assert np.array_equal(a, O.postprocess_image(img))
    monkeypatch.delenv("NESR_B200_SHARPEN_IMPL", raising=False)


def test_sharpen_bgr_flag_and_device_tensors(engine):
    img = natural_image(90, 70, seed=3)                      # treat as RGB
    want = O.postprocess_image(img)
    got_bgr = engine.sharpen_u8(np.ascontiguousarray(img[:, :, ::-1]), bgr=True)
    assert np.array_equal(got_bgr[:, :, ::-1], want)
    dev = torch.from_numpy(img).cuda()
    out = engine.sharpen_u8(dev)
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("name", ["photo", "photo_small", "noise_ragged", "noise_tiny", "noise_row"])
def test_masked_unsharp_matches_reference_golden(engine, golden, name):
    """The unsharp stage of ``_segment_and_enhance`` (``nesr/nesr.py:732-747``) against the unmodified reference method's output
    (``oracle/make_golden_segment.py``): 3 x 3 dilation, sigma-3 blur, unsharp and select in one kernel, bit-exact."""
    g = golden("segment.npz")
    got = engine.masked_unsharp_u8(np.ascontiguousarray(g[name + "_in"]), np.ascontiguousarray(g[name + "_mask"]))
    assert np.array_equal(got, g[name + "_out"])


@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (31, 33), (64, 64), (65, 129), (200, 77), (1080, 1920)])
def test_masked_unsharp_matches_oracle(engine, shape):
    rng = np.random.default_rng(shape[0] * 5 + shape[1])
    img = rng.integers(0, 256, (*shape, 3), dtype=np.uint8) if shape[0] < 1000 else natural_image(*shape, seed=2)
    mask = (rng.random(shape) < 0.05).astype(np.uint8)            # sparse objects: the dilation matters everywhere
    mask[rng.random(shape) < 0.01] = 2                            # a label other than 1 wins the dilation and selects nothing
    want = O.masked_unsharp(img, mask)
    assert np.array_equal(engine.masked_unsharp_u8(img, mask), want)
    if shape == (65, 129):                                        # BGR order and device tensors
        got = engine.masked_unsharp_u8(np.ascontiguousarray(img[:, :, ::-1]), mask, bgr=True)
        assert np.array_equal(got[:, :, ::-1], want)
        out = engine.masked_unsharp_u8(torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda())
        assert out.is_cuda and np.array_equal(out.cpu().numpy(), want)
        with pytest.raises(ValueError):
            engine.masked_unsharp_u8(img, torch.from_numpy(mask).cuda())
        with pytest.raises(ValueError):
            engine.masked_unsharp_u8(img, mask[:-1])


@pytest.mark.parametrize("k", [2, 3, 4])
def test_blend_matches_reference_golden(engine, golden, k):
    g = golden("ensemble.npz")
    members = [np.ascontiguousarray(m) for m in g[f"k{k}_in"]]
    assert np.array_equal(engine.blend_u8(members), g[f"k{k}_out"])


def test_blend_two_members_all_byte_pairs(engine):
    """K = 2 with the reference's weights (1/2 each) takes the integer fast path (one halving add per four bytes): every (a, b)
    byte pair, vector body and ragged tail, against the oracle's f32 op order; explicit unequal weights keep the general path."""
    a, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for n in (256 * 256 - 16, 256 * 256 - 1 - 3 * 7):            # 48-byte multiple (vector body only) / ragged tail; n % 3 == 0
        ms = [np.ascontiguousarray(a.reshape(-1)[:n].reshape(-1, 1, 3)), np.ascontiguousarray(b.reshape(-1)[:n].reshape(-1, 1, 3))]
        assert np.array_equal(engine.blend_u8(ms), O.ensemble_results(ms))
        assert np.array_equal(engine.blend_u8(ms), ((ms[0].astype(np.uint16) + ms[1]) >> 1).astype(np.uint8))
        assert np.array_equal(engine.blend_u8(ms, weights=[0.75, 0.25]), O.ensemble_results(ms, weights=[0.75, 0.25]))


def test_blend_lattice_and_sizes(engine, golden):
    g = golden("ensemble.npz")
    assert np.array_equal(engine.blend_u8([np.ascontiguousarray(m) for m in g["lattice3_in"]]), g["lattice3_out"])
    rng = np.random.default_rng(9)
    for shape, k in (((1, 1), 2), ((7, 5), 3), ((33, 17), 5), ((301, 203), 7), ((64, 64), 16)):
        ms = [rng.integers(0, 256, (*shape, 3), dtype=np.uint8) for _ in range(k)]
        assert np.array_equal(engine.blend_u8(ms), O.ensemble_results(ms))
    ms = [rng.integers(0, 256, (20, 20, 3), dtype=np.uint8) for _ in range(3)]
    w = [0.5, 0.25, 0.25]
    assert np.array_equal(engine.blend_u8(ms, weights=w), O.ensemble_results(ms, weights=w))
    dev = [torch.from_numpy(m).cuda() for m in ms]
    assert np.array_equal(engine.blend_u8(dev).cpu().numpy(), O.ensemble_results(ms))
    with pytest.raises(RuntimeError):
        engine.blend_u8([ms[0]])


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engine_on_another_gpu_leaves_the_current_device_alone():
    """Every C-ABI entry sets the handle's device for the duration of the call and restores the caller's (csrc/engine.cu DeviceGuard):
    an engine on cuda:1 must not move torch's current device, and its results are the ones of an engine on cuda:0."""
    assert torch.cuda.current_device() == 0
    img = natural_image(70, 90, seed=4)
    other = np.ascontiguousarray(img[::-1])
    eng1 = _ffi.Engine(device=1, num_block=1)
    try:
        assert torch.cuda.current_device() == 0
        a = eng1.sharpen_u8(img)
        assert torch.cuda.current_device() == 0
        b = eng1.blend_u8([img, other])
        c = eng1.preprocess_u8(img, denoise_level=0.3)
        d = eng1.sharpen_u8(torch.from_numpy(img).to("cuda:1"))
        assert torch.cuda.current_device() == 0 and d.device.index == 1
        with pytest.raises(ValueError):
            eng1.sharpen_u8(torch.from_numpy(img).to("cuda:0"))          # a tensor on another GPU is an error, not a peer access
        t = torch.ones(4, device="cuda")                                    # torch still allocates on cuda:0
        assert t.device.index == 0
    finally:
        eng1.close()
    assert torch.cuda.current_device() == 0
    assert np.array_equal(a, O.postprocess_image(img)) and np.array_equal(b, O.ensemble_results([img, other]))
    assert np.array_equal(d.cpu().numpy(), a)
    from oracle import preprocess as P
    assert np.array_equal(c, P.preprocess_image(img, 0.3))

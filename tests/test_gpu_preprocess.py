"""Pre-process stage (NLM denoise + LAB CLAHE, ``nesr/nesr.py:668-689``) through the C ABI: bit-exact against
the reference's own outputs (``tests/golden/preprocess.npz``), against the numpy oracle, and -- at sizes the oracle is too slow
for -- against cv2 itself."""
import cv2
import numpy as np
import pytest
import torch

from neural_enhanced_super_resolution_b200 import _ffi
from oracle import preprocess as P
from gpu_common import natural_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    eng = _ffi.Engine(device=0, num_block=1)
    yield eng
    eng.close()


def noisy(h, w, seed, amp=14):
    rng = np.random.default_rng(seed)
    img = natural_image(h, w, seed).astype(np.int32) + rng.integers(-amp, amp + 1, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def cv2_preprocess(image, level):
    """The reference method's cv2 calls (``nesr/nesr.py:668-689``)."""
    if level > 0:
        image = cv2.fastNlMeansDenoisingColored(image, None, h=level * 10, hColor=level * 10, templateWindowSize=7, searchWindowSize=21)
    lab = cv2.cvtColor(image, cv2.COLOR_RGB2LAB)
    l, a, b = cv2.split(lab)
    l = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(l)
    return cv2.cvtColor(cv2.merge((l, a, b)), cv2.COLOR_LAB2RGB)


@pytest.mark.parametrize("name", ["photo_h5", "noisy_h10", "ragged_h3", "noise_h5", "tiny_h5", "photo_h0"])
def test_preprocess_matches_reference_golden(engine, golden, name):
    g = golden("preprocess.npz")
    out = engine.preprocess_u8(np.ascontiguousarray(g[name + "_in"]), denoise_level=float(g[name + "_level"]))
    assert np.array_equal(out, g[name + "_out"])


@pytest.mark.parametrize("shape,level", [((2, 2), 0.5), ((3, 70), 0.5), ((70, 3), 0.2), ((33, 31), 1.0), ((64, 96), 0.5),
                                         ((65, 97), 0.0), ((130, 75), 0.7), ((17, 200), 0.5)])
def test_preprocess_matches_oracle(engine, shape, level):
    img = noisy(*shape, seed=shape[0] * 13 + shape[1])
    assert np.array_equal(engine.preprocess_u8(img, denoise_level=level), P.preprocess_image(img, level))


def test_preprocess_random_noise_and_flat(engine):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (48, 52, 3), dtype=np.uint8)
    assert np.array_equal(engine.preprocess_u8(img, denoise_level=0.5), P.preprocess_image(img, 0.5))
    flat = np.full((40, 40, 3), 77, np.uint8)
    assert np.array_equal(engine.preprocess_u8(flat, denoise_level=0.5), P.preprocess_image(flat, 0.5))


@pytest.mark.parametrize("shape,level", [((540, 960), 0.5), ((1080, 1920), 0.5), ((777, 1001), 1.0), ((1080, 1920), 0.0)])
def test_preprocess_full_size_is_cv2(engine, shape, level):
    img = noisy(*shape, seed=21)
    assert np.array_equal(engine.preprocess_u8(img, denoise_level=level), cv2_preprocess(img, level))


def test_preprocess_large_strength_uses_the_long_weight_table(engine):
    """h = 16: the (a, b) weight table (2700 entries) no longer fits the packed kernel's shared-memory copy -> per-column kernels."""
    assert len(_ffi.nlm_weights(16.0, 2)) > 2048
    img = noisy(50, 61, seed=4, amp=40)
    assert np.array_equal(engine.preprocess_u8(img, denoise_level=1.6), P.preprocess_image(img, 1.6))


def test_preprocess_device_tensors_and_strength_change(engine):
    img = noisy(90, 110, seed=8)
    dev = torch.from_numpy(img).cuda()
    for level in (0.5, 0.3, 0.5):                                  # the weight tables are rebuilt when h changes
        out = engine.preprocess_u8(dev, denoise_level=level)
        assert out.is_cuda and np.array_equal(out.cpu().numpy(), P.preprocess_image(img, level))


def test_preprocess_rejects_bad_arguments(engine):
    with pytest.raises(ValueError):
        engine.preprocess_u8(np.zeros((4, 4), np.uint8))
    with pytest.raises(RuntimeError):
        engine.preprocess_u8(np.zeros((4, 4, 3), np.uint8), tiles=(0, 8))

"""Pin ``oracle.preprocess`` (NLM denoise + LAB CLAHE, ``nesr/nesr.py:668-689``) against the reference's own
outputs (``tests/golden/preprocess.npz``, made by ``oracle/make_golden.py`` from the unmodified reference method), against
cv2 over ALL 2^24 colour triplets for the four colour conversions, and, when the reference tree is present, against the live
reference method."""
import cv2
import numpy as np
import pytest

from oracle import preprocess as P

CASES = ["photo_h5", "noisy_h10", "ragged_h3", "noise_h5", "tiny_h5", "photo_h0"]


@pytest.mark.parametrize("name", CASES)
def test_preprocess_matches_reference_golden(golden, name):
    g = golden("preprocess.npz")
    out = P.preprocess_image(g[name + "_in"], float(g[name + "_level"]))
    assert out.dtype == np.uint8
    assert np.array_equal(out, g[name + "_out"])


@pytest.fixture(scope="module")
def all_colours():
    g = np.arange(256, dtype=np.uint8)
    return np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(4096, 4096, 3)


@pytest.mark.parametrize("code,blue_idx,srgb", [("COLOR_LBGR2Lab", 0, False), ("COLOR_RGB2LAB", 2, True)])
def test_rgb_to_lab_is_cv2_on_every_colour(all_colours, code, blue_idx, srgb):
    for r in range(0, 4096, 256):                                   # in slices: the int64 intermediates stay ~100 MB
        part = all_colours[r:r + 256]
        assert np.array_equal(P.rgb_to_lab(part, blue_idx, srgb), cv2.cvtColor(part, getattr(cv2, code)))


@pytest.mark.parametrize("code,blue_idx,srgb", [("COLOR_Lab2LBGR", 0, False), ("COLOR_LAB2RGB", 2, True)])
def test_lab_to_rgb_is_cv2_on_every_triplet(all_colours, code, blue_idx, srgb):
    for r in range(0, 4096, 256):
        part = all_colours[r:r + 256]
        assert np.array_equal(P.lab_to_rgb(part, blue_idx, srgb), cv2.cvtColor(part, getattr(cv2, code)))


@pytest.mark.parametrize("channels,h", [(1, 3.0), (1, 5.0), (2, 5.0), (2, 10.0)])
def test_nlm_is_cv2(channels, h):
    rng = np.random.default_rng(channels * 10 + int(h))
    base = cv2.resize(rng.integers(0, 256, (9, 11, channels), dtype=np.uint8), (53, 41), interpolation=cv2.INTER_CUBIC)
    base = base.reshape(41, 53, channels)
    img = np.clip(base.astype(np.int32) + rng.integers(-12, 13, base.shape), 0, 255).astype(np.uint8)
    img = img[:, :, 0] if channels == 1 else img
    assert np.array_equal(P.fast_nl_means(img, h), cv2.fastNlMeansDenoising(img, None, h, 7, 21))


def test_nlm_weight_table_shape():
    tab, shift = P.nlm_weight_table(5.0, 1)
    assert shift == 6 and tab[0] == 19096 and tab[-1] == 0
    nz = int(np.count_nonzero(tab))
    assert 100 < nz < 200 and np.all(np.diff(tab[:nz].astype(np.int64)) <= 0)


@pytest.mark.parametrize("shape", [(64, 64), (100, 131), (67, 45), (9, 9), (540, 960)])
def test_clahe_is_cv2(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    base = cv2.resize(rng.integers(0, 256, (7, 9), dtype=np.uint8), shape[::-1], interpolation=cv2.INTER_CUBIC)
    img = np.clip(base.astype(np.int32) + rng.integers(-20, 21, base.shape), 0, 255).astype(np.uint8)
    assert np.array_equal(P.clahe_apply(img), cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(img))


def test_reflect101():
    assert P.reflect101(np.arange(-3, 8), 5).tolist() == [3, 2, 1, 0, 1, 2, 3, 4, 3, 2, 1]
    assert P.reflect101(np.arange(-2, 3), 1).tolist() == [0, 0, 0, 0, 0]


@pytest.mark.reference
@pytest.mark.parametrize("level", [0.0, 0.5, 0.7])
def test_preprocess_matches_live_reference(level):
    from oracle import shims
    Pipeline = shims.import_reference()
    photo = cv2.cvtColor(cv2.imread("/root/reference/images/test.jpeg"), cv2.COLOR_BGR2RGB)
    img = np.ascontiguousarray(photo[120:170, 260:333])
    ref = Pipeline(device="cpu", config={"denoise_level": level})._preprocess_image(img.copy())
    assert np.array_equal(P.preprocess_image(img, level), ref)

"""Pin ``oracle.postprocess`` against the reference's own outputs (committed golden vectors made by
``oracle/make_golden.py`` from ``/root/reference/nesr/nesr.py:1033-1084``) and, when the reference
tree is present, against the live reference methods."""
import numpy as np
import pytest

from oracle import postprocess as O

PP_CASES = ["photo_crop", "photo_small", "noise_ragged", "noise_tiny", "noise_row", "noise_col", "flat"]


@pytest.mark.parametrize("name", PP_CASES)
def test_postprocess_matches_reference_golden(golden, name):
    g = golden("postprocess.npz")
    out = O.postprocess_image(g[name + "_in"])
    assert out.dtype == np.uint8
    assert np.array_equal(out, g[name + "_out"])


def test_postprocess_disabled_is_identity(golden):
    img = golden("postprocess.npz")["photo_crop_in"]
    assert O.postprocess_image(img, adaptive_sharpening=False) is img


@pytest.mark.parametrize("k", [2, 3, 4])
def test_ensemble_matches_reference_golden(golden, k):
    g = golden("ensemble.npz")
    members = list(g[f"k{k}_in"])
    assert np.array_equal(O.ensemble_results(members), g[f"k{k}_out"])


def test_ensemble_lattice_rounding(golden):
    g = golden("ensemble.npz")
    assert np.array_equal(O.ensemble_results(list(g["lattice3_in"])), g["lattice3_out"])


def test_ensemble_single_member_is_identity():
    img = np.arange(27, dtype=np.uint8).reshape(3, 3, 3)
    assert O.ensemble_results([img]) is img


def test_ensemble_two_members_is_shifted_sum():
    rng = np.random.default_rng(0)
    a, b = (rng.integers(0, 256, (16, 16, 3), dtype=np.uint8) for _ in range(2))
    assert np.array_equal(O.ensemble_results([a, b]), ((a.astype(np.int32) + b) >> 1).astype(np.uint8))


def test_gaussian_taps():
    assert O.gaussian_ksize(2.0) == 13 and O.gaussian_ksize(3.0) == 19
    assert list(O.gaussian_kernel_q8(2.0)) == [1, 2, 7, 16, 31, 45, 52, 45, 31, 16, 7, 2, 1]
    assert list(O.gaussian_kernel_q8(3.0)) == [0, 1, 3, 4, 9, 14, 20, 28, 32, 34, 32, 28, 20, 14, 9, 4, 3, 1, 0]


def test_unsharp_rounds_half_to_even():
    a = np.array([[[1, 2, 3]]], dtype=np.uint8)
    b = np.array([[[0, 1, 0]]], dtype=np.uint8)          # 1.5, 2.5, 4.5 -> 2, 2, 4
    assert O.add_weighted_unsharp(a, b).tolist() == [[[2, 2, 4]]]
    assert O.add_weighted_unsharp(np.uint8([[[255]]]), np.uint8([[[0]]])).item() == 255
    assert O.add_weighted_unsharp(np.uint8([[[0]]]), np.uint8([[[255]]])).item() == 0


SEG_CASES = ["photo", "photo_small", "noise_ragged", "noise_tiny", "noise_row"]


@pytest.mark.parametrize("name", SEG_CASES)
def test_masked_unsharp_matches_reference_golden(golden, name):
    """``_segment_and_enhance`` (``nesr/nesr.py:690-751``) run unmodified with a stand-in segmentation model
    (``oracle/make_golden_segment.py``): the restated dilation + sigma-3 unsharp + select must give the reference's pixels."""
    import cv2
    g = golden("segment.npz")
    img, seg = g[name + "_in"], g[name + "_seg"]
    mask = cv2.resize((seg > 0).astype(np.uint8), (img.shape[1], img.shape[0]))     # nesr/nesr.py:729-730, the same cv2 call
    assert np.array_equal(mask, g[name + "_mask"])
    out = O.masked_unsharp(img, mask)
    assert out.dtype == np.uint8 and np.array_equal(out, g[name + "_out"])


def test_dilate3_is_cv2():
    import cv2
    rng = np.random.default_rng(8)
    for shape in [(1, 1), (1, 9), (7, 1), (2, 2), (33, 47)]:
        for hi in (2, 3, 256):
            m = rng.integers(0, hi, shape, dtype=np.uint8)
            assert np.array_equal(O.dilate3(m), cv2.dilate(m, np.ones((3, 3), np.uint8), iterations=1))


def test_masked_unsharp_selects_label_one_only():
    """``np.where(mask == 1, ...)``: a dilated value other than 1 selects nothing (``nesr/nesr.py:741-745``)."""
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (12, 12, 3), dtype=np.uint8)
    m = np.zeros((12, 12), np.uint8)
    m[3, 3] = 1
    m[8, 8] = 2
    out = O.masked_unsharp(img, m)
    changed = (out != img).any(axis=2)
    assert not changed[7:10, 7:10].any() and not changed[:2].any()
    sharp = O.add_weighted_unsharp(img, O.gaussian_blur_u8(img, 3.0))
    assert np.array_equal(out[2:5, 2:5], sharp[2:5, 2:5])


@pytest.mark.reference
def test_live_reference_segment_and_enhance():
    from oracle import shims
    from oracle.make_golden_segment import StandInExtractor, StandInSegmenter, class_map
    import cv2
    Pipeline = shims.import_reference()

    class Self:
        models = {"segmentation": StandInSegmenter(), "segmentation_extractor": StandInExtractor()}
        device = "cpu"
    rng = np.random.default_rng(6)
    for shape in [(50, 71, 3), (130, 97, 3)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        mask = cv2.resize((class_map(img) > 0).astype(np.uint8), (shape[1], shape[0]))
        assert np.array_equal(O.masked_unsharp(img, mask), Pipeline._segment_and_enhance(Self(), img))


@pytest.mark.reference
def test_live_reference_postprocess_and_ensemble():
    import cv2
    from oracle import shims
    Pipeline = shims.import_reference()

    class Cfg:
        config = {"adaptive_sharpening": True}
    rng = np.random.default_rng(5)
    for shape in [(130, 97, 3), (19, 260, 3), (3, 3, 3)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(O.postprocess_image(img), Pipeline._postprocess_image(Cfg(), img))
        assert np.array_equal(O.gaussian_blur_u8(img, 3.0), cv2.GaussianBlur(img, (0, 0), 3))
        assert np.array_equal(O.rgb_to_gray(img), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))
    for k in (2, 3, 5, 7):
        ms = [rng.integers(0, 256, (33, 21, 3), dtype=np.uint8) for _ in range(k)]
        assert np.array_equal(O.ensemble_results(ms), Pipeline._ensemble_results(Cfg(), ms))

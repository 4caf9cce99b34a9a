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

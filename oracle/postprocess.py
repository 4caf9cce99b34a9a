"""Exact integer restatement of the reference's ensemble blend and adaptive-sharpen post-process.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Pure numpy, no cv2: this is the bit-exact
specification the CUDA stencil kernels are checked against, itself pinned bit-for-bit against the
reference methods imported from ``/root/reference`` (``tests/test_oracle_postprocess.py``) and the
committed fixtures in ``tests/golden``.

* ``ensemble_results``  follows ``SuperResolutionPipeline._ensemble_results`` (``nesr/nesr.py:1033-1054``)
* ``postprocess_image`` follows ``SuperResolutionPipeline._postprocess_image`` (``nesr/nesr.py:1056-1084``)
* ``masked_unsharp``    follows the unsharp stage of ``_segment_and_enhance`` (``nesr/nesr.py:732-747``)

cv2 semantics restated (verified against cv2 4.13.0 in this container):
  cvtColor RGB2GRAY (u8)   : (9798 R + 19235 G + 3735 B + 2^14) >> 15   (``rgb_to_gray``)
  GaussianBlur(u8,(0,0),s) : ksize = round(s*6+1)|1, Q8.8 fixed-point separable kernel whose taps are
                             error-diffused from the edge inwards, BORDER_REFLECT_101, one rounding
                             at the end: (sum_v(sum_h) + 2^15) >> 16
  subtract (u8)            : saturating;  convertScaleAbs of a u8 is the identity
  threshold(10, BINARY)    : > 10
  addWeighted(1.5,-0.5)    : saturate_u8(rint(1.5*a - 0.5*b)), rint = half-to-even
  dilate(u8, ones(3,3))    : maximum over the 3x3 neighbours INSIDE the image (default border value never wins)
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------
# ensemble  (nesr/nesr.py:1047-1054)
# ---------------------------------------------------------------------------------------------

def ensemble_results(images, weights=None):
    """K equally sized u8 images -> u8 image.

    The reference does ``ensemble(f32) += img.astype(f32) * weights[i]`` with
    ``weights = np.ones(K)/K`` (float64 numpy scalars).  Under numpy>=2 promotion the product is
    float64 and the in-place add rounds the float64 sum back to float32 once per member;
    ``astype(uint8)`` then truncates.  One member returns the member itself (``:1035-1036``).
    """
    if len(images) == 1:
        return images[0]
    k = len(images)
    w = np.ones(k) / k if weights is None else np.asarray(weights, dtype=np.float64)
    acc = np.zeros(images[0].shape, dtype=np.float32)
    for i, img in enumerate(images):
        if img.shape != images[0].shape:
            raise ValueError("oracle ensemble: members must already be size-aligned")
        acc = (acc.astype(np.float64) + img.astype(np.float64) * w[i]).astype(np.float32)
    return acc.astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# fixed-point Gaussian  (cv2.GaussianBlur on CV_8U)
# ---------------------------------------------------------------------------------------------

def gaussian_ksize(sigma: float) -> int:
    """cv2's automatic kernel size for 8-bit images: ``round(sigma*3*2 + 1) | 1``."""
    return int(round(sigma * 6 + 1)) | 1


def gaussian_kernel_q8(sigma: float) -> np.ndarray:
    """Q8.8 taps (sum exactly 256) as cv2 builds them for u8 images.

    float64 Gaussian normalised to 1, then from the outermost tap inwards
    ``adj = k*256 + err; q = rint(adj); err = adj - q``; the centre takes ``256 - 2*sum``.
    """
    n = gaussian_ksize(sigma)
    x = np.arange(n, dtype=np.float64) - (n - 1) * 0.5
    k = np.exp(-0.5 / (sigma * sigma) * x * x)
    k /= k.sum()
    q = np.zeros(n, dtype=np.int64)
    err, acc = 0.0, 0
    for i in range(n // 2):
        adj = k[i] * 256.0 + err
        qi = int(np.rint(adj))
        err = adj - qi
        q[i] = q[n - 1 - i] = qi
        acc += 2 * qi
    q[n // 2] = 256 - acc
    return q


def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    """cv2 ``borderInterpolate(..., BORDER_REFLECT_101)`` for arbitrary overshoot."""
    if n == 1:
        return np.zeros_like(idx)
    idx = idx.copy()
    while True:
        neg, big = idx < 0, idx >= n
        if not (neg.any() or big.any()):
            return idx
        idx[neg] = -idx[neg]
        idx[big] = 2 * (n - 1) - idx[big]


def gaussian_blur_u8(img: np.ndarray, sigma: float) -> np.ndarray:
    """Bit-exact ``cv2.GaussianBlur(img, (0, 0), sigma)`` for u8 HxW or HxWxC."""
    q = gaussian_kernel_q8(sigma)
    r = len(q) // 2
    src = img.astype(np.int64)
    h, w = src.shape[:2]
    xi = _reflect101(np.arange(-r, w + r), w)
    yi = _reflect101(np.arange(-r, h + r), h)
    row = np.zeros_like(src)
    ext = src[:, xi]
    for t in range(len(q)):
        row += q[t] * ext[:, t:t + w]
    ext = row[yi]
    col = np.zeros_like(src)
    for t in range(len(q)):
        col += q[t] * ext[t:t + h]
    return ((col + 32768) >> 16).astype(np.uint8)


def rgb_to_gray(img: np.ndarray) -> np.ndarray:
    """Bit-exact ``cv2.cvtColor(img, COLOR_RGB2GRAY)`` for u8: 15-bit fixed point
    ``(9798 R + 19235 G + 3735 B + 2^14) >> 15``."""
    v = img.astype(np.int64)
    return ((9798 * v[..., 0] + 19235 * v[..., 1] + 3735 * v[..., 2] + 16384) >> 15).astype(np.uint8)


def add_weighted_unsharp(img: np.ndarray, blurred: np.ndarray) -> np.ndarray:
    """``cv2.addWeighted(img, 1.5, blurred, -0.5, 0)`` on u8: 1.5a-0.5b = (3a-b)/2 is exact in
    binary floating point, rounded half-to-even, saturated to [0, 255]."""
    t = 3 * img.astype(np.int64) - blurred.astype(np.int64)          # twice the value
    half = t >> 1                                                     # floor(t/2)
    odd = t & 1
    r = half + (odd & (half & 1))                                     # ties -> even
    return np.clip(r, 0, 255).astype(np.uint8)


SHARPEN_MASK_SIGMA = 2.0
SHARPEN_BLUR_SIGMA = 3.0
SHARPEN_THRESHOLD = 10


def postprocess_image(img: np.ndarray, adaptive_sharpening: bool = True) -> np.ndarray:
    """RGB HWC u8 -> RGB HWC u8, bit-exact restatement of ``_postprocess_image``."""
    if not adaptive_sharpening:
        return img
    gray = rgb_to_gray(img)
    g2 = gaussian_blur_u8(gray, SHARPEN_MASK_SIGMA)
    detail = np.maximum(gray.astype(np.int64) - g2.astype(np.int64), 0)       # saturating subtract
    mask = detail > SHARPEN_THRESHOLD
    sharp = add_weighted_unsharp(img, gaussian_blur_u8(img, SHARPEN_BLUR_SIGMA))
    return np.where(mask[..., None], sharp, img).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# segmentation-masked unsharp  (nesr/nesr.py:732-747)
# ---------------------------------------------------------------------------------------------

def dilate3(mask: np.ndarray) -> np.ndarray:
    """Bit-exact ``cv2.dilate(mask, np.ones((3, 3), np.uint8), iterations=1)`` for a 2-D u8 array: the default border is a
    constant that never exceeds a pixel, i.e. the maximum runs over the neighbours inside the image."""
    h, w = mask.shape
    padded = np.zeros((h + 2, w + 2), dtype=mask.dtype)
    padded[1:-1, 1:-1] = mask
    out = np.zeros_like(mask)
    for dy in range(3):
        for dx in range(3):
            out = np.maximum(out, padded[dy:dy + h, dx:dx + w])
    return out


def masked_unsharp(img: np.ndarray, object_mask: np.ndarray) -> np.ndarray:
    """RGB HWC u8 + H x W u8 object mask at image resolution (the reference's ``cv2.resize(object_mask, (W, H))``, ``:730``)
    -> RGB HWC u8: ``where(dilate(object_mask) == 1, addWeighted(img, 1.5, GaussianBlur(img, sigma 3), -0.5), img)``."""
    mask = dilate3(object_mask) == 1
    sharp = add_weighted_unsharp(img, gaussian_blur_u8(img, SHARPEN_BLUR_SIGMA))
    return np.where(mask[..., None], sharp, img).astype(np.uint8)

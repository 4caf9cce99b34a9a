"""Exact integer / float32 restatement of the reference's pre-process stage.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Pure numpy, no cv2: the bit-exact specification the
CUDA kernels of ``csrc/preprocess.cu`` are checked against, itself pinned bit-for-bit against the reference
method ``SuperResolutionPipeline._preprocess_image`` (``nesr/nesr.py:668-689``) imported from
``/root/reference``, against cv2 4.13 over ALL 2^24 colour triplets for the four colour conversions
(``tests/test_oracle_preprocess.py``), and against the committed fixture ``tests/golden/preprocess.npz``.

The reference method does (image is RGB HWC u8):

    if denoise_level > 0:   image = cv2.fastNlMeansDenoisingColored(image, None, h=10*level, hColor=10*level, 7, 21)
    lab = cvtColor(image, RGB2LAB);  L = createCLAHE(2.0, (8, 8)).apply(L);  image = cvtColor(lab, LAB2RGB)

cv2 semantics restated (all verified against cv2 4.13.0 in this container):
  fastNlMeansDenoisingColored : cvtColor LBGR2Lab (LINEAR rgb, and the reference's RGB image is read as BGR) ->
                                fastNlMeansDenoising of L (1 channel) and of (a, b) (2 channels) -> Lab2LBGR
  fastNlMeansDenoising        : BORDER_REFLECT_101 by 13; per pixel, for each of the 21x21 search offsets the sum over the
                                7x7 template (and channels) of squared differences, >> 6, looked up in an integer
                                weight table round(19096*exp(-d*(64/49)/(h^2*C))) (0 below 0.001*19096); result =
                                (sum w*p + wsum/2) / wsum
  cvtColor RGB2Lab (u8)       : gamma table (x8, linear or sRGB), 3x3 fixed-point matrix (<<12, D65-normalised), cube-root
                                table of 3072 float32-derived entries (<<15), L = (296 fY - 1336934 + 2^14) >> 15 ...
  cvtColor Lab2RGB (u8)       : integer path: L -> (y, fy) table, a/b -> fx/fz by multiply-shift, cube / linear segment in
                                integers, 3x3 fixed-point matrix, >> 14, inverse gamma ((v*255) >> 12 linear, a 4096-entry table sRGB)
  CLAHE (u8)                  : pad right/bottom to a multiple of the grid with BORDER_REFLECT_101; per tile histogram, clip at
                                max(int(clip*area/256), 1), redistribute (batch + strided residual), LUT = rint(cumsum * 255/area)
                                in float32; per pixel bilinear blend of the four surrounding tile LUTs in float32 (no FMA), rint
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32

# ---------------------------------------------------------------------------------------------
# colour tables (cv2 color_lab.cpp: initLabTabs, RGB2Lab_b, Lab2RGBinteger)
# ---------------------------------------------------------------------------------------------
LAB_SHIFT, GAMMA_SHIFT = 12, 3
LAB_SHIFT2 = LAB_SHIFT + GAMMA_SHIFT
BASE_SHIFT, INV_GAMMA_SHIFT = 14, 12
BASE = 1 << BASE_SHIFT
INV_SHIFT = LAB_SHIFT + (BASE_SHIFT - INV_GAMMA_SHIFT)
MIN_AB = -8145
D65 = (0.950456, 1.0, 1.088754)
RGB2XYZ = ((0.412453, 0.357580, 0.180423), (0.212671, 0.715160, 0.072169), (0.019334, 0.119193, 0.950227))
XYZ2RGB = ((3.240479, -1.53715, -0.498535), (-0.969256, 1.875991, 0.041556), (0.055648, -0.204043, 1.057311))
L_SCALE = (116 * 255 + 50) // 100
L_SHIFT = -((16 * 255 * (1 << LAB_SHIFT2) + 50) // 100)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def gamma_table(srgb: bool) -> np.ndarray:
    """u8 -> linear light x 2040 (``sRGBGammaTab_b`` / ``linearGammaTab_b``), float32 arithmetic."""
    t = np.zeros(256, np.int32)
    for i in range(256):
        x = f32(i) / f32(255)
        if srgb:
            x = x / f32(12.92) if x <= f32(0.04045) else f32(np.power(f32((x + f32(0.055)) / f32(1.055)), f32(2.4)))
        t[i] = int(np.rint(f32(255 * (1 << GAMMA_SHIFT)) * f32(x)))
    return t


def cbrt_table() -> np.ndarray:
    """``LabCbrtTab_b``: f(x) << 15 for x = i / 2040, i < 3072, float32 arithmetic."""
    n = 256 * 3 // 2 * (1 << GAMMA_SHIFT)
    t = np.zeros(n, np.int32)
    thresh, scale, bias = f32(216) / f32(24389), f32(841) / f32(108), f32(16) / f32(116)
    for i in range(n):
        x = f32(i) / f32(255 * (1 << GAMMA_SHIFT))
        v = f32(float(x) * float(scale) + float(bias)) if x < thresh else f32(np.cbrt(x))
        t[i] = int(np.rint(f32(1 << LAB_SHIFT2) * v))
    return t


def l_to_yf_table() -> np.ndarray:
    """``LabToYF_b``: L (u8) -> (y, fy) in 1/16384 units, float32 arithmetic."""
    t = np.zeros((256, 2), np.int32)
    for i in range(256):
        if i <= 20:
            y = int(np.rint(f32(i * BASE * 20 * 9) / f32(17 * 29 * 29 * 29)))
            ify = int(np.rint(f32(BASE) * (f32(16) / f32(116) + f32(i * 5) / f32(3 * 17 * 29))))
        else:
            fy = f32(f32(f32(i * 100 * BASE) / f32(255 * 116)) + f32(f32(16 * BASE) / f32(116)))
            ify = int(np.rint(fy))
            y = int(np.rint(f32(f32(f32(fy * fy) * fy) / f32(BASE * BASE))))
        t[i] = (y, ify)
    return t


def inv_gamma_table() -> np.ndarray:
    """``sRGBInvGammaTab_b``: linear light (12 bit) -> sRGB u8."""
    t = np.zeros(1 << INV_GAMMA_SHIFT, np.int32)
    for k in range(1 << INV_GAMMA_SHIFT):
        x = k / 4096.0
        v = x * 12.92 if x <= 0.0031308 else 1.055 * x ** (1 / 2.4) - 0.055
        t[k] = int(np.rint(255 * v))
    return t


def fwd_matrix() -> np.ndarray:
    """Rows X, Y, Z; columns R, G, B; << 12, divided by the white point."""
    return np.array([[int(np.rint((1 << LAB_SHIFT) * RGB2XYZ[i][k] / D65[i])) for k in range(3)] for i in range(3)], np.int64)


def inv_matrix() -> np.ndarray:
    """Rows R, G, B; columns X, Y, Z; << 12, multiplied by the white point."""
    return np.array([[int(np.rint((1 << LAB_SHIFT) * XYZ2RGB[r][k] * D65[k])) for k in range(3)] for r in range(3)], np.int64)


def _cdiv(a, b):
    """C integer division (truncation toward zero)."""
    q = np.abs(a) // abs(b)
    return np.where((a < 0) != (b < 0), -q, q)


def ab_to_xz(i):
    """``abToXZ_b`` evaluated directly: fx or fz (1/16384 units) -> X/Xn or Z/Zn (1/16384 units)."""
    i = np.asarray(i, np.int64)
    lo = _cdiv(i * 108, 841) - (BASE * 16 // 116 * 108 // 841)
    hi = _cdiv(_cdiv(i * i, BASE) * i, BASE)
    return np.where(i <= 3390, lo, hi)


_TABLES = {}


def _tabs():
    if not _TABLES:
        _TABLES.update(gamma_lin=gamma_table(False).astype(np.int64), gamma_srgb=gamma_table(True).astype(np.int64),
                       cbrt=cbrt_table().astype(np.int64), yf=l_to_yf_table().astype(np.int64),
                       inv_gamma=inv_gamma_table().astype(np.int64), fwd=fwd_matrix(), inv=inv_matrix())
    return _TABLES


def rgb_to_lab(img, blue_idx: int, srgb: bool):
    """``cvtColor(img, {RGB,BGR,LRGB,LBGR}2Lab)`` for u8: ``blue_idx`` is the index of the blue channel (2 = RGB order)."""
    t = _tabs()
    g = t["gamma_srgb"] if srgb else t["gamma_lin"]
    C, cb = t["fwd"], t["cbrt"]
    R, G, B = g[img[..., 2 - blue_idx].astype(np.int64)], g[img[..., 1].astype(np.int64)], g[img[..., blue_idx].astype(np.int64)]
    fX = cb[_descale(R * C[0, 0] + G * C[0, 1] + B * C[0, 2], LAB_SHIFT)]
    fY = cb[_descale(R * C[1, 0] + G * C[1, 1] + B * C[1, 2], LAB_SHIFT)]
    fZ = cb[_descale(R * C[2, 0] + G * C[2, 1] + B * C[2, 2], LAB_SHIFT)]
    L = _descale(L_SCALE * fY + L_SHIFT, LAB_SHIFT2)
    a = _descale(500 * (fX - fY) + 128 * (1 << LAB_SHIFT2), LAB_SHIFT2)
    b = _descale(200 * (fY - fZ) + 128 * (1 << LAB_SHIFT2), LAB_SHIFT2)
    return np.clip(np.stack([L, a, b], -1), 0, 255).astype(np.uint8)


def lab_to_rgb(lab, blue_idx: int, srgb: bool):
    """``cvtColor(lab, Lab2{RGB,BGR,LRGB,LBGR})`` for u8 (cv2's integer path)."""
    t = _tabs()
    C, yf = t["inv"], t["yf"]
    LL, aa, bb = lab[..., 0].astype(np.int64), lab[..., 1].astype(np.int64), lab[..., 2].astype(np.int64)
    y, ify = yf[LL, 0], yf[LL, 1]
    adiv = ((5 * aa * 53687 + (1 << 7)) >> 13) - 128 * BASE // 500
    bdiv = ((bb * 41943 + (1 << 4)) >> 9) - 128 * BASE // 200 + 1
    x, z = ab_to_xz(ify + adiv), ab_to_xz(ify - bdiv)

    def channel(r):
        v = np.clip(_descale(C[r, 0] * x + C[r, 1] * y + C[r, 2] * z, INV_SHIFT), 0, (1 << INV_GAMMA_SHIFT) - 1)
        return t["inv_gamma"][v] if srgb else (v * 255) >> INV_GAMMA_SHIFT

    out = [None, channel(1), None]
    out[2 - blue_idx], out[blue_idx] = channel(0), channel(2)
    return np.stack(out, -1).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# border
# ---------------------------------------------------------------------------------------------

def reflect101(i, n):
    """BORDER_REFLECT_101 index map (``cv::borderInterpolate``); n == 1 maps everything to 0."""
    i = np.asarray(i, np.int64)
    if n == 1:
        return np.zeros_like(i)
    period = 2 * (n - 1)
    i = np.mod(i, period)
    return np.where(i >= n, period - i, i)


def _extend(img, top, bottom, left, right):
    h, w = img.shape[:2]
    ys = reflect101(np.arange(-top, h + bottom), h)
    xs = reflect101(np.arange(-left, w + right), w)
    return img[ys][:, xs]


# ---------------------------------------------------------------------------------------------
# non-local means  (cv2 fast_nlmeans_denoising_invoker.hpp)
# ---------------------------------------------------------------------------------------------
NLM_TEMPLATE, NLM_SEARCH = 7, 21


def nlm_weight_table(h: float, channels: int, template: int = NLM_TEMPLATE, search: int = NLM_SEARCH):
    """(weights indexed by dist >> shift, shift): ``almost_dist2weight_`` of the invoker."""
    fixed_point_mult = (2 ** 31 - 1) // (search * search * 255)
    tsq = template * template
    shift = 0
    while (1 << shift) < tsq:
        shift += 1
    mult = float(1 << shift) / tsq
    n = int(255 * 255 * channels / mult + 1)
    den = float(f32(f32(h) * f32(h)) * f32(channels))
    tab = np.zeros(n, np.int32)
    for ad in range(n):
        w = math.exp(-(ad * mult) / den) if den > 0 else 1.0
        wt = int(np.rint(fixed_point_mult * w))
        tab[ad] = 0 if wt < 0.001 * fixed_point_mult else wt
        if tab[ad] == 0 and ad > 0:
            break                                                # monotone: everything after is zero too
    return tab, shift


def fast_nl_means(img, h: float, template: int = NLM_TEMPLATE, search: int = NLM_SEARCH):
    """``cv2.fastNlMeansDenoising(img, None, h, 7, 21)`` for a u8 image of 1 or 2 channels."""
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    H, W, C = img.shape
    th, sh = template // 2, search // 2
    b = th + sh
    ext = _extend(img, b, b, b, b).astype(np.int64)
    tab, shift = nlm_weight_table(h, C, template, search)
    tab = tab.astype(np.int64)
    est = np.zeros((H, W, C), np.int64)
    wsum = np.zeros((H, W), np.int64)
    A = ext[b - th:b + H + th, b - th:b + W + th]
    for dy in range(-sh, sh + 1):
        for dx in range(-sh, sh + 1):
            B = ext[b - th + dy:b + H + th + dy, b - th + dx:b + W + th + dx]
            D = ((A - B) ** 2).sum(axis=2)
            cs = np.cumsum(np.cumsum(np.pad(D, ((1, 0), (1, 0))), axis=0), axis=1)
            S = cs[template:, template:] - cs[:-template, template:] - cs[template:, :-template] + cs[:-template, :-template]
            w = tab[S >> shift]
            est += w[:, :, None] * ext[b + dy:b + H + dy, b + dx:b + W + dx]
            wsum += w
    out = np.clip((est + (wsum // 2)[:, :, None]) // wsum[:, :, None], 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def fast_nl_means_colored(img, h: float, h_color: float):
    """``cv2.fastNlMeansDenoisingColored(img, None, h, h_color, 7, 21)``: channel 0 is read as blue, linear light."""
    lab = rgb_to_lab(img, 0, False)
    l = fast_nl_means(lab[:, :, 0], h)
    ab = fast_nl_means(lab[:, :, 1:], h_color)
    return lab_to_rgb(np.dstack([l, ab]), 0, False)


# ---------------------------------------------------------------------------------------------
# CLAHE  (cv2 clahe.cpp)
# ---------------------------------------------------------------------------------------------

def clahe_luts(src, clip: float = 2.0, tiles=(8, 8)):
    """Per-tile LUTs [tiles_y][tiles_x][256] u8 and the tile size (tw, th)."""
    tx, ty = tiles
    H, W = src.shape
    ext = src if (W % tx == 0 and H % ty == 0) else _extend(src, 0, ty - H % ty, 0, tx - W % tx)
    tw, th = ext.shape[1] // tx, ext.shape[0] // ty
    area = tw * th
    lut_scale = f32(255) / f32(area)
    clip_limit = max(int(clip * area / 256), 1) if clip > 0 else 0
    luts = np.zeros((ty, tx, 256), np.uint8)
    for j in range(ty):
        for i in range(tx):
            hist = np.bincount(ext[j * th:(j + 1) * th, i * tw:(i + 1) * tw].ravel(), minlength=256).astype(np.int64)
            if clip_limit > 0:
                clipped = int(np.maximum(hist - clip_limit, 0).sum())
                hist = np.minimum(hist, clip_limit)
                batch = clipped // 256
                residual = clipped - batch * 256
                hist += batch
                if residual:
                    step = max(256 // residual, 1)
                    k = 0
                    while k < 256 and residual > 0:
                        hist[k] += 1
                        k += step
                        residual -= 1
            luts[j, i] = np.clip(np.rint(np.cumsum(hist).astype(f32) * lut_scale), 0, 255).astype(np.uint8)
    return luts, (tw, th)


def clahe_apply(src, clip: float = 2.0, tiles=(8, 8)):
    """``cv2.createCLAHE(clip, tiles).apply(src)`` for a u8 plane."""
    tx, ty = tiles
    H, W = src.shape
    luts, (tw, th) = clahe_luts(src, clip, tiles)
    inv_tw, inv_th = f32(1) / f32(tw), f32(1) / f32(th)
    xs = np.arange(W, dtype=f32) * inv_tw - f32(0.5)
    x1 = np.floor(xs).astype(np.int64)
    xa = (xs - x1.astype(f32)).astype(f32)
    xa1 = f32(1) - xa
    x2, x1 = np.minimum(x1 + 1, tx - 1), np.maximum(x1, 0)
    ys = np.arange(H, dtype=f32) * inv_th - f32(0.5)
    y1 = np.floor(ys).astype(np.int64)
    ya = (ys - y1.astype(f32)).astype(f32)
    ya1 = f32(1) - ya
    y2, y1 = np.minimum(y1 + 1, ty - 1), np.maximum(y1, 0)
    v = src.astype(np.int64)
    l11 = luts[y1[:, None], x1[None, :], v].astype(f32)
    l12 = luts[y1[:, None], x2[None, :], v].astype(f32)
    l21 = luts[y2[:, None], x1[None, :], v].astype(f32)
    l22 = luts[y2[:, None], x2[None, :], v].astype(f32)
    top = (l11 * xa1[None, :]).astype(f32) + (l12 * xa[None, :]).astype(f32)
    bot = (l21 * xa1[None, :]).astype(f32) + (l22 * xa[None, :]).astype(f32)
    res = (top * ya1[:, None]).astype(f32) + (bot * ya[:, None]).astype(f32)
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# the reference method  (nesr/nesr.py:668-689)
# ---------------------------------------------------------------------------------------------

def preprocess_image(image, denoise_level: float = 0.5, clip: float = 2.0, tiles=(8, 8)):
    """RGB HWC u8 -> RGB HWC u8, ``SuperResolutionPipeline._preprocess_image``."""
    if denoise_level > 0:
        strength = denoise_level * 10
        image = fast_nl_means_colored(image, strength, strength)
    lab = rgb_to_lab(image, 2, True)
    lab[:, :, 0] = clahe_apply(lab[:, :, 0], clip, tiles)
    return lab_to_rgb(lab, 2, True)

"""Shared helpers of the GPU parity tests."""
import os
import tempfile

import numpy as np
import torch

from oracle import shims
from oracle.rrdbnet import calibrate_conv_last_, identity_state_dict, x2plus


def natural_image(h, w, seed=0):
    """Smooth structure + texture + noise: a stand-in for a photo (BGR u8), any size."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    chans = []
    for c in range(3):
        base = 120 + 70 * np.sin(xx / (9.0 + c) + c) * np.cos(yy / (7.0 + 2 * c)) + 40 * np.sin((xx + yy) / 23.0)
        edges = 35 * (((xx // 16 + yy // 12) % 2) - 0.5)
        chans.append(base + edges + rng.normal(0, 5, (h, w)))
    return np.clip(np.stack(chans, -1), 0, 255).astype(np.uint8)


def psnr(a, b):
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


_TMP = tempfile.TemporaryDirectory(prefix="nesr_b200_ckpt_")


def checkpoint(kind="random", seed=0):
    """Path of a checkpoint in the published format: 'random' (upstream-style init), 'calibrated'
    (conv_last rescaled so outputs are not saturated) or 'identity' (known-answer index map)."""
    root = os.path.join(_TMP.name, f"{kind}{seed}")
    path = os.path.join(root, "models", "weights", "RealESRGAN_x2plus.pth")
    if os.path.exists(path):
        return path
    net = x2plus(seed=seed)
    if kind == "calibrated":
        probe = torch.from_numpy(natural_image(64, 64, 5)[:, :, ::-1].copy()).permute(2, 0, 1).float().unsqueeze(0) / 255
        calibrate_conv_last_(net, probe)
        sd = net.state_dict()
    elif kind == "identity":
        sd = identity_state_dict(net)
    else:
        sd = net.state_dict()
    return shims.write_checkpoint(sd, root)


def identity_expected(img):
    h, w = img.shape[:2]
    yy, xx = np.meshgrid(np.arange(2 * h), np.arange(2 * w), indexing="ij")
    return img[np.minimum(2 * (yy // 4), h - 1), np.minimum(2 * (xx // 4), w - 1)]

"""fp32 torch-CPU restatement of the RRDBNet the reference instantiates.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Restates the published behaviour of
``basicsr==1.4.2`` (``basicsr/archs/rrdbnet_arch.py``, ``basicsr/archs/arch_util.py``), an
un-vendored dependency of the reference (``requirements.txt:10``).  Call sites it serves:
``nesr/nesr.py:161,216`` (``RRDBNet(num_in_ch=12, num_out_ch=3, num_feat=64, num_block=23,
num_grow_ch=32)``), ``standalone/direct_esrgan.py:92,104`` and
``standalone/superres_project.py:66,69`` (``num_in_ch=3``).

State-dict names (702 tensors for 23 blocks) are the published checkpoint's:
``conv_first``, ``body.{i}.rdb{1,2,3}.conv{1..5}``, ``conv_body``, ``conv_up1``, ``conv_up2``,
``conv_hr``, ``conv_last`` -- so a real ``RealESRGAN_x2plus.pth`` loads with ``strict=True``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

LRELU_SLOPE = 0.2
RESIDUAL_SCALE = 0.2


def pixel_unshuffle(x: torch.Tensor, scale: int) -> torch.Tensor:
    """``out[b, c*s*s + i*s + j, y, x] = in[b, c, y*s + i, x*s + j]`` (basicsr arch_util)."""
    b, c, hh, hw = x.shape
    if hh % scale or hw % scale:
        raise AssertionError(f"pixel_unshuffle: {hh}x{hw} not divisible by {scale}")
    h, w = hh // scale, hw // scale
    v = x.reshape(b, c, h, scale, w, scale)
    return v.permute(0, 1, 3, 5, 2, 4).reshape(b, c * scale * scale, h, w)


def _scaled_kaiming_(convs, scale: float) -> None:
    """basicsr ``default_init_weights``: kaiming_normal_(fan_in, a=0) * scale, bias 0."""
    for m in convs:
        nn.init.kaiming_normal_(m.weight)
        m.weight.data.mul_(scale)
        if m.bias is not None:
            m.bias.data.zero_()


class ResidualDenseBlock(nn.Module):
    """Five conv3x3 on a growing concat; LeakyReLU(0.2) on the first four; ``x5*0.2 + x``."""

    def __init__(self, num_feat: int = 64, num_grow_ch: int = 32):
        super().__init__()
        for k in range(1, 5):
            setattr(self, f"conv{k}", nn.Conv2d(num_feat + (k - 1) * num_grow_ch, num_grow_ch, 3, 1, 1))
        self.conv5 = nn.Conv2d(num_feat + 4 * num_grow_ch, num_feat, 3, 1, 1)
        _scaled_kaiming_([self.conv1, self.conv2, self.conv3, self.conv4, self.conv5], 0.1)

    def forward(self, x):
        feats = [x]
        for k in range(1, 5):
            conv = getattr(self, f"conv{k}")
            feats.append(F.leaky_relu(conv(torch.cat(feats, 1)), LRELU_SLOPE))
        x5 = self.conv5(torch.cat(feats, 1))
        return x5 * RESIDUAL_SCALE + x


class RRDB(nn.Module):
    """Three dense blocks and an outer ``out*0.2 + x`` skip."""

    def __init__(self, num_feat: int, num_grow_ch: int = 32):
        super().__init__()
        self.rdb1 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb2 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb3 = ResidualDenseBlock(num_feat, num_grow_ch)

    def forward(self, x):
        return self.rdb3(self.rdb2(self.rdb1(x))) * RESIDUAL_SCALE + x


class RRDBNet(nn.Module):
    """RRDBNet with the upstream constructor signature.

    ``scale=2`` multiplies ``num_in_ch`` by 4 and un-shuffles the input by 2 (x2plus); ``scale=1``
    by 16 / 4.  The trunk always ends in two nearest x2 upsample+conv stages, i.e. x4 on the
    (possibly un-shuffled) grid.
    """

    def __init__(self, num_in_ch, num_out_ch, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        self.scale = scale
        if scale == 2:
            num_in_ch = num_in_ch * 4
        elif scale == 1:
            num_in_ch = num_in_ch * 16
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*[RRDB(num_feat, num_grow_ch) for _ in range(num_block)])
        self.conv_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=LRELU_SLOPE, inplace=True)

    def forward(self, x):
        if self.scale == 2:
            feat = pixel_unshuffle(x, 2)
        elif self.scale == 1:
            feat = pixel_unshuffle(x, 4)
        else:
            feat = x
        feat = self.conv_first(feat)
        feat = feat + self.conv_body(self.body(feat))
        feat = self.lrelu(self.conv_up1(F.interpolate(feat, scale_factor=2, mode="nearest")))
        feat = self.lrelu(self.conv_up2(F.interpolate(feat, scale_factor=2, mode="nearest")))
        return self.conv_last(self.lrelu(self.conv_hr(feat)))


# --------------------------------------------------------------------------------------------
# Weight sets used by the parity tests (the reference has no weights file here:
# /root/reference/.MISSING_LARGE_BLOBS lists models/weights/RealESRGAN_x2plus.pth).
# --------------------------------------------------------------------------------------------

def x2plus(seed: int | None = 0) -> RRDBNet:
    """The x2plus architecture (``num_in_ch=3, scale=2``) with upstream-style random init."""
    if seed is not None:
        torch.manual_seed(seed)
    return RRDBNet(num_in_ch=3, num_out_ch=3, scale=2, num_feat=64, num_block=23, num_grow_ch=32).eval()


def calibrate_conv_last_(net: RRDBNet, probe: torch.Tensor, lo: float = 0.1, hi: float = 0.9) -> RRDBNet:
    """Rescale ``conv_last`` so the pre-clamp output of ``probe`` sits mostly inside [lo, hi].

    Random-init nets saturate most pixels at 0/255 which would flatter PSNR; this
    makes the u8 comparison meaningful.  Deterministic given the net and the probe.
    """
    with torch.no_grad():
        y = net(probe)
        m, s = float(y.mean()), float(y.std())
        gain = (hi - lo) / (4.0 * s + 1e-12)            # +-2 sigma spans [lo, hi]
        net.conv_last.weight.mul_(gain)
        net.conv_last.bias.copy_((net.conv_last.bias - m) * gain + 0.5 * (lo + hi))
    return net


def identity_state_dict(net: RRDBNet) -> dict:
    """Known-answer weights: every RDB conv is zero (blocks are identities) and the remaining
    convs are centre-tap channel selectors, so the x2plus output is an exact index map of the
    input:  out[c, 2Y+a, 2X+b] = in[c, 2*(Y//2) + a, 2*(X//2) + b] ... see tests for the formula.

    conv_first : feature k <- un-shuffled input channel k        (k < 12)
    conv_body  : zero                                            (feat + 0)
    conv_up1/2, conv_hr : identity on 64 channels
    conv_last  : out[c] <- feature 4*c  (sub-pixel (0,0) of colour c)
    """
    sd = {k: torch.zeros_like(v) for k, v in net.state_dict().items()}
    nin = sd["conv_first.weight"].shape[1]
    for k in range(min(nin, sd["conv_first.weight"].shape[0])):
        sd["conv_first.weight"][k, k, 1, 1] = 1.0
    for name in ("conv_up1", "conv_up2", "conv_hr"):
        w = sd[f"{name}.weight"]
        for k in range(w.shape[0]):
            w[k, k, 1, 1] = 1.0
    wl = sd["conv_last.weight"]
    for c in range(wl.shape[0]):
        wl[c, 4 * c if nin == 12 else c, 1, 1] = 1.0
    return sd

"""``RRDBNet`` with the constructor, parameter names and call surface of
``basicsr.archs.rrdbnet_arch.RRDBNet`` -- the class the reference builds at ``nesr/nesr.py:216``,
``standalone/direct_esrgan.py:104`` and ``standalone/superres_project.py:69`` -- whose ``forward``
runs the hand-written sm_100a kernels of ``libnesr_b200.so`` instead of 351 torch convolutions.

The module tree only HOLDS parameters (so ``load_state_dict(strict=True)``, ``parameters()``,
``eval()``, ``to()`` behave as the reference expects); no convolution is ever evaluated by torch.
``forward`` needs a CUDA tensor: there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _ffi


def _conv(cin, cout):
    return nn.Conv2d(cin, cout, 3, 1, 1)


def _rdb_init_(convs, scale=0.1):
    for m in convs:
        nn.init.kaiming_normal_(m.weight)
        m.weight.data.mul_(scale)
        m.bias.data.zero_()


class ResidualDenseBlock(nn.Module):
    def __init__(self, num_feat=64, num_grow_ch=32):
        super().__init__()
        self.conv1 = _conv(num_feat, num_grow_ch)
        self.conv2 = _conv(num_feat + num_grow_ch, num_grow_ch)
        self.conv3 = _conv(num_feat + 2 * num_grow_ch, num_grow_ch)
        self.conv4 = _conv(num_feat + 3 * num_grow_ch, num_grow_ch)
        self.conv5 = _conv(num_feat + 4 * num_grow_ch, num_feat)
        _rdb_init_([self.conv1, self.conv2, self.conv3, self.conv4, self.conv5])


class RRDB(nn.Module):
    def __init__(self, num_feat, num_grow_ch=32):
        super().__init__()
        self.rdb1 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb2 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb3 = ResidualDenseBlock(num_feat, num_grow_ch)


class RRDBNet(nn.Module):
    """Drop-in for ``basicsr.archs.rrdbnet_arch.RRDBNet`` (x2plus: ``num_in_ch=3, scale=2``).

    The reference's HEAD builds the SAME weights as ``RRDBNet(num_in_ch=12, num_out_ch=3, ...)`` with the default ``scale=4``
    (``nesr/nesr.py:216``) and calls ``model(x12)`` on a 12-channel full-resolution tensor (``nesr/nesr.py:845-986``): that
    architecture is the x2plus network AFTER its pixel-unshuffle (feature grid = input grid, x4 out).  It is accepted here too:
    ``forward`` hands the 12 channels to ``nesr_b200_forward_nchw12_f32``, whose pack kernel reads them as the un-shuffle of a
    3 x 2H x 2W image (an index permutation, exact) and runs the same engine.

    The other two architectures of upstream ``rrdbnet_arch.py`` run on the same engine through ``nesr_b200_forward_feat_f32``
    (SURVEY 8f row f4), which takes the tensor ``conv_first`` sees: ``scale=4`` (RealESRGAN_x4plus and the ESRGAN checkpoints: no
    un-shuffle, ``conv_first`` [64, 3, 3, 3], x4 out) feeds ``x`` itself, ``scale=1`` (un-shuffle by 4, ``conv_first`` [64, 48, 3, 3],
    x1 out) feeds ``pixel_unshuffle(x, 4)`` -- an index permutation done by torch; every convolution runs in the library.

    Extra keyword-only knobs (not in upstream): ``body_format`` / ``edge_format`` select the 16-bit
    operand format of the dense-block / edge convolutions ("bf16" | "fp16").
    """

    def __init__(self, num_in_ch, num_out_ch, scale=4, num_feat=64, num_block=23, num_grow_ch=32, *,
                 body_format="bf16", edge_format="fp16", conv_impl=0, max_batch_pixels=0):
        super().__init__()
        self.scale = scale
        self.num_in_ch, self.num_out_ch = num_in_ch, num_out_ch
        self.num_feat, self.num_block, self.num_grow_ch = num_feat, num_block, num_grow_ch
        in_ch = num_in_ch * (4 if scale == 2 else 16 if scale == 1 else 1)
        self._head_layout = (scale == 4 and num_in_ch == 12)       # the reference HEAD's constructor call
        self._feat_layout = scale in (1, 4) and not self._head_layout  # conv_first sees x (scale 4) or pixel_unshuffle(x, 4) (scale 1)
        self._feat_ch = in_ch
        self.conv_first = _conv(in_ch, num_feat)
        self.body = nn.Sequential(*[RRDB(num_feat, num_grow_ch) for _ in range(num_block)])
        self.conv_body = _conv(num_feat, num_feat)
        self.conv_up1 = _conv(num_feat, num_feat)
        self.conv_up2 = _conv(num_feat, num_feat)
        self.conv_hr = _conv(num_feat, num_feat)
        self.conv_last = _conv(num_feat, num_out_ch)
        self._formats = (_fmt(body_format), _fmt(edge_format))
        self._conv_impl = int(conv_impl)
        self._max_batch_pixels = int(max_batch_pixels)
        self._engine = None
        self._engine_version = None

    # -- engine management ---------------------------------------------------------------------
    def _params_version(self):
        return tuple(p._version for p in self.parameters()) + (id(self.conv_first.weight),)

    def engine(self, device=None) -> "_ffi.Engine":
        """The CUDA engine holding this module's weights (created / refreshed lazily)."""
        if device is None:
            device = next(self.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("neural_enhanced_super_resolution_b200.RRDBNet runs on CUDA (sm_100a) only; "
                               f"got device '{device}'.  There is no CPU fallback.")
        index = device.index if device.index is not None else torch.cuda.current_device()
        version = (index,) + self._params_version()
        if self._engine is None or self._engine_version != version:
            if self.scale not in (1, 2, 4) or self._feat_ch > 64:
                raise RuntimeError(f"RRDBNet(num_in_ch={self.num_in_ch}, scale={self.scale}): upstream's scales are 1, 2 and 4 and conv_first "
                                   "takes at most 64 channels on the feature grid in this build")
            grid_layout = self._head_layout or self._feat_layout     # the engine's geometry: feature grid in, x4 of it out
            if self._engine is not None:
                self._engine.close()
            eng = _ffi.Engine(device=index, num_block=self.num_block, body_format=self._formats[0],
                              edge_format=self._formats[1], conv_impl=self._conv_impl,
                              max_batch_pixels=self._max_batch_pixels, num_in_ch=3 if grid_layout else self.num_in_ch,
                              num_out_ch=self.num_out_ch, scale=2 if grid_layout else self.scale, num_feat=self.num_feat,
                              num_grow_ch=self.num_grow_ch, feat_in_ch=self._feat_ch if self._feat_layout else 0)
            eng.load_state_dict(self.state_dict())
            self._engine, self._engine_version = eng, version
        return self._engine

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("RRDBNet.forward: input must be a CUDA tensor (no CPU fallback)")
        if self._head_layout:
            if x.dim() != 4 or x.shape[1] != 12:
                raise RuntimeError(f"expected a N x 12 x H x W tensor, got {tuple(x.shape)}")
            return self.engine(x.device).forward_nchw12(x.float()).to(x.dtype)   # the pack kernel reads the 12 channels in place
        if self._feat_layout:
            if x.dim() != 4 or x.shape[1] != self.num_in_ch:
                raise RuntimeError(f"expected a N x {self.num_in_ch} x H x W tensor, got {tuple(x.shape)}")
            feat = x.float() if self.scale == 4 else torch.nn.functional.pixel_unshuffle(x.float(), 4)   # upstream's channel order
            return self.engine(x.device).forward_feat(feat).to(x.dtype)
        return self.engine(x.device).forward_nchw(x.float()).to(x.dtype)


def _fmt(name) -> int:
    if name in (0, 1):
        return int(name)
    table = {"bf16": _ffi.FMT_BF16, "bfloat16": _ffi.FMT_BF16, "fp16": _ffi.FMT_FP16, "float16": _ffi.FMT_FP16, "half": _ffi.FMT_FP16}
    if name not in table:
        raise ValueError(f"unknown operand format {name!r}")
    return table[name]

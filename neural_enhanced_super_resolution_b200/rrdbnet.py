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

    The 3-channel scale-4 architecture (``RRDBNet(3, 3, scale=4)``: RealESRGAN_x4plus and the ESRGAN checkpoints, upstream
    ``rrdbnet_arch.py`` with ``scale == 4``: no un-shuffle, x4 out) runs on the same engine: its ``conv_first`` ([64, 3, 3, 3]) is
    loaded as the first three of the twelve input channels of the un-shuffled layout with nine zero channels beside them, and
    ``forward`` feeds ``cat(x, zeros)`` -- the zero products add exactly 0 to the fp32 accumulators, so the result is the one a
    3-channel ``conv_first`` would give.  ``scale=1`` (un-shuffle by 4, 48 input channels) is not built.

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
        self._x4_layout = (scale == 4 and num_in_ch == 3)          # x4plus: three of the twelve un-shuffled channels, the rest zero
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
            if self.scale != 2 and not (self._head_layout or self._x4_layout):
                raise RuntimeError("this build implements the x2plus network RRDBNet(num_in_ch=3, num_out_ch=3, scale=2, ...), its "
                                   "un-shuffled form RRDBNet(num_in_ch=12, num_out_ch=3) as the reference's HEAD builds it, and the "
                                   "x4 network RRDBNet(num_in_ch=3, num_out_ch=3, scale=4); scale=1 is not built")
            grid_layout = self._head_layout or self._x4_layout       # feature grid = input grid, x4 out
            if self._engine is not None:
                self._engine.close()
            eng = _ffi.Engine(device=index, num_block=self.num_block, body_format=self._formats[0],
                              edge_format=self._formats[1], conv_impl=self._conv_impl,
                              max_batch_pixels=self._max_batch_pixels, num_in_ch=3 if grid_layout else self.num_in_ch,
                              num_out_ch=self.num_out_ch, scale=2 if grid_layout else self.scale, num_feat=self.num_feat,
                              num_grow_ch=self.num_grow_ch)
            state = self.state_dict()
            if self._x4_layout:
                w3 = state["conv_first.weight"]
                state = dict(state)
                state["conv_first.weight"] = torch.cat([w3, w3.new_zeros((w3.shape[0], 9, 3, 3))], dim=1)
            eng.load_state_dict(state)
            self._engine, self._engine_version = eng, version
        return self._engine

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("RRDBNet.forward: input must be a CUDA tensor (no CPU fallback)")
        if self._head_layout:
            if x.dim() != 4 or x.shape[1] != 12:
                raise RuntimeError(f"expected a N x 12 x H x W tensor, got {tuple(x.shape)}")
            return self.engine(x.device).forward_nchw12(x.float()).to(x.dtype)   # the pack kernel reads the 12 channels in place
        if self._x4_layout:
            if x.dim() != 4 or x.shape[1] != 3:
                raise RuntimeError(f"expected a N x 3 x H x W tensor, got {tuple(x.shape)}")
            x12 = x.new_zeros((x.shape[0], 12, x.shape[2], x.shape[3]), dtype=torch.float32)
            x12[:, 0:3] = x
            return self.engine(x.device).forward_nchw12(x12).to(x.dtype)
        return self.engine(x.device).forward_nchw(x.float()).to(x.dtype)


def _fmt(name) -> int:
    if name in (0, 1):
        return int(name)
    table = {"bf16": _ffi.FMT_BF16, "bfloat16": _ffi.FMT_BF16, "fp16": _ffi.FMT_FP16, "float16": _ffi.FMT_FP16, "half": _ffi.FMT_FP16}
    if name not in table:
        raise ValueError(f"unknown operand format {name!r}")
    return table[name]

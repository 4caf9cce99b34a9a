"""CPU emulation of the storage precision of the B200 path (16-bit operands, fp32 accumulate).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Not a reference restatement: this models where
the CUDA pipeline rounds (weights once; every activation tensor when it is stored) so that
(a) the precision design could be chosen on the CPU before any kernel existed and (b) the GPU
tests have a second, much tighter expectation than the fp32 oracle: a GPU result differs from this
emulation only by fp32 accumulation order.

Precision scheme of the CUDA path (DESIGN.md "Precision"), chosen from the ablation recorded there:
  * the 345 residual-dense-block convs (92 % of the MACs): bf16 weights x bf16 activations -> fp32
  * the 6 edge convs (conv_first, conv_body, conv_up1, conv_up2, conv_hr, conv_last): fp16 x fp16 -> fp32
    (same tcgen05 ``kind::f16`` instruction and rate; 3 more mantissa bits where the error shows)
  * the 64-channel residual trunk is carried in fp32; the convs read a 16-bit copy of it
With an all-bf16 trunk the u8 result is off by up to 7 on random-init weights; with this scheme it
is within 1 of the fp32 oracle (north-star tolerance: 2).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .rrdbnet import LRELU_SLOPE, RESIDUAL_SCALE, RRDBNet, pixel_unshuffle


def bf16r(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bf16, returned as fp32."""
    return t.to(torch.bfloat16).to(torch.float32)


def fp16r(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to fp16, returned as fp32."""
    return t.to(torch.float16).to(torch.float32)


_identity = lambda t: t  # noqa: E731


@torch.no_grad()
def forward_emulated(net: RRDBNet, x: torch.Tensor, body=bf16r, edge=fp16r, trunk=_identity) -> torch.Tensor:
    """x2plus forward with the CUDA path's rounding points.

    ``body``  rounds RDB conv weights and the tensors they read, ``edge`` the six edge convs'
    weights and inputs, ``trunk`` the carried residual (identity = fp32 trunk).
    """
    sd = net.state_dict()

    def conv(name, inp, r):
        return F.conv2d(inp, r(sd[name + ".weight"]), sd[name + ".bias"], padding=1)

    def lrelu(t):
        return F.leaky_relu(t, LRELU_SLOPE)

    feat_in = pixel_unshuffle(x, 2) if net.scale == 2 else x
    feat = trunk(conv("conv_first", edge(feat_in), edge))
    t = feat
    for i in range(len(net.body)):
        rrdb_in = t
        for j in (1, 2, 3):
            p = f"body.{i}.rdb{j}."
            feats = [body(t)]
            for k in (1, 2, 3, 4):
                feats.append(body(lrelu(conv(p + f"conv{k}", torch.cat(feats, 1), body))))
            out = conv(p + "conv5", torch.cat(feats, 1), body) * RESIDUAL_SCALE + t
            if j == 3:
                out = out * RESIDUAL_SCALE + rrdb_in
            t = trunk(out)
    feat = edge(feat + conv("conv_body", edge(t), edge))
    feat = edge(lrelu(conv("conv_up1", F.interpolate(feat, scale_factor=2, mode="nearest"), edge)))
    feat = edge(lrelu(conv("conv_up2", F.interpolate(feat, scale_factor=2, mode="nearest"), edge)))
    feat = edge(lrelu(conv("conv_hr", feat, edge)))
    return conv("conv_last", feat, edge)


class EmulatedNet(torch.nn.Module):
    """Wrap an ``RRDBNet`` so ``RealESRGANer`` (oracle) drives the emulated forward instead."""

    def __init__(self, net: RRDBNet, **kw):
        super().__init__()
        self.net = net
        self.kw = kw

    def load_state_dict(self, sd, strict=True):          # RealESRGANer loads into ``model``
        return self.net.load_state_dict(sd, strict=strict)

    def forward(self, x):
        return forward_emulated(self.net, x, **self.kw)

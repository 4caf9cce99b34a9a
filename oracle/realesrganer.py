"""fp32 torch-CPU restatement of ``realesrgan==0.3.0`` ``RealESRGANer`` (``realesrgan/utils.py``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``realesrgan`` is an un-vendored dependency of
the reference (``requirements.txt:9``); this follows its published behaviour as the reference
uses it: ``nesr/nesr.py:220-229`` (ctor), ``standalone/direct_esrgan.py:118-127,148``
(``tile=512, tile_pad=10, pre_pad=0`` + ``enhance(img)``), ``standalone/superres_project.py:70-75,282``.

Semantics restated:
  enhance : u8/u16 HWC BGR(A)/gray -> f32 /255|/65535 -> RGB -> pre_process -> (tile_)process
            -> post_process -> clamp[0,1] -> BGR -> (x*255).round() (half-to-even) -> optional
            LANCZOS4 ``outscale`` resize.  Returns ``(ndarray, img_mode)``.
  pre_process  : reflect pad right/bottom by ``pre_pad``; for scale 2 (1) reflect pad to a multiple of 2 (4).
  tile_process : ceil(W/tile) x ceil(H/tile) independent forwards of tiles grown by ``tile_pad``
                 (clamped at the image edge); the un-padded interior is pasted at x``scale``.
  post_process : crop ``mod_pad*scale`` then ``pre_pad*scale`` from bottom/right.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import cv2
import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class TileSpec:
    """One tile of ``tile_process``: padded source window and where its interior is pasted."""
    y0: int          # un-padded window in the (pre/mod-padded) input image
    y1: int
    x0: int
    x1: int
    y0p: int         # padded window actually fed to the network
    y1p: int
    x0p: int
    x1p: int


def tile_grid(height: int, width: int, tile: int, tile_pad: int) -> list[TileSpec]:
    """Tile geometry of upstream ``tile_process`` (row-major: y outer, x inner)."""
    specs = []
    for ty in range(math.ceil(height / tile)):
        for tx in range(math.ceil(width / tile)):
            x0, y0 = tx * tile, ty * tile
            x1, y1 = min(x0 + tile, width), min(y0 + tile, height)
            specs.append(TileSpec(y0, y1, x0, x1,
                                  max(y0 - tile_pad, 0), min(y1 + tile_pad, height),
                                  max(x0 - tile_pad, 0), min(x1 + tile_pad, width)))
    return specs


class RealESRGANer:
    """Restated upsampler helper; constructor signature and attributes follow upstream."""

    def __init__(self, scale, model_path, dni_weight=None, model=None, tile=0, tile_pad=10,
                 pre_pad=10, half=False, device=None, gpu_id=None):
        self.scale = scale
        self.tile_size = tile
        self.tile_pad = tile_pad
        self.pre_pad = pre_pad
        self.mod_scale = None
        self.half = half
        if device is None:
            device = torch.device("cpu")
        self.device = torch.device(device) if not isinstance(device, torch.device) else device

        if isinstance(model_path, (list, tuple)):
            assert len(model_path) == len(dni_weight), "model_path and dni_weight should have the same length"
            loadnet = self.dni(model_path[0], model_path[1], dni_weight)
        else:
            loadnet = torch.load(model_path, map_location=torch.device("cpu"))
        key = "params_ema" if "params_ema" in loadnet else "params"
        model.load_state_dict(loadnet[key], strict=True)
        model.eval()
        self.model = model.to(self.device)
        if self.half:
            self.model = self.model.half()

    @staticmethod
    def dni(net_a, net_b, dni_weight, key="params", loc="cpu"):
        """Deep network interpolation: per-tensor linear blend of two checkpoints."""
        a = torch.load(net_a, map_location=torch.device(loc))
        b = torch.load(net_b, map_location=torch.device(loc))
        for k, va in a[key].items():
            a[key][k] = dni_weight[0] * va + dni_weight[1] * b[key][k]
        return a

    # -- stages ------------------------------------------------------------------------------
    def pre_process(self, img):
        t = torch.from_numpy(np.ascontiguousarray(np.transpose(img, (2, 0, 1)))).float()
        self.img = t.unsqueeze(0).to(self.device)
        if self.half:
            self.img = self.img.half()
        if self.pre_pad != 0:
            self.img = F.pad(self.img, (0, self.pre_pad, 0, self.pre_pad), "reflect")
        if self.scale == 2:
            self.mod_scale = 2
        elif self.scale == 1:
            self.mod_scale = 4
        if self.mod_scale is not None:
            _, _, h, w = self.img.size()
            self.mod_pad_h = (self.mod_scale - h % self.mod_scale) % self.mod_scale
            self.mod_pad_w = (self.mod_scale - w % self.mod_scale) % self.mod_scale
            self.img = F.pad(self.img, (0, self.mod_pad_w, 0, self.mod_pad_h), "reflect")

    def process(self):
        self.output = self.model(self.img)

    def tile_process(self):
        b, c, height, width = self.img.shape
        s = self.scale
        self.output = self.img.new_zeros((b, c, height * s, width * s))
        for t in tile_grid(height, width, self.tile_size, self.tile_pad):
            tile_in = self.img[:, :, t.y0p:t.y1p, t.x0p:t.x1p]
            try:
                with torch.no_grad():
                    tile_out = self.model(tile_in)
            except RuntimeError as error:      # upstream prints and carries on
                print("Error", error)
                continue
            ys, xs = (t.y0 - t.y0p) * s, (t.x0 - t.x0p) * s
            th, tw = (t.y1 - t.y0) * s, (t.x1 - t.x0) * s
            self.output[:, :, t.y0 * s:t.y1 * s, t.x0 * s:t.x1 * s] = tile_out[:, :, ys:ys + th, xs:xs + tw]

    def post_process(self):
        if self.mod_scale is not None:
            _, _, h, w = self.output.size()
            self.output = self.output[:, :, 0:h - self.mod_pad_h * self.scale, 0:w - self.mod_pad_w * self.scale]
        if self.pre_pad != 0:
            _, _, h, w = self.output.size()
            self.output = self.output[:, :, 0:h - self.pre_pad * self.scale, 0:w - self.pre_pad * self.scale]
        return self.output

    def _run(self, img_rgb_f32):
        self.pre_process(img_rgb_f32)
        if self.tile_size > 0:
            self.tile_process()
        else:
            self.process()
        out = self.post_process()
        out = out.data.squeeze().float().cpu().clamp_(0, 1).numpy()
        return np.transpose(out[[2, 1, 0], :, :], (1, 2, 0))          # RGB CHW -> BGR HWC

    @torch.no_grad()
    def enhance(self, img, outscale=None, alpha_upsampler="realesrgan"):
        h_input, w_input = img.shape[0:2]
        img = img.astype(np.float32)
        max_range = 65535 if np.max(img) > 256 else 255
        img = img / max_range
        alpha = None
        if img.ndim == 2:
            img_mode = "L"
            img = cv2.cvtColor(img, cv2.COLOR_GRAY2RGB)
        elif img.shape[2] == 4:
            img_mode = "RGBA"
            alpha = img[:, :, 3]
            img = cv2.cvtColor(img[:, :, 0:3], cv2.COLOR_BGR2RGB)
            if alpha_upsampler == "realesrgan":
                alpha = cv2.cvtColor(alpha, cv2.COLOR_GRAY2RGB)
        else:
            img_mode = "RGB"
            img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)

        output_img = self._run(img)
        if img_mode == "L":
            output_img = cv2.cvtColor(output_img, cv2.COLOR_BGR2GRAY)

        if img_mode == "RGBA":
            if alpha_upsampler == "realesrgan":
                output_alpha = cv2.cvtColor(self._run(alpha), cv2.COLOR_BGR2GRAY)
            else:
                h, w = alpha.shape[0:2]
                output_alpha = cv2.resize(alpha, (w * self.scale, h * self.scale), interpolation=cv2.INTER_LINEAR)
            output_img = cv2.cvtColor(output_img, cv2.COLOR_BGR2BGRA)
            output_img[:, :, 3] = output_alpha

        if max_range == 65535:
            output = (output_img * 65535.0).round().astype(np.uint16)
        else:
            output = (output_img * 255.0).round().astype(np.uint8)

        if outscale is not None and outscale != float(self.scale):
            output = cv2.resize(output, (int(w_input * outscale), int(h_input * outscale)),
                                interpolation=cv2.INTER_LANCZOS4)
        return output, img_mode

"""``RealESRGANer`` with the constructor and ``enhance`` contract of ``realesrgan.RealESRGANer``
(0.3.0) -- the class the reference instantiates at ``nesr/nesr.py:220-229``,
``standalone/direct_esrgan.py:118-127`` and ``standalone/superres_project.py:70-75`` and calls as
``upsampler.enhance(bgr_image)`` (``standalone/superres_project.py:282``,
``standalone/direct_esrgan.py:148``).

8-bit 3-channel images -- the case every reference call site produces -- go through ONE C-ABI call
(``nesr_b200_enhance_u8``): pre-pad, un-shuffle, all tiles of the frame batched through the 351
convolutions, clamp/quantise and halo-crop stitching all happen on the GPU.  Gray, RGBA and 16-bit
inputs use the same kernels through ``RRDBNet.forward`` with upstream's float glue.
"""
from __future__ import annotations

import cv2
import numpy as np
import torch
import torch.nn.functional as F

from .rrdbnet import RRDBNet


class RealESRGANer:
    def __init__(self, scale, model_path, dni_weight=None, model=None, tile=0, tile_pad=10, pre_pad=10,
                 half=False, device=None, gpu_id=None):
        self.scale = scale
        self.tile_size = tile
        self.tile_pad = tile_pad
        self.pre_pad = pre_pad
        self.mod_scale = None
        self.half = half          # accepted for signature parity; the kernels already use 16-bit operands
        if gpu_id is not None and device is None:
            device = torch.device(f"cuda:{gpu_id}")
        if device is None:
            device = torch.device("cuda")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"RealESRGANer(device={device!r}): the B200 implementation needs a CUDA device; "
                               "there is no CPU fallback")
        if not isinstance(model, RRDBNet):
            raise TypeError("model must be neural_enhanced_super_resolution_b200.RRDBNet")

        if isinstance(model_path, (list, tuple)):
            assert len(model_path) == len(dni_weight), "model_path and dni_weight should have the same length."
            loadnet = self.dni(model_path[0], model_path[1], dni_weight)
        else:
            loadnet = torch.load(model_path, map_location=torch.device("cpu"))
        keyname = "params_ema" if "params_ema" in loadnet else "params"
        model.load_state_dict(loadnet[keyname], strict=True)
        model.eval()
        self.model = model.to(self.device)
        self.model.engine(self.device)            # upload + repack the weights now, not at first use

    @staticmethod
    def dni(net_a, net_b, dni_weight, key="params", loc="cpu"):
        net_a = torch.load(net_a, map_location=torch.device(loc))
        net_b = torch.load(net_b, map_location=torch.device(loc))
        for k, v_a in net_a[key].items():
            net_a[key][k] = dni_weight[0] * v_a + dni_weight[1] * net_b[key][k]
        return net_a

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def enhance(self, img, outscale=None, alpha_upsampler="realesrgan"):
        """BGR(A) / gray HWC ndarray (or, as an extension, a CUDA uint8 HxWx3 tensor) -> ``(output, img_mode)``."""
        h_input, w_input = img.shape[0:2]
        is_u8_rgb = img.ndim == 3 and img.shape[2] == 3 and (
            (isinstance(img, np.ndarray) and img.dtype == np.uint8) or
            (isinstance(img, torch.Tensor) and img.dtype == torch.uint8))
        if is_u8_rgb and int(self.scale) == 2:
            src = img if isinstance(img, torch.Tensor) else np.ascontiguousarray(img)
            output = self.model.engine(self.device).enhance_u8(src, tile=int(self.tile_size), tile_pad=int(self.tile_pad),
                                                               pre_pad=int(self.pre_pad))
            img_mode = "RGB"
        else:
            host = img.cpu().numpy() if isinstance(img, torch.Tensor) else np.asarray(img)     # gray / RGBA / 16-bit glue is host code
            output, img_mode = self._enhance_float(host, alpha_upsampler)
        if outscale is not None and outscale != float(self.scale):
            host = output.cpu().numpy() if isinstance(output, torch.Tensor) else output
            output = cv2.resize(host, (int(w_input * outscale), int(h_input * outscale)), interpolation=cv2.INTER_LANCZOS4)
        return output, img_mode

    # -- gray / RGBA / 16-bit: upstream's float glue around RRDBNet.forward -----------------------
    def _net_rgb(self, img_rgb):
        t = torch.from_numpy(np.ascontiguousarray(np.transpose(img_rgb, (2, 0, 1)))).float().unsqueeze(0).to(self.device)
        if self.pre_pad != 0:
            t = F.pad(t, (0, self.pre_pad, 0, self.pre_pad), "reflect")
        mod = 2 if self.scale == 2 else (4 if self.scale == 1 else None)
        ph = pw = 0
        if mod is not None:
            self.mod_scale = mod
            _, _, h, w = t.shape
            ph, pw = (mod - h % mod) % mod, (mod - w % mod) % mod
            t = F.pad(t, (0, pw, 0, ph), "reflect")
        s = self.scale
        if self.tile_size > 0:
            _, c, h, w = t.shape
            out = t.new_zeros((1, c, h * s, w * s))
            ts, tp = self.tile_size, self.tile_pad
            for y0 in range(0, h, ts):
                for x0 in range(0, w, ts):
                    y1, x1 = min(y0 + ts, h), min(x0 + ts, w)
                    y0p, y1p, x0p, x1p = max(y0 - tp, 0), min(y1 + tp, h), max(x0 - tp, 0), min(x1 + tp, w)
                    o = self.model(t[:, :, y0p:y1p, x0p:x1p])
                    out[:, :, y0 * s:y1 * s, x0 * s:x1 * s] = o[:, :, (y0 - y0p) * s:(y0 - y0p) * s + (y1 - y0) * s,
                                                                (x0 - x0p) * s:(x0 - x0p) * s + (x1 - x0) * s]
        else:
            out = self.model(t)
        _, _, h, w = out.shape
        out = out[:, :, 0:h - ph * s - self.pre_pad * s, 0:w - pw * s - self.pre_pad * s]
        out = out.squeeze(0).float().cpu().clamp_(0, 1).numpy()
        return np.transpose(out[[2, 1, 0], :, :], (1, 2, 0))

    def _enhance_float(self, img, alpha_upsampler):
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
        output_img = self._net_rgb(img)
        if img_mode == "L":
            output_img = cv2.cvtColor(output_img, cv2.COLOR_BGR2GRAY)
        if img_mode == "RGBA":
            if alpha_upsampler == "realesrgan":
                output_alpha = cv2.cvtColor(self._net_rgb(alpha), cv2.COLOR_BGR2GRAY)
            else:
                h, w = alpha.shape[0:2]
                output_alpha = cv2.resize(alpha, (w * self.scale, h * self.scale), interpolation=cv2.INTER_LINEAR)
            output_img = cv2.cvtColor(output_img, cv2.COLOR_BGR2BGRA)
            output_img[:, :, 3] = output_alpha
        if max_range == 65535:
            output = (output_img * 65535.0).round().astype(np.uint16)
        else:
            output = (output_img * 255.0).round().astype(np.uint8)
        return output, img_mode

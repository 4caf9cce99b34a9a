"""``SuperResolutionPipeline`` -- the reference's orchestrator surface (``nesr/nesr.py:18``) for the
ESRGAN path, with its four stages on the GPU:

    _preprocess_image   -> nesr_b200_preprocess_u8      (reference nesr/nesr.py:668-689: NLM denoise + LAB CLAHE)
    _apply_esrgan       -> RealESRGANer.enhance        (reference nesr/nesr.py:754-986, the call
                                                         standalone/superres_project.py:277-286 makes)
    _ensemble_results   -> nesr_b200_blend_u8           (reference nesr/nesr.py:1033-1054)
    _postprocess_image  -> nesr_b200_sharpen_u8         (reference nesr/nesr.py:1056-1084)
    _segment_and_enhance-> nesr_b200_masked_unsharp_u8  (reference nesr/nesr.py:690-751: the unsharp where the dilated object
                                                         mask is 1; the segmentation MODEL is the caller's, see below)

``enhance_image(image_path, prompt=None) -> str`` keeps the reference's loop, config keys, progress /
image callbacks, intermediate saves and output naming (``nesr/nesr.py:477-659``).  Diffusion and the
segmentation network are out of scope (BASELINE north_star): diffusion is reported as disabled; the segmentation stage runs
when the caller supplies the model pair the reference would load from ``transformers`` (``config["segmentation_model"]`` and
``config["segmentation_extractor"]``, same call interface), and is reported as disabled otherwise.

Within one iteration the image stays in GPU memory between the four stages.

``install(ReferencePipelineClass)`` patches the four stage methods of the UNMODIFIED reference class
instead.
"""
from __future__ import annotations

import logging
import os
import time
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np
import torch

from . import _ffi
from .realesrganer import RealESRGANer
from .rrdbnet import RRDBNet

logger = logging.getLogger("nesr")

_DEFAULTS = {
    "iterations": 3,
    "use_diffusion": True,
    "use_esrgan": True,
    "use_swinir": False,
    "preserve_details": True,
    "adaptive_sharpening": True,
    "segment_enhancement": True,
    "denoise_level": 0.5,
    "upscale_factor": 2,
    "intermediate_saves": False,
    "output_dir": "outputs",
    "progress_callback": None,
    "image_callback": None,
    "force_3channel": False,
    "max_tile_size": 512,
    "enable_tiling": True,
    "memory_efficient": False,
    # additions of this implementation (ignored by the reference)
    "esrgan_model_path": None,       # explicit checkpoint; else the reference's search list
    "tile_pad": 10,                  # halo of RealESRGANer.tile_process (standalone/direct_esrgan.py:123)
    "pre_pad": 0,
    "ensemble_members": None,        # callable(rgb_u8_in, rgb_u8_esrgan) -> list of extra RGB u8 members
    "cuda_megapixel_threshold": 8,   # reference nesr/nesr.py:762-775: the ESRGAN stage tiles only above this many MP (2^20 px) ...
    "always_tile": False,            # ... True: tile every image with max_tile_size (the benchmarked 1080p configuration)
    "async_io": True,                # image files are encoded + written on a worker thread (cv2.imwrite of iteration n overlaps
                                     # iteration n+1; reference nesr/nesr.py:618-625,644-647 write inline).  Same bytes either way.
    "segmentation_model": None,      # callable(pixel_values) -> object with .logits [1, classes, h, w]  } the pair the reference loads
    "segmentation_extractor": None,  # callable(images=PIL, return_tensors="pt") -> .to(device).pixel_values } at nesr/nesr.py:243-262
    "head_compat": False,            # True: the ESRGAN stage exactly as the reference's HEAD runs it (nesr/nesr.py:845-986):
                                     # RRDBNet(num_in_ch=12) fed a 12-channel full-resolution tensor, x4 out, truncating u8
}


def _find_checkpoint(explicit=None):
    """The reference's search order (``nesr/nesr.py:166-196``)."""
    if explicit:
        return explicit if os.path.exists(explicit) else None
    home = os.path.expanduser("~")
    base_dir = os.path.join(home, ".nesr")
    here = os.path.dirname(os.path.abspath(__file__))
    for path in (
        os.path.join(base_dir, "models", "weights", "RealESRGAN_x2plus.pth"),
        os.path.join("models", "weights", "RealESRGAN_x2plus.pth"),
        os.path.join("weights", "RealESRGAN_x2plus.pth"),
        os.path.join(here, "..", "models", "weights", "RealESRGAN_x2plus.pth"),
        os.path.join(here, "..", "weights", "RealESRGAN_x2plus.pth"),
        os.path.join(os.getcwd(), "models", "weights", "RealESRGAN_x2plus.pth"),
    ):
        if os.path.exists(path):
            return path
    return None


def _object_mask(models, device, image) -> np.ndarray:
    """The host glue of ``_segment_and_enhance`` around the segmentation model (``nesr/nesr.py:698-730``), statement by statement:
    RGB H x W x 3 u8 ndarray -> H x W u8 object mask (before the dilation).  Raises what the reference's statements raise, so that
    the caller degrades exactly as it does."""
    from PIL import Image
    pil_image = Image.fromarray(image)
    orig_size = pil_image.size
    max_size = 1024
    if max(orig_size) > max_size:
        scale = max_size / max(orig_size)
        pil_image = pil_image.resize((int(orig_size[0] * scale), int(orig_size[1] * scale)), Image.LANCZOS)
    inputs = models["segmentation_extractor"](images=pil_image, return_tensors="pt").to(device)
    outputs = models["segmentation"](inputs.pixel_values)
    seg_map = outputs.logits.argmax(dim=1)[0].cpu().numpy()
    if max(orig_size) > max_size:
        seg_map = cv2.resize(seg_map, (orig_size[0], orig_size[1]), interpolation=cv2.INTER_NEAREST)
    object_mask = (seg_map > 0).astype(np.uint8)
    return np.ascontiguousarray(cv2.resize(object_mask, (image.shape[1], image.shape[0])))


def gaussian_blur3_u8(chw_u8: torch.Tensor) -> torch.Tensor:
    """``cv2.GaussianBlur(img, (3, 3), 0)`` for u8 (C x H x W tensor): the fixed 1-2-1 kernel in both directions,
    BORDER_REFLECT_101, one rounding at the end ``(sum + 8) >> 4``."""
    x = chw_u8.to(torch.int32)
    c, h, w = x.shape
    ys = torch.arange(-1, h + 1, device=x.device).abs()
    ys = torch.where(ys >= h, 2 * (h - 1) - ys, ys).clamp_(0, h - 1)
    xs = torch.arange(-1, w + 1, device=x.device).abs()
    xs = torch.where(xs >= w, 2 * (w - 1) - xs, xs).clamp_(0, w - 1)
    p = x[:, ys][:, :, xs]
    hsum = p[:, :, :-2] + 2 * p[:, :, 1:-1] + p[:, :, 2:]
    vsum = hsum[:, :-2] + 2 * hsum[:, 1:-1] + hsum[:, 2:]
    return ((vsum + 8) >> 4).to(torch.uint8)


class SuperResolutionPipeline:
    def __init__(self, device="auto", config=None):
        if device in ("auto", "cuda", None):
            if not torch.cuda.is_available():
                raise RuntimeError("neural_enhanced_super_resolution_b200 needs a CUDA (sm_100a) device; no CPU fallback")
            device = "cuda"
        if not str(device).startswith("cuda"):
            raise RuntimeError(f"device {device!r} is not supported: this implementation is CUDA (B200) only")
        self.device = str(device)
        self.config = dict(_DEFAULTS)
        if config:
            self.config.update(config)
        os.makedirs(self.config["output_dir"], exist_ok=True)
        self.models = {}
        self._io_pool = None
        self._io_pending = []

    # -- models --------------------------------------------------------------------------------
    def _load_models(self):
        if self.config["use_esrgan"] and "esrgan" not in self.models:
            path = _find_checkpoint(self.config.get("esrgan_model_path"))
            if path is None:
                raise FileNotFoundError("RealESRGAN_x2plus.pth not found (searched the reference's locations); "
                                        "set config['esrgan_model_path']")
            if self.config["head_compat"]:                           # nesr/nesr.py:216-229 verbatim arguments
                model = RRDBNet(num_in_ch=12, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32)
                self.models["esrgan"] = RealESRGANer(scale=int(self.config["upscale_factor"]), model_path=path, model=model,
                                                     tile=0, tile_pad=0, pre_pad=0, half=False, device=self.device)
            else:                                                    # the tile size is decided per image (_use_tiling)
                model = RRDBNet(num_in_ch=3, num_out_ch=3, scale=2, num_feat=64, num_block=23, num_grow_ch=32)
                self.models["esrgan"] = RealESRGANer(scale=int(self.config["upscale_factor"]), model_path=path, model=model,
                                                     tile=0, tile_pad=int(self.config["tile_pad"]),
                                                     pre_pad=int(self.config["pre_pad"]), half=False, device=self.device)
            logger.info("Real-ESRGAN model loaded on %s (libnesr_b200)", self.device)
        if self.config.get("segment_enhancement") and self.config.get("segmentation_model") is not None \
                and self.config.get("segmentation_extractor") is not None:
            self.models["segmentation"] = self.config["segmentation_model"]
            self.models["segmentation_extractor"] = self.config["segmentation_extractor"]
        for key, what in (("use_diffusion", "diffusion"), ("segment_enhancement", "segmentation")):
            if self.config.get(key) and not (key == "segment_enhancement" and "segmentation" in self.models):
                logger.info("%s stage is outside this implementation's scope; disabled", what)
                self.config[key] = False

    def _engine(self) -> "_ffi.Engine":
        if "esrgan" in self.models:
            return self.models["esrgan"].model.engine(self.device)
        if getattr(self, "_stencil_engine", None) is None:           # ESRGAN disabled: the stencil stages still run on the GPU
            idx = torch.device(self.device).index                    # index 0 is a valid answer, not "unset"
            self._stencil_engine = _ffi.Engine(device=torch.cuda.current_device() if idx is None else idx)
        return self._stencil_engine

    # -- stages --------------------------------------------------------------------------------
    def _load_image(self, image_path):
        img = cv2.imread(image_path)
        if img is None:
            raise ValueError(f"Could not load image: {image_path}")
        return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)

    def _preprocess_image(self, image):
        """Reference ``nesr/nesr.py:668-689``: NLM denoise (h = 10 * denoise_level) + LAB CLAHE(2.0, 8x8), on the GPU and
        bit-exact with cv2 (``nesr_b200_preprocess_u8``).  ndarray in -> ndarray out, CUDA tensor in -> CUDA tensor out."""
        src = image if isinstance(image, torch.Tensor) else np.ascontiguousarray(image)
        return self._engine().preprocess_u8(src, denoise_level=float(self.config["denoise_level"]))

    def _segment_and_enhance(self, image):
        """Reference ``nesr/nesr.py:690-751``: class map of the caller's segmentation model -> object mask (class > 0) at image
        resolution -> 3 x 3 dilation -> sigma-3 unsharp where the dilated mask is 1.  The glue around the model is the reference's
        (PIL LANCZOS shrink above 1024 pixels, ``argmax``, ``cv2.resize``); dilation, blur, unsharp and select are ONE kernel
        (``nesr_b200_masked_unsharp_u8``), bit-exact with the reference's cv2 calls.  ndarray in -> ndarray out, CUDA tensor in ->
        CUDA tensor out; any failure returns the image unchanged, as the reference does."""
        if "segmentation" not in self.models or "segmentation_extractor" not in self.models:
            return image
        try:
            mask = _object_mask(self.models, self.device, image.cpu().numpy() if isinstance(image, torch.Tensor) else image)
            if isinstance(image, torch.Tensor):
                return self._engine().masked_unsharp_u8(image.contiguous(), torch.from_numpy(mask).to(image.device))
            return self._engine().masked_unsharp_u8(np.ascontiguousarray(image), mask)
        except Exception as exc:                                     # noqa: BLE001 -- reference nesr/nesr.py:749-751
            logger.warning("Segmentation enhancement failed: %s", exc)
            return image

    def _use_tiling(self, h, w):
        """The reference's tiling policy on CUDA (``nesr/nesr.py:761-790``): tile when tiling is enabled and the image is larger
        than ``cuda_megapixel_threshold`` MP (default 8), and always above 16 MP.  ``always_tile`` is this implementation's opt-in
        to tile every image (tiles are independent forwards: with halo < receptive field, tiling changes border pixels)."""
        mp = (h * w) / (1024 * 1024)
        use = bool(self.config["enable_tiling"]) and (bool(self.config.get("always_tile")) or
                                                      mp > self.config.get("cuda_megapixel_threshold", 8))
        return use or mp > 16

    def _apply_esrgan(self, image):
        """RGB HWC u8 (ndarray or CUDA tensor) -> RGB HWC u8 at x2, same container kind."""
        if not self.config["use_esrgan"] or "esrgan" not in self.models:
            return None
        h, w = image.shape[:2]
        if self.config["head_compat"]:
            if self._use_tiling(h, w):
                return self._process_with_tiling(self._apply_esrgan_head, image, tile_size=int(self.config["max_tile_size"]), padding=16)
            return self._apply_esrgan_head(image)
        self.models["esrgan"].tile_size = int(self.config["max_tile_size"]) if self._use_tiling(h, w) else 0
        if isinstance(image, torch.Tensor):
            out, _ = self.models["esrgan"].enhance(image.flip(-1).contiguous())
            return out.flip(-1).contiguous()
        out, _ = self.models["esrgan"].enhance(cv2.cvtColor(image, cv2.COLOR_RGB2BGR))
        return cv2.cvtColor(out, cv2.COLOR_BGR2RGB)

    def _apply_esrgan_head(self, image):
        """The reference HEAD's ESRGAN stage (``_apply_esrgan_12channel`` / ``_apply_esrgan_3channel``, ``nesr/nesr.py:845-986``),
        untiled, in ONE library call (``nesr_b200_enhance_head_u8``): BGR / 255 -> 12 channels (image, x1.1, x0.9 clamped, 3x3 Gaussian
        blur -- or four copies with ``force_3channel``), built inside the input pack kernel -> ``model(x12)`` -> ``clip(out * 255, 0, 255)``
        TRUNCATED to u8 in conv_last's epilogue -> RGB.  x4 per call (the 12-channel network is the x2plus network without its
        un-shuffle).  Images above ``cuda_megapixel_threshold`` go through ``_process_with_tiling`` below, tile by tile.
        (``head_reference_glue`` below is the same stage written with torch expressions, kept as the test's second opinion.)"""
        src = image.contiguous() if isinstance(image, torch.Tensor) else np.ascontiguousarray(image)
        return self.models["esrgan"].model.engine(self.device).enhance_head_u8(src, force_3channel=bool(self.config["force_3channel"]))

    def head_reference_glue(self, image):
        """``_apply_esrgan_head`` as round 1 ran it: the 12-channel tensor and the u8 quantisation as torch expressions around
        ``model(x12)`` -- the statement-by-statement mirror of ``nesr/nesr.py:859-899``.  Test cross-check only."""
        host = not isinstance(image, torch.Tensor)
        rgb = torch.from_numpy(np.ascontiguousarray(image)).to(self.device) if host else image
        bgr = rgb.flip(-1).permute(2, 0, 1).contiguous()                                     # 3 x H x W u8
        t = bgr.float() / 255.0
        if self.config["force_3channel"]:
            x12 = torch.cat([t, t, t, t], 0)
        else:
            x12 = torch.cat([t, torch.clamp(t * 1.1, 0, 1), torch.clamp(t * 0.9, 0, 1), gaussian_blur3_u8(bgr).float() / 255.0], 0)
        out = self.models["esrgan"].model(x12.unsqueeze(0)).squeeze(0)                       # 3 x 4H x 4W float
        out = torch.clamp(out.permute(1, 2, 0) * 255.0, 0, 255).to(torch.uint8)              # truncation, as astype(np.uint8)
        out = out.flip(-1).contiguous()
        return out.cpu().numpy() if host else out

    def _process_with_tiling(self, processor_func, image, tile_size=512, padding=10):
        """The reference HEAD's own tiler (``nesr/nesr.py:311-475``), used by ``head_compat`` above the megapixel threshold.
        Tiles of ``tile_size`` extended by ``padding`` (clamped at the image edge) are processed one by one; the un-padded
        interior of each result is pasted at ``upscale_factor`` times its input position.  The processor's real scale (x4 in HEAD)
        need not equal ``upscale_factor``: the interior is cut with the tile's measured scale, ``int()``-truncated, and brought to
        the destination size with LANCZOS4 -- exactly the reference's arithmetic, including its probe forward on the top-left
        ``min(256, tile_size)`` corner (whose result it discards) and its per-tile bicubic fallback."""
        host = not isinstance(image, torch.Tensor)
        img = image if host else image.cpu().numpy()
        h, w, c = img.shape
        if h <= tile_size and w <= tile_size:
            return processor_func(image)
        f = self.config["upscale_factor"]
        out_h, out_w = int(h * f), int(w * f)
        output = np.zeros((out_h, out_w, c), np.uint8)
        probe = min(256, tile_size)
        try:
            processor_func(np.ascontiguousarray(img[:probe, :probe]))
            works = True
        except Exception as exc:                                     # noqa: BLE001 -- the reference degrades to bicubic tiles
            logger.warning("Tile processor test failed: %s", exc)
            works = False

        def bicubic(tile):
            return cv2.resize(tile, (int(tile.shape[1] * f), int(tile.shape[0] * f)), interpolation=cv2.INTER_CUBIC)

        for i in range(-(-h // tile_size)):
            for j in range(-(-w // tile_size)):
                y0, y1 = max(0, i * tile_size - padding), min(h, (i + 1) * tile_size + padding)
                x0, x1 = max(0, j * tile_size - padding), min(w, (j + 1) * tile_size + padding)
                tile = np.ascontiguousarray(img[y0:y1, x0:x1])
                try:
                    done = processor_func(tile) if works else bicubic(tile)
                    oy0, oy1, ox0, ox1 = int(y0 * f), int(y1 * f), int(x0 * f), int(x1 * f)
                    if padding > 0:
                        pad_up = int(padding * f)
                        oy0 += pad_up if y0 > 0 else 0
                        oy1 -= pad_up if y1 < h else 0
                        ox0 += pad_up if x0 > 0 else 0
                        ox1 -= pad_up if x1 < w else 0
                    th, tw = done.shape[:2]
                    sy, sx = th / tile.shape[0], tw / tile.shape[1]
                    ty0 = 0 if y0 == 0 else int(padding * sy)
                    ty1 = th if y1 == h else int(th - padding * sy)
                    tx0 = 0 if x0 == 0 else int(padding * sx)
                    tx1 = tw if x1 == w else int(tw - padding * sx)
                    ty0 = max(0, min(ty0, th - 1)); ty1 = max(ty0 + 1, min(ty1, th))
                    tx0 = max(0, min(tx0, tw - 1)); tx1 = max(tx0 + 1, min(tx1, tw))
                    if oy1 - oy0 <= 0 or ox1 - ox0 <= 0:
                        continue
                    region = done[ty0:ty1, tx0:tx1]
                    if region.shape[0] != oy1 - oy0 or region.shape[1] != ox1 - ox0:
                        region = cv2.resize(region, (ox1 - ox0, oy1 - oy0), interpolation=cv2.INTER_LANCZOS4)
                    output[oy0:oy1, ox0:ox1] = region
                except Exception as exc:                             # noqa: BLE001 -- reference: bicubic for this tile only
                    logger.warning("Error processing tile (%d,%d): %s", i, j, exc)
                    oy0, oy1 = int(i * tile_size * f), min(int(h * f), int((i + 1) * tile_size * f))
                    ox0, ox1 = int(j * tile_size * f), min(int(w * f), int((j + 1) * tile_size * f))
                    if oy1 > oy0 and ox1 > ox0:
                        output[oy0:oy1, ox0:ox1] = cv2.resize(bicubic(tile), (ox1 - ox0, oy1 - oy0), interpolation=cv2.INTER_CUBIC)
        return output if host else torch.from_numpy(output).to(image.device)

    def _ensemble_results(self, upscaled_images):
        if len(upscaled_images) == 1:
            return upscaled_images[0]
        first = upscaled_images[0]
        as_tensor = isinstance(first, torch.Tensor)
        shapes = [tuple(i.shape[:2]) for i in upscaled_images]
        target_h, target_w = max(shapes)
        members = []
        for img in upscaled_images:
            if tuple(img.shape[:2]) != (target_h, target_w):        # reference: LANCZOS4 align (host)
                host = img.cpu().numpy() if isinstance(img, torch.Tensor) else img
                img = cv2.resize(host, (target_w, target_h), interpolation=cv2.INTER_LANCZOS4)
            if as_tensor and not isinstance(img, torch.Tensor):
                img = torch.from_numpy(np.ascontiguousarray(img)).to(first.device)
            if not as_tensor and isinstance(img, torch.Tensor):
                img = img.cpu().numpy()
            members.append(img.contiguous() if as_tensor else np.ascontiguousarray(img))
        return self._engine().blend_u8(members)

    def _postprocess_image(self, image):
        if not self.config["adaptive_sharpening"]:
            return image
        img = image.contiguous() if isinstance(image, torch.Tensor) else np.ascontiguousarray(image)
        return self._engine().sharpen_u8(img, bgr=False)

    # -- file output (reference nesr/nesr.py:618-625, 644-647: cv2.imwrite inline) ------------------
    def _save_image(self, path, rgb):
        """Encode + write ``rgb`` (H x W x 3 u8 ndarray, not modified afterwards) to ``path``.  With ``async_io`` the PNG / JPEG
        encode -- by now the slowest step of an iteration -- runs on a worker thread while the GPU works on the next
        iteration; ``_flush_io`` joins the writes before ``enhance_image`` returns.  The file bytes do not depend on the mode."""
        def write():
            if not cv2.imwrite(path, cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR)):
                raise IOError(f"could not write {path}")
        if not self.config.get("async_io", True):
            write()
            return
        if self._io_pool is None:
            self._io_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="nesr-io")   # one writer: files appear in order
        self._io_pending.append(self._io_pool.submit(write))

    def _flush_io(self):
        pending, self._io_pending = self._io_pending, []
        for fut in pending:
            fut.result()                                           # re-raises a failed write

    # -- the loop (reference nesr/nesr.py:477-659) ------------------------------------------------
    def _progress(self, stage, it, msg):
        cb = self.config.get("progress_callback")
        if cb:
            cb(stage, it, self.config["iterations"], msg)

    def enhance_image(self, image_path, prompt=None):
        self._load_models()
        image = self._load_image(image_path)
        original_h, original_w = image.shape[:2]
        current = image
        self._progress("Starting enhancement", 0, f"Image size: {original_w}x{original_h}")
        n_iter = self.config["iterations"]
        for iteration in range(n_iter):
            t0 = time.time()
            self._progress("Enhancement", iteration, f"Starting iteration {iteration + 1}/{n_iter}")
            self._progress("Preprocessing", iteration, "Applying denoising and contrast enhancement")
            dev_in = self._preprocess_image(torch.from_numpy(np.ascontiguousarray(current)).to(self.device))
            if self.config["segment_enhancement"] and "segmentation" in self.models:
                self._progress("Segmentation", iteration, "Performing region-based analysis and enhancement")
                dev_in = self._segment_and_enhance(dev_in)
            upscaled = []
            if self.config["use_esrgan"] and "esrgan" in self.models:
                self._progress("ESRGAN", iteration, "Applying Real-ESRGAN upscaling")
                res = self._apply_esrgan(dev_in)
                if res is not None:
                    upscaled.append(res)
            extra = self.config.get("ensemble_members")
            if extra and upscaled:
                for m in extra(dev_in.cpu().numpy(), upscaled[0]):
                    upscaled.append(m if isinstance(m, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(m)).to(self.device))
            self._progress("Ensemble", iteration, "Combining results from multiple models")
            if upscaled:
                dev = self._ensemble_results(upscaled)
            else:
                logger.warning("All models failed, falling back to bicubic upscaling")
                current = dev_in.cpu().numpy()
                h, w = current.shape[:2]
                f = self.config["upscale_factor"]
                dev = torch.from_numpy(cv2.resize(current, (int(w * f), int(h * f)), interpolation=cv2.INTER_CUBIC)).to(self.device)
            self._progress("Postprocessing", iteration, "Applying final enhancements")
            dev = self._postprocess_image(dev)
            current = dev.cpu().numpy()
            if self.config["intermediate_saves"]:
                self._save_image(os.path.join(self.config["output_dir"], f"intermediate_iter{iteration + 1}.png"), current)
            if self.config.get("image_callback"):
                self.config["image_callback"](current)
            logger.info(f"Completed iteration {iteration + 1} in {time.time() - t0:.1f}s")
        final_h, final_w = current.shape[:2]
        scale_achieved = round(final_h / original_h, 1)
        base_name, ext = os.path.splitext(os.path.basename(image_path))
        final_path = os.path.join(self.config["output_dir"], f"{base_name}_enhanced_x{scale_achieved}{ext}")
        self._save_image(final_path, current)
        self._flush_io()                                           # the returned path exists, complete, when we return (a failed write raises here)
        self._progress("Complete", n_iter, f"Enhancement complete: {original_w}x{original_h} → {final_w}x{final_h} (x{scale_achieved})")
        return final_path


# ------------------------------------------------------------------------------------------------
# patching the unmodified reference
# ------------------------------------------------------------------------------------------------

def install_shims() -> None:
    """Register this package as ``basicsr.archs.rrdbnet_arch`` / ``realesrgan`` so reference code that
    does ``from basicsr.archs.rrdbnet_arch import RRDBNet; from realesrgan import RealESRGANer``
    (``nesr/nesr.py:161-162``, ``standalone/*.py``) gets the B200 classes."""
    import importlib.machinery
    import sys
    import types

    def mod(name, pkg):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=pkg)
        if pkg:
            m.__path__ = []
        sys.modules[name] = m
        return m

    basicsr, archs, arch = mod("basicsr", True), mod("basicsr.archs", True), mod("basicsr.archs.rrdbnet_arch", False)
    arch.RRDBNet = RRDBNet
    basicsr.archs, archs.rrdbnet_arch = archs, arch
    mod("realesrgan", True).RealESRGANer = RealESRGANer


def install(reference_cls, engine_getter=None) -> None:
    """Swap the GPU stages into the reference's ``SuperResolutionPipeline`` class in place.

    ``_ensemble_results`` / ``_postprocess_image`` keep their numpy-in / numpy-out contract; the
    ESRGAN stage becomes ``self.models['esrgan'].enhance(bgr)`` -- the call of the reference's
    previous revision and of ``standalone/superres_project.py:282``.
    """
    shared = {}

    def _engine(self):
        up = getattr(self, "models", {}).get("esrgan")
        if up is not None and isinstance(getattr(up, "model", None), RRDBNet):
            return up.model.engine(up.device)
        if engine_getter is not None:
            return engine_getter()
        if "eng" not in shared:
            shared["eng"] = _ffi.Engine(device=torch.cuda.current_device())
        return shared["eng"]

    def _apply_esrgan(self, image):
        if not self.config["use_esrgan"] or "esrgan" not in self.models:
            return None
        out, _ = self.models["esrgan"].enhance(cv2.cvtColor(image, cv2.COLOR_RGB2BGR))
        return cv2.cvtColor(out, cv2.COLOR_BGR2RGB)

    def _ensemble_results(self, upscaled_images):
        if len(upscaled_images) == 1:
            return upscaled_images[0]
        target_h, target_w = max([(img.shape[0], img.shape[1]) for img in upscaled_images])
        aligned = [np.ascontiguousarray(img if img.shape[:2] == (target_h, target_w)
                                        else cv2.resize(img, (target_w, target_h), interpolation=cv2.INTER_LANCZOS4))
                   for img in upscaled_images]
        return _engine(self).blend_u8(aligned)

    def _postprocess_image(self, image):
        if not self.config["adaptive_sharpening"]:
            return image
        return _engine(self).sharpen_u8(np.ascontiguousarray(image), bgr=False)

    def _preprocess_image(self, image):
        return _engine(self).preprocess_u8(np.ascontiguousarray(image), denoise_level=float(self.config["denoise_level"]))

    def _segment_and_enhance(self, image):
        if "segmentation" not in self.models or "segmentation_extractor" not in self.models:
            return image
        try:
            return _engine(self).masked_unsharp_u8(np.ascontiguousarray(image), _object_mask(self.models, self.device, image))
        except Exception as exc:                                     # noqa: BLE001 -- reference nesr/nesr.py:749-751
            logger.warning("Segmentation enhancement failed: %s", exc)
            return image

    reference_cls._preprocess_image = _preprocess_image
    reference_cls._segment_and_enhance = _segment_and_enhance
    reference_cls._apply_esrgan = _apply_esrgan
    reference_cls._ensemble_results = _ensemble_results
    reference_cls._postprocess_image = _postprocess_image

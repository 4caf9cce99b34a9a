"""ctypes binding of ``libnesr_b200.so`` (C ABI: ``include/nesr_b200.h``).

This is the whole Python<->CUDA boundary: plain pointers and sizes, no torch types.  torch is used
only by callers for device memory hand-off (``tensor.data_ptr()``).  There is no CPU fallback: a
missing library or a missing CUDA device raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

# NESR_B200_LIB: another build of the same library (the timing-experiment variant libnesr_b200_prof.so, tools/ only)
_LIB_PATH = os.environ.get("NESR_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnesr_b200.so")
ABI_VERSION = 1
FMT_BF16, FMT_FP16 = 0, 1
PTR_IN_DEVICE, PTR_OUT_DEVICE = 1, 2

EXPORTS = (
    "nesr_b200_default_config", "nesr_b200_create", "nesr_b200_destroy", "nesr_b200_last_error",
    "nesr_b200_load_weight", "nesr_b200_finalize_weights", "nesr_b200_enhance_u8",
    "nesr_b200_enhance_batch_u8", "nesr_b200_tile_count", "nesr_b200_debug_plan", "nesr_b200_enhance_tiles_u8",
    "nesr_b200_forward_nchw_f32", "nesr_b200_forward_feat_f32", "nesr_b200_blend_u8", "nesr_b200_sharpen_u8", "nesr_b200_masked_unsharp_u8", "nesr_b200_get_stats",
    "nesr_b200_synchronize", "nesr_b200_debug_conv", "nesr_b200_preprocess_u8", "nesr_b200_debug_lab_table",
    "nesr_b200_debug_nlm_weights", "nesr_b200_forward_nchw12_f32", "nesr_b200_enhance_tiles_packed_u8",
    "nesr_b200_unpack_tiles_u8", "nesr_b200_enhance_tile_list_packed_u8", "nesr_b200_enhance_head_u8", "nesr_b200_unpack_tile_list_u8",
)


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("num_in_ch", C.c_int32),
                ("num_out_ch", C.c_int32), ("scale", C.c_int32), ("num_feat", C.c_int32),
                ("num_block", C.c_int32), ("num_grow_ch", C.c_int32), ("body_format", C.c_int32),
                ("edge_format", C.c_int32), ("conv_impl", C.c_int32), ("feat_in_ch", C.c_int32),
                ("max_batch_pixels", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("conv_launches", C.c_int64), ("tiles_processed", C.c_int64),
                ("last_device_ms", C.c_double), ("last_conv_ms", C.c_double), ("arena_bytes", C.c_int64),
                ("last_trunk_ms", C.c_double), ("last_trunk_launches", C.c_int64)]


_lib = None
_lib_lock = threading.Lock()


def library_path() -> str:
    return _LIB_PATH


def load_library() -> C.CDLL:
    """Load the shared library and declare every prototype.  Raises if it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} not found: build it with `python -m neural_enhanced_super_resolution_b200._build` "
                "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
        lib = C.CDLL(_LIB_PATH)
        H = C.c_void_p
        u8p, f32p, i64p = C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)
        lib.nesr_b200_default_config.argtypes = [C.POINTER(Config), C.c_int32]
        lib.nesr_b200_default_config.restype = None
        lib.nesr_b200_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
        lib.nesr_b200_destroy.argtypes = [H]
        lib.nesr_b200_last_error.argtypes = [H]
        lib.nesr_b200_last_error.restype = C.c_char_p
        lib.nesr_b200_load_weight.argtypes = [H, C.c_char_p, f32p, i64p, C.c_int32]
        lib.nesr_b200_finalize_weights.argtypes = [H]
        lib.nesr_b200_enhance_u8.argtypes = [H, u8p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                             C.c_int32, u8p, C.c_int64, C.c_int32]
        lib.nesr_b200_enhance_batch_u8.argtypes = [H, u8p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                                   C.c_int32, C.c_int32, C.c_int32, u8p, C.c_int64, C.c_int64,
                                                   C.c_int32]
        lib.nesr_b200_tile_count.argtypes = [C.c_int32] * 5
        lib.nesr_b200_debug_plan.argtypes = [C.c_int32] * 8 + [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
        lib.nesr_b200_enhance_tiles_u8.argtypes = [H, u8p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                                   C.c_int32, C.c_int32, C.c_int32, u8p, C.c_int64, C.c_int32]
        lib.nesr_b200_enhance_tiles_packed_u8.argtypes = [H, u8p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                                          C.c_int32, C.c_int32, u8p, C.c_int32, C.c_int32, C.c_int32]
        lib.nesr_b200_enhance_tile_list_packed_u8.argtypes = [H, u8p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                                              C.POINTER(C.c_int32), C.c_int32, u8p, C.c_int32, C.c_int32, C.c_int32]
        lib.nesr_b200_unpack_tile_list_u8.argtypes = [H, u8p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                      C.POINTER(C.c_int32), C.c_int32, u8p, C.c_int64]
        lib.nesr_b200_enhance_head_u8.argtypes = [H, u8p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, u8p, C.c_int64, C.c_int32]
        lib.nesr_b200_unpack_tiles_u8.argtypes = [H, u8p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                  C.c_int32, u8p, C.c_int64]
        lib.nesr_b200_forward_nchw_f32.argtypes = [H, f32p, C.c_int32, C.c_int32, C.c_int32, f32p, C.c_void_p]
        lib.nesr_b200_forward_nchw12_f32.argtypes = [H, f32p, C.c_int32, C.c_int32, C.c_int32, f32p, C.c_void_p]
        lib.nesr_b200_forward_feat_f32.argtypes = [H, f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, f32p, C.c_void_p]
        lib.nesr_b200_blend_u8.argtypes = [H, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32,
                                           C.POINTER(C.c_double), u8p, C.c_int32]
        lib.nesr_b200_sharpen_u8.argtypes = [H, u8p, C.c_int32, C.c_int32, C.c_int32, u8p, C.c_int32]
        lib.nesr_b200_masked_unsharp_u8.argtypes = [H, u8p, u8p, C.c_int32, C.c_int32, C.c_int32, u8p, C.c_int32]
        lib.nesr_b200_preprocess_u8.argtypes = [H, u8p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32,
                                                C.c_int32, u8p, C.c_int32]
        lib.nesr_b200_debug_lab_table.argtypes = [C.c_int32, C.c_void_p, C.c_int32]
        lib.nesr_b200_debug_nlm_weights.argtypes = [C.c_float, C.c_int32, C.c_void_p, C.c_int32]
        lib.nesr_b200_get_stats.argtypes = [H, C.POINTER(Stats)]
        lib.nesr_b200_synchronize.argtypes = [H]
        lib.nesr_b200_debug_conv.argtypes = [H, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             f32p, f32p, f32p, C.c_int32, f32p]
        for name in EXPORTS:
            fn = getattr(lib, name)
            if name not in ("nesr_b200_default_config", "nesr_b200_last_error"):
                fn.restype = C.c_int
        _lib = lib
        return lib


def lab_table(which: int) -> np.ndarray:
    """The library's committed 8-bit Lab table ``which`` (0 sRGB gamma, 1 cube root, 2 L->(y, fy), 3 inverse sRGB gamma)."""
    lib = load_library()
    nbytes = lib.nesr_b200_debug_lab_table(which, None, 0)
    if nbytes < 0:
        raise ValueError("no such table")
    buf = np.empty(nbytes, np.uint8)
    lib.nesr_b200_debug_lab_table(which, buf.ctypes.data, nbytes)
    return buf if which == 3 else buf.view(np.uint16)


def nlm_weights(h: float, channels: int) -> np.ndarray:
    """The library's NLM weight table (non-zero prefix) for strength ``h`` and 1 or 2 channels."""
    lib = load_library()
    n = lib.nesr_b200_debug_nlm_weights(float(h), channels, None, 0)
    if n < 0:
        raise ValueError("bad arguments")
    buf = np.empty(n, np.int32)
    lib.nesr_b200_debug_nlm_weights(float(h), channels, buf.ctypes.data, n)
    return buf


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def _image_ptr(img, device=None):
    """(address, is_device) of a contiguous u8 image held in numpy or in a torch CUDA tensor.  ``device``: the engine's
    device ordinal -- a tensor on another GPU is an error, not a peer access."""
    if _is_torch_cuda(img):
        if not img.is_contiguous() or str(img.dtype) != "torch.uint8":
            raise ValueError("device images must be contiguous torch.uint8")
        if device is not None and img.device.index != device:
            raise ValueError(f"image is on cuda:{img.device.index} but this engine runs on cuda:{device}")
        import torch
        torch.cuda.current_stream(img.device).synchronize()   # the library works on its own stream
        return img.data_ptr(), True
    arr = img
    if not isinstance(arr, np.ndarray) or arr.dtype != np.uint8 or not arr.flags["C_CONTIGUOUS"]:
        raise ValueError("host images must be C-contiguous numpy uint8 arrays")
    return arr.ctypes.data, False


class Engine:
    """One ``nesr_b200_handle``: weights + arenas + stream on one GPU.  Not thread-safe."""

    def __init__(self, device: int = 0, num_block: int = 23, body_format: int = FMT_BF16,
                 edge_format: int = FMT_FP16, conv_impl: int = 0, max_batch_pixels: int = 0,
                 num_in_ch: int = 3, num_out_ch: int = 3, scale: int = 2, num_feat: int = 64,
                 num_grow_ch: int = 32, feat_in_ch: int = 0):
        self._lib = load_library()
        cfg = Config()
        self._lib.nesr_b200_default_config(C.byref(cfg), int(device))
        cfg.num_block, cfg.body_format, cfg.edge_format = int(num_block), int(body_format), int(edge_format)
        cfg.conv_impl, cfg.max_batch_pixels = int(conv_impl), int(max_batch_pixels)
        cfg.num_in_ch, cfg.num_out_ch, cfg.scale = int(num_in_ch), int(num_out_ch), int(scale)
        cfg.num_feat, cfg.num_grow_ch = int(num_feat), int(num_grow_ch)
        cfg.feat_in_ch = int(feat_in_ch)
        self.config = cfg
        self.device = int(device)
        self.scale = int(scale)
        self._h = C.c_void_p()
        rc = self._lib.nesr_b200_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self._lib.nesr_b200_last_error(None).decode()
            self._h = C.c_void_p()
            raise RuntimeError(f"nesr_b200_create failed ({rc}): {msg}")

    # -- plumbing ----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.nesr_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self._lib.nesr_b200_last_error(self._h).decode()}")

    # -- weights -----------------------------------------------------------------------------
    def load_state_dict(self, state_dict) -> None:
        """strict load of a checkpoint ``state_dict`` (torch tensors or numpy arrays, fp32 OIHW)."""
        for name, t in state_dict.items():
            a = t.detach().cpu().float().contiguous().numpy() if hasattr(t, "detach") else np.ascontiguousarray(t, np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            self._check(self._lib.nesr_b200_load_weight(self._h, name.encode(), a.ctypes.data, shape, a.ndim),
                        f"load_weight({name})")
        self._check(self._lib.nesr_b200_finalize_weights(self._h), "finalize_weights")

    # -- RealESRGANer.enhance ----------------------------------------------------------------
    def _alloc_like(self, img, shape):
        if _is_torch_cuda(img):
            import torch
            return torch.empty(shape, dtype=torch.uint8, device=img.device)
        return np.empty(shape, dtype=np.uint8)

    def enhance_u8(self, img, tile: int = 0, tile_pad: int = 10, pre_pad: int = 0, out=None):
        h, w = img.shape[:2]
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("enhance_u8 expects H x W x 3")
        if out is None:
            out = self._alloc_like(img, (h * self.scale, w * self.scale, 3))
        ip, idev = _image_ptr(img, self.device)
        op, odev = _image_ptr(out, self.device)
        flags = (PTR_IN_DEVICE if idev else 0) | (PTR_OUT_DEVICE if odev else 0)
        self._check(self._lib.nesr_b200_enhance_u8(self._h, ip, h, w, w * 3, tile, tile_pad, pre_pad, op,
                                                   w * self.scale * 3, flags), "enhance_u8")
        return out

    def enhance_batch_u8(self, frames, tile: int = 0, tile_pad: int = 10, pre_pad: int = 0, out=None):
        n, h, w = frames.shape[:3]
        if out is None:
            out = self._alloc_like(frames, (n, h * self.scale, w * self.scale, 3))
        ip, idev = _image_ptr(frames, self.device)
        op, odev = _image_ptr(out, self.device)
        flags = (PTR_IN_DEVICE if idev else 0) | (PTR_OUT_DEVICE if odev else 0)
        s = self.scale
        self._check(self._lib.nesr_b200_enhance_batch_u8(self._h, ip, n, h, w, w * 3, h * w * 3, tile, tile_pad,
                                                         pre_pad, op, w * s * 3, h * s * w * s * 3, flags),
                    "enhance_batch_u8")
        return out

    def enhance_head_u8(self, rgb, force_3channel: bool = False, out=None):
        """The reference HEAD's ESRGAN stage (``nesr/nesr.py:845-986``): RGB H x W x 3 u8 -> RGB 4H x 4W x 3 u8 (12-channel input built in
        the pack kernel, truncating u8 quantisation).  Same container kind out as in."""
        h, w = rgb.shape[:2]
        if rgb.ndim != 3 or rgb.shape[2] != 3:
            raise ValueError("enhance_head_u8 expects H x W x 3")
        if out is None:
            out = self._alloc_like(rgb, (4 * h, 4 * w, 3))
        ip, idev = _image_ptr(rgb, self.device)
        op, odev = _image_ptr(out, self.device)
        flags = (PTR_IN_DEVICE if idev else 0) | (PTR_OUT_DEVICE if odev else 0)
        self._check(self._lib.nesr_b200_enhance_head_u8(self._h, ip, h, w, w * 3, int(bool(force_3channel)), op, 4 * w * 3, flags),
                    "enhance_head_u8")
        return out

    def tile_count(self, h: int, w: int, tile: int, pre_pad: int = 0) -> int:
        n = self._lib.nesr_b200_tile_count(h, w, tile, pre_pad, self.scale)
        if n < 0:
            raise ValueError("bad tile_count arguments")
        return n

    def enhance_tiles_u8(self, img, out, tile: int, tile_pad: int, pre_pad: int, first: int, count: int):
        """Process tiles [first, first+count) of the tile grid into ``out`` (full-size output)."""
        h, w = img.shape[:2]
        ip, idev = _image_ptr(img, self.device)
        op, odev = _image_ptr(out, self.device)
        flags = (PTR_IN_DEVICE if idev else 0) | (PTR_OUT_DEVICE if odev else 0)
        self._check(self._lib.nesr_b200_enhance_tiles_u8(self._h, ip, h, w, w * 3, tile, tile_pad, pre_pad, first,
                                                         count, op, w * self.scale * 3, flags), "enhance_tiles_u8")
        return out

    def tile_costs(self, h: int, w: int, tile: int, tile_pad: int, pre_pad: int = 0):
        """Padded feature pixels of every tile of upstream's row-major tile_process grid (the unit of conv work)."""
        s = self.scale
        mod = 2 if s == 2 else 1
        hp, wp = -(-(h + pre_pad) // mod) * mod, -(-(w + pre_pad) // mod) * mod
        if tile <= 0:
            return [(hp // 2) * (wp // 2)]
        costs = []
        for y0 in range(0, hp, tile):
            for x0 in range(0, wp, tile):
                th = min(min(y0 + tile, hp) + tile_pad, hp) - max(y0 - tile_pad, 0)
                tw = min(min(x0 + tile, wp) + tile_pad, wp) - max(x0 - tile_pad, 0)
                costs.append((th // 2) * (tw // 2))
        return costs

    def slot_shape(self, h: int, w: int, tile: int):
        """(slot_h, slot_w) of the tile-major exchange buffer: the largest output rectangle a tile pastes."""
        s = self.scale
        return ((min(tile, h) if tile > 0 else h) * s, (min(tile, w) if tile > 0 else w) * s)

    def enhance_tiles_packed_u8(self, img, slots, tile: int, tile_pad: int, pre_pad: int, first: int, count: int):
        """Tiles [first, first+count) into ``slots`` (CUDA uint8 [count, slot_h, slot_w, 3]), tile k at slot k's origin."""
        h, w = img.shape[:2]
        ip, idev = _image_ptr(img, self.device)
        sp, sdev = _image_ptr(slots, self.device)
        if not sdev or slots.shape[0] < count:
            raise ValueError("slots must be a CUDA uint8 tensor with one slot per tile of the range")
        flags = (PTR_IN_DEVICE if idev else 0) | PTR_OUT_DEVICE
        self._check(self._lib.nesr_b200_enhance_tiles_packed_u8(self._h, ip, h, w, w * 3, tile, tile_pad, pre_pad, first, count, sp,
                                                                slots.shape[2], slots.shape[1], flags), "enhance_tiles_packed_u8")
        return slots

    def enhance_tile_list_packed_u8(self, img, slots, tile: int, tile_pad: int, pre_pad: int, tile_ids):
        """Tiles ``tile_ids`` (any subset of the grid) into ``slots`` (CUDA uint8 [>= len, slot_h, slot_w, 3]), tile ``tile_ids[k]`` at slot k."""
        h, w = img.shape[:2]
        ip, idev = _image_ptr(img, self.device)
        sp, sdev = _image_ptr(slots, self.device)
        n = len(tile_ids)
        if not sdev or slots.shape[0] < n:
            raise ValueError("slots must be a CUDA uint8 tensor with one slot per listed tile")
        ids = (C.c_int32 * n)(*[int(t) for t in tile_ids])
        flags = (PTR_IN_DEVICE if idev else 0) | PTR_OUT_DEVICE
        self._check(self._lib.nesr_b200_enhance_tile_list_packed_u8(self._h, ip, h, w, w * 3, tile, tile_pad, pre_pad, ids, n, sp,
                                                                    slots.shape[2], slots.shape[1], flags), "enhance_tile_list_packed_u8")
        return slots

    def unpack_tiles_u8(self, slots, out, h: int, w: int, tile: int, pre_pad: int, first: int, count: int):
        """Paste the slots of tiles [first, first+count) into the full 2h x 2w frame ``out`` (both CUDA uint8)."""
        sp, sdev = _image_ptr(slots, self.device)
        op, odev = _image_ptr(out, self.device)
        if not (sdev and odev):
            raise ValueError("unpack_tiles_u8 works on CUDA tensors")
        self._check(self._lib.nesr_b200_unpack_tiles_u8(self._h, sp, slots.shape[2], slots.shape[1], h, w, tile, pre_pad, first, count,
                                                        op, out.shape[1] * 3), "unpack_tiles_u8")
        return out

    def unpack_tile_list_u8(self, slots, out, h: int, w: int, tile: int, pre_pad: int, tile_ids):
        """Paste ``slots`` (CUDA uint8 [len(tile_ids), slot_h, slot_w, 3]; slot k holds tile ``tile_ids[k]``, < 0: empty) into ``out``."""
        sp, sdev = _image_ptr(slots, self.device)
        op, odev = _image_ptr(out, self.device)
        n = len(tile_ids)
        if not (sdev and odev) or slots.shape[0] < n:
            raise ValueError("unpack_tile_list_u8 works on CUDA tensors with one slot per list entry")
        ids = (C.c_int32 * n)(*[int(t) for t in tile_ids])
        self._check(self._lib.nesr_b200_unpack_tile_list_u8(self._h, sp, slots.shape[2], slots.shape[1], h, w, tile, pre_pad, ids, n, op,
                                                            out.shape[1] * 3), "unpack_tile_list_u8")
        return out

    # -- RRDBNet.forward ---------------------------------------------------------------------
    def forward_nchw(self, x):
        import torch
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
            raise ValueError("forward_nchw expects a CUDA float32 NCHW tensor")
        x = x.contiguous()
        n, c, h, w = x.shape
        if c != self.config.num_in_ch:
            raise RuntimeError(f"expected {self.config.num_in_ch} input channels, got {c}")
        y = torch.empty((n, self.config.num_out_ch, h * self.scale, w * self.scale), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        self._check(self._lib.nesr_b200_forward_nchw_f32(self._h, x.data_ptr(), n, h, w, y.data_ptr(),
                                                         C.c_void_p(stream)), "forward_nchw_f32")
        return y

    def forward_nchw12(self, x12):
        """The reference HEAD's ``model(x12)``: CUDA float32 [n, 12, H, W] -> [n, num_out_ch, 4H, 4W]."""
        import torch
        if not (x12.is_cuda and x12.dtype == torch.float32 and x12.dim() == 4 and x12.shape[1] == 12):
            raise ValueError("forward_nchw12 expects a CUDA float32 N x 12 x H x W tensor")
        x12 = x12.contiguous()
        n, _, h, w = x12.shape
        y = torch.empty((n, self.config.num_out_ch, 4 * h, 4 * w), dtype=torch.float32, device=x12.device)
        stream = torch.cuda.current_stream(x12.device).cuda_stream
        self._check(self._lib.nesr_b200_forward_nchw12_f32(self._h, x12.data_ptr(), n, h, w, y.data_ptr(),
                                                           C.c_void_p(stream)), "forward_nchw12_f32")
        return y

    def forward_feat(self, feat):
        """conv_first .. conv_last on a tensor that is already on the feature grid (scale-4 nets: x itself; scale-1 nets:
        ``pixel_unshuffle(x, 4)``): CUDA float32 [n, feat_in_ch, h, w] -> [n, num_out_ch, 4h, 4w]."""
        import torch
        if not (feat.is_cuda and feat.dtype == torch.float32 and feat.dim() == 4):
            raise ValueError("forward_feat expects a CUDA float32 N x C x H x W tensor")
        feat = feat.contiguous()
        n, c, h, w = feat.shape
        y = torch.empty((n, self.config.num_out_ch, 4 * h, 4 * w), dtype=torch.float32, device=feat.device)
        stream = torch.cuda.current_stream(feat.device).cuda_stream
        self._check(self._lib.nesr_b200_forward_feat_f32(self._h, feat.data_ptr(), n, c, h, w, y.data_ptr(), C.c_void_p(stream)),
                    "forward_feat_f32")
        return y

    # -- post-process ------------------------------------------------------------------------
    def blend_u8(self, members, weights=None, out=None):
        k = len(members)
        h, w = members[0].shape[:2]
        ptrs, dev = [], None
        for m in members:
            if tuple(m.shape) != (h, w, 3):
                raise ValueError("blend members must share one H x W x 3 shape")
            p, d = _image_ptr(m, self.device)
            if dev is not None and d != dev:
                raise ValueError("blend members must all be host or all be device images")
            dev = d
            ptrs.append(p)
        if out is None:
            out = self._alloc_like(members[0], (h, w, 3))
        op, odev = _image_ptr(out, self.device)
        arr = (C.c_void_p * k)(*ptrs)
        wts = None if weights is None else (C.c_double * k)(*[float(v) for v in weights])
        flags = (PTR_IN_DEVICE if dev else 0) | (PTR_OUT_DEVICE if odev else 0)
        self._check(self._lib.nesr_b200_blend_u8(self._h, arr, k, h, w, wts, op, flags), "blend_u8")
        return out

    def sharpen_u8(self, img, bgr: bool = False, out=None):
        h, w = img.shape[:2]
        if out is None:
            out = self._alloc_like(img, (h, w, 3))
        ip, idev = _image_ptr(img, self.device)
        op, odev = _image_ptr(out, self.device)
        flags = (PTR_IN_DEVICE if idev else 0) | (PTR_OUT_DEVICE if odev else 0)
        self._check(self._lib.nesr_b200_sharpen_u8(self._h, ip, h, w, int(bool(bgr)), op, flags), "sharpen_u8")
        return out

    def masked_unsharp_u8(self, img, object_mask, bgr: bool = False, out=None):
        """The unsharp stage of ``SuperResolutionPipeline._segment_and_enhance`` (``nesr/nesr.py:728-747``): ``object_mask`` (H x W u8, on
        the same side as ``img``) is dilated 3 x 3 and the sigma-3 unsharp replaces the pixels where the dilated mask is 1."""
        h, w = img.shape[:2]
        if img.ndim != 3 or img.shape[2] != 3 or tuple(object_mask.shape) != (h, w):
            raise ValueError("masked_unsharp_u8 expects an H x W x 3 image and an H x W mask")
        if out is None:
            out = self._alloc_like(img, (h, w, 3))
        ip, idev = _image_ptr(img, self.device)
        mp, mdev = _image_ptr(object_mask, self.device)
        if mdev != idev:
            raise ValueError("image and mask must both be host arrays or both be device tensors")
        op, odev = _image_ptr(out, self.device)
        flags = (PTR_IN_DEVICE if idev else 0) | (PTR_OUT_DEVICE if odev else 0)
        self._check(self._lib.nesr_b200_masked_unsharp_u8(self._h, ip, mp, h, w, int(bool(bgr)), op, flags), "masked_unsharp_u8")
        return out

    def preprocess_u8(self, img, denoise_level: float = 0.5, clip: float = 2.0, tiles=(8, 8), out=None):
        """``SuperResolutionPipeline._preprocess_image`` (``nesr/nesr.py:668-689``): RGB H x W x 3 u8 -> same, bit-exact with
        cv2 4.13.  ``denoise_level`` is the reference's config value (h = hColor = 10 * level; 0 skips the denoiser)."""
        h, w = img.shape[:2]
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("preprocess_u8 expects H x W x 3")
        if out is None:
            out = self._alloc_like(img, (h, w, 3))
        ip, idev = _image_ptr(img, self.device)
        op, odev = _image_ptr(out, self.device)
        flags = (PTR_IN_DEVICE if idev else 0) | (PTR_OUT_DEVICE if odev else 0)
        strength = float(denoise_level) * 10
        self._check(self._lib.nesr_b200_preprocess_u8(self._h, ip, h, w, strength, strength, float(clip), int(tiles[0]),
                                                      int(tiles[1]), op, flags), "preprocess_u8")
        return out

    # -- misc --------------------------------------------------------------------------------
    def stats(self) -> dict:
        s = Stats()
        self._check(self._lib.nesr_b200_get_stats(self._h, C.byref(s)), "get_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def synchronize(self) -> None:
        self._check(self._lib.nesr_b200_synchronize(self._h), "synchronize")

    def debug_conv(self, x, weight, bias, lrelu: bool = False, impl: int = 0, fmt: int = FMT_BF16):
        """One 3x3 conv on host fp32 arrays (x: C x H x W, weight: O x C x 3 x 3) -> O x H x W fp32."""
        x = np.ascontiguousarray(x, np.float32)
        weight = np.ascontiguousarray(weight, np.float32)
        bias = np.ascontiguousarray(bias, np.float32)
        cin, h, w = x.shape
        cout = weight.shape[0]
        y = np.empty((cout, h, w), np.float32)
        self._check(self._lib.nesr_b200_debug_conv(self._h, impl, fmt, h, w, cin, cout, weight.ctypes.data,
                                                   bias.ctypes.data, x.ctypes.data, int(bool(lrelu)), y.ctypes.data),
                    "debug_conv")
        return y

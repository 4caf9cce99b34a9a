"""Achieved HBM bandwidth of the bandwidth-bound kernels (blend, sharpen) on device-resident images, CUDA-event timed
inside the engine (stats()["last_device_ms"]).  Algorithmic bytes: blend 3*(K+1) B per pixel, sharpen 6 B per pixel.

  python tools/stencil_bench.py            (results of r1: profiles/r1_stencil_bandwidth.txt)
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_enhanced_super_resolution_b200 as pkg  # noqa: E402

peak = 6544.3
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
torch.manual_seed(0)
eng = pkg.RRDBNet(3, 3, scale=2, num_block=1).cuda().eval().engine()
rng = np.random.default_rng(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (H, W) in ((2160, 3840), (4320, 7680)):
    imgs = [torch.from_numpy(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).cuda() for _ in range(3)]
    out = torch.empty_like(imgs[0])
    for name, fn, bytes_px in (("blend K=2", lambda: eng.blend_u8(imgs[:2], out=out), 9), ("blend K=3", lambda: eng.blend_u8(imgs, out=out), 12),
                               ("sharpen", lambda: eng.sharpen_u8(imgs[0], out=out), 6)):
        ms = []
        for it in range(8):
            flush.fill_(it)                      # evict L2 between iterations
            fn()
            ms.append(eng.stats()["last_device_ms"])
        t = float(np.median(ms[3:]))
        gbs = bytes_px * H * W / (t * 1e-3) / 1e9
        print(f"{name:10s} {W}x{H}: {t:.3f} ms  {H * W / t / 1e3:.0f} Mpix/s  algorithmic {gbs:.0f} GB/s = {100 * gbs / peak:.1f} % of the measured HBM peak ({peak:.0f} GB/s)")

# the sharpen kernel skips the colour blurs of 16 x 8 blocks without a selected pixel: a photo-like image (smooth structure + mild noise) beside the noise above
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_common import natural_image  # noqa: E402
H, W = 2160, 3840
import cv2  # noqa: E402
base = natural_image(H, W, seed=1)
for name, host in (("photo-like", base), ("smooth (the photo-like image blurred, sigma 3)", cv2.GaussianBlur(base, (0, 0), 3))):
    photo = torch.from_numpy(host).cuda()
    out = torch.empty_like(photo)
    ms = []
    for it in range(8):
        flush.fill_(it)
        eng.sharpen_u8(photo, out=out)
        ms.append(eng.stats()["last_device_ms"])
    t = float(np.median(ms[3:]))
    sel = float((out != photo).any(dim=2).float().mean())
    print(f"sharpen    {W}x{H} {name} ({100 * sel:.1f} % of the pixels changed): {t:.3f} ms  algorithmic {6 * H * W / (t * 1e-3) / 1e9:.0f} GB/s = "
          f"{100 * 6 * H * W / (t * 1e-3) / 1e9 / peak:.1f} % of the measured HBM peak")

"""Device time of the pre-process stage (nesr_b200_preprocess_u8) and the host cv2 time beside it.

  python tools/preprocess_bench.py [H W level steps]
"""
import os
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_enhanced_super_resolution_b200 import _ffi  # noqa: E402

H, W, level, steps = ([float(a) for a in sys.argv[1:5]] + [1080, 1920, 0.5, 5][len(sys.argv) - 1:])[:4]
H, W, steps = int(H), int(W), int(steps)
rng = np.random.default_rng(0)
yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
img = np.clip(np.stack([120 + 70 * np.sin(xx / (9 + c)) * np.cos(yy / (7 + c)) for c in range(3)], -1) + rng.normal(0, 8, (H, W, 3)), 0, 255).astype(np.uint8)
eng = _ffi.Engine(device=0, num_block=1)
dev = torch.from_numpy(img).cuda()
out = torch.empty_like(dev)
for _ in range(2):
    eng.preprocess_u8(dev, denoise_level=level, out=out)
ms = []
for _ in range(steps):
    eng.preprocess_u8(dev, denoise_level=level, out=out)
    ms.append(eng.stats()["last_device_ms"])
t0 = time.time()
if level > 0:
    ref = cv2.fastNlMeansDenoisingColored(img, None, level * 10, level * 10, 7, 21)
else:
    ref = img
lab = cv2.cvtColor(ref, cv2.COLOR_RGB2LAB)
l, a, b = cv2.split(lab)
ref = cv2.cvtColor(cv2.merge((cv2.createCLAHE(2.0, (8, 8)).apply(l), a, b)), cv2.COLOR_LAB2RGB)
cpu_ms = (time.time() - t0) * 1e3
print(f"{W}x{H} level {level}: GPU {np.median(ms):.3f} ms ({H * W / np.median(ms) / 1e3:.1f} Mpix/s), host cv2 {cpu_ms:.0f} ms on {os.cpu_count()} cores, "
      f"bit-exact {np.array_equal(out.cpu().numpy(), ref)}")

"""Device-resident timing of one enhance call (tools for kernel work; bench.py is the contract).

  python tools/quick_bench.py [H W tile tile_pad steps]   (NESR_B200_DEBUG_FLAGS selects timing experiments)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_enhanced_super_resolution_b200 as pkg  # noqa: E402

H, W, tile, pad, steps = ([int(a) for a in sys.argv[1:6]] + [1080, 1920, 512, 10, 5][len(sys.argv) - 1:])[:5]
torch.manual_seed(0)
net = pkg.RRDBNet(3, 3, scale=2, num_block=int(os.environ.get("NESR_NUM_BLOCK", "23")), conv_impl=int(os.environ.get("NESR_CONV_IMPL", "0")), max_batch_pixels=int(os.environ.get("NESR_MAX_BATCH_PX", "0"))).cuda().eval()
eng = net.engine()
rng = np.random.default_rng(0)
img = torch.from_numpy(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).cuda()
out = torch.empty((2 * H, 2 * W, 3), dtype=torch.uint8, device="cuda")
for _ in range(int(os.environ.get('NESR_WARMUP', '3'))):
    eng.enhance_u8(img, tile=tile, tile_pad=pad, out=out)
ms, cms = [], []
for _ in range(steps):
    eng.enhance_u8(img, tile=tile, tile_pad=pad, out=out)
    st = eng.stats()
    ms.append(st["last_device_ms"]); cms.append(st["last_conv_ms"])
mpix = 4 * H * W / 1e6
print(f"flags={os.environ.get('NESR_B200_DEBUG_FLAGS', '0')} impl={os.environ.get('NESR_CONV_IMPL', '0')} "
      f"{W}x{H} tile={tile}: device {np.median(ms):.2f} ms  conv {np.median(cms):.2f} ms  "
      f"{mpix / np.median(ms) * 1e3:.1f} Mpix/s  util(burst 1636.7)={2241504 * mpix * 1e6 / (np.median(cms) * 1e-3) / 1636.7e12:.3f}")

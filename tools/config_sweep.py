"""Device-timed throughput of the other BASELINE.json configs on ONE GPU (bench.py stays on C2, the config the metric is quoted on).

  python tools/config_sweep.py [steps]      -> one JSON line per config on stdout

C1  512x512 -> 1024x1024, untiled                         (enhance_u8)
C3  3840x2160 -> 7680x4320, tile 512 / halo 10            (enhance_u8, 40 tiles, several L2-resident tile groups)
C4  32 frames of 512x512 (one GPU's share of 256 frames)  (enhance_batch_u8)
C5  256 -> 512 -> 1024 -> 2048: ESRGAN + 2-member blend + adaptive sharpen per iteration (device-resident stages), without and
    with the NLM + CLAHE pre-process in front of every iteration

Inputs are resident in HBM; times are CUDA events around the calls (median of `steps`), after 3 warm-up calls.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_enhanced_super_resolution_b200 as pkg  # noqa: E402

FLOP_PER_OUT_PX = 2241504
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.manual_seed(0)
net = pkg.RRDBNet(3, 3, scale=2, max_batch_pixels=int(os.environ.get("NESR_MAX_BATCH_PX", "0"))).cuda().eval()
eng = net.engine()
rng = np.random.default_rng(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def report(name, out_px, ms, **extra):
    line = {"config": name, "ms": round(ms, 3), "out_mpix_per_s": round(out_px / ms / 1e3, 1),
            "conv_tflops_algorithmic": round(out_px * FLOP_PER_OUT_PX / ms / 1e9, 1), **extra}
    print(json.dumps(line), flush=True)


def u8(*shape):
    return torch.from_numpy(rng.integers(0, 256, shape, dtype=np.uint8)).cuda()


# C1
img = u8(512, 512, 3)
out = torch.empty((1024, 1024, 3), dtype=torch.uint8, device="cuda")
report("C1 512x512->1024x1024 untiled", 1024 * 1024, timed(lambda: eng.enhance_u8(img, tile=0, out=out)))

# C3
img = u8(2160, 3840, 3)
out = torch.empty((4320, 7680, 3), dtype=torch.uint8, device="cuda")
ms = timed(lambda: eng.enhance_u8(img, tile=512, tile_pad=10, out=out))
report("C3 3840x2160->7680x4320 tile 512 halo 10", 4320 * 7680, ms, trunk_launches=eng.stats().get("last_trunk_launches"))

# C4 (one GPU's 32 of 256 frames)
frames = u8(32, 512, 512, 3)
out = torch.empty((32, 1024, 1024, 3), dtype=torch.uint8, device="cuda")
ms = timed(lambda: eng.enhance_batch_u8(frames, tile=0, out=out))
report("C4 32 x (512x512->1024x1024) batched", 32 * 1024 * 1024, ms, trunk_launches=eng.stats().get("last_trunk_launches"))

# C5: three iterations, each ESRGAN x2 + blend with a second member + sharpen, all on the device
cur0 = u8(256, 256, 3)


def c5(level=0.0):
    cur = cur0
    for _ in range(3):
        if level > 0:
            cur = eng.preprocess_u8(cur, denoise_level=level)
        up = eng.enhance_u8(cur, tile=0)
        other = up.flip(0).contiguous()                          # a second ensemble member of the same size
        ens = eng.blend_u8([up, other])
        cur = eng.sharpen_u8(ens)
    return cur


report("C5 256->512->1024->2048 (ESRGAN + blend K=2 + sharpen per iteration)", 512 * 512 + 1024 * 1024 + 2048 * 2048, timed(c5))
report("C5 with the pre-process stage (NLM h=5 + CLAHE) in front of every iteration, as enhance_image runs it",
       512 * 512 + 1024 * 1024 + 2048 * 2048, timed(lambda: c5(0.5)))

"""Bit-identity of two builds of the library on the same frame (one process per build): python tools/ab_bits.py <lib.so> <out.npy> [H W tile pad]"""
import os
import sys

import numpy as np
import torch

os.environ["NESR_B200_LIB"] = os.path.abspath(sys.argv[1])
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_enhanced_super_resolution_b200 as pkg  # noqa: E402

H, W, tile, pad = ([int(a) for a in sys.argv[3:7]] + [1080, 1920, 512, 10][len(sys.argv) - 3:])[:4]
torch.manual_seed(0)
net = pkg.RRDBNet(3, 3, scale=2).cuda().eval()
img = np.random.default_rng(0).integers(0, 256, (H, W, 3), dtype=np.uint8)
out = net.engine().enhance_u8(img, tile=tile, tile_pad=pad)
np.save(sys.argv[2], out)
print(sys.argv[1], out.shape, int(out.astype(np.int64).sum()))

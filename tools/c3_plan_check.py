"""4K -> 8K frame under the default plan (the larger-cap plan is taken when it needs fewer tile groups) and under an explicit
200k-pixel cap: launches, device time, and that the two outputs are bit-identical.  python tools/c3_plan_check.py"""
import sys, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import neural_enhanced_super_resolution_b200 as pkg
from oracle.rrdbnet import x2plus
sd = x2plus(seed=0).state_dict()
rng = np.random.default_rng(1)
img = torch.from_numpy(rng.integers(0, 256, (2160, 3840, 3), dtype=np.uint8)).cuda()
outs = []
for cap in (0, 200000):
    net = pkg.RRDBNet(3, 3, scale=2, max_batch_pixels=cap); net.load_state_dict(sd); net = net.cuda().eval()
    eng = net.engine()
    outs.append(eng.enhance_u8(img, tile=512, tile_pad=10)); print(cap, eng.stats()["last_trunk_launches"], eng.stats()["last_device_ms"])
print("4K default plan == cap 200k plan:", torch.equal(outs[0], outs[1]))

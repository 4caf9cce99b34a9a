"""BASELINE config 4 in full on one GPU: 256 frames of 512x512 -> 1024x1024 through nesr_b200_enhance_batch_u8 (device-resident),
checked frame by frame against single calls.  python tools/c4_full.py"""
import sys, time, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import neural_enhanced_super_resolution_b200 as pkg
torch.manual_seed(0)
net = pkg.RRDBNet(3, 3, scale=2).cuda().eval()
eng = net.engine()
rng = np.random.default_rng(0)
frames = torch.from_numpy(rng.integers(0, 256, (256, 512, 512, 3), dtype=np.uint8)).cuda()
out = torch.empty((256, 1024, 1024, 3), dtype=torch.uint8, device="cuda")
eng.enhance_batch_u8(frames, tile=0, out=out)
t0 = time.time(); eng.enhance_batch_u8(frames, tile=0, out=out); torch.cuda.synchronize(); dt = time.time() - t0
st = eng.stats()
ok = all(torch.equal(out[i], eng.enhance_u8(frames[i], tile=0)) for i in (0, 1, 2, 100, 254, 255))
print(f"C4 full: 256 frames in {dt*1e3:.1f} ms = {256*1024*1024/dt/1e6:.1f} Mpix/s, device {st['last_device_ms']:.1f} ms, trunk launches {st['last_trunk_launches']}, arena {st['arena_bytes']/2**30:.1f} GiB, frames match single calls: {ok}")

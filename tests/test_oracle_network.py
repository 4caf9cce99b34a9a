"""Self-checks and regression pins of the network half of the oracle (``oracle.rrdbnet``,
``oracle.realesrganer``).  The reference ships no tests or vectors for this path and its
``basicsr``/``realesrgan`` dependencies are not installable here, so parity is UNPINNED by the
reference; these tests pin structure (published checkpoint layout), index semantics (against
torch's own ops) and a committed regression output."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import shims
from oracle.realesrganer import RealESRGANer, tile_grid
from oracle.rrdbnet import RRDBNet, identity_state_dict, pixel_unshuffle, x2plus


def _digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module")
def net0():
    return x2plus(seed=0)


@pytest.fixture(scope="module")
def ckpt0(net0, tmp_path_factory):
    return shims.write_checkpoint(net0.state_dict(), str(tmp_path_factory.mktemp("ckpt")))


def test_structure_matches_published_x2plus_checkpoint(net0):
    sd = net0.state_dict()
    assert len(sd) == 702
    assert sum(v.numel() for v in sd.values()) == 16_703_171
    assert sum(1 for k in sd if k.endswith(".weight")) == 351
    assert tuple(sd["conv_first.weight"].shape) == (64, 12, 3, 3)
    assert tuple(sd["body.22.rdb3.conv5.weight"].shape) == (64, 192, 3, 3)
    assert tuple(sd["body.0.rdb1.conv2.weight"].shape) == (32, 96, 3, 3)
    assert tuple(sd["conv_last.weight"].shape) == (3, 64, 3, 3)
    expect = {"conv_first", "conv_body", "conv_up1", "conv_up2", "conv_hr", "conv_last"}
    expect |= {f"body.{i}.rdb{j}.conv{k}" for i in range(23) for j in (1, 2, 3) for k in range(1, 6)}
    assert {k.rsplit(".", 1)[0] for k in sd} == expect


def test_constructor_variants_used_by_reference():
    # nesr/nesr.py:216 -- 12 input channels, default scale 4: no un-shuffle, x4 output
    head = RRDBNet(num_in_ch=12, num_out_ch=3, num_feat=64, num_block=1, num_grow_ch=32).eval()
    with torch.no_grad():
        assert tuple(head(torch.zeros(1, 12, 8, 6)).shape) == (1, 3, 32, 24)
    # x2plus: 3 channels, scale 2: x2 output, odd sizes rejected by the un-shuffle
    x2 = RRDBNet(num_in_ch=3, num_out_ch=3, scale=2, num_feat=64, num_block=1, num_grow_ch=32).eval()
    with torch.no_grad():
        assert tuple(x2(torch.zeros(1, 3, 8, 6)).shape) == (1, 3, 16, 12)
        with pytest.raises(AssertionError):
            x2(torch.zeros(1, 3, 7, 6))


def test_pixel_unshuffle_matches_torch_and_formula():
    x = torch.arange(2 * 3 * 6 * 8, dtype=torch.float32).reshape(2, 3, 6, 8)
    u = pixel_unshuffle(x, 2)
    assert torch.equal(u, torch.nn.functional.pixel_unshuffle(x, 2))
    for c in range(3):
        for i in range(2):
            for j in range(2):
                assert torch.equal(u[:, c * 4 + i * 2 + j], x[:, c, i::2, j::2])


def test_init_scheme(net0):
    sd = net0.state_dict()
    assert float(sd["body.3.rdb2.conv4.bias"].abs().max()) == 0.0
    w = sd["body.3.rdb2.conv4.weight"]
    fan_in = w.shape[1] * 9
    assert abs(float(w.std()) / (0.1 * (2.0 / fan_in) ** 0.5) - 1.0) < 0.05
    assert float(sd["conv_up1.bias"].abs().max()) > 0.0          # torch default init untouched


def test_tile_grid_baseline_configs():
    # BASELINE config 2: 1920x1080, tile 512, halo 10  
    g = tile_grid(1080, 1920, 512, 10)
    assert len(g) == 12
    assert [t.x1p - t.x0p for t in g[:4]] == [522, 532, 532, 394]
    assert [g[i * 4].y1p - g[i * 4].y0p for i in range(3)] == [522, 532, 66]
    # config 3: 3840x2160 -> 8 x 5 tiles
    g = tile_grid(2160, 3840, 512, 10)
    assert len(g) == 40
    assert [t.x1p - t.x0p for t in g[:8]] == [522] + [532] * 6 + [266]
    assert [g[i * 8].y1p - g[i * 8].y0p for i in range(5)] == [522, 532, 532, 532, 122]
    # interiors partition the image
    cover = np.zeros((2160, 3840), np.int32)
    for t in g:
        cover[t.y0:t.y1, t.x0:t.x1] += 1
    assert cover.min() == 1 and cover.max() == 1


def test_regression_golden(golden, ckpt0, net0):
    g = golden("esrgan_x2.npz")
    assert _digest(net0.state_dict()) == str(g["weights_sha256"]), "seeded init changed: regenerate goldens"
    full, mode = RealESRGANer(2, ckpt0, model=x2plus(None), tile=0, tile_pad=10, pre_pad=0).enhance(g["crop_bgr"])
    assert mode == "RGB" and full.dtype == np.uint8 and full.shape == (128, 160, 3)
    assert np.abs(full.astype(int) - g["full"]).max() <= 1          # fp32 summation order may differ by host
    assert (full != g["full"]).mean() < 1e-3
    tiled, _ = RealESRGANer(2, ckpt0, model=x2plus(None), tile=32, tile_pad=4, pre_pad=0).enhance(g["crop_bgr"])
    assert np.abs(tiled.astype(int) - g["tiled"]).max() <= 1
    assert (tiled != full).any()                                       # halo < receptive field: tiling changes pixels
    odd, _ = RealESRGANer(2, ckpt0, model=x2plus(None), tile=0, tile_pad=10, pre_pad=10).enhance(g["odd_bgr"])
    assert odd.shape == (74, 102, 3)
    assert np.abs(odd.astype(int) - g["odd"]).max() <= 1


def test_identity_weights_give_exact_index_map(tmp_path, photo_bgr):
    net = x2plus(None)
    path = shims.write_checkpoint(identity_state_dict(net), str(tmp_path))
    for tile in (0, 32):
        out, _ = RealESRGANer(2, path, model=x2plus(None), tile=tile, tile_pad=4, pre_pad=0).enhance(photo_bgr)
        h, w = photo_bgr.shape[:2]
        yy, xx = np.meshgrid(np.arange(2 * h), np.arange(2 * w), indexing="ij")
        assert np.array_equal(out, photo_bgr[2 * (yy // 4), 2 * (xx // 4)])


def test_enhance_outscale_and_modes(ckpt0):
    rng = np.random.default_rng(3)
    up = RealESRGANer(2, ckpt0, model=x2plus(None), tile=0, tile_pad=10, pre_pad=0)
    img = rng.integers(0, 256, (12, 10, 3), dtype=np.uint8)
    out, _ = up.enhance(img, outscale=3)
    assert out.shape == (36, 30, 3)
    out, mode = up.enhance(img[:, :, 0])
    assert mode == "L" and out.shape == (24, 20)
    out, mode = up.enhance(np.dstack([img, img[:, :, :1]]))
    assert mode == "RGBA" and out.shape == (24, 20, 4)


@pytest.mark.reference
def test_unmodified_reference_pipeline_runs_through_shims(golden, tmp_path):
    """The reference's ``enhance_image`` (nesr/nesr.py:477) with the oracle standing in for its
    un-installable dependencies reproduces the committed output (HEAD: 12-channel replicate, x4)."""
    import cv2
    g = golden("pipeline.npz")
    Pipeline = shims.import_reference()
    shims.install_shims()
    cwd = os.getcwd()
    try:
        os.chdir(tmp_path)
        torch.manual_seed(1)
        head = RRDBNet(num_in_ch=12, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32)
        assert _digest(head.state_dict()) == str(g["weights_sha256"])
        shims.write_checkpoint(head.state_dict(), str(tmp_path))
        src = str(tmp_path / "in.png")
        cv2.imwrite(src, cv2.cvtColor(g["small_rgb"], cv2.COLOR_RGB2BGR))
        pipe = Pipeline(device="cpu", config={"iterations": 1, "use_diffusion": False, "segment_enhancement": False,
                                              "denoise_level": 0, "output_dir": str(tmp_path / "out")})
        path = pipe.enhance_image(src)
        assert os.path.basename(path) == str(g["result_name"]) == "in_enhanced_x4.0.png"
        out = cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB)
        assert out.shape == (128, 160, 3)
        assert np.abs(out.astype(int) - g["head_out"]).max() <= 1
    finally:
        os.chdir(cwd)
        shims.remove_shims()

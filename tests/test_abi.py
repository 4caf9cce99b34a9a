"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, fails loudly without a GPU (no CPU fallback), and the Python mirrors present the
reference's class surface.  No compute runs here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import neural_enhanced_super_resolution_b200 as pkg
from neural_enhanced_super_resolution_b200 import _ffi
from oracle.rrdbnet import x2plus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = torch.cuda.is_available()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "nesr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nesr_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _ffi.load_library()
    declared = _declared_functions()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nesr_b200.h but not exported"
    assert sorted(_ffi.EXPORTS) == declared


def test_config_struct_matches_header_defaults():
    lib = _ffi.load_library()
    cfg = _ffi.Config()
    lib.nesr_b200_default_config(ctypes.byref(cfg), 0)
    assert (cfg.abi_version, cfg.num_in_ch, cfg.num_out_ch, cfg.scale) == (_ffi.ABI_VERSION, 3, 3, 2)
    assert (cfg.num_feat, cfg.num_block, cfg.num_grow_ch) == (64, 23, 32)
    assert (cfg.body_format, cfg.edge_format, cfg.conv_impl) == (_ffi.FMT_BF16, _ffi.FMT_FP16, 0)
    assert ctypes.sizeof(_ffi.Config) == 56 and ctypes.sizeof(_ffi.Stats) == 64


def test_tile_count_follows_upstream_grid():
    lib = _ffi.load_library()
    assert lib.nesr_b200_tile_count(1080, 1920, 512, 0, 2) == 12
    assert lib.nesr_b200_tile_count(2160, 3840, 512, 0, 2) == 40
    assert lib.nesr_b200_tile_count(512, 512, 0, 0, 2) == 1
    assert lib.nesr_b200_tile_count(513, 512, 512, 0, 2) == 2       # mod-pad to 514 rows -> 2 tile rows
    assert lib.nesr_b200_tile_count(512, 512, 512, 10, 2) == 4      # pre_pad grows the padded image


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
        _ffi.Engine(device=0)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.RRDBNet(3, 3, scale=2, num_block=1)(torch.zeros(1, 3, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.SuperResolutionPipeline(device="cpu", config={"output_dir": "/tmp/nesr_b200_test_out"})


def test_rrdbnet_mirror_has_the_checkpoint_layout():
    ours = pkg.RRDBNet(num_in_ch=3, num_out_ch=3, scale=2, num_feat=64, num_block=23, num_grow_ch=32)
    ref = x2plus(seed=None)
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert list(sd_o.keys()) == list(sd_r.keys())
    assert all(sd_o[k].shape == sd_r[k].shape for k in sd_r)
    ours.load_state_dict(sd_r, strict=True)
    ours.eval()
    assert next(ours.parameters()).device.type == "cpu"
    bad = dict(sd_r)
    bad.pop("conv_hr.bias")
    with pytest.raises(RuntimeError):
        ours.load_state_dict(bad, strict=True)


def test_realesrganer_mirror_signature(tmp_path):
    import inspect
    from oracle.realesrganer import RealESRGANer as OracleUp
    assert list(inspect.signature(pkg.RealESRGANer.__init__).parameters) == list(inspect.signature(OracleUp.__init__).parameters)
    assert list(inspect.signature(pkg.RealESRGANer.enhance).parameters) == list(inspect.signature(OracleUp.enhance).parameters)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.RealESRGANer(2, str(tmp_path / "none.pth"), model=pkg.RRDBNet(3, 3, scale=2, num_block=1), device="cpu")


def test_product_package_does_not_import_the_oracle():
    import subprocess
    import sys
    code = ("import sys; import neural_enhanced_super_resolution_b200; "
            "bad=[m for m in sys.modules if m=='oracle' or m.startswith('oracle.')]; print(bad); sys.exit(1 if bad else 0)")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_image_pointer_validation():
    with pytest.raises(ValueError):
        _ffi._image_ptr(np.zeros((4, 4, 3), np.float32))
    with pytest.raises(ValueError):
        _ffi._image_ptr(np.zeros((4, 8, 3), np.uint8)[:, ::2])
    addr, dev = _ffi._image_ptr(np.zeros((4, 4, 3), np.uint8))
    assert addr != 0 and dev is False


def test_lab_tables_match_oracle():
    """The Lab tables compiled into the library (csrc/lab_tables.inc) are the oracle's, which equal cv2's on all 2^24 colours."""
    from oracle import preprocess as P
    assert np.array_equal(_ffi.lab_table(0), P.gamma_table(True))
    assert np.array_equal(_ffi.lab_table(1), P.cbrt_table())
    assert np.array_equal(_ffi.lab_table(2), P.l_to_yf_table().reshape(-1))
    assert np.array_equal(_ffi.lab_table(3), P.inv_gamma_table())


def test_nlm_weight_tables_match_oracle():
    from oracle import preprocess as P
    for h in (1.0, 2.5, 3.0, 5.0, 7.0, 10.0):
        for channels in (1, 2):
            got = _ffi.nlm_weights(h, channels)
            want, _ = P.nlm_weight_table(h, channels)
            assert len(got) == np.count_nonzero(want) and np.array_equal(got, want[:len(got)])

"""``SuperResolutionPipeline.enhance_image`` on the GPU: API contract of the reference
(nesr/nesr.py:477-659: callbacks, naming, iteration chain) and exact composition of the stages."""
import os

import cv2
import numpy as np
import pytest

import neural_enhanced_super_resolution_b200 as pkg
from oracle import postprocess as O
from oracle import preprocess as P
from gpu_common import checkpoint, natural_image

pytestmark = pytest.mark.gpu


def _bicubic_member(rgb_in, rgb_esrgan):
    h, w = rgb_in.shape[:2]
    return [cv2.resize(rgb_in, (2 * w, 2 * h), interpolation=cv2.INTER_CUBIC)]


def test_enhance_image_chain_and_contract(tmp_path):
    rgb = natural_image(40, 48, seed=12)[:, :, ::-1].copy()
    src = str(tmp_path / "frame.png")
    cv2.imwrite(src, cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR))
    stages, images = [], []
    pipe = pkg.SuperResolutionPipeline(device="cuda", config={
        "iterations": 2, "use_diffusion": False, "segment_enhancement": False, "denoise_level": 0.5,
        "output_dir": str(tmp_path / "out"), "esrgan_model_path": checkpoint("calibrated"),
        "max_tile_size": 32, "tile_pad": 4, "ensemble_members": _bicubic_member, "intermediate_saves": True,
        "progress_callback": lambda stage, it, total, msg: stages.append(stage),
        "image_callback": lambda im: images.append(im.copy())})
    path = pipe.enhance_image(src)
    assert os.path.basename(path) == "frame_enhanced_x4.0.png"
    out = cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB)
    assert out.shape == (160, 192, 3) and len(images) == 2 and np.array_equal(images[-1], out)
    assert os.path.exists(tmp_path / "out" / "intermediate_iter1.png")
    for s in ("Starting enhancement", "Preprocessing", "ESRGAN", "Ensemble", "Postprocessing", "Complete"):
        assert s in stages
    # composition: iteration = preprocess -> esrgan -> blend(esrgan, bicubic) -> sharpen, stages exact
    cur = rgb
    for _ in range(2):
        cur = P.preprocess_image(cur, 0.5)
        e = pipe._apply_esrgan(cur)
        cur = O.postprocess_image(O.ensemble_results([e, _bicubic_member(cur, e)[0]]))
    assert np.array_equal(cur, out)


def test_install_patches_a_reference_like_class(tmp_path):
    class RefLike:                                   # the four hooks of nesr/nesr.py, numpy in/out
        def __init__(self):
            self.config = {"use_esrgan": True, "adaptive_sharpening": True, "denoise_level": 0.4}
            self.models = {"esrgan": pkg.RealESRGANer(2, checkpoint("calibrated"), model=pkg.RRDBNet(3, 3, scale=2),
                                                      tile=0, pre_pad=0, device="cuda")}
    pkg.install(RefLike)
    r = RefLike()
    rgb = natural_image(32, 40, seed=4)
    up = r._apply_esrgan(rgb)
    assert up.shape == (64, 80, 3)
    other = np.ascontiguousarray(up[::-1])
    assert np.array_equal(r._ensemble_results([up, other]), O.ensemble_results([up, other]))
    assert r._ensemble_results([up]) is up
    assert np.array_equal(r._postprocess_image(up), O.postprocess_image(up))
    assert np.array_equal(r._preprocess_image(rgb), P.preprocess_image(rgb, 0.4))

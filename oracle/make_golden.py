"""Generate ``tests/golden/*.npz`` by running the REFERENCE's own code in the authoring container.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Run from the repo root:

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

Needs ``/root/reference`` (absent on the GPU box, which only reads the committed fixtures).
What is produced by the reference itself (imported unmodified from ``/root/reference``):
  * ``postprocess.npz``  -- inputs and outputs of ``SuperResolutionPipeline._postprocess_image``
                            (``nesr/nesr.py:1056-1084``, cv2 4.13) incl. tiny / ragged sizes
  * ``ensemble.npz``     -- inputs and outputs of ``_ensemble_results`` (``nesr/nesr.py:1033-1054``), K = 2, 3, 4
  * ``preprocess.npz``   -- inputs and outputs of ``_preprocess_image`` (``nesr/nesr.py:668-689``: NLM denoise + LAB CLAHE,
                            cv2 4.13) at denoise levels 0 / 0.3 / 0.5 / 1.0 incl. ragged and tiny sizes
                            (``python -m oracle.make_golden preprocess`` regenerates only this file)
  * ``pipeline.npz``     -- the unmodified ``enhance_image`` run end to end (1 iteration, diffusion and
                            segmentation off) with the oracle registered as ``basicsr``/``realesrgan``:
                            HEAD behaviour (12-channel replicate, x4) -- pins shims + glue
What is produced by the oracle restatement (the reference cannot: its deps are not installable):
  * ``esrgan_x2.npz``    -- ``RealESRGANer.enhance`` on a natural crop, untiled and tiled, seeded weights
                            (a regression pin for the restatement; parity unpinned by the reference)
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile

import cv2
import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def state_dict_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def make_preprocess() -> None:
    """``preprocess.npz``: the reference's ``_preprocess_image`` run on crops of its own test image and on noise."""
    from oracle import shims
    Pipeline = shims.import_reference(REF)
    photo = cv2.cvtColor(cv2.imread(os.path.join(REF, "images", "test.jpeg")), cv2.COLOR_BGR2RGB)
    rng = np.random.default_rng(20261019)
    noisy = np.clip(photo[100:164, 300:396].astype(np.int32) + rng.integers(-25, 26, (64, 96, 3)), 0, 255).astype(np.uint8)
    cases = {
        "photo_h5": (np.ascontiguousarray(photo[200:248, 180:240]), 0.5),      # 48 x 60, the GUI default level
        "noisy_h10": (noisy, 1.0),                                              # 64 x 96 (divisible by the CLAHE grid)
        "ragged_h3": (np.ascontiguousarray(photo[300:337, 100:145]), 0.3),     # 37 x 45
        "noise_h5": (rng.integers(0, 256, (33, 29, 3), dtype=np.uint8), 0.5),
        "tiny_h5": (rng.integers(0, 256, (9, 11, 3), dtype=np.uint8), 0.5),
        "photo_h0": (cv2.resize(photo, (100, 76), interpolation=cv2.INTER_AREA), 0.0),   # CLAHE only
    }
    out = {}
    for name, (img, level) in cases.items():
        pipe = Pipeline(device="cpu", config={"denoise_level": level})
        out[name + "_in"] = img
        out[name + "_level"] = np.array(level)
        out[name + "_out"] = pipe._preprocess_image(img.copy())
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **out)
    print("preprocess.npz", os.path.getsize(os.path.join(OUT, "preprocess.npz")))


def main() -> None:
    sys.dont_write_bytecode = True
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["preprocess"]:
        make_preprocess()
        return
    make_preprocess()
    from oracle import shims
    from oracle.realesrganer import RealESRGANer
    from oracle.rrdbnet import RRDBNet, x2plus

    Pipeline = shims.import_reference(REF)

    class _Cfg:
        config = {"adaptive_sharpening": True}

    photo = cv2.cvtColor(cv2.imread(os.path.join(REF, "images", "test.jpeg")), cv2.COLOR_BGR2RGB)
    rng = np.random.default_rng(20261018)

    # ---- post-process -------------------------------------------------------------------
    pp_in = {
        "photo_crop": np.ascontiguousarray(photo[200:296, 180:308]),        # 96 x 128
        "photo_small": cv2.resize(photo, (100, 76), interpolation=cv2.INTER_AREA),
        "noise_ragged": rng.integers(0, 256, (67, 45, 3), dtype=np.uint8),
        "noise_tiny": rng.integers(0, 256, (5, 7, 3), dtype=np.uint8),
        "noise_row": rng.integers(0, 256, (1, 33, 3), dtype=np.uint8),
        "noise_col": rng.integers(0, 256, (29, 2, 3), dtype=np.uint8),
        "flat": np.full((24, 24, 3), 200, dtype=np.uint8),
    }
    pp = {}
    for name, img in pp_in.items():
        pp[name + "_in"] = img
        pp[name + "_out"] = Pipeline._postprocess_image(_Cfg(), img)
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **pp)

    # ---- ensemble -----------------------------------------------------------------------
    ens = {}
    for k in (2, 3, 4):
        members = [rng.integers(0, 256, (40, 56, 3), dtype=np.uint8) for _ in range(k)]
        ens[f"k{k}_in"] = np.stack(members)
        ens[f"k{k}_out"] = Pipeline._ensemble_results(_Cfg(), members)
    # every (a, b, c) triple of a coarse lattice: exercises the f64-product / f32-accumulate rounding
    lat = np.arange(0, 256, 5, dtype=np.uint8)
    a, b, c = np.meshgrid(lat, lat, lat, indexing="ij")
    tri = [np.repeat(v.reshape(52, -1, 1), 3, axis=2) for v in (a, b, c)]
    ens["lattice3_in"] = np.stack(tri)
    ens["lattice3_out"] = Pipeline._ensemble_results(_Cfg(), tri)
    np.savez_compressed(os.path.join(OUT, "ensemble.npz"), **ens)

    # ---- oracle network regression pin ----------------------------------------------------
    torch.set_num_threads(os.cpu_count() or 1)
    net = x2plus(seed=0)
    digest = state_dict_digest(net.state_dict())
    crop_bgr = np.ascontiguousarray(photo[200:264, 180:260, ::-1])          # 64 x 80 BGR
    odd_bgr = np.ascontiguousarray(photo[100:137, 60:111, ::-1])            # 37 x 51 (odd: mod-pad path)
    with tempfile.TemporaryDirectory() as td:
        ckpt = shims.write_checkpoint(net.state_dict(), td)
        full, _ = RealESRGANer(2, ckpt, model=x2plus(None), tile=0, tile_pad=10, pre_pad=0).enhance(crop_bgr)
        tiled, _ = RealESRGANer(2, ckpt, model=x2plus(None), tile=32, tile_pad=4, pre_pad=0).enhance(crop_bgr)
        odd, _ = RealESRGANer(2, ckpt, model=x2plus(None), tile=0, tile_pad=10, pre_pad=10).enhance(odd_bgr)

        # ---- unmodified reference pipeline through the shims (HEAD behaviour) --------------
        shims.install_shims()
        cwd = os.getcwd()
        try:
            os.chdir(td)                                       # reference searches ./models/weights
            torch.manual_seed(1)
            head_net = RRDBNet(num_in_ch=12, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32)
            shims.write_checkpoint(head_net.state_dict(), td)
            src = os.path.join(td, "in.png")
            small = np.ascontiguousarray(photo[220:252, 200:240])           # 32 x 40 RGB
            cv2.imwrite(src, cv2.cvtColor(small, cv2.COLOR_RGB2BGR))
            pipe = Pipeline(device="cpu", config={
                "iterations": 1, "use_diffusion": False, "segment_enhancement": False,
                "denoise_level": 0, "output_dir": os.path.join(td, "out")})
            result_path = pipe.enhance_image(src)
            head_out = cv2.cvtColor(cv2.imread(result_path), cv2.COLOR_BGR2RGB)
            head_digest = state_dict_digest(head_net.state_dict())
        finally:
            os.chdir(cwd)
            shims.remove_shims()

    np.savez_compressed(os.path.join(OUT, "esrgan_x2.npz"),
                        crop_bgr=crop_bgr, full=full, tiled=tiled, odd_bgr=odd_bgr, odd=odd,
                        weights_sha256=np.array(digest))
    np.savez_compressed(os.path.join(OUT, "pipeline.npz"),
                        small_rgb=small, head_out=head_out, weights_sha256=np.array(head_digest),
                        result_name=np.array(os.path.basename(result_path)))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

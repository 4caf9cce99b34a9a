"""Generate ``tests/golden/pipeline_tiled.npz``: the UNMODIFIED reference ``enhance_image`` (HEAD) on an image above its
tiling threshold, so that ``_apply_esrgan`` -> ``_process_with_tiling`` (``nesr/nesr.py:761-806, 311-475``) runs: padded
tiles, the x4 processor against ``upscale_factor`` 2, int-truncated coordinates and the LANCZOS4 resize of every interior.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Run from the repo root (needs ``/root/reference``):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_tiled

The oracle network stands in for ``basicsr``/``realesrgan`` through the shims, exactly as for ``pipeline.npz``
(``oracle/make_golden.py``); the other fixtures are not touched.
"""
from __future__ import annotations

import os
import tempfile

import cv2
import numpy as np
import torch

from oracle import shims
from oracle.make_golden import OUT, REF, state_dict_digest
from oracle.rrdbnet import RRDBNet

CONFIG = {"iterations": 1, "use_diffusion": False, "segment_enhancement": False, "denoise_level": 0,
          "max_tile_size": 24, "cpu_megapixel_threshold": 0.0005}          # 40 x 56 = 0.0021 MP -> 2 x 3 tiles, padding 16


def main() -> None:
    Pipeline = shims.import_reference(REF)
    photo = cv2.cvtColor(cv2.imread(os.path.join(REF, "images", "test.jpeg")), cv2.COLOR_BGR2RGB)
    torch.set_num_threads(os.cpu_count() or 1)
    with tempfile.TemporaryDirectory() as td:
        shims.install_shims()
        cwd = os.getcwd()
        try:
            os.chdir(td)                                                   # the reference searches ./models/weights
            torch.manual_seed(1)
            head_net = RRDBNet(num_in_ch=12, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32)
            shims.write_checkpoint(head_net.state_dict(), td)
            small = np.ascontiguousarray(photo[216:256, 196:252])          # 40 x 56 RGB
            src = os.path.join(td, "in.png")
            cv2.imwrite(src, cv2.cvtColor(small, cv2.COLOR_RGB2BGR))
            pipe = Pipeline(device="cpu", config={**CONFIG, "output_dir": os.path.join(td, "out")})
            path = pipe.enhance_image(src)
            out = cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB)
        finally:
            os.chdir(cwd)
            shims.remove_shims()
    np.savez_compressed(os.path.join(OUT, "pipeline_tiled.npz"), small_rgb=small, head_out=out,
                        weights_sha256=np.array(state_dict_digest(head_net.state_dict())), result_name=np.array(os.path.basename(path)),
                        max_tile_size=np.array(CONFIG["max_tile_size"]), megapixel_threshold=np.array(CONFIG["cpu_megapixel_threshold"]))
    print("pipeline_tiled.npz", out.shape, os.path.basename(path))


if __name__ == "__main__":
    main()

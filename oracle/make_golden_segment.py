"""Generate ``tests/golden/segment.npz``: the UNMODIFIED reference ``SuperResolutionPipeline._segment_and_enhance``
(``nesr/nesr.py:690-751``) run with a deterministic stand-in for the segmentation model (the reference loads SegFormer through
``transformers``; its weights are not in this container and the model is out of scope -- what is pinned here is everything the
method does AROUND it: ``argmax`` map -> ``> 0`` object mask -> ``cv2.resize`` to the image -> 3 x 3 dilation -> sigma-3 unsharp
where the mask is 1).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Run from the repo root (needs ``/root/reference``):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_segment

Per case the fixture holds the image, the stand-in's class map (so that the GPU box needs no model), the object mask at image
resolution exactly as the reference computes it (``cv2.resize((seg_map > 0).astype(np.uint8), (W, H))``, ``:729-730``) and the
reference method's output.
"""
from __future__ import annotations

import os

import cv2
import numpy as np
import torch

from oracle import shims
from oracle.make_golden import OUT, REF


class StandInExtractor:
    """``extractor(images=pil, return_tensors="pt").to(device).pixel_values``: the image at 64 x 64, CHW float in [0, 1]."""

    class _Batch:
        def __init__(self, pixel_values):
            self.pixel_values = pixel_values

        def to(self, device):
            return self

    def __call__(self, images, return_tensors="pt"):
        arr = np.asarray(images.convert("RGB").resize((64, 64)), dtype=np.float32) / 255.0
        return self._Batch(torch.from_numpy(arr).permute(2, 0, 1)[None])


class StandInSegmenter:
    """``model(pixel_values).logits``: three classes at a quarter of the input grid (as SegFormer's head) -- background, "bright"
    and "reddish" regions of the 4 x 4-averaged image.  Deterministic; any class > 0 is an object for the reference."""

    class _Out:
        def __init__(self, logits):
            self.logits = logits

    def __call__(self, pixel_values):
        p = torch.nn.functional.avg_pool2d(pixel_values, 4)
        lum = p.mean(dim=1, keepdim=True)
        return self._Out(torch.cat([torch.full_like(lum, 0.45), lum, p[:, 0:1] - p[:, 2:3] + 0.3], dim=1))


def class_map(image_rgb: np.ndarray) -> np.ndarray:
    """The stand-in's ``argmax`` map for an image <= 1024 pixels on its longer side (what ``:712-715`` hand to the rest)."""
    from PIL import Image
    inputs = StandInExtractor()(images=Image.fromarray(image_rgb))
    return StandInSegmenter()(inputs.pixel_values).logits.argmax(dim=1)[0].cpu().numpy()


def main() -> None:
    Pipeline = shims.import_reference(REF)

    class _Self:
        models = {"segmentation": StandInSegmenter(), "segmentation_extractor": StandInExtractor()}
        device = "cpu"

    photo = cv2.cvtColor(cv2.imread(os.path.join(REF, "images", "test.jpeg")), cv2.COLOR_BGR2RGB)
    rng = np.random.default_rng(20261019)
    cases = {
        "photo": np.ascontiguousarray(photo[120:312, 100:356]),                   # 192 x 256
        "photo_small": cv2.resize(photo, (90, 70), interpolation=cv2.INTER_AREA),
        "noise_ragged": rng.integers(0, 256, (67, 45, 3), dtype=np.uint8),
        "noise_tiny": rng.integers(0, 256, (5, 7, 3), dtype=np.uint8),
        "noise_row": rng.integers(0, 256, (1, 33, 3), dtype=np.uint8),
    }
    out = {}
    for name, img in cases.items():
        res = Pipeline._segment_and_enhance(_Self(), img)
        seg = class_map(img)
        mask = cv2.resize((seg > 0).astype(np.uint8), (img.shape[1], img.shape[0]))
        out[name + "_in"], out[name + "_seg"], out[name + "_mask"], out[name + "_out"] = img, seg.astype(np.int64), mask, res
        print(name, img.shape, "objects", float(mask.mean()), "changed", float((res != img).any(axis=2).mean()))
        assert res is not img, "the reference method took its failure path"
    np.savez_compressed(os.path.join(OUT, "segment.npz"), **out)


if __name__ == "__main__":
    main()

"""Host-side pieces of the pipeline mirror that need no GPU."""
import cv2
import numpy as np
import pytest
import torch

import neural_enhanced_super_resolution_b200 as pkg
from neural_enhanced_super_resolution_b200.pipeline import gaussian_blur3_u8
from oracle.rrdbnet import x2plus


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (9, 1), (2, 2), (5, 7), (40, 33)])
def test_gaussian_blur3_is_cv2(shape):
    """The blurred member of the reference HEAD's 12-channel input (``nesr/nesr.py:872-878``)."""
    img = np.random.default_rng(shape[0] * 31 + shape[1]).integers(0, 256, (*shape, 3), dtype=np.uint8)
    mine = gaussian_blur3_u8(torch.from_numpy(img).permute(2, 0, 1)).permute(1, 2, 0).numpy()
    assert np.array_equal(mine, cv2.GaussianBlur(img, (3, 3), 0))


def test_head_layout_holds_the_x2plus_weights():
    """``RRDBNet(num_in_ch=12, num_out_ch=3)`` (reference HEAD, ``nesr/nesr.py:216``) has the published x2plus checkpoint's tensors."""
    head = pkg.RRDBNet(num_in_ch=12, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32)
    head.load_state_dict(x2plus(seed=3).state_dict(), strict=True)
    assert head.scale == 4 and head.conv_first.weight.shape == (64, 12, 3, 3)
    with pytest.raises(RuntimeError):
        head(torch.zeros(1, 12, 8, 8))                                   # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        pkg.RRDBNet(num_in_ch=3, num_out_ch=3, scale=4).engine("cuda:0")  # a true x4 network is not this build

"""The tcgen05/TMEM/TMA conv kernel, one layer at a time, through the C ABI (nesr_b200_debug_conv):
against torch's fp32 conv2d on operands rounded to the kernel's 16-bit format, and against the SIMT
validation kernel that shares its buffers and epilogue."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from neural_enhanced_super_resolution_b200 import _ffi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    eng = _ffi.Engine(device=0, num_block=1)
    yield eng
    eng.close()


def _round(t, fmt):
    return t.to(torch.bfloat16 if fmt == _ffi.FMT_BF16 else torch.float16).float()


CASES = [  # cin, cout, H, W
    (16, 64, 9, 13),      # conv_first shape class (Cin padded to 16, single k-step)
    (12, 64, 8, 8),       # true conv_first channel count
    (64, 32, 20, 37),     # RDB conv1
    (96, 32, 11, 29),     # conv2: half-used second chunk
    (128, 32, 17, 16),
    (160, 32, 13, 21),
    (192, 64, 19, 23),    # conv5
    (64, 64, 40, 67),     # conv_body / up / hr, several M-blocks per row group
    (64, 3, 15, 33),      # conv_last (N padded to 16)
    (64, 32, 1, 1),       # degenerate
    (64, 32, 3, 300),     # wide: pitch > 128, three column strips
    (192, 64, 70, 140),   # conv5 as two folded passes, several bands per strip, ring wraps
    (64, 64, 45, 129),    # Cout 64 ring (8 slots) wraps many times; 1-pixel second strip
    (160, 32, 37, 128),   # exactly one full strip
]


FOLD, SIMT = 0, 1      # conv_impl: row-folded tcgen05 (product), SIMT validation kernel


@pytest.mark.parametrize("impl", [FOLD])
@pytest.mark.parametrize("fmt", [_ffi.FMT_BF16, _ffi.FMT_FP16])
@pytest.mark.parametrize("cin,cout,h,w", CASES)
def test_tc_conv_matches_torch(engine, cin, cout, h, w, fmt, impl):
    g = torch.Generator().manual_seed(cin * 1000 + cout * 10 + h)
    x = torch.randn(cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    want = F.conv2d(_round(x, fmt).unsqueeze(0).double(), _round(wt, fmt).double(), b.double(), padding=1)[0]
    want = F.leaky_relu(want, 0.2).float().numpy()
    got = engine.debug_conv(x.numpy(), wt.numpy(), b.numpy(), lrelu=True, impl=impl, fmt=fmt)
    assert got.shape == want.shape
    err = np.abs(got - want).max()
    assert err < 2e-4 * max(1.0, np.abs(want).max()), f"max err {err}"


@pytest.mark.parametrize("cin,cout,h,w", [(96, 32, 11, 29), (192, 64, 19, 23), (64, 3, 15, 33)])
def test_tc_conv_matches_simt_validation_kernel(engine, cin, cout, h, w):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(cin, h, w, generator=g).numpy()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).numpy()
    b = torch.randn(cout, generator=g).numpy()
    simt = engine.debug_conv(x, wt, b, impl=SIMT)
    assert np.abs(engine.debug_conv(x, wt, b, impl=FOLD) - simt).max() < 1e-4


def test_conv_is_linear_and_translation_consistent(engine):
    """Size-independent properties: conv(a x) = a conv(x) for power-of-two a (exact in fp), and an
    impulse reproduces the flipped kernel around its position (tap order / shift direction)."""
    cin, cout, h, w = 64, 32, 9, 11
    wt = np.zeros((cout, cin, 3, 3), np.float32)
    vals = np.arange(1, 10, dtype=np.float32).reshape(3, 3) / 16.0
    wt[5, 7] = vals
    x = np.zeros((cin, h, w), np.float32)
    x[7, 4, 6] = 1.0
    y = engine.debug_conv(x, wt, np.zeros(cout, np.float32))
    want = np.zeros((h, w), np.float32)
    for ky in range(3):
        for kx in range(3):
            want[4 - (ky - 1), 6 - (kx - 1)] = vals[ky, kx]      # y[p] = sum_k w[k] x[p + k - 1]
    assert np.array_equal(y[5], want)
    assert np.count_nonzero(np.delete(y, 5, axis=0)) == 0
    assert np.array_equal(engine.debug_conv(4 * x, wt, np.zeros(cout, np.float32)), 4 * y)

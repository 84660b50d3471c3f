"""GPU: the flat-shift halo-tile kernel with shared-memory-resident weights (csrc/conv_flat.cu) - the default path of
the 3x3 / stride-1 layers with 33..64 output channels at training batch sizes (fprop and dgrad) - against the numpy
oracle at the north-star TF32 tolerance, on shapes that exercise ragged last tiles, a non-square image whose padded
row is not a multiple of 8 pixels, bias, and fewer than 64 output channels."""
import ctypes

import numpy as np
import pytest

from gpu_util import assert_close, require_gpu
from oracle import tortto_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu():
    require_gpu()


CASES = [  # n, c, h, w, k, pad, bias   (the kernel is taken from 4 tiles of 128 padded pixels per SM: 592 tiles)
    (80, 64, 32, 32, 64, 1, False),   # the layer-1 shape of preact_resnet18 (reduced batch)
    (130, 64, 20, 26, 48, 1, True),   # Wp = 28, 48 output channels, bias, ragged last tile
    (90, 32, 30, 30, 64, 1, False),   # one 32-channel slab
    (240, 64, 17, 19, 40, 0, False),  # no padding (P = H - 2)
]


@pytest.mark.parametrize("n,c,h,w,k,pad,bias", CASES)
def test_flat_resident_fprop_dgrad_vs_oracle(n, c, h, w, k, pad, bias):
    import pytortto_b200 as tt
    from pytortto_b200 import _cabi, ops
    tt.set_math_mode("tf32")
    rng = np.random.default_rng(n + h)
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    wt = (rng.standard_normal((k, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32)
    b = rng.standard_normal(k).astype(np.float32) if bias else None
    d = ops.conv_desc(x.shape, wt.shape, (1, 1), (pad, pad), (1, 1), 1)
    lib = _cabi.load()
    assert lib.ttb_conv2d_kernel_variant(ctypes.byref(d), 0) == 2, "fprop of this shape must take the flat-shift kernel"
    # dgrad is the same correlation with the channel roles swapped: it reduces over the OUTPUT channels (whole 32-channel
    # slabs needed) and its "output channels" are the conv's input channels (33..64 for the resident variant)
    assert lib.ttb_conv2d_kernel_variant(ctypes.byref(d), 1) == (2 if (k % 32 == 0 and 33 <= c <= 64) else 1)
    yo = O.conv2d_forward(x, wt, b, 1, pad, 1)
    dy = rng.standard_normal(yo.shape).astype(np.float32)
    dxo, dwo, _ = O.conv2d_backward(x, wt, dy, 1, pad, 1)
    xin = tt.nn.Parameter(tt.tensor(x).cuda())
    wp = tt.nn.Parameter(tt.tensor(wt).cuda())
    bp = None if b is None else tt.nn.Parameter(tt.tensor(b).cuda())
    y = tt.nn.functional.conv2d(xin, wp, bp, (1, 1), (pad, pad), (1, 1), 1)
    y.backward(tt.tensor(dy).cuda())
    assert_close("flat y", y.data.get(), yo, 2e-3)
    assert_close("flat dx", xin.grad.get(), dxo, 2e-3)
    assert_close("flat dw", wp.grad.get(), dwo, 2e-3)


def test_flat_kernel_selection_rules():
    """small batches (fewer than 4 tiles per SM), strided / dilated / wide-output layers stay on the im2col kernel"""
    import pytortto_b200 as tt
    from pytortto_b200 import _cabi, ops
    tt.set_math_mode("tf32")
    lib = _cabi.load()

    def variant(n, c, h, k, ks=3, stride=1, pad=1, dil=1):
        d = ops.conv_desc((n, c, h, h), (k, c, ks, ks), (stride, stride), (pad, pad), (dil, dil), 1)
        return [lib.ttb_conv2d_kernel_variant(ctypes.byref(d), p) for p in (0, 1, 2)]
    assert variant(256, 64, 32, 64) == [2, 2, 1]
    assert variant(8, 64, 32, 64) == [1, 1, 1]
    assert variant(256, 64, 32, 128) == [1, 1, 1]
    assert variant(256, 64, 32, 64, stride=2) == [1, 1, 1]
    assert variant(256, 64, 32, 64, ks=1, pad=0) == [1, 1, 1]
    assert variant(256, 128, 16, 128) == [1, 1, 1]
    tt.set_math_mode("fp32")
    assert variant(256, 64, 32, 64) == [0, 0, 0]
    tt.set_math_mode("tf32")

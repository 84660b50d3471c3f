"""GPU: the flat-shift halo-tile kernel (csrc/conv_flat.cu) is an experiment that lives in the TUNING build only
(libtortto_b200_tuning.so, -DTTB_TUNING); the release library routes every problem to the im2col kernels.  This test
runs scripts/flat_check.py in a subprocess against the tuning library with the resident-weight variant switched on
(TTB_FLAT=-1) and with every eligible problem on the flat kernels (TTB_FLAT=1): fprop and dgrad against the exact fp32
kernels at the north-star TF32 tolerance (2e-3), and checks that the release library never selects it."""
import ctypes
import os
import subprocess
import sys

import pytest

from gpu_util import require_gpu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TUNING = os.path.join(ROOT, "pytortto_b200", "libtortto_b200_tuning.so")


@pytest.fixture(autouse=True)
def _gpu():
    require_gpu()


@pytest.mark.parametrize("mode", ["-1", "1"])
def test_flat_kernels_in_the_tuning_build(mode):
    if not os.path.exists(TUNING):
        pytest.skip("tuning library not built (python -m pytortto_b200.build --tuning)")
    env = dict(os.environ, TORTTO_B200_LIB="tuning", TTB_FLAT=mode)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "flat_check.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0, "flat-shift kernels disagree with the exact fp32 kernels"
    assert "variants fprop/dgrad [2, 2]" in r.stdout, "the flat-shift kernel was not exercised"
    assert "FAIL" not in r.stdout


def test_release_library_never_selects_the_flat_kernel():
    import pytortto_b200 as tt
    from pytortto_b200 import _cabi, ops
    if _cabi.LIB_PATH.endswith("_tuning.so"):
        pytest.skip("running against the tuning library")
    tt.set_math_mode("tf32")
    lib = _cabi.load()

    def variant(n, c, h, k, ks=3, stride=1, pad=1, dil=1):
        d = ops.conv_desc((n, c, h, h), (k, c, ks, ks), (stride, stride), (pad, pad), (dil, dil), 1)
        return [lib.ttb_conv2d_kernel_variant(ctypes.byref(d), p) for p in (0, 1, 2)]
    assert variant(256, 64, 32, 64) == [1, 1, 1]
    assert variant(8, 64, 32, 64) == [1, 1, 1]
    assert variant(256, 128, 16, 128) == [1, 1, 1]
    tt.set_math_mode("fp32")
    assert variant(256, 64, 32, 64) == [0, 0, 0]
    tt.set_math_mode("tf32")

"""Parity + timing of the flat-shift halo-tile kernels (csrc/conv_flat.cu) on a B200:

    python scripts/flat_check.py                                        # release library: im2col kernels (the A/B baseline)
    TORTTO_B200_LIB=tuning TTB_FLAT=-1 python scripts/flat_check.py     # resident-weight flat-shift variant where it applies
    TORTTO_B200_LIB=tuning TTB_FLAT=1 python scripts/flat_check.py      # flat-shift everywhere eligible (streamed variant too)

Every case is compared with the exact fp32 direct kernels on the same inputs (tolerance 2e-3 of the tensor max) and
timed from a replayed CUDA graph; the kernel variant each pass took is printed (1 im2col, 2 flat-shift resident)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytortto_b200 as tt
from pytortto_b200 import _cabi, ops
from pytortto_b200.xparray import cparray, current_stream_ptr
from scripts.conv_sweep import graph_time_us

CASES = [  # n, c, h, w, k, ks, pad, bias
    (2, 32, 8, 8, 32, 3, 1, False),
    (3, 64, 12, 10, 64, 3, 1, True),      # W + 2 = 12: rows not a multiple of 8 pixels, ragged last tile
    (4, 64, 32, 32, 64, 3, 1, False),     # layer-1 shape at a small batch
    (4, 128, 16, 16, 128, 3, 1, False),   # two strip groups per tile
    (2, 96, 9, 7, 72, 3, 1, True),        # 3 slabs (a 1-slab last group), K not a multiple of 32
    (2, 64, 10, 10, 64, 5, 2, False),     # 5x5
    (2, 64, 8, 8, 40, 3, 0, False),       # no padding
    (256, 64, 32, 32, 64, 3, 1, False),   # the real layer-1 problem
    (256, 128, 16, 16, 128, 3, 1, False),
]


def main():
    print("library:", _cabi.LIB_PATH, " TTB_FLAT =", os.environ.get("TTB_FLAT"))
    rng = np.random.default_rng(0)
    worst = 0.0
    for n, c, h, w, k, ks, pad, bias in CASES:
        x = cparray.from_numpy(rng.standard_normal((n, c, h, w)).astype(np.float32))
        wt = cparray.from_numpy((rng.standard_normal((k, c, ks, ks)) / np.sqrt(c * ks * ks)).astype(np.float32))
        b = cparray.from_numpy(rng.standard_normal((k,)).astype(np.float32)) if bias else None
        tt.set_math_mode("fp32")
        d32 = ops.conv_desc(x.shape, wt.shape, (1, 1), (pad, pad), (1, 1), 1)
        ref = ops.conv2d_fprop(x, wt, b, d32).get()
        tt.set_math_mode("tf32")
        d = ops.conv_desc(x.shape, wt.shape, (1, 1), (pad, pad), (1, 1), 1)
        y = ops.conv2d_fprop(x, wt, b, d).get()
        err = float(np.abs(y - ref).max() / np.abs(ref).max())
        worst = max(worst, err)
        t = graph_time_us(lambda: ops.conv2d_fprop(x, wt, b, d))
        gf = 2.0 * n * d.p * d.q * k * c * ks * ks / 1e9
        var = [_cabi.load().ttb_conv2d_kernel_variant(ctypes.byref(d), p_) for p_ in (0, 1)]
        print(f"n{n} c{c} {h}x{w} k{k} f{ks} p{pad} bias={int(bias)} variants fprop/dgrad {var}: rel-err {err:.2e} {'OK' if err < 2e-3 else 'FAIL'}"
              f"   {t:8.1f} us  {gf / t * 1e3:6.0f} TF/s", flush=True)
        # dgrad through the pre-packed entry points (what a backward sweep calls; the prototype hooks in there)
        lib = _cabi.load()
        if k % 32 == 0 and lib.ttb_conv2d_dgrad_prepacked_supported(ctypes.byref(d)):
            dy = cparray.from_numpy(rng.standard_normal((n, k, d.p, d.q)).astype(np.float32))
            tt.set_math_mode("fp32")
            dref = ops.conv2d_dgrad(dy, wt, d32).get()
            tt.set_math_mode("tf32")
            packed = torch.empty(wt.t.numel(), dtype=torch.float32, device="cuda")
            descs = (ctypes.POINTER(_cabi.ConvDesc) * 1)(ctypes.pointer(d))
            src = (ctypes.c_void_p * 1)(wt.t.data_ptr())
            dst = (ctypes.c_void_p * 1)(packed.data_ptr())
            _cabi.call("ttb_conv2d_dgrad_pack_weights", 1, descs, src, dst, current_stream_ptr())
            dx = cparray(torch.empty_like(x.t))

            def run_dgrad():
                _cabi.call("ttb_conv2d_dgrad_prepacked", ctypes.byref(d), dy.t.data_ptr(), packed.data_ptr(), None, dx.t.data_ptr(),
                           current_stream_ptr())
            run_dgrad()
            derr = float(np.abs(dx.get() - dref).max() / np.abs(dref).max())
            worst = max(worst, derr)
            td = graph_time_us(run_dgrad)
            print(f"    dgrad: rel-err {derr:.2e} {'OK' if derr < 2e-3 else 'FAIL'}   {td:8.1f} us  {gf / td * 1e3:6.0f} TF/s", flush=True)
    print("worst rel-err", worst)
    return 0 if worst < 2e-3 else 1


if __name__ == "__main__":
    sys.exit(main())

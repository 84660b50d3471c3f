"""Does an HBM-bound BatchNorm backward overlap with a tensor-bound wgrad when they run on two streams?
Serial vs concurrent device time for the layer-1 shapes (batch 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytortto_b200 as tt
from pytortto_b200 import ops
from pytortto_b200.xparray import cparray

tt.set_math_mode("tf32")
rng = np.random.default_rng(0)
for (n, c, h, k) in [(256, 64, 32, 64), (256, 128, 16, 128), (256, 256, 8, 256)]:
    x = cparray.from_numpy(rng.standard_normal((n, c, h, h)).astype(np.float32))
    w = cparray.from_numpy((rng.standard_normal((k, c, 3, 3)) * 0.05).astype(np.float32))
    d = ops.conv_desc(x.shape, w.shape, (1, 1), (1, 1), (1, 1), 1)
    dy = cparray.from_numpy(rng.standard_normal((n, k, h, h)).astype(np.float32))
    a = torch.randn(n, c, h, h, device="cuda").contiguous(memory_format=torch.channels_last)
    b = torch.empty_like(a)
    side = torch.cuda.Stream()

    def hbm_work():  # stand-in for BN backward: a few streaming passes over an activation-sized tensor
        for _ in range(3):
            torch.add(a, 1.0, out=b)

    def timed(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters

    def serial():
        ops.conv2d_wgrad(x, dy, d)
        hbm_work()

    def concurrent():
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.conv2d_wgrad(x, dy, d)
        hbm_work()
        main.wait_stream(side)

    def concurrent_dgrad():
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.conv2d_wgrad(x, dy, d)
        ops.conv2d_dgrad(dy, w, d)
        main.wait_stream(side)

    t_w = timed(lambda: ops.conv2d_wgrad(x, dy, d))
    t_h = timed(hbm_work)
    t_d = timed(lambda: ops.conv2d_dgrad(dy, w, d))
    print(f"C={c} H={h}: wgrad {t_w:.1f} us, hbm passes {t_h:.1f} us, dgrad {t_d:.1f} us | serial w+hbm {timed(serial):.1f} | "
          f"concurrent w||hbm {timed(concurrent):.1f} | concurrent w||dgrad {timed(concurrent_dgrad):.1f}", flush=True)

"""Per-pass timing of the conv kernels on the preact_resnet18 layer shapes (batch 256), CUDA events, TF32.
    python scripts/bench_conv.py [variant ...]      (TTB_WGRAD_VARIANT values to compare)
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytortto_b200 as tt
from pytortto_b200 import ops
from pytortto_b200.xparray import cparray

LAYERS = [  # name, N, C, H, W, K, ks, stride, pad
    ("stem 3->64 3x3 32", 256, 3, 32, 32, 64, 3, 1, 1),
    ("L1 64->64 3x3 32", 256, 64, 32, 32, 64, 3, 1, 1),
    ("L2 64->128 3x3 s2", 256, 64, 32, 32, 128, 3, 2, 1),
    ("L2 64->128 1x1 s2", 256, 64, 32, 32, 128, 1, 2, 0),
    ("L2 128->128 3x3 16", 256, 128, 16, 16, 128, 3, 1, 1),
    ("L3 128->256 3x3 s2", 256, 128, 16, 16, 256, 3, 2, 1),
    ("L3 256->256 3x3 8", 256, 256, 8, 8, 256, 3, 1, 1),
    ("L4 256->512 3x3 s2", 256, 256, 8, 8, 512, 3, 2, 1),
    ("L4 512->512 3x3 4", 256, 512, 4, 4, 512, 3, 1, 1),
]


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    variants = sys.argv[1:] or ["0"]
    tt.set_math_mode(os.environ.get("MATH", "tf32"))
    rng = np.random.default_rng(0)
    only = os.environ.get("ONLY")
    for name, n, c, h, w, k, ks, s, p in LAYERS:
        if only and only not in name:
            continue
        x = cparray.from_numpy(rng.standard_normal((n, c, h, w)).astype(np.float32))
        wt = cparray.from_numpy((rng.standard_normal((k, c, ks, ks)) * 0.05).astype(np.float32))
        d = ops.conv_desc(x.shape, wt.shape, (s, s), (p, p), (1, 1), 1)
        dy = cparray.from_numpy(rng.standard_normal((n, k, d.p, d.q)).astype(np.float32))
        gf = 2.0 * n * d.p * d.q * k * c * ks * ks / 1e9
        t_f = timeit(lambda: ops.conv2d_fprop(x, wt, None, d))
        t_d = timeit(lambda: ops.conv2d_dgrad(dy, wt, d))
        line = f"{name:22s} {gf:6.2f} GF | fprop {t_f:7.1f} us {gf / t_f * 1e3:6.0f} TF/s | dgrad {t_d:7.1f} us {gf / t_d * 1e3:6.0f} TF/s |"
        for v in variants:
            os.environ["TTB_WGRAD_VARIANT"] = v
            t_w = timeit(lambda: ops.conv2d_wgrad(x, dy, d))
            line += f" wgrad[v{v}] {t_w:7.1f} us {gf / t_w * 1e3:5.0f} TF/s |"
        os.environ.pop("TTB_WGRAD_VARIANT", None)
        print(line, flush=True)


if __name__ == "__main__":
    main()

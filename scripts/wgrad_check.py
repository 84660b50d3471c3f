"""wgrad of single layers: parity against a float64 torch reference + device time from a replayed CUDA graph.
    [TORTTO_B200_LIB=tuning TTB_WGRAD_HALO=0|1|2 TTB_HALO_KP=64|128] python scripts/wgrad_check.py [tf32|bf16] [time]
Used for the A/B of the haloed-tile wgrad (conv_wgrad_halo.cu) against the im2col wgrad (conv_igemm.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytortto_b200 as tt
from pytortto_b200 import ops
from pytortto_b200.xparray import cparray
from scripts.conv_sweep import graph_time_us

# (n, c, h, w, k, r, s, pad)
SHAPES = [
    (4, 32, 8, 8, 32, 3, 3, 1),
    (3, 64, 16, 16, 64, 3, 3, 1),
    (2, 64, 12, 20, 32, 3, 3, 1),      # ragged box rows / columns
    (2, 32, 9, 24, 64, 3, 2, 0),       # no padding, 2-wide filter
    (2, 128, 16, 16, 128, 3, 3, 1),
    (2, 64, 70, 70, 64, 3, 3, 1),      # rows wider than a box
    (2, 64, 8, 8, 256, 3, 3, 1),
    (256, 64, 32, 32, 64, 3, 3, 1),    # preact_resnet18 stage 1
    (256, 128, 16, 16, 128, 3, 3, 1),
    (256, 256, 8, 8, 256, 3, 3, 1),
    (8, 32, 512, 512, 32, 3, 3, 1),    # UNet level 1
    (8, 64, 256, 256, 64, 3, 3, 1),
    (8, 128, 128, 128, 128, 3, 3, 1),
    (256, 64, 56, 56, 64, 3, 3, 1),    # ResNet-50 stage 1
    (256, 128, 28, 28, 128, 3, 3, 1),
]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    timing = len(sys.argv) > 2
    tt.set_math_mode(mode)
    torch.backends.cudnn.allow_tf32 = False  # (the large shapes use an fp32 cuDNN reference)
    tol = 2e-3 if mode == "tf32" else 1e-2
    rng = np.random.default_rng(0)
    print(f"mode {mode}  TTB_WGRAD_HALO={os.environ.get('TTB_WGRAD_HALO', '-')} TTB_HALO_KP={os.environ.get('TTB_HALO_KP', '-')}")
    bad = 0
    only = os.environ.get("ONLY")
    for (n, c, h, w, k, r, s, pad) in SHAPES:
        if only and only != f"{n},{c},{h},{k}":
            continue
        if mode == "bf16" and (c % 64 or k % 64):
            continue
        big = n * c * h * w > 3e7
        x = torch.from_numpy(rng.standard_normal((n, c, h, w)).astype(np.float32)).cuda()
        wt_shape = (k, c, r, s)
        d = ops.conv_desc((n, c, h, w), wt_shape, (1, 1), (pad, pad), (1, 1), 1)
        dy = torch.from_numpy(rng.standard_normal((n, k, d.p, d.q)).astype(np.float32)).cuda()
        xa = cparray(x.contiguous(memory_format=torch.channels_last))
        dya = cparray(dy.contiguous(memory_format=torch.channels_last))
        dw = ops.conv2d_wgrad(xa, dya, d)
        torch.cuda.synchronize()
        got = dw.t.double()
        ref = torch.nn.grad.conv2d_weight(x.double() if not big else x.float(), wt_shape, dy.double() if not big else dy.float(),
                                          stride=1, padding=pad)
        err = float((got - ref.double()).abs().max() / ref.double().abs().max())
        gf = 2.0 * n * d.p * d.q * k * c * r * s / 1e9
        line = f"n{n:3d} c{c:3d} {h:3d}x{w:3d} k{k:3d} f{r}x{s} p{pad} | rel-err {err:.2e} {'ok ' if err < tol else 'BAD'}"
        bad += err >= tol
        if timing:
            t = graph_time_us(lambda: ops.conv2d_wgrad(xa, dya, d))
            line += f" | {t:8.1f} us {gf / t * 1e3:6.0f} TF/s"
        print(line, flush=True)
    print("FAILED" if bad else "all ok")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

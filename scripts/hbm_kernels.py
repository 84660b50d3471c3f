"""One launch of every HBM-bound kernel at the sizes of the BASELINE configurations - the command profiled with
`ncu --set full` for the per-kernel DRAM-byte / achieved-GB/s table under profiles/ (VERDICT r1 item 7):

    python scripts/hbm_kernels.py            # launches only (under ncu)
    python scripts/hbm_kernels.py --time     # device time from a replayed CUDA graph + algorithmic GB/s, no profiler

BatchNorm statistics / apply(+ReLU) / backward reduce / backward apply(+ReLU mask, + residual accumulate) on the
preact_resnet18 layer-1 activation (256 x 64 x 32 x 32 fp32 = 67 MB, larger than... the 126 MB L2 holds at most one
operand, every pass streams from HBM), ReLU fwd / bwd, add, MaxPool2d(3, 2, 1) fwd / bwd on the ResNet-50 stem output
(64 x 64 x 112 x 112), channel concatenation (UNet), global mean, bias gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytortto_b200 as tt
from pytortto_b200 import ops
from pytortto_b200.xparray import cparray, new_f32
from scripts.conv_sweep import graph_time_us

tt.set_math_mode("tf32")
rng = np.random.default_rng(0)


def arr(*shape):
    return cparray(torch.randn(shape, device="cuda").contiguous(memory_format=torch.channels_last) if len(shape) == 4
                   else torch.randn(shape, device="cuda"))


x, dy, res = arr(256, 64, 32, 32), arr(256, 64, 32, 32), arr(256, 64, 32, 32)
g, b = arr(64), arr(64)
rm, rv = new_f32((64,)), new_f32((64,))
y, stats, count = ops.bn_forward_train(x, g, b, rm, rv, 0.1, 1e-5, relu=True)
n_el = x.size
px = arr(64, 64, 112, 112)
pd = ops.pool_desc(px.shape, (3, 3), (2, 2), (1, 1), (1, 1), False)
py, pidx = ops.maxpool2d_fwd(px, pd)
pdy = arr(*py.shape)
ca, cb = arr(8, 64, 256, 256), arr(8, 64, 256, 256)
cases = [  # name, callable, algorithmic bytes (what the pass must read + write once)
    ("bn_stats (col_reduce<0>)", lambda: ops.bn_sums(x), 4 * n_el),
    ("bn_stats+finalize+apply+relu (forward)", lambda: ops.bn_forward_train(x, g, b, rm, rv, 0.1, 1e-5, relu=True), 12 * n_el),
    ("bn_backward: reduce+finalize+apply, relu mask recomputed, residual accumulate", lambda: ops.bn_backward(
        dy, x, g, stats, count, fused_relu=True, accum=res), 24 * n_el),
    ("relu_fwd", lambda: ops.relu_fwd(x), 8 * n_el),
    ("relu_bwd", lambda: ops.relu_bwd(dy, y), 12 * n_el),
    ("add", lambda: ops.add_arrays(x, dy), 12 * n_el),
    ("maxpool2d_fwd 3x3 s2 p1 (64x64x112x112)", lambda: ops.maxpool2d_fwd(px, pd), 4 * px.size + 5 * py.size),
    ("maxpool2d_bwd", lambda: ops.maxpool2d_bwd(pdy, pidx, pd), 4 * px.size + 5 * py.size),
    ("cat_channels 64+64 (8x256x256)", lambda: ops.cat_channels([ca, cb]), 16 * ca.size),
    ("mean_hw (256x512x4x4)", None, 0),
    ("bias_grad (8x64x512x512 rows)", None, 0),
]
mx = arr(256, 512, 4, 4)
cases[9] = ("mean_hw (256x512x4x4)", lambda: ops.mean_hw(mx, True), 4 * mx.size)
bx = arr(8, 64, 256, 256)
cases[10] = ("bias_grad (8x64x256x256)", lambda: ops.bias_grad(bx), 4 * bx.size)

if "--time" in sys.argv:
    peak = 6552.6
    print(f"{'kernel(s)':78s} {'us':>9s} {'GB/s (algorithmic)':>20s} {'of 6552.6':>10s}")
    for name, fn, nbytes in cases:
        t = graph_time_us(fn)
        print(f"{name:78s} {t:9.1f} {nbytes / t / 1e3:20.0f} {nbytes / t / 1e3 / peak:10.3f}")
else:
    for name, fn, _ in cases:
        fn()
    torch.cuda.synchronize()
    print("ok")

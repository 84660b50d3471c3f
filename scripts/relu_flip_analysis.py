"""Why whole-network gradients of preact_resnet18 (batch 256, at initialisation) cannot agree to 2e-3 between two
arithmetics - measured on the CPU with the numpy oracle alone (no GPU involved):

    python scripts/relu_flip_analysis.py [--batch 256] > profiles/r2_relu_flip_analysis.txt

1. the reference's fp32 algorithm vs a float64 evaluation of the same formulas: where the first difference appears
   (one ReLU decision), how large it is, how it propagates;
2. the same algorithm with conv operands rounded to TF32 (10-bit mantissa) / bf16 (7-bit), i.e. what ANY tensor-core
   implementation of the reference computes: number of ReLU decisions that differ and the resulting gradient error.
TEST INFRASTRUCTURE / analysis only (imports oracle/)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import pytortto_b200 as tt  # noqa: E402  (host-side module system only: same initial parameters as the tests)
from oracle import resnet_oracle as R  # noqa: E402
from oracle import tortto_oracle as O  # noqa: E402
from pytortto_b200.examples import make_models  # noqa: E402

L, C = [2, 2, 2, 2], [64, 128, 256, 512]


def round_mantissa(a, bits):
    """round-to-nearest-even to `bits` explicit mantissa bits (10 = TF32, 7 = bf16), keeping float32 storage"""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    u = u + ((1 << (drop - 1)) - 1) + ((u >> drop) & 1)
    u = (u >> drop) << drop
    return (u & 0xFFFFFFFF).astype(np.uint32).view(np.float32).reshape(a.shape)


def run(x, lab, p0, dtype=np.float32, bits=None):
    """one forward + backward; returns grads, the ReLU masks (in call order) and the loss"""
    masks = []
    f_fwd, f_bwd, f_relu = O.conv2d_forward, O.conv2d_backward, O.relu_forward
    if bits is not None:
        def fwd(x_, w_, b_, *a, **k):
            return f_fwd(round_mantissa(x_, bits), round_mantissa(w_, bits), b_, *a, **k)

        def bwd(x_, w_, dy_, *a, **k):
            return f_bwd(round_mantissa(x_, bits), round_mantissa(w_, bits), round_mantissa(dy_, bits), *a, **k)
        O.conv2d_forward, O.conv2d_backward = fwd, bwd

    def relu(y):
        out = f_relu(y)
        masks.append(out > 0)
        return out
    O.relu_forward = relu
    try:
        o = R.StepOracle(L, C, {k: v.astype(dtype) for k, v in p0.items()}, dtype=dtype)
        loss, _, grads = o.forward_backward(x.astype(dtype), lab)
    finally:
        O.conv2d_forward, O.conv2d_backward, O.relu_forward = f_fwd, f_bwd, f_relu
    return grads, masks, float(loss)


def compare(tag, g, m, g64, m64):
    flips = [int((a != b).sum()) for a, b in zip(m, m64)]
    total = sum(a.size for a in m64)
    worst, worst_l2, name = 0.0, 0.0, ""
    for k in g64:
        d = g[k].astype(np.float64) - g64[k]
        rel = float(np.abs(d).max() / np.abs(g64[k]).max())
        l2 = float(np.linalg.norm(d) / np.linalg.norm(g64[k]))
        if l2 > worst_l2:
            worst_l2, name = l2, k
        worst = max(worst, rel)
    print(f"{tag}: ReLU decisions that differ from float64: {sum(flips)} of {total} ({sum(flips) / total:.2e}); per BN+ReLU layer "
          f"(forward order) {flips}")
    print(f"{tag}: worst gradient error vs float64: max-abs/tensor-max {worst:.3e}, rel-L2 {worst_l2:.3e} ({name})")
    return flips


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((a.batch, 3, 32, 32)).astype(np.float32)
    lab = rng.integers(0, 10, a.batch).astype(np.int64)
    tt.manual_seed(0)
    M = make_models(tt)
    net = M["PreactResNet"](M["BasicBlock"], L, C)
    p0 = {k: np.array(p.data, copy=True) for k, p in net.named_parameters()}
    t = time.time()
    g64, m64, l64 = run(x, lab, p0, np.float64)
    print(f"preact_resnet18, batch {a.batch}, seed 0 (the inputs of tests/test_gpu_fullsize.py); float64 loss {l64:.8f} "
          f"[{time.time() - t:.0f} s]")
    g32, m32, l32 = run(x, lab, p0, np.float32)
    print(f"fp32 loss {l32:.8f}")
    compare("reference algorithm, fp32", g32, m32, g64, m64)
    names = list(g64)
    print("  per tensor (backward order), fp32 vs float64 rel-L2: " +
          ", ".join(f"{k}={np.linalg.norm(g32[k] - g64[k]) / np.linalg.norm(g64[k]):.1e}" for k in names[:12]) + ", ...")
    for tag, bits in (("TF32 operands (10-bit mantissa)", 10), ("bf16 operands (7-bit mantissa)", 7)):
        g, m, l = run(x, lab, p0, np.float32, bits)
        print(f"{tag}: loss {l:.8f} (rel. to float64 {abs(l - l64) / abs(l64):.2e})")
        compare("reference algorithm, " + tag, g, m, g64, m64)


if __name__ == "__main__":
    main()

"""Per-layer device time of every convolution call of one training step of a bench model (eager pass, CUDA events around
each C-ABI call, GPU parked behind a spin kernel so no launch latency is inside an interval; wgrad overlap off).
    python scripts/per_layer.py [model] [batch] [math]        model: bench.py's names
Prints one row per distinct (entry point, problem): calls, mean us, TFLOP/s."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import pytortto_b200 as tt
from pytortto_b200 import _cabi


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "preact_resnet18"
    cfg = bench.MODELS[model]
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["batch"]
    math = sys.argv[3] if len(sys.argv) > 3 else cfg["math"]

    class A:
        pass
    args = A()
    args.model, args.batch, args.math = model, batch, math
    torch.cuda.set_device(0)
    _cabi.load()
    tt.set_math_mode(math)
    net, crit, opt, x_host, y_host, ydt = bench.build_workload(tt, args, 0)
    x = tt.tensor(x_host.numpy()).cuda()
    y = tt.tensor(y_host.numpy(), dtype=ydt).cuda()

    def step():
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    tt.set_wgrad_overlap(False)
    step()
    records = []
    orig = _cabi.call

    def wrapped(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        key = None
        if name.startswith("ttb_conv2d_") and hasattr(a[0], "_obj") and isinstance(a[0]._obj, _cabi.ConvDesc):
            d = a[0]._obj
            key = (d.n, d.c, d.h, d.w, d.k, d.r, d.s, d.stride_h, d.pad_h, d.groups, d.p, d.q)
        records.append((name, key, e0, e1))
    _cabi.call = wrapped
    torch.cuda.synchronize()
    torch.cuda._sleep(int(2.0e9 * 0.3))
    step()
    torch.cuda.synchronize()
    _cabi.call = orig
    agg = collections.OrderedDict()
    other = collections.OrderedDict()
    for name, key, e0, e1 in records:
        t = e0.elapsed_time(e1) * 1e3
        if key is None:
            o = other.setdefault(name, [0, 0.0])
            o[0] += 1
            o[1] += t
            continue
        a = agg.setdefault((name.replace("ttb_conv2d_", ""), key), [0, 0.0])
        a[0] += 1
        a[1] += t
    print(f"{model} batch {batch} {math}")
    tot = 0.0
    for (name, k), (cnt, t) in agg.items():
        n, c, h, w, kk, r, s, st, pd, g, p, q = k
        gf = 2.0 * n * p * q * kk * (c // g) * r * s / 1e9
        tot += t
        print(f"{name:18s} n{n} c{c:4d} {h:3d}x{w:3d} k{kk:4d} f{r}x{s} s{st} p{pd} g{g} -> {p}x{q} | {cnt:2d}x {t / cnt:8.1f} us "
              f"{gf:7.2f} GF {gf / (t / cnt) * 1e3:6.0f} TF/s")
    print(f"conv total {tot / 1e3:.3f} ms")
    for name, (cnt, t) in sorted(other.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"{name:28s} {cnt:4d}x {t / 1e3:8.3f} ms total {t / cnt:8.1f} us each")


if __name__ == "__main__":
    main()

"""Single-layer Conv2d sweep (SURVEY.md §8(d) cfg 5, the reference's `examples/conv2d_result_speed_comparison.ipynb`):
`nn.Conv2d(C, C, k, stride=s, padding=k//2, bias=False)`, C in {16..512}, (k, s) in {(3,1),(3,2),(1,1),(1,2),(7,2)},
N = 128, H = W mirroring the nets (32,32,32,16,8,4).  Per pass (fprop / dgrad / wgrad), device time of

  * this library (C ABI, TF32 tensor path or the exact direct kernels where the tensor path does not apply),
  * an explicit im2col + cuBLAS GEMM (`torch.nn.functional.unfold` + `matmul`, TF32 on) - the stand-in for the
    reference's CuPy path (`_conv2d` cupy branch = window view + einsum, grad_nn.py:623-642), which cannot run here,
  * the numpy oracle (port of the reference's algorithm) on the host cores at batch 8, scaled to the sweep's batch,
  * cuDNN through `torch.nn.functional.conv2d` / `torch.nn.grad.*` (channels_last, TF32 on) as a library yard-stick.

Every candidate is captured in a CUDA graph holding REPS back-to-back calls and replayed, so host dispatch is
excluded for all three alike.  Writes a markdown table (default gpurun_out/conv_sweep.md).

    python scripts/conv_sweep.py [out.md]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
import pytortto_b200 as tt
from pytortto_b200 import ops
from pytortto_b200.xparray import cparray
from oracle import tortto_oracle as O  # (the numpy column only: the reference's CPU algorithm as a baseline)

REPS = 10
N = 128
SIZES = {16: 32, 32: 32, 64: 32, 128: 16, 256: 8, 512: 4}
KS = [(3, 1), (3, 2), (1, 1), (1, 2), (7, 2)]


def graph_time_us(fn):
    """device microseconds per call of fn, from a replayed CUDA graph of REPS calls"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = float("inf")
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / REPS)
    return best


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_sweep.md"
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    tt.set_math_mode("tf32")
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    rng = np.random.default_rng(0)
    rows = []
    for c, hw in SIZES.items():
        for k, s in KS:
            h = 224 if k == 7 else hw  # 7x7/s2 is the ImageNet stem shape
            n = 16 if k == 7 else N
            cin = 3 if k == 7 else c
            p = k // 2
            xn = rng.standard_normal((n, cin, h, h)).astype(np.float32)
            wn = (rng.standard_normal((c, cin, k, k)) * 0.05).astype(np.float32)
            x = cparray.from_numpy(xn)
            w = cparray.from_numpy(wn)
            d = ops.conv_desc(x.shape, w.shape, (s, s), (p, p), (1, 1), 1)
            dyn = rng.standard_normal((n, c, d.p, d.q)).astype(np.float32)
            dy = cparray.from_numpy(dyn)
            gf = 2.0 * n * d.p * d.q * c * cin * k * k / 1e9
            ours = [graph_time_us(lambda: ops.conv2d_fprop(x, w, None, d)),
                    graph_time_us(lambda: ops.conv2d_dgrad(dy, w, d)),
                    graph_time_us(lambda: ops.conv2d_wgrad(x, dy, d))]
            # cuDNN (channels_last tensors = the same physical layout this library uses)
            xt, wt_, dyt = x.t, w.t, dy.t
            cud = [graph_time_us(lambda: F.conv2d(xt, wt_, None, s, p)),
                   graph_time_us(lambda: torch.nn.grad.conv2d_input(xt.shape, wt_, dyt, s, p)),
                   graph_time_us(lambda: torch.nn.grad.conv2d_weight(xt, wt_.shape, dyt, s, p))]
            # explicit im2col + GEMM (forward and weight gradient share the unfolded matrix; dgrad = GEMM + fold)
            xc = xt.contiguous()
            w2 = wt_.contiguous().view(c, -1)
            dy2 = dyt.contiguous().view(n, c, -1)

            def im2col_f():
                return torch.matmul(w2, F.unfold(xc, k, 1, p, s))

            def im2col_d():
                return F.fold(torch.matmul(w2.t(), dy2), (h, h), k, 1, p, s)

            def im2col_w():
                return torch.matmul(dy2, F.unfold(xc, k, 1, p, s).transpose(1, 2)).sum(0)

            try:
                i2c = [graph_time_us(im2col_f), graph_time_us(im2col_d), graph_time_us(im2col_w)]
            except Exception as e:  # e.g. out of memory for an unfolded 7x7 matrix
                print("im2col skipped:", type(e).__name__, e)
                i2c = [float("nan")] * 3
            tensor_path = bool(ops.tensor_path_supported(d)) if hasattr(ops, "tensor_path_supported") else None
            # the reference's own numpy algorithm (oracle port) on the host cores, at a reduced batch, scaled to n images
            nn_ = 4 if k == 7 else 8
            cpu = []
            import time as _t
            for fn_ in (lambda: O.conv2d_forward(xn[:nn_], wn, None, s, p), lambda: O.conv2d_backward_input(dyn[:nn_], wn, (h, h), s, p),
                        lambda: O.conv2d_backward_weight(xn[:nn_], dyn[:nn_], wn.shape, s, p)):
                t0 = _t.perf_counter()
                fn_()
                cpu.append((_t.perf_counter() - t0) * 1e6 * n / nn_)
            rows.append((c, cin, k, s, h, n, gf, ours, i2c, cud, tensor_path, cpu))
            print(f"C={cin}->{c} k{k} s{s} H{h} N{n} {gf:7.2f} GF  ours {ours[0]:7.1f}/{ours[1]:7.1f}/{ours[2]:7.1f} us  "
                  f"im2col {i2c[0]:7.1f}/{i2c[1]:7.1f}/{i2c[2]:7.1f}  cudnn {cud[0]:7.1f}/{cud[1]:7.1f}/{cud[2]:7.1f}", flush=True)
            del x, w, dy, xt, wt_, dyt, xc, w2, dy2
            torch.cuda.empty_cache()
    with open(out_path, "w") as f:
        f.write("# Conv2d single-layer sweep (cfg 5), B200, TF32, N=128 (7x7: 3->C, 224x224, N=16); device time per call\n\n")
        f.write("`ours` = this library through the C ABI; `im2col` = torch unfold + cuBLAS TF32 matmul (stand-in for the\n"
                "reference's CuPy window-view + einsum path); `cudnn` = torch.nn.functional.conv2d / torch.nn.grad.* on\n"
                "channels_last tensors with TF32 allowed.  TF/s = 2*N*P*Q*K*C*R*S / time.  fprop / dgrad / wgrad.\n\n")
        f.write("| Cin->Cout | k | s | H | GFLOP/pass | ours us (f/d/w) | ours TF/s (f/d/w) | im2col us (f/d/w) | cudnn us (f/d/w) | "
                "ours vs im2col (f/d/w) | ours vs cudnn (f/d/w) | numpy oracle ms (f/d/w; host cores, batch 8 scaled) |\n"
                "|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for c, cin, k, s, h, n, gf, o, i, q, tp, cpu in rows:
            fmt = lambda v: "/".join(f"{a:.1f}" for a in v)
            tf = "/".join(f"{gf / a * 1e3:.0f}" for a in o)
            r1 = "/".join(f"{b / a:.2f}x" for a, b in zip(o, i))
            r2 = "/".join(f"{b / a:.2f}x" for a, b in zip(o, q))
            cp = "/".join(f"{a / 1e3:.0f}" for a in cpu)
            f.write(f"| {cin}->{c} | {k} | {s} | {h} | {gf:.2f} | {fmt(o)} | {tf} | {fmt(i)} | {fmt(q)} | {r1} | {r2} | {cp} |\n")
    print("wrote", out_path)


if __name__ == "__main__":
    main()

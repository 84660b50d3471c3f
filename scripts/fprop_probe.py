"""Device time of fprop (and optionally dgrad / wgrad) on the preact_resnet18 layer shapes from a replayed CUDA graph
(no host dispatch in the number).  Used with the TTB_IGEMM_DBG experiment switch of conv_igemm.cu.
    TTB_IGEMM_DBG=1 python scripts/fprop_probe.py [batch]
PASSES: f fprop, s fprop with epilogue statistics, d dgrad (pre-packed weights, as inside a backward sweep), w wgrad;
ONLY=<substring> restricts the layers.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytortto_b200 as tt
from pytortto_b200 import ops
from pytortto_b200.xparray import cparray
from scripts.conv_sweep import graph_time_us
from scripts.bench_conv import LAYERS


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    passes = os.environ.get("PASSES", "f")
    tt.set_math_mode(os.environ.get("MATH", "tf32"))
    rng = np.random.default_rng(0)
    print("TTB_IGEMM_DBG =", os.environ.get("TTB_IGEMM_DBG", "0"))
    only = os.environ.get("ONLY")
    for name, n, c, h, w, k, ks, s, p in LAYERS:
        if only and only not in name:
            continue
        n = batch
        x = cparray.from_numpy(rng.standard_normal((n, c, h, w)).astype(np.float32))
        wt = cparray.from_numpy((rng.standard_normal((k, c, ks, ks)) * 0.05).astype(np.float32))
        d = ops.conv_desc(x.shape, wt.shape, (s, s), (p, p), (1, 1), 1)
        dy = cparray.from_numpy(rng.standard_normal((n, k, d.p, d.q)).astype(np.float32))
        gf = 2.0 * n * d.p * d.q * k * c * ks * ks / 1e9
        line = f"{name:22s} {gf:6.2f} GF |"
        if "f" in passes:
            t = graph_time_us(lambda: ops.conv2d_fprop(x, wt, None, d))
            line += f" fprop {t:7.1f} us {gf / t * 1e3:6.0f} TF/s |"
        if "s" in passes:
            t = graph_time_us(lambda: ops.conv2d_fprop(x, wt, None, d, stats=True))
            line += f" fprop+stats {t:7.1f} us {gf / t * 1e3:6.0f} TF/s |"
        if "d" in passes:
            ops.register_dgrad_weight(wt, d)
            ops.begin_backward_sweep()
            ops.conv2d_dgrad(dy, wt, d)  # (packs the weights once, outside the timed graph)
            t = graph_time_us(lambda: ops.conv2d_dgrad(dy, wt, d))
            ops.end_backward_sweep()
            line += f" dgrad {t:7.1f} us {gf / t * 1e3:6.0f} TF/s |"
        if "w" in passes:
            t = graph_time_us(lambda: ops.conv2d_wgrad(x, dy, d))
            line += f" wgrad {t:7.1f} us {gf / t * 1e3:6.0f} TF/s |"
        print(line, flush=True)


if __name__ == "__main__":
    main()

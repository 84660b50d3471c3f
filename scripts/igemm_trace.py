"""clock64() timeline of CTA 0 of one persistent fprop igemm launch (diagnostics hook ttb_debug_set_igemm_trace).
    python scripts/igemm_trace.py [layer-substring ...]
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytortto_b200 as tt
from pytortto_b200 import ops, _cabi
from pytortto_b200.xparray import cparray
from scripts.bench_conv import LAYERS


def main():
    want = sys.argv[1:] or ["L1 64->64", "L3 256->256", "L4 512->512"]
    tt.set_math_mode("tf32")
    lib = _cabi.load()
    lib.ttb_debug_set_igemm_trace.argtypes = [ctypes.c_void_p]
    lib.ttb_debug_set_igemm_trace.restype = ctypes.c_int
    rng = np.random.default_rng(0)
    for name, n, c, h, w, k, ks, s, p in LAYERS:
        if not any(t in name for t in want):
            continue
        x = cparray.from_numpy(rng.standard_normal((n, c, h, w)).astype(np.float32))
        wt = cparray.from_numpy((rng.standard_normal((k, c, ks, ks)) * 0.05).astype(np.float32))
        d = ops.conv_desc(x.shape, wt.shape, (s, s), (p, p), (1, 1), 1)
        for _ in range(3):
            ops.conv2d_fprop(x, wt, None, d)
        buf = torch.zeros(4200, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        lib.ttb_debug_set_igemm_trace(ctypes.c_void_p(buf.data_ptr()))
        ops.conv2d_fprop(x, wt, None, d)
        torch.cuda.synchronize()
        lib.ttb_debug_set_igemm_trace(ctypes.c_void_p(0))
        t = buf.cpu().numpy()
        t0 = t[0]
        nkb = ks * ks * max(c // 32, 1)
        f = lambda a: " ".join(str(int(v - t0)) if v else "-" for v in a)
        print(f"== {name}: {nkb} k-blocks per tile; clocks relative to CTA entry; prologue done {t[1]-t0}; CTA end {t[2]-t0}")
        ntile = int(np.count_nonzero(t[16:516:2]))
        print(f"   tiles by CTA 0: {ntile}")
        print("   MMA tile start      :", f(t[16:16 + 2 * min(ntile, 8):2]))
        print("   MMA tile issued     :", f(t[17:17 + 2 * min(ntile, 8):2]))
        print("   epilogue start      :", f(t[528:528 + 2 * min(ntile, 8):2]))
        print("   epilogue end        :", f(t[529:529 + 2 * min(ntile, 8):2]))
        print("   producer issue (all 0..39; '-' = other CTA-0 producers do not stamp):", f(t[2048:2088]))
        print("   MMA full-wait passed 0..39:", f(t[3072:3112]))
        cons = t[3072:3072 + min(1000, ntile * nkb)]
        dc = np.diff(cons[4:])
        if len(dc):
            print(f"   MMA k-block interval: median {np.median(dc):.0f} mean {dc.mean():.0f} min {dc.min()} max {dc.max()}")


if __name__ == "__main__":
    main()

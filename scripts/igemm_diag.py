"""GPU diagnostic: tcgen05 implicit-GEMM conv (tf32) vs the exact fp32 direct kernels, pass by pass, with error
structure printed and the first failing case dumped to gpurun_out/ for offline analysis.
    python scripts/igemm_diag.py
"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pytortto_b200 as tt  # noqa: E402
from pytortto_b200 import _cabi, ops  # noqa: E402
from pytortto_b200.xparray import cparray  # noqa: E402

MODE = os.environ.get("MATH", "tf32")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def run_pass(mode, which, x, w, dy, s, p, d):
    tt.set_math_mode(mode)
    desc = ops.conv_desc(x.shape, w.shape, (s, s), (p, p), (d, d), 1)
    sup = _cabi.load().ttb_conv2d_tensor_path_supported(ctypes.byref(desc), {"fprop": 0, "dgrad": 1, "wgrad": 2}[which])
    if which == "fprop":
        out = ops.conv2d_fprop(x, w, None, desc)
    elif which == "dgrad":
        out = ops.conv2d_dgrad(dy, w, desc)
    else:
        out = ops.conv2d_wgrad(x, dy, desc)
    torch.cuda.synchronize()
    return out.get(), sup


def structure(got, ref, name):
    diff = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    denom = max(np.abs(ref).max(), 1e-30)
    bad = diff > 5e-3 * denom
    print(f"    {name}: {bad.sum()} / {bad.size} elements off; nonfinite={np.sum(~np.isfinite(got))}")
    if bad.any():
        for ax in range(got.ndim):
            other = tuple(a for a in range(got.ndim) if a != ax)
            frac = bad.mean(axis=other)
            idx = np.nonzero(frac > 0)[0]
            print(f"      axis {ax} (size {got.shape[ax]}): bad indices {idx[:24].tolist()}{'...' if idx.size > 24 else ''} "
                  f"frac range {frac.min():.3f}..{frac.max():.3f}")
        i = np.unravel_index(np.argmax(diff), got.shape)
        print(f"      worst @{i}: got {got[i]:.6g} ref {ref[i]:.6g}; ratio got/ref median over bad: "
              f"{np.median(got[bad] / np.where(ref[bad] == 0, 1, ref[bad])):.4g}")


CASES = [  # n, cin, h, w, cout, k, s, p, d
    (2, 32, 8, 8, 32, 1, 1, 0, 1),      # plain GEMM, one K block, M = 128 exactly
    (2, 64, 8, 8, 64, 1, 1, 0, 1),      # two K blocks
    (4, 128, 8, 8, 128, 1, 1, 0, 1),    # BN tile 128
    (2, 32, 8, 8, 32, 3, 1, 1, 1),      # 3x3 taps + padding halo
    (4, 64, 16, 16, 64, 3, 1, 1, 1),
    (3, 64, 15, 17, 96, 3, 1, 1, 1),
    (4, 64, 16, 16, 128, 3, 2, 1, 1),
    (4, 64, 16, 16, 128, 1, 2, 0, 1),
    (2, 128, 9, 9, 128, 3, 2, 1, 1),
    (2, 256, 8, 8, 256, 3, 1, 1, 1),
    (8, 512, 4, 4, 512, 3, 1, 1, 1),
    (2, 64, 12, 12, 64, 3, 1, 2, 2),
    (32, 64, 32, 32, 64, 3, 1, 1, 1),   # many tiles
]


def main():
    failed = 0
    for case in CASES:
        n, ci, h, w_, co, k, s, p, d = case
        rng = np.random.default_rng(abs(hash(case)) % (2 ** 31))
        x = cparray.from_numpy(rng.standard_normal((n, ci, h, w_)).astype(np.float32))
        w = cparray.from_numpy((rng.standard_normal((co, ci, k, k)) / np.sqrt(ci * k * k)).astype(np.float32))
        ho, wo = ops.conv_out_hw(h, w_, k, k, (s, s), (p, p), (d, d))
        dy = cparray.from_numpy(rng.standard_normal((n, co, ho, wo)).astype(np.float32))
        print(f"case n{n} c{ci} {h}x{w_} -> k{co} f{k} s{s} p{p} d{d}  (M={n * ho * wo})", flush=True)
        for which in ("fprop", "dgrad", "wgrad"):
            ref, _ = run_pass("fp32", which, x, w, dy, s, p, d)
            t0 = time.time()
            got, sup = run_pass(MODE, which, x, w, dy, s, p, d)
            denom = max(np.abs(ref).max(), 1e-30)
            rel = float(np.abs(got.astype(np.float64) - ref).max() / denom) if np.isfinite(got).all() else float("inf")
            ok = rel < (2e-3 if MODE == "tf32" else 1e-2)
            print(f"  {which}: tensor_path={sup} rel={rel:.3e} {'OK' if ok else 'FAIL'} ({time.time() - t0:.2f}s)", flush=True)
            if not ok:
                failed += 1
                structure(got, ref, which)
                if failed <= 3:
                    np.savez_compressed(os.path.join(OUT, f"diag_fail_{which}_{'_'.join(map(str, case))}.npz"), got=got,
                                        ref=ref, x=x.get(), w=w.get(), dy=dy.get())
    print("FAILED passes:", failed)
    if failed and MODE == "tf32":
        wgrad_variants()
    return 1 if failed else 0


def wgrad_variants():
    """bring-up aid: try the MN-major descriptor / swizzle variants on one small wgrad problem"""
    case = (4, 64, 16, 16, 64, 3, 1, 1, 1)
    n, ci, h, w_, co, k, s, p, d = case
    rng = np.random.default_rng(1)
    x = cparray.from_numpy(rng.standard_normal((n, ci, h, w_)).astype(np.float32))
    w = cparray.from_numpy((rng.standard_normal((co, ci, k, k)) / np.sqrt(ci * k * k)).astype(np.float32))
    ho, wo = ops.conv_out_hw(h, w_, k, k, (s, s), (p, p), (d, d))
    dy = cparray.from_numpy(rng.standard_normal((n, co, ho, wo)).astype(np.float32))
    ref, _ = run_pass("fp32", "wgrad", x, w, dy, s, p, d)
    for swz in (4, 3):            # CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B = 4, _128B = 3
        for layout in (1, 2):
            for lbo, sbo in ((4096, 512), (512, 4096), (4096, 1024), (1024, 4096), (4096, 256)):
                os.environ.update(TTB_WGRAD_SWIZZLE=str(swz), TTB_WGRAD_LAYOUT=str(layout), TTB_WGRAD_LBO=str(lbo),
                                  TTB_WGRAD_SBO=str(sbo))
                try:
                    got, _ = run_pass("tf32", "wgrad", x, w, dy, s, p, d)
                    rel = float(np.abs(got.astype(np.float64) - ref).max() / np.abs(ref).max())
                    print(f"  variant swz={swz} layout={layout} lbo={lbo} sbo={sbo}: rel={rel:.3e} nonzero={np.mean(got != 0):.3f}", flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"  variant swz={swz} layout={layout} lbo={lbo} sbo={sbo}: EXC {e}", flush=True)
                    return
    for k_ in ("TTB_WGRAD_SWIZZLE", "TTB_WGRAD_LAYOUT", "TTB_WGRAD_LBO", "TTB_WGRAD_SBO"):
        os.environ.pop(k_, None)


if __name__ == "__main__":
    sys.exit(main())

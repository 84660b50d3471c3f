"""Helpers shared by the -m gpu tests."""
import numpy as np
import pytest


def require_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytortto_b200 import _cabi
    _cabi.load()  # fail loudly if the extension is missing on a GPU box


def report(name, got, ref):
    """max-abs and relative error (to the tensor max) + where the worst element is; returned as a string."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if got.shape != ref.shape:
        return f"{name}: SHAPE {got.shape} vs {ref.shape}", float("inf")
    if got.size == 0:
        return f"{name}: empty", 0.0
    diff = np.abs(got - ref)
    bad = ~np.isfinite(got)
    denom = max(float(np.max(np.abs(ref))), 1e-30)
    rel = float(np.max(np.where(bad, np.inf, diff))) / denom
    idx = np.unravel_index(int(np.argmax(np.where(bad, np.inf, diff))), got.shape)
    frac = float(np.mean(diff > 1e-3 * denom))
    return (f"{name}: rel={rel:.3e} maxabs={float(np.max(diff)):.3e} worst@{idx} got={got[idx]:.6g} ref={ref[idx]:.6g} "
            f"frac>1e-3={frac:.4f} nonfinite={int(bad.sum())}"), rel


def assert_close(name, got, ref, tol):
    msg, rel = report(name, got, ref)
    print(msg)
    assert rel <= tol, msg

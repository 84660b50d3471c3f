"""CPU, this container only: run the REAL reference (imported from /root/reference) side by side with the numpy
oracle on fresh random inputs.  Skipped where /root/reference is absent (the GPU box).  Runs in a subprocess
because the reference must be imported before torch/scipy."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/oracle")
import ref_import
tt = ref_import.import_reference()
import numpy as np
from oracle import tortto_oracle as O
rng = np.random.default_rng(1234)
def rel(a, b): return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / max(np.max(np.abs(b)), 1e-30))
worst = 0.0
for trial in range(12):
    n, g = int(rng.integers(1, 4)), int(rng.choice([1, 1, 2]))
    ci, co = g * int(rng.integers(1, 5)), g * int(rng.integers(1, 5))
    k = (int(rng.integers(1, 4)), int(rng.integers(1, 4))); s = (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
    d = (int(rng.integers(1, 3)), int(rng.integers(1, 3))); p = (int(rng.integers(0, 3)), int(rng.integers(0, 3)))
    h, w = int(rng.integers(6, 14)), int(rng.integers(6, 14))
    x = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = rng.standard_normal((co, ci // g, *k)).astype(np.float32)
    b = rng.standard_normal(co).astype(np.float32)
    xt = tt.tensor(x, requires_grad=True); wtt = tt.tensor(wt, requires_grad=True); bt = tt.tensor(b, requires_grad=True)
    y = tt.nn.functional.conv2d(xt, wtt, bt, s, p, d, g)
    dy = rng.standard_normal(y.shape).astype(np.float32)
    y.backward(tt.tensor(dy))
    yo = O.conv2d_forward(x, wt, b, s, p, d, g)
    dxo, dwo, dbo = O.conv2d_backward(x, wt, dy, s, p, d, g, has_bias=True)
    worst = max(worst, rel(yo, y.data), rel(dxo, xt.grad), rel(dwo, wtt.grad), rel(dbo, bt.grad))
    # max pool on the same input
    kk = (int(rng.integers(2, 4)),) * 2; ss = (int(rng.integers(1, 3)),) * 2; pp = (int(rng.integers(0, 2)),) * 2
    ceil = bool(rng.integers(0, 2))
    xt2 = tt.tensor(x, requires_grad=True)
    yp = tt.nn.functional.max_pool2d(xt2, kk, ss, pp, (1, 1), ceil)
    dyp = rng.standard_normal(yp.shape).astype(np.float32)
    yp.backward(tt.tensor(dyp))
    ypo, idx = O.max_pool2d_forward(x, kk, ss, pp, 1, ceil)
    assert np.array_equal(ypo, yp.data), "pool fwd"
    assert np.array_equal(O.max_pool2d_backward(dyp, idx, x.shape, kk, ss, pp, 1, ceil), xt2.grad), "pool bwd"
print("WORST", worst)
assert worst < 2e-5
# Adam / AdamW (optim/adam.py, adamw.py + _functional.py:25-115) vs the oracle's restatement, 4 steps, with weight decay / amsgrad
for cls_name, decoupled, amsgrad in (("Adam", False, False), ("Adam", False, True), ("AdamW", True, False)):
    shapes = [(4, 3, 3, 3), (7,), (5, 6)]
    p0 = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    ref_p = [tt.nn.Parameter(tt.tensor(p.copy())) for p in p0]
    opt = getattr(tt.optim, cls_name)(ref_p, lr=1e-2, weight_decay=1e-2, amsgrad=amsgrad)
    op = [p.copy() for p in p0]; m = [np.zeros_like(p) for p in p0]; v = [np.zeros_like(p) for p in p0]
    vm = [np.zeros_like(p) for p in p0] if amsgrad else None
    for step in range(1, 5):
        gs = [rng.standard_normal(s).astype(np.float32) for s in shapes]
        for p, g in zip(ref_p, gs):
            p.grad = g.copy()
        opt.step()
        O.adam_step(op, gs, m, v, [step] * 3, lr=1e-2, weight_decay=1e-2, max_exp_avg_sqs=vm, decoupled=decoupled)
        for a, b in zip(op, ref_p):
            assert rel(a, b.data) < 2e-6, (cls_name, amsgrad, step, rel(a, b.data))
print("ADAM OK")
# loss functions with reductions / ignore_index (autograd/grad_nn.py:236-349) vs the oracle's restatements
F = tt.nn.functional
for red in ("mean", "sum", "none"):
    x = rng.standard_normal((6, 3, 5, 4)).astype(np.float32) * 3; t = (rng.random((6, 3, 5, 4)) < 0.5).astype(np.float32)
    xt = tt.tensor(x, requires_grad=True)
    y = F.binary_cross_entropy_with_logits(xt, tt.tensor(t), reduction=red)
    g = rng.standard_normal(np.shape(y.data)).astype(np.float32)
    y.backward(tt.tensor(g))
    assert rel(O.bce_with_logits_forward(x, t, red), y.data) < 1e-6 and rel(O.bce_with_logits_backward(g, x, t, red), xt.grad) < 1e-6, red
    for ign in (-100, 2):
        lp = O.log_softmax_forward(rng.standard_normal((9, 5)).astype(np.float32)); tg = rng.integers(0, 5, 9)
        lt = tt.tensor(lp, requires_grad=True)
        y = F.nll_loss(lt, tt.tensor(tg, dtype=np.int64), ignore_index=ign, reduction=red)
        g = rng.standard_normal(np.shape(y.data)).astype(np.float32)
        y.backward(tt.tensor(g.copy()))
        yo, n = O.nll_loss_forward_ex(lp, tg, ign, red)
        assert rel(yo, y.data) < 1e-6 and rel(O.nll_loss_backward_ex(g, lp, tg, ign, red, n), lt.grad) < 1e-6, (red, ign)
print("LOSSES OK")
'''


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/tortto"), reason="reference not mounted")
def test_oracle_matches_live_reference():
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "WORST" in r.stdout and "ADAM OK" in r.stdout and "LOSSES OK" in r.stdout

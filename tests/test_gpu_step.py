"""GPU: whole training step of the reduced PreactResNet through the public API vs the real reference's fixture
(loss, log-probs, every parameter gradient, updated parameters, BN running stats), and the full-size
preact_resnet18 step checked through size-independent properties."""
import numpy as np
import pytest

from conftest import load_golden
from gpu_util import assert_close, require_gpu

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu():
    require_gpu()


def _build(tt, layers, channels):
    from pytortto_b200.examples import make_models
    M = make_models(tt)
    return M["PreactResNet"](M["BasicBlock"], layers, channels)


def _rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# fp32 path: every gradient within 5e-4 (max-abs / tensor max) of the REAL reference - this is the parity gate for
# the whole pipeline.  tf32 path: loss / log-probs within the north-star 2e-3; whole-network gradients are bounded
# loosely because this 8-image, 2x2-spatial toy net is chaotic under TF32 operand rounding: emulating TF32 operand
# truncation inside the numpy oracle moves its own gradients by up to 3.3e-1 (DESIGN.md "TF32 whole-step error").
@pytest.mark.parametrize("mode,tol_grad", [("fp32", 5e-4), ("tf32", 0.5)])
def test_preact_step_golden(mode, tol_grad):
    import pytortto_b200 as tt
    tt.set_math_mode(mode)
    g = load_golden("preact_step.npz")
    names = [str(n) for n in g["param_names"]]
    tt.manual_seed(7)
    net = _build(tt, [1, 1, 1, 1], [32, 32, 64, 64])
    for k, p in net.named_parameters():  # same seed -> bit-identical init (nn/init.py draws from np.random like the reference)
        np.testing.assert_array_equal(p.data, g[f"init/{k}"])
    net.cuda()
    crit = tt.nn.NLLLoss()
    opt = tt.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    net.train()
    worst = 0.0
    for step in range(2):
        opt.zero_grad()
        logp = net(tt.tensor(g[f"step{step}/x"]).cuda())
        loss = crit(logp, tt.tensor(g[f"step{step}/labels"], dtype=np.int64).cuda())
        loss.backward()
        tol_out = 1e-5 if mode == "fp32" else 2e-3
        assert abs(loss.item() - float(g[f"step{step}/loss"])) <= tol_out * (1 if step == 0 else 20) * max(1.0, abs(float(g[f"step{step}/loss"])))
        assert_close(f"step{step} logp", logp.data.get(), g[f"step{step}/logp"], tol_out * (1 if step == 0 else 20))
        params = dict(net.named_parameters())
        for n in names:
            from gpu_util import report
            msg, rel = report(f"step{step} grad {n}", params[n].grad.get(), g[f"step{step}/grad/{n}"])
            worst = max(worst, rel)
            if mode == "fp32" or step == 0:  # the tf32 run diverges chaotically after the first update (see above)
                assert rel <= tol_grad * (1 if step == 0 else 5), msg
            assert np.isfinite(params[n].grad.get()).all()
        opt.step()
    print(f"[{mode}] worst gradient rel-err over both steps: {worst:.3e}")
    if mode != "fp32":
        return
    params = dict(net.named_parameters())
    for n in names:
        assert_close(f"param {n}", params[n].data.get(), g[f"step1/param/{n}"], 1e-4)
    sd = net.state_dict()
    for k in g.files:
        if k.startswith("final/"):
            assert_close(k, sd[k[len("final/"):]], g[k], 1e-4)


def test_preact_resnet18_full_size_properties():
    """BASELINE config 2 shape (batch 256, 3x32x32): finite loss near ln(10) at init, every parameter receives a
    gradient, the TF32 tensor path agrees with the exact-fp32 path on the same weights, and the step is
    deterministic (two runs bit-identical)."""
    import pytortto_b200 as tt
    rng = np.random.default_rng(0)
    x = rng.standard_normal((256, 3, 32, 32)).astype(np.float32)
    lab = rng.integers(0, 10, 256).astype(np.int64)
    results = {}
    for mode in ("tf32", "fp32", "tf32"):
        tt.set_math_mode(mode)
        tt.manual_seed(0)
        net = _build(tt, [2, 2, 2, 2], [64, 128, 256, 512]).cuda()
        net.train()
        logp = net(tt.tensor(x).cuda())
        loss = tt.nn.NLLLoss()(logp, tt.tensor(lab, dtype=np.int64).cuda())
        loss.backward()
        grads = {k: p.grad.get() for k, p in net.named_parameters()}
        assert all(np.isfinite(v).all() for v in grads.values())
        key = mode if mode not in results else mode + "_again"
        results[key] = (loss.item(), grads)
    l_tf32, g_tf32 = results["tf32"]
    l_fp32, g_fp32 = results["fp32"]
    assert abs(l_tf32 - np.log(10)) < 0.5
    assert abs(l_tf32 - l_fp32) < 2e-3 * abs(l_fp32)
    worst, worst_l2, worst_name = 0.0, 0.0, ""
    for k in g_fp32:
        denom = max(float(np.abs(g_fp32[k]).max()), 1e-30)
        rel = float(np.abs(g_tf32[k] - g_fp32[k]).max()) / denom
        l2 = _rel_l2(g_tf32[k], g_fp32[k])
        print(f"  {k}: max-abs rel {rel:.3e}, rel-L2 {l2:.3e}")
        if rel > worst:
            worst, worst_name = rel, k
        worst_l2 = max(worst_l2, l2)
    print(f"preact_resnet18 N=256: worst grad rel-err tf32 vs fp32 path = {worst:.3e} ({worst_name}), worst rel-L2 {worst_l2:.3e}")
    assert worst_l2 < 0.25 and worst < 0.5
    l2, g2 = results["tf32_again"]
    assert l2 == l_tf32 and all(np.array_equal(g2[k], g_tf32[k]) for k in g2), "step is not deterministic"


def test_residual_gradient_is_folded_into_bn_backward():
    """x feeds BatchNorm->ReLU->conv AND an identity shortcut: the shortcut gradient that is already pending for x is
    added inside the BatchNorm dx pass (Tensor._sweep / ttb_bn_bwd_apply accum) - same numbers as the separate add."""
    import pytortto_b200 as tt
    from pytortto_b200 import ops
    tt.set_math_mode("fp32")
    rng = np.random.default_rng(5)
    xn = rng.standard_normal((4, 32, 8, 8)).astype(np.float32)
    results = []
    for fold in (True, False):
        tt.manual_seed(11)
        bn = tt.nn.Sequential(tt.nn.BatchNorm2d(32), tt.nn.ReLU()).cuda()
        conv = tt.nn.Conv2d(32, 32, 3, padding=1, bias=False).cuda()
        tt.autograd.grad_nn._BatchNormBase._accumulates_input0 = fold
        calls0 = ops._cabi.launch_count
        leaf = tt.tensor(xn, requires_grad=True)
        x = leaf.cuda() * 1.5          # non-leaf x with two consumers
        out = conv(bn(x)) + x
        (out * out).sum().backward()
        results.append((np.asarray(leaf.grad), ops._cabi.launch_count - calls0))
    tt.autograd.grad_nn._BatchNormBase._accumulates_input0 = True
    np.testing.assert_allclose(results[0][0], results[1][0], rtol=1e-6, atol=1e-6)
    assert results[0][1] < results[1][1], "folding must save the separate add launch"


def test_preact_step_bf16_loss_and_first_step_gradients():
    """bf16 mode end to end (shadows co-written by BatchNorm, weights re-packed once per step) on the golden step
    fixture of the REAL reference: loss / log-probs within the north-star bf16 tolerance (1e-2) at both steps."""
    import pytortto_b200 as tt
    tt.set_math_mode("bf16")
    g = load_golden("preact_step.npz")
    tt.manual_seed(7)
    net = _build(tt, [1, 1, 1, 1], [32, 32, 64, 64]).cuda()
    crit = tt.nn.NLLLoss()
    opt = tt.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    net.train()
    for step in range(2):
        opt.zero_grad()
        logp = net(tt.tensor(g[f"step{step}/x"]).cuda())
        loss = crit(logp, tt.tensor(g[f"step{step}/labels"], dtype=np.int64).cuda())
        loss.backward()
        ref = float(g[f"step{step}/loss"])
        assert abs(loss.item() - ref) <= 1e-2 * (1 if step == 0 else 5) * max(1.0, abs(ref)), (step, loss.item(), ref)
        assert all(np.isfinite(p.grad.get()).all() for p in net.parameters())
        opt.step()
    tt.set_math_mode("tf32")

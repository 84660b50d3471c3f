"""GPU: the ops between the conv stacks and the loss as own kernels (csrc/head.cu) - global average pool, the fused
Linear node, LogSoftmax, NLL loss (reductions, ignore_index), BCE-with-logits, NHWC channel cat / split, bias add -
against the numpy oracle (pinned to the live reference by tests/test_oracle_vs_reference.py).  fp32 CUDA-core
arithmetic: tolerance 2e-5 of the tensor max."""
import numpy as np
import pytest

from gpu_util import assert_close, require_gpu
from oracle import tortto_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu():
    require_gpu()


def _tt():
    import pytortto_b200 as tt
    tt.set_math_mode("tf32")
    return tt


def test_classifier_head_chain_vs_oracle():
    """mean over (H, W) -> flatten -> Linear -> LogSoftmax -> NLLLoss(mean), forward and every gradient; no library
    kernel is launched on the way (every call is a C-ABI launch)."""
    tt = _tt()
    from pytortto_b200 import _cabi
    rng = np.random.default_rng(51)
    n, c, h, w, k = 37, 96, 4, 4, 10
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    wt = (rng.standard_normal((k, c)) / 10).astype(np.float32)
    b = rng.standard_normal(k).astype(np.float32)
    lab = rng.integers(0, k, n).astype(np.int64)
    pooled = x.mean(axis=(-1, -2), dtype=np.float32).reshape(n, c)
    logits = O.linear_forward(pooled, wt, b)
    logp = O.log_softmax_forward(logits)
    loss = O.nll_loss_forward(logp, lab)
    dlogits = O.log_softmax_backward(O.nll_loss_backward(np.float32(1.0), logp, lab), logp)
    dpool, dw, db = O.linear_backward(dlogits, pooled, wt)
    dx = np.broadcast_to((dpool / (h * w)).reshape(n, c, 1, 1), x.shape)
    xin = tt.nn.Parameter(tt.tensor(x).cuda())
    fc = tt.nn.Linear(c, k)
    fc.weight.data[...] = wt
    fc.bias.data[...] = b
    fc.cuda()
    n0 = _cabi.launch_count
    feat = tt.flatten(tt.mean(xin, (-1, -2), True), 1)
    out = tt.nn.LogSoftmax(dim=-1)(fc(feat))
    l = tt.nn.NLLLoss()(out, tt.tensor(lab, dtype=np.int64).cuda())
    l.backward()
    assert _cabi.launch_count - n0 >= 9  # mean, linear, lsm, nll fwd + nll, lsm, 3 x linear, mean bwd
    assert_close("logp", out.data.get(), logp, 2e-5)
    assert abs(l.item() - float(loss)) < 2e-5 * max(1.0, abs(float(loss)))
    assert_close("dx", xin.grad.get(), dx, 2e-5)
    assert_close("dW", fc.weight.grad.get(), dw, 2e-5)
    assert_close("db", fc.bias.grad.get(), db, 2e-5)


@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
@pytest.mark.parametrize("ignore_index", [-100, 3])
def test_nll_loss_reductions(reduction, ignore_index):
    tt = _tt()
    rng = np.random.default_rng(52)
    lp = O.log_softmax_forward(rng.standard_normal((300, 7)).astype(np.float32))
    tg = rng.integers(0, 7, 300).astype(np.int64)
    yo, n = O.nll_loss_forward_ex(lp, tg, ignore_index, reduction)
    g = rng.standard_normal(np.shape(yo)).astype(np.float32)
    dxo = O.nll_loss_backward_ex(g, lp, tg, ignore_index, reduction, n)
    lt = tt.nn.Parameter(tt.tensor(lp).cuda())
    y = tt.nn.functional.nll_loss(lt, tt.tensor(tg, dtype=np.int64).cuda(), ignore_index=ignore_index, reduction=reduction)
    y.backward(tt.tensor(g).cuda())
    assert_close("nll", np.asarray(y.data.get()), np.asarray(yo), 2e-6)
    assert_close("nll dx", lt.grad.get(), dxo, 2e-6)


@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_bce_with_logits(reduction):
    tt = _tt()
    rng = np.random.default_rng(53)
    x = (rng.standard_normal((5, 1, 33, 17)) * 4).astype(np.float32)
    t = (rng.random(x.shape) < 0.5).astype(np.float32)
    yo = O.bce_with_logits_forward(x, t, reduction)
    g = rng.standard_normal(np.shape(yo)).astype(np.float32)
    dxo = O.bce_with_logits_backward(g, x, t, reduction)
    xt = tt.nn.Parameter(tt.tensor(x).cuda())
    y = tt.nn.functional.binary_cross_entropy_with_logits(xt, tt.tensor(t).cuda(), reduction=reduction)
    y.backward(tt.tensor(g).cuda())
    assert_close("bce", np.asarray(y.data.get()), np.asarray(yo), 5e-6)
    assert_close("bce dx", xt.grad.get(), dxo, 5e-6)


def test_channel_cat_and_split_nhwc():
    """tt.cat(dim=1) of NHWC activations (UNet skip connections, grad_fcn.py:881-904) incl. channel counts that are not a
    multiple of 4 (scalar path) and an input that needs no gradient."""
    tt = _tt()
    rng = np.random.default_rng(54)
    for cs in ([32, 64], [3, 5, 8], [64, 64]):
        arrs = [rng.standard_normal((3, c, 9, 7)).astype(np.float32) for c in cs]
        ts = [tt.tensor(a, requires_grad=(i != 1)).cuda() for i, a in enumerate(arrs)]
        leaves = [tt.nn.Parameter(t) if i != 1 else t for i, t in enumerate(ts)]
        y = tt.cat(leaves, dim=1)
        want = np.concatenate(arrs, axis=1)
        np.testing.assert_array_equal(y.data.get(), want)
        g = rng.standard_normal(want.shape).astype(np.float32)
        y.backward(tt.tensor(g).cuda())
        off = 0
        for i, c in enumerate(cs):
            if i != 1:
                np.testing.assert_array_equal(leaves[i].grad.get(), g[:, off:off + c])
            off += c


def test_conv_transpose_bias_add_kernel():
    tt = _tt()
    rng = np.random.default_rng(55)
    x = rng.standard_normal((2, 64, 6, 6)).astype(np.float32)
    m = tt.nn.ConvTranspose2d(64, 32, kernel_size=2, stride=2).cuda()
    y = m(tt.tensor(x).cuda())
    yo = O.conv_transpose2d_forward(x, m.weight.data.get(), m.bias.data.get(), 2, 0, 0, 1, 1)
    assert_close("convT + bias", y.data.get(), yo, 2e-3)

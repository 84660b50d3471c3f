"""CPU: pins the numpy oracle (oracle/tortto_oracle.py) to fixtures produced by the real reference
(oracle/make_golden.py).  Tolerance: fp32 BLAS summation order differs, so 2e-5 relative to the tensor max
(the reference's own known-answer check uses atol=rtol=1e-5 / rtol 1e-3,
examples/conv2d_result_speed_comparison.ipynb:120-123)."""
import numpy as np
import pytest

from conftest import cases_of, load_golden, rel_err
from oracle import tortto_oracle as O

TOL = 2e-5


def _conv_cases():
    return cases_of(load_golden("conv2d.npz"))


@pytest.mark.parametrize("name", _conv_cases())
def test_conv2d(name):
    g = load_golden("conv2d.npz")
    n, ci, h, w, co, kh, kw, sh, sw, ph, pw, dh, dw, groups, bias = g[f"{name}/cfg"]
    x, wt, dy = g[f"{name}/x"], g[f"{name}/w"], g[f"{name}/dy"]
    b = g[f"{name}/b"] if bias else None
    y = O.conv2d_forward(x, wt, b, (sh, sw), (ph, pw), (dh, dw), int(groups))
    assert y.shape == g[f"{name}/y"].shape
    assert rel_err(y, g[f"{name}/y"]) < TOL
    dx, dwt, db = O.conv2d_backward(x, wt, dy, (sh, sw), (ph, pw), (dh, dw), int(groups), has_bias=bool(bias))
    assert rel_err(dx, g[f"{name}/dx"]) < TOL
    assert rel_err(dwt, g[f"{name}/dw"]) < TOL
    if bias:
        assert rel_err(db, g[f"{name}/db"]) < TOL


@pytest.mark.parametrize("name", cases_of(load_golden("conv_transpose2d.npz")))
def test_conv_transpose2d(name):
    g = load_golden("conv_transpose2d.npz")
    n, ci, h, w, co, kh, kw, sh, sw, ph, pw, oph, opw, dh, dw, groups, bias = g[f"{name}/cfg"]
    x, wt, dy = g[f"{name}/x"], g[f"{name}/w"], g[f"{name}/dy"]
    b = g[f"{name}/b"] if bias else None
    # nn.ConvTranspose2d hands F.conv_transpose2d `_single(output_padding)` == output_padding[:1]
    # (nn/modules/conv.py:129, utils.py:5-10), so the W output padding silently equals the H one.
    y = O.conv_transpose2d_forward(x, wt, b, (sh, sw), (ph, pw), (int(oph),), int(groups), (dh, dw))
    assert y.shape == g[f"{name}/y"].shape
    assert rel_err(y, g[f"{name}/y"]) < TOL
    dx, dwt, db = O.conv_transpose2d_backward(x, wt, dy, (sh, sw), (ph, pw), (dh, dw), int(groups), bool(bias))
    assert rel_err(dx, g[f"{name}/dx"]) < TOL
    assert rel_err(dwt, g[f"{name}/dw"]) < TOL
    if bias:
        assert rel_err(db, g[f"{name}/db"]) < TOL


@pytest.mark.parametrize("name", cases_of(load_golden("batch_norm.npz")))
def test_batch_norm(name):
    g = load_golden("batch_norm.npz")
    affine, track, mom, training, steps, eps = g[f"{name}/cfg"]
    gamma = g[f"{name}/gamma"] if affine else None
    beta = g[f"{name}/beta"] if affine else None
    rm = g[f"{name}/rm0"] if track else None
    rv = g[f"{name}/rv0"] if track else None
    nbt = 0
    for st in range(int(steps)):
        x, dy = g[f"{name}/x{st}"], g[f"{name}/dy{st}"]
        momentum = mom
        if training and track:
            nbt += 1
            momentum = 1.0 / nbt if mom < 0 else mom  # nn/modules/batchnorm.py:64-70
        y, rm, rv, saved = O.batch_norm_forward(x, gamma, beta, rm, rv, bool(training), momentum, eps)
        assert rel_err(y, g[f"{name}/y{st}"]) < TOL
        dx, dgamma, dbeta = O.batch_norm_backward(dy, x, gamma, saved, (True, bool(affine), bool(affine)))
        assert rel_err(dx, g[f"{name}/dx{st}"]) < 5e-5
        if affine:
            assert rel_err(dgamma, g[f"{name}/dgamma{st}"]) < TOL
            assert rel_err(dbeta, g[f"{name}/dbeta{st}"]) < TOL
        if track:
            assert rel_err(rm, g[f"{name}/rm{st + 1}"]) < TOL
            assert rel_err(rv, g[f"{name}/rv{st + 1}"]) < TOL
            assert float(g[f"{name}/nbt{st + 1}"].reshape(-1)[0]) == (nbt if training else 0)


def test_batch_norm_large_mean():
    """|mean| / sd ~ 1e3: the oracle (two-pass, like the reference) stays as close to the float64 truth as the reference
    itself does (each tensor within max(2e-5, 3 x the reference's own distance))."""
    g = load_golden("batch_norm_large_mean.npz")
    c = g["x"].shape[1]
    y, rm, rv, saved = O.batch_norm_forward(g["x"], g["gamma"], g["beta"], np.zeros(c, np.float32), np.ones(c, np.float32),
                                            True, 0.1, float(g["eps"][0]))
    dx, dgamma, dbeta = O.batch_norm_backward(g["dy"], g["x"], g["gamma"], saved)
    for k, v in {"y": y, "dx": dx, "dgamma": dgamma, "dbeta": dbeta, "rv": rv, "rm": rm}.items():
        assert rel_err(v, g[k + "64"]) <= max(2e-5, 3.0 * rel_err(g[k], g[k + "64"])), k


def test_relu():
    g = load_golden("relu.npz")
    y = O.relu_forward(g["x"])
    np.testing.assert_array_equal(y, g["y"])  # includes the NaN element (np.maximum propagates NaN)
    np.testing.assert_array_equal(O.relu_backward(g["dy"], y), g["dx"])
    assert int(g["inplace_version_bump"][0]) == 1  # helper.py:10-16
    np.testing.assert_array_equal(O.relu_forward(g["x"] * 2.0), g["y_inplace"])
    np.testing.assert_array_equal(O.relu_backward(g["dy"], g["y_inplace"]) * 2.0, g["dx_inplace"])


@pytest.mark.parametrize("name", cases_of(load_golden("max_pool2d.npz")))
def test_max_pool2d(name):
    g = load_golden("max_pool2d.npz")
    kh, kw, sh, sw, ph, pw, dh, dw, ceil = g[f"{name}/cfg"]
    x, dy = g[f"{name}/x"], g[f"{name}/dy"]
    y, idx = O.max_pool2d_forward(x, (kh, kw), (sh, sw), (ph, pw), (dh, dw), bool(ceil))
    np.testing.assert_array_equal(y, g[f"{name}/y"])
    dx = O.max_pool2d_backward(dy, idx, x.shape, (kh, kw), (sh, sw), (ph, pw), (dh, dw), bool(ceil))
    np.testing.assert_array_equal(dx, g[f"{name}/dx"])  # bit-exact incl. last-writer-wins overlap semantics


def test_conv_error_messages():
    x = np.zeros((2, 3, 8, 8), np.float32)
    with pytest.raises(RuntimeError, match="expected input"):
        O.conv2d_forward(x, np.zeros((4, 2, 3, 3), np.float32))
    with pytest.raises(RuntimeError, match="Expected 3D"):
        O.conv2d_forward(x[0], np.zeros((4, 3, 3, 3), np.float32))
    with pytest.raises(RuntimeError, match="pad should be smaller"):
        O.max_pool2d_forward(x, 2, 2, 2)
    with pytest.raises(ValueError, match="more than 1 value"):
        O.batch_norm_forward(np.zeros((1, 3, 1, 1), np.float32), None, None, None, None, True, 0.1, 1e-5)


def test_preact_step():
    """Whole training step (2 SGD steps) of the reduced PreactResNet vs the real reference."""
    from oracle.resnet_oracle import StepOracle
    g = load_golden("preact_step.npz")
    names = [str(n) for n in g["param_names"]]
    params = {n: g[f"init/{n}"].copy() for n in names}
    net = StepOracle([1, 1, 1, 1], [32, 32, 64, 64], params)
    assert list(params) == names
    for step in range(2):
        loss, logp, grads = net.train_step(g[f"step{step}/x"], g[f"step{step}/labels"])
        assert abs(float(loss) - float(g[f"step{step}/loss"])) < 1e-5
        assert rel_err(logp, g[f"step{step}/logp"]) < 1e-5
        for n in names:
            assert rel_err(grads[n], g[f"step{step}/grad/{n}"]) < 2e-4, n
    for n in names:
        assert rel_err(net.params[n], g[f"step1/param/{n}"]) < 1e-5, n
    for k in g.files:
        if k.startswith("final/"):
            assert rel_err(net.buffers[k[len("final/"):]], g[k]) < 1e-5, k

"""GPU parity (through the public API -> autograd Functions -> C ABI) against the committed golden fixtures of the
REAL reference and against the numpy oracle on seeded random inputs.

Tolerances (BASELINE.json north_star): TF32 tensor path rel-err <= 2e-3 of the tensor max; the exact-fp32 path is
held to 2e-5 (fp32 summation-order noise only); ReLU and MaxPool are bit-exact."""
import numpy as np
import pytest

from conftest import cases_of, load_golden
from gpu_util import assert_close, report, require_gpu
from oracle import tortto_oracle as O

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-5, "tf32": 2e-3, "bf16": 1e-2}


@pytest.fixture(autouse=True)
def _gpu():
    require_gpu()


def _tt(mode):
    import pytortto_b200 as tt
    tt.set_math_mode(mode)
    return tt


def _run_conv(tt, x, w, b, dy, stride, padding, dilation, groups):
    conv_w = tt.nn.Parameter(tt.tensor(w).cuda())
    conv_b = None if b is None else tt.nn.Parameter(tt.tensor(b).cuda())
    xin = tt.nn.Parameter(tt.tensor(x).cuda())
    y = tt.nn.functional.conv2d(xin, conv_w, conv_b, stride, padding, dilation, groups)
    y.backward(tt.tensor(dy).cuda())
    return (y.data.get(), xin.grad.get(), conv_w.grad.get(), None if b is None else conv_b.grad.get())


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("name", cases_of(load_golden("conv2d.npz")))
def test_conv2d_golden(name, mode):
    tt = _tt(mode)
    g = load_golden("conv2d.npz")
    n, ci, h, w, co, kh, kw, sh, sw, ph, pw, dh, dw, groups, bias = [int(v) for v in g[f"{name}/cfg"]]
    b = g[f"{name}/b"] if bias else None
    y, dx, dwt, db = _run_conv(tt, g[f"{name}/x"], g[f"{name}/w"], b, g[f"{name}/dy"], (sh, sw), (ph, pw), (dh, dw), groups)
    tol = TOL[mode]
    assert_close(f"{name}[{mode}] y", y, g[f"{name}/y"], tol)
    assert_close(f"{name}[{mode}] dx", dx, g[f"{name}/dx"], tol)
    assert_close(f"{name}[{mode}] dw", dwt, g[f"{name}/dw"], tol)
    if bias:
        assert_close(f"{name}[{mode}] db", db, g[f"{name}/db"], 2e-5)


# tensor-core shaped problems (Cin, Cout multiples of 32) vs the oracle: (N, Cin, H, W, Cout, k, s, p, d)
TC_CASES = [
    (4, 32, 16, 16, 32, 3, 1, 1, 1),
    (4, 64, 16, 16, 64, 3, 1, 1, 1),
    (3, 64, 15, 17, 96, 3, 1, 1, 1),      # ragged M, Cout not a power of two
    (4, 64, 16, 16, 128, 3, 2, 1, 1),
    (4, 64, 16, 16, 128, 1, 2, 0, 1),
    (2, 128, 9, 9, 128, 3, 2, 1, 1),      # odd size, stride 2 (remainder rows)
    (2, 256, 8, 8, 256, 3, 1, 1, 1),
    (8, 512, 4, 4, 512, 3, 1, 1, 1),
    (2, 64, 12, 12, 64, 3, 1, 2, 2),      # dilation 2
    (2, 32, 10, 10, 64, 5, 1, 2, 1),      # 5x5
    (2, 96, 7, 7, 40, 1, 1, 0, 1),        # 1x1, odd channel counts (multiples of 8)
    (1, 32, 5, 5, 32, 3, 1, 1, 1),        # M < one tile
    (2, 64, 14, 14, 64, 3, 3, 1, 1),      # stride 3
]


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("case", TC_CASES, ids=[f"n{c[0]}_c{c[1]}_{c[2]}x{c[3]}_k{c[4]}_f{c[5]}s{c[6]}p{c[7]}d{c[8]}" for c in TC_CASES])
def test_conv2d_tensor_path_vs_oracle(case, mode):
    tt = _tt(mode)
    n, ci, h, w, co, k, s, p, d = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = (rng.standard_normal((co, ci, k, k)) / np.sqrt(ci * k * k)).astype(np.float32)
    yo = O.conv2d_forward(x, wt, None, s, p, d)
    dy = rng.standard_normal(yo.shape).astype(np.float32)
    dxo, dwo, _ = O.conv2d_backward(x, wt, dy, s, p, d)
    y, dx, dw, _ = _run_conv(tt, x, wt, None, dy, (s, s), (p, p), (d, d), 1)
    from pytortto_b200 import ops, _cabi
    import ctypes
    desc = ops.conv_desc(x.shape, wt.shape, (s, s), (p, p), (d, d), 1)
    used = [_cabi.load().ttb_conv2d_tensor_path_supported(ctypes.byref(desc), i) for i in range(3)]
    print("tensor path used (fprop, dgrad, wgrad):", used)
    # every pass pads its reduction channels to whole K-blocks (32 tf32 / 64 bf16) in a staged copy when needed, and
    # dgrad a narrow dx to 8 channels, so all three run on the tensor path
    assert used == [1, 1, 1], "this case is meant to exercise the tcgen05 path"
    tol = TOL[mode]
    assert_close("y", y, yo, tol)
    assert_close("dx", dx, dxo, tol)
    assert_close("dw", dw, dwo, tol)


# stride-1 layers with <= 128 filters: wgrad runs on the haloed-tile kernel (conv_wgrad_halo.cu) - ragged boxes, rows wider
# than a box, no padding, 2- and 4-wide filters, non-square filters, 1 / 2 channel slabs and 1 / 3 filter rows per CTA
# (N, Cin, H, W, Cout, kh, kw, ph, pw)
HALO_CASES = [
    (3, 64, 12, 20, 32, 3, 3, 1, 1),
    (2, 32, 9, 24, 64, 3, 2, 0, 0),
    (2, 64, 70, 70, 64, 3, 3, 1, 1),
    (2, 128, 16, 16, 128, 3, 3, 1, 1),
    (2, 64, 10, 16, 128, 5, 3, 2, 1),
    (2, 96, 9, 11, 64, 1, 4, 0, 2),
    (2, 192, 8, 16, 64, 3, 3, 1, 1),
    (5, 64, 33, 8, 64, 3, 3, 1, 1),
]


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("case", HALO_CASES, ids=[f"n{c[0]}_c{c[1]}_{c[2]}x{c[3]}_k{c[4]}_f{c[5]}x{c[6]}p{c[7]}{c[8]}" for c in HALO_CASES])
def test_wgrad_haloed_tile_vs_oracle(case, mode):
    tt = _tt(mode)
    n, ci, h, w, co, kh, kw, ph, pw = case
    if mode == "bf16" and (ci % 64 or co % 64):
        pytest.skip("bf16 operands need 64-channel blocks: this shape runs as TF32 (covered by the tf32 case)")
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = (rng.standard_normal((co, ci, kh, kw)) / np.sqrt(ci * kh * kw)).astype(np.float32)
    yo = O.conv2d_forward(x, wt, None, (1, 1), (ph, pw), (1, 1))
    dy = rng.standard_normal(yo.shape).astype(np.float32)
    dxo, dwo, _ = O.conv2d_backward(x, wt, dy, (1, 1), (ph, pw), (1, 1))
    y, dx, dw, _ = _run_conv(tt, x, wt, None, dy, (1, 1), (ph, pw), (1, 1), 1)
    tol = TOL[mode]
    assert_close("y", y, yo, tol)
    assert_close("dx", dx, dxo, tol)
    assert_close("dw", dw, dwo, tol)


# 1 x 1 convolutions with <= 4 filters (a segmentation head): HBM-streaming exact-fp32 kernels in every math mode
@pytest.mark.parametrize("case", [(3, 32, 17, 19, 1, True), (2, 64, 8, 8, 3, False), (2, 8, 5, 7, 4, True), (1, 128, 9, 9, 2, True)],
                         ids=lambda c: f"n{c[0]}_c{c[1]}_{c[2]}x{c[3]}_k{c[4]}")
def test_pointwise_conv_with_few_filters(case):
    tt = _tt("tf32")
    n, ci, h, w, co, bias = case
    rng = np.random.default_rng(11)
    x = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = (rng.standard_normal((co, ci, 1, 1)) / np.sqrt(ci)).astype(np.float32)
    b = rng.standard_normal(co).astype(np.float32) if bias else None
    yo = O.conv2d_forward(x, wt, b, 1, 0, 1)
    dy = rng.standard_normal(yo.shape).astype(np.float32)
    dxo, dwo, dbo = O.conv2d_backward(x, wt, dy, 1, 0, 1)
    y, dx, dw, db = _run_conv(tt, x, wt, b, dy, (1, 1), (0, 0), (1, 1), 1)
    assert_close("y", y, yo, 2e-5)
    assert_close("dx", dx, dxo, 2e-5)
    assert_close("dw", dw, dwo, 2e-5)
    if bias:
        assert_close("db", db, dy.sum((0, 2, 3)), 2e-5)


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
@pytest.mark.parametrize("name", cases_of(load_golden("conv_transpose2d.npz")))
def test_conv_transpose2d_golden(name, mode):
    tt = _tt(mode)
    g = load_golden("conv_transpose2d.npz")
    n, ci, h, w, co, kh, kw, sh, sw, ph, pw, oph, opw, dh, dw, groups, bias = [int(v) for v in g[f"{name}/cfg"]]
    m = tt.nn.ConvTranspose2d(ci, co, (kh, kw), stride=(sh, sw), padding=(ph, pw), output_padding=(oph, opw),
                              groups=groups, bias=bool(bias), dilation=(dh, dw))
    m.weight.data[...] = g[f"{name}/w"]
    if bias:
        m.bias.data[...] = g[f"{name}/b"]
    m.cuda()
    xin = tt.nn.Parameter(tt.tensor(g[f"{name}/x"]).cuda())
    y = m(xin)
    assert y.shape == g[f"{name}/y"].shape
    y.backward(tt.tensor(g[f"{name}/dy"]).cuda())
    tol = TOL[mode]
    assert_close(f"{name} y", y.data.get(), g[f"{name}/y"], tol)
    assert_close(f"{name} dx", xin.grad.get(), g[f"{name}/dx"], tol)
    assert_close(f"{name} dw", m.weight.grad.get(), g[f"{name}/dw"], tol)
    if bias:
        assert_close(f"{name} db", m.bias.grad.get(), g[f"{name}/db"], 2e-5)


@pytest.mark.parametrize("name", cases_of(load_golden("batch_norm.npz")))
def test_batch_norm_golden(name):
    tt = _tt("tf32")
    g = load_golden("batch_norm.npz")
    affine, track, mom, training, steps, eps = g[f"{name}/cfg"]
    c = g[f"{name}/x0"].shape[1]
    bn = tt.nn.BatchNorm2d(c, eps=float(eps), momentum=None if mom < 0 else float(mom), affine=bool(affine),
                           track_running_stats=bool(track))
    if affine:
        bn.weight.data[...] = g[f"{name}/gamma"]
        bn.bias.data[...] = g[f"{name}/beta"]
    if track:
        bn.running_mean.data[...] = g[f"{name}/rm0"]
        bn.running_var.data[...] = g[f"{name}/rv0"]
    bn.cuda()
    bn.train(bool(training))
    for st in range(int(steps)):
        xin = tt.nn.Parameter(tt.tensor(g[f"{name}/x{st}"]).cuda())
        if affine:
            bn.weight.grad = None
            bn.bias.grad = None
        y = bn(xin)
        y.backward(tt.tensor(g[f"{name}/dy{st}"]).cuda())
        assert_close(f"{name} y{st}", y.data.get(), g[f"{name}/y{st}"], 2e-5)
        assert_close(f"{name} dx{st}", xin.grad.get(), g[f"{name}/dx{st}"], 1e-4)
        if affine:
            assert_close(f"{name} dgamma{st}", bn.weight.grad.get(), g[f"{name}/dgamma{st}"], 2e-5)
            assert_close(f"{name} dbeta{st}", bn.bias.grad.get(), g[f"{name}/dbeta{st}"], 2e-5)
        if track:
            sd = bn.state_dict()
            assert_close(f"{name} rm{st + 1}", sd["running_mean"], g[f"{name}/rm{st + 1}"], 2e-5)
            assert_close(f"{name} rv{st + 1}", sd["running_var"], g[f"{name}/rv{st + 1}"], 2e-5)
            assert float(sd["num_batches_tracked"]) == float(g[f"{name}/nbt{st + 1}"].reshape(-1)[0])


def test_batch_norm_large_mean_fixture():
    """|mean| / sd ~ 1e3 (tests/golden/batch_norm_large_mean.npz, made by the REAL reference + a float64 evaluation of
    the same formulas).  In this regime two fp32 implementations cannot agree to 2e-5 - the reference's own y is
    4.3e-5 from the float64 truth, its dgamma 7.6e-5 - so each tensor is held to max(2e-5, 3 x the reference's own
    distance from the truth).  A one-pass fp32 sum(x^2) variance (round 1) fails this by orders of magnitude."""
    tt = _tt("tf32")
    g = load_golden("batch_norm_large_mean.npz")
    c = g["x"].shape[1]
    bn = tt.nn.BatchNorm2d(c, eps=float(g["eps"][0]))
    bn.weight.data[...] = g["gamma"]
    bn.bias.data[...] = g["beta"]
    bn.cuda().train()
    xin = tt.nn.Parameter(tt.tensor(g["x"]).cuda())
    y = bn(xin)
    y.backward(tt.tensor(g["dy"]).cuda())
    sd = bn.state_dict()
    got = {"y": y.data.get(), "dx": xin.grad.get(), "dgamma": bn.weight.grad.get(), "dbeta": bn.bias.grad.get(),
           "rv": sd["running_var"], "rm": sd["running_mean"]}
    for k, v in got.items():
        truth = g[k + "64"]
        _, ours = report(f"large-mean {k} vs float64", v, truth)
        _, ref = report(f"large-mean {k} reference vs float64", g[k], truth)
        bound = max(2e-5, 3.0 * ref)
        print(f"large-mean BN {k}: ours {ours:.3e}  reference {ref:.3e}  bound {bound:.3e}")
        assert ours <= bound, (k, ours, ref)


def test_batch_norm_vs_oracle_large():
    """resnet-sized BN (C=64, 256x32x32 would be 67 MB; use N=32) incl. non-multiple-of-4 channel fallback."""
    tt = _tt("tf32")
    rng = np.random.default_rng(11)
    for shape in [(32, 64, 32, 32), (5, 7, 9, 11), (16, 512, 4, 4)]:
        x = (rng.standard_normal(shape) * 2 + 0.5).astype(np.float32)
        dy = rng.standard_normal(shape).astype(np.float32)
        c = shape[1]
        gamma = rng.standard_normal(c).astype(np.float32)
        beta = rng.standard_normal(c).astype(np.float32)
        yo, rm, rv, saved = O.batch_norm_forward(x, gamma, beta, np.zeros(c, np.float32), np.ones(c, np.float32), True, 0.1, 1e-5)
        dxo, dgo, dbo = O.batch_norm_backward(dy, x, gamma, saved)
        bn = tt.nn.BatchNorm2d(c)
        bn.weight.data[...] = gamma
        bn.bias.data[...] = beta
        bn.cuda()
        xin = tt.nn.Parameter(tt.tensor(x).cuda())
        y = bn(xin)
        y.backward(tt.tensor(dy).cuda())
        assert_close(f"bn{shape} y", y.data.get(), yo, 2e-5)
        assert_close(f"bn{shape} dx", xin.grad.get(), dxo, 1e-4)
        assert_close(f"bn{shape} dgamma", bn.weight.grad.get(), dgo, 5e-5)
        assert_close(f"bn{shape} dbeta", bn.bias.grad.get(), dbo, 5e-5)
        assert_close(f"bn{shape} running_var", bn.state_dict()["running_var"], rv, 2e-5)


def test_stem_conv_runs_on_tensor_path_via_channel_padding():
    """Cin = 3 (network stems) is zero-padded to 32 channels for fprop + wgrad, and dgrad writes an 8-channel dx that
    is cropped to 3: all three passes run on tcgen05."""
    tt = _tt("tf32")
    import ctypes
    from pytortto_b200 import _cabi, ops
    rng = np.random.default_rng(21)
    x = rng.standard_normal((8, 3, 32, 32)).astype(np.float32)
    wt = (rng.standard_normal((64, 3, 3, 3)) / np.sqrt(27)).astype(np.float32)
    yo = O.conv2d_forward(x, wt, None, 1, 1, 1)
    dy = rng.standard_normal(yo.shape).astype(np.float32)
    dxo, dwo, _ = O.conv2d_backward(x, wt, dy, 1, 1, 1)
    desc = ops.conv_desc(x.shape, wt.shape, (1, 1), (1, 1), (1, 1), 1)
    assert [_cabi.load().ttb_conv2d_tensor_path_supported(ctypes.byref(desc), i) for i in range(3)] == [1, 1, 1]
    y, dx, dw, _ = _run_conv(tt, x, wt, None, dy, (1, 1), (1, 1), (1, 1), 1)
    assert_close("stem y", y, yo, 2e-3)
    assert_close("stem dx", dx, dxo, 2e-3)
    assert_close("stem dw", dw, dwo, 2e-3)


def test_fused_bn_relu_equals_separate_nodes():
    """nn.Sequential(BatchNorm2d, ReLU) lowers to one fused node; values and gradients must equal the two-node form."""
    tt = _tt("tf32")
    rng = np.random.default_rng(31)
    x = rng.standard_normal((6, 32, 9, 9)).astype(np.float32)
    dy = rng.standard_normal(x.shape).astype(np.float32)
    outs = []
    for fuse in (True, False):
        tt.nn.set_bn_relu_fusion(fuse)
        seq = tt.nn.Sequential(tt.nn.BatchNorm2d(32), tt.nn.ReLU())
        seq[0].weight.data[...] = np.linspace(0.5, 1.5, 32, dtype=np.float32)
        seq[0].bias.data[...] = np.linspace(-0.3, 0.3, 32, dtype=np.float32)
        seq.cuda()
        xin = tt.nn.Parameter(tt.tensor(x).cuda())
        y = seq(xin)
        assert (y.grad_fn.__class__.__name__ == "BatchNormReluBackward") == fuse
        y.backward(tt.tensor(dy).cuda())
        outs.append((y.data.get(), xin.grad.get(), seq[0].weight.grad.get(), seq[0].bias.grad.get(),
                     seq[0].state_dict()["running_var"]))
    tt.nn.set_bn_relu_fusion(True)
    for a, b in zip(*outs):
        np.testing.assert_array_equal(a, b)


def test_relu_golden():
    tt = _tt("tf32")
    g = load_golden("relu.npz")
    xin = tt.nn.Parameter(tt.tensor(g["x"]).cuda())
    y = tt.nn.functional.relu(xin)
    y.backward(tt.tensor(g["dy"]).cuda())
    np.testing.assert_array_equal(y.data.get(), g["y"])
    np.testing.assert_array_equal(xin.grad.get(), g["dx"])
    # in-place on a non-leaf bumps the version counter (helper.py:10-16) and still back-propagates
    xin2 = tt.nn.Parameter(tt.tensor(g["x"]).cuda())
    h = xin2 * 2.0
    v0 = h._version
    y2 = tt.nn.functional.relu(h, inplace=True)
    assert h._version - v0 == int(g["inplace_version_bump"][0]) and y2 is h
    y2.backward(tt.tensor(g["dy"]).cuda())
    np.testing.assert_array_equal(y2.data.get(), g["y_inplace"])
    np.testing.assert_array_equal(xin2.grad.get(), g["dx_inplace"])
    with pytest.raises(RuntimeError, match="leaf Variable"):
        tt.nn.functional.relu(xin2, inplace=True)


@pytest.mark.parametrize("name", cases_of(load_golden("max_pool2d.npz")))
def test_max_pool2d_golden(name):
    tt = _tt("tf32")
    g = load_golden("max_pool2d.npz")
    kh, kw, sh, sw, ph, pw, dh, dw, ceil = [int(v) for v in g[f"{name}/cfg"]]
    xin = tt.nn.Parameter(tt.tensor(g[f"{name}/x"]).cuda())
    y = tt.nn.functional.max_pool2d(xin, (kh, kw), (sh, sw), (ph, pw), (dh, dw), bool(ceil))
    y.backward(tt.tensor(g[f"{name}/dy"]).cuda())
    np.testing.assert_array_equal(y.data.get(), g[f"{name}/y"])
    np.testing.assert_array_equal(xin.grad.get(), g[f"{name}/dx"])  # incl. last-writer-wins overlap semantics


def test_max_pool2d_large_vs_oracle():
    tt = _tt("tf32")
    rng = np.random.default_rng(5)
    x = rng.standard_normal((4, 64, 56, 56)).astype(np.float32)
    yo, idx = O.max_pool2d_forward(x, 3, 2, 1)
    dy = rng.standard_normal(yo.shape).astype(np.float32)
    dxo = O.max_pool2d_backward(dy, idx, x.shape, 3, 2, 1)
    xin = tt.nn.Parameter(tt.tensor(x).cuda())
    y = tt.nn.MaxPool2d(3, 2, 1)(xin)
    y.backward(tt.tensor(dy).cuda())
    np.testing.assert_array_equal(y.data.get(), yo)
    np.testing.assert_array_equal(xin.grad.get(), dxo)


def test_error_messages_match_reference():
    tt = _tt("tf32")
    x = tt.tensor(np.zeros((2, 3, 8, 8), np.float32)).cuda()
    with pytest.raises(RuntimeError, match="expected input"):
        tt.nn.functional.conv2d(x, tt.tensor(np.zeros((4, 2, 3, 3), np.float32)).cuda())
    with pytest.raises(RuntimeError, match="should be the same"):
        tt.nn.functional.conv2d(x, tt.tensor(np.zeros((4, 3, 3, 3), np.float32)))
    with pytest.raises(RuntimeError, match="pad should be smaller"):
        tt.nn.functional.max_pool2d(x, (2, 2), (2, 2), (2, 2))
    with pytest.raises(ValueError, match="more than 1 value"):
        tt.nn.BatchNorm2d(3).cuda()(tt.tensor(np.zeros((1, 3, 1, 1), np.float32)).cuda())
    with pytest.raises(ValueError, match="expected 4D input"):
        tt.nn.BatchNorm2d(3).cuda()(tt.tensor(np.zeros((3, 3), np.float32)).cuda())


def test_multi_tensor_backward_helpers_match_per_layer_calls():
    """ttb_conv2d_dgrad_pack_weights + ttb_conv2d_dgrad_prepacked and ttb_conv2d_wgrad_partial + ttb_sum_splits_multi
    (one launch for several layers) give bit-identical results to the per-layer ttb_conv2d_dgrad / ttb_conv2d_wgrad."""
    tt = _tt("tf32")
    import ctypes
    import torch
    from pytortto_b200 import _cabi, ops
    from pytortto_b200.xparray import cparray, current_stream_ptr
    rng = np.random.default_rng(77)
    layers = []
    for (n, c, h, k, ks, s) in [(8, 64, 16, 64, 3, 1), (8, 32, 16, 64, 3, 2), (4, 128, 8, 96, 1, 1)]:
        x = cparray.from_numpy(rng.standard_normal((n, c, h, h)).astype(np.float32))
        w = cparray.from_numpy((rng.standard_normal((k, c, ks, ks)) * 0.1).astype(np.float32))
        d = ops.conv_desc(x.shape, w.shape, (s, s), (ks // 2, ks // 2), (1, 1), 1)
        dy = cparray.from_numpy(rng.standard_normal((n, k, d.p, d.q)).astype(np.float32))
        layers.append((x, w, d, dy))
    lib = _cabi.load()
    st = current_stream_ptr()
    n = len(layers)
    assert all(lib.ttb_conv2d_dgrad_prepacked_supported(ctypes.byref(d)) for _, _, d, _ in layers)
    packed = [torch.empty(w.t.numel(), dtype=torch.float32, device="cuda") for _, w, _, _ in layers]
    descs = (ctypes.POINTER(_cabi.ConvDesc) * n)(*[ctypes.pointer(d) for _, _, d, _ in layers])
    src = (ctypes.c_void_p * n)(*[w.t.data_ptr() for _, w, _, _ in layers])
    dst = (ctypes.c_void_p * n)(*[p.data_ptr() for p in packed])
    _cabi.call("ttb_conv2d_dgrad_pack_weights", n, descs, src, dst, st)
    sums, keep = [], []
    for (x, w, d, dy), pk in zip(layers, packed):
        dx_ref = ops.conv2d_dgrad(dy, w, d).get()
        dx = cparray(torch.empty_like(x.t))
        _cabi.call("ttb_conv2d_dgrad_prepacked", ctypes.byref(d), dy.t.data_ptr(), pk.data_ptr(), None, dx.t.data_ptr(), st)
        np.testing.assert_array_equal(dx.get(), dx_ref)
        dw_ref = ops.conv2d_wgrad(x, dy, d).get()
        dw = cparray(torch.zeros_like(w.t))
        nb = lib.ttb_conv2d_workspace_size(ctypes.byref(d), 2)
        ws = torch.empty(max(nb, 1), dtype=torch.uint8, device="cuda")
        splits, partials = ctypes.c_int(0), ctypes.c_void_p(0)
        _cabi.call("ttb_conv2d_wgrad_partial", ctypes.byref(d), x.t.data_ptr(), dy.t.data_ptr(), dw.t.data_ptr(), ws.data_ptr(), nb,
                   ctypes.byref(splits), ctypes.byref(partials), st)
        keep.append((ws, dw, dw_ref))
        if splits.value > 1:
            sums.append((partials.value, splits.value, dw.size, dw.t.data_ptr()))
    assert sums, "at least one layer is expected to split its pixel range"
    m = len(sums)
    _cabi.call("ttb_sum_splits_multi", m, (ctypes.c_void_p * m)(*[t[0] for t in sums]), (ctypes.c_int * m)(*[t[1] for t in sums]),
               (ctypes.c_int64 * m)(*[t[2] for t in sums]), (ctypes.c_void_p * m)(*[t[3] for t in sums]), st)
    for ws, dw, dw_ref in keep:
        np.testing.assert_array_equal(dw.get(), dw_ref)


@pytest.mark.parametrize("cls_name,decoupled,amsgrad", [("Adam", False, False), ("Adam", False, True), ("AdamW", True, False)])
def test_fused_adam_vs_oracle(cls_name, decoupled, amsgrad):
    """ttb_adam_step_multi (one launch for all tensors, step count on the device) against the oracle's restatement of
    optim/_functional.py:25-115 (pinned to the live reference by tests/test_oracle_vs_reference.py), 4 steps."""
    tt = _tt("tf32")
    rng = np.random.default_rng(41)
    shapes = [(8, 4, 3, 3), (7,), (5, 6), (1030,)]
    p0 = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    params = [tt.nn.Parameter(tt.tensor(p.copy()).cuda()) for p in p0]
    opt = getattr(tt.optim, cls_name)(params, lr=1e-2, weight_decay=1e-2, amsgrad=amsgrad)
    op = [p.copy() for p in p0]
    m = [np.zeros_like(p) for p in p0]
    v = [np.zeros_like(p) for p in p0]
    vm = [np.zeros_like(p) for p in p0] if amsgrad else None
    for step in range(1, 5):
        gs = [rng.standard_normal(s).astype(np.float32) for s in shapes]
        for p, g in zip(params, gs):
            p.grad = tt.tensor(g).cuda().data
        opt.step()
        O.adam_step(op, gs, m, v, [step] * len(op), lr=1e-2, weight_decay=1e-2, max_exp_avg_sqs=vm, decoupled=decoupled)
        for a, b in zip(op, params):
            assert_close(f"{cls_name} amsgrad={amsgrad} step {step}", b.data.get(), a, 2e-6)
    assert opt.state[params[0]]['step'] == 4


def test_non_leaf_conv_weight_gradient_is_ordered():
    """conv2d(x, w * 0.5): dW goes to Mul.backward on the main stream, not to a leaf - the wgrad kernel must not be forked
    to the side stream.  Same numbers as the leaf-weight run scaled by 0.5, bit for bit, every repetition."""
    tt = _tt("tf32")
    rng = np.random.default_rng(43)
    x = rng.standard_normal((64, 64, 32, 32)).astype(np.float32)
    w = (rng.standard_normal((64, 64, 3, 3)) / 24).astype(np.float32)
    xd = tt.tensor(x).cuda()
    wl = tt.nn.Parameter(tt.tensor(w * 0.5).cuda())
    y = tt.nn.functional.conv2d(xd, wl, None, (1, 1), (1, 1), (1, 1), 1)
    dy = tt.tensor(rng.standard_normal(y.shape).astype(np.float32)).cuda()
    y.backward(dy)
    want = wl.grad.get() * 0.5
    for _ in range(5):
        wp = tt.nn.Parameter(tt.tensor(w).cuda())
        y2 = tt.nn.functional.conv2d(xd, wp * 0.5, None, (1, 1), (1, 1), (1, 1), 1)
        y2.backward(dy)
        np.testing.assert_array_equal(wp.grad.get(), want)


def test_bf16_weight_copies_follow_weight_updates():
    """bf16 mode keeps bf16 copies of the conv weights ([K][R][S][C] and [C][R][S][K], one multi-tensor re-pack per
    weight change): an optimizer step, an in-place array op and copy_ must all be seen by the next forward / backward."""
    tt = _tt("bf16")
    rng = np.random.default_rng(44)
    x = rng.standard_normal((4, 64, 8, 8)).astype(np.float32)
    conv = tt.nn.Conv2d(64, 64, 3, padding=1, bias=False).cuda()
    xin = tt.nn.Parameter(tt.tensor(x).cuda())

    def run():
        xin.grad = None
        conv.weight.grad = None
        y = conv(xin)
        (y * y).sum().backward()
        return y.data.get(), xin.grad.get()

    y0, dx0 = run()
    conv.weight.data *= 2.0                       # cparray in-place op
    y1, dx1 = run()
    np.testing.assert_allclose(y1, 2 * y0, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(dx1, 4 * dx0, rtol=1e-5, atol=1e-5)
    opt = tt.optim.SGD(conv.parameters(), lr=1.0)
    conv.weight.grad = conv.weight.data.copy()    # p <- p - 1.0 * p = 0
    opt.step()
    y2, _ = run()
    assert float(np.abs(y2).max()) == 0.0
    conv.weight.copy_(tt.tensor(np.asarray(rng.standard_normal((64, 64, 3, 3)) / 24, np.float32)))
    y3, _ = run()
    want = O.conv2d_forward(x, conv.weight.data.get(), None, 1, 1, 1)
    assert_close("bf16 after copy_", y3, want, 1e-2)


def test_bf16_shadow_path_matches_staged_conversion():
    """BatchNorm+ReLU co-writes the bf16 copy the next conv reads: same numbers as converting the fp32 output (round to
    nearest even both ways), so conv(bn_relu(x)) in bf16 mode is bit-identical with and without the shadow."""
    tt = _tt("bf16")
    rng = np.random.default_rng(45)
    x = rng.standard_normal((8, 64, 16, 16)).astype(np.float32)
    tt.manual_seed(3)
    bn = tt.nn.Sequential(tt.nn.BatchNorm2d(64), tt.nn.ReLU()).cuda()
    conv = tt.nn.Conv2d(64, 128, 3, padding=1, bias=False).cuda()
    xin = tt.tensor(x).cuda()
    a = bn(xin)
    assert a.data._h is not None, "bf16 mode: BatchNorm+ReLU must co-write the bf16 shadow"
    y_shadow = conv(a).data.get()
    a.data._h = None                              # force the conversion kernel
    y_conv = conv(a).data.get()
    np.testing.assert_array_equal(y_shadow, y_conv)
    yo = O.conv2d_forward(a.data.get(), conv.weight.data.get(), None, 1, 1, 1)
    assert_close("bf16 conv after BN+ReLU", y_shadow, yo, 1e-2)


# grouped convolutions on the tensor path: (N, Cin, H, W, Cout, k, stride, pad, groups)
#   aligned groups (Cin/g a multiple of 32): the TMA maps read each group's channel slice in place;
#   <= 4 channels per group: tap-packed per group (the reference notebook's known-answer layer is groups = 2, Cin/g = 2)
GROUPED_CASES = [
    (4, 128, 16, 16, 128, 3, 1, 1, 2),
    (4, 128, 16, 16, 128, 3, 2, 1, 4),
    (2, 64, 9, 11, 96, 3, 1, 1, 2),
    (3, 192, 8, 8, 192, 1, 1, 0, 6),      # more groups than one multi-problem launch holds
    (8, 4, 20, 20, 16, 3, 1, 1, 2),       # 2 channels per group: tap-packed fprop / wgrad, exact dgrad
    (4, 8, 12, 12, 32, 3, 2, 1, 2),       # 4 channels per group, stride 2
]


@pytest.mark.parametrize("case", GROUPED_CASES, ids=[f"n{c[0]}_c{c[1]}_{c[2]}x{c[3]}_k{c[4]}_f{c[5]}s{c[6]}g{c[8]}" for c in GROUPED_CASES])
def test_grouped_conv2d_tensor_path_vs_oracle(case):
    import ctypes
    tt = _tt("tf32")
    from pytortto_b200 import ops, _cabi
    n, ci, h, w, co, k, s, p, g = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = (rng.standard_normal((co, ci // g, k, k)) / np.sqrt(ci // g * k * k)).astype(np.float32)
    b = rng.standard_normal(co).astype(np.float32)
    yo = O.conv2d_forward(x, wt, b, s, p, 1, g)
    dy = rng.standard_normal(yo.shape).astype(np.float32)
    dxo, dwo, dbo = O.conv2d_backward(x, wt, dy, s, p, 1, g, has_bias=True)
    desc = ops.conv_desc(x.shape, wt.shape, (s, s), (p, p), (1, 1), g)
    used = [_cabi.load().ttb_conv2d_tensor_path_supported(ctypes.byref(desc), i) for i in range(3)]
    print("tensor path used (fprop, dgrad, wgrad):", used)
    assert used[0] == 1 and used[2] == 1, "fprop and wgrad of this grouped layer are meant to run on the tcgen05 path"
    assert used[1] == (1 if (co // g) % 32 == 0 and (ci // g) % 8 == 0 else 0)
    y, dx, dw, db = _run_conv(tt, x, wt, b, dy, (s, s), (p, p), (1, 1), g)
    assert_close("y", y, yo, 2e-3)
    assert_close("dx", dx, dxo, 2e-3)
    assert_close("dw", dw, dwo, 2e-3)
    assert_close("db", db, dbo, 2e-5)


# <= 4 input channels (the network stems): tap-packed staging - R*S*C taps of a pixel in one row, ONE K-block for 3x3x3
STEM_CASES = [  # (N, Cin, H, W, Cout, kh, kw, stride, pad)
    (16, 3, 32, 32, 64, 3, 3, 1, 1),
    (4, 3, 56, 56, 64, 7, 7, 2, 3),      # the ImageNet stem (147 taps -> 160 columns)
    (2, 3, 33, 29, 32, 3, 3, 1, 1),      # UNet's first layer, ragged
    (3, 1, 28, 28, 16, 5, 5, 1, 2),
    (2, 4, 17, 17, 24, 3, 2, 2, 1),
    (2, 4, 30, 26, 16, 7, 5, 2, 3),      # tall filter, stride 2: ROW-packed staging (7 K-blocks of round_up(5*4) columns)
    (3, 3, 21, 40, 32, 7, 7, 2, 3),      # the stem again on a ragged, non-square image (odd H: the last filter rows hang over)
]


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("case", STEM_CASES, ids=[f"n{c[0]}_c{c[1]}_{c[2]}x{c[3]}_k{c[4]}_f{c[5]}x{c[6]}s{c[7]}" for c in STEM_CASES])
def test_stem_conv2d_tap_packed_vs_oracle(case, mode):
    tt = _tt(mode)
    n, ci, h, w, co, kh, kw, s, p = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = (rng.standard_normal((co, ci, kh, kw)) / np.sqrt(ci * kh * kw)).astype(np.float32)
    b = rng.standard_normal(co).astype(np.float32)
    yo = O.conv2d_forward(x, wt, b, s, p, 1)
    dy = rng.standard_normal(yo.shape).astype(np.float32)
    dxo, dwo, dbo = O.conv2d_backward(x, wt, dy, s, p, 1, 1, has_bias=True)
    y, dx, dw, db = _run_conv(tt, x, wt, b, dy, (s, s), (p, p), (1, 1), 1)
    tol = TOL["tf32"]  # (bf16 mode runs <= 4-channel layers on the TF32 tensor path)
    assert_close("y", y, yo, tol)
    assert_close("dx", dx, dxo, tol)
    assert_close("dw", dw, dwo, tol)
    assert_close("db", db, dbo, 2e-5)


@pytest.mark.parametrize("mode", ["tf32", "fp32"])
def test_empty_batch_flows_through_conv_relu_pool(mode):
    """A batch of zero images (the last, short batch of a data loader can be empty after filtering): shapes follow the
    reference's formulas (numpy on zero-size arrays), nothing is launched on zero elements, and the weight / bias
    gradients of the convolution are zeros of the parameter's shape."""
    tt = _tt(mode)
    np.random.seed(0)
    conv = tt.nn.Conv2d(8, 16, 3, 2, 1).cuda()
    x = tt.nn.Parameter(tt.tensor(np.zeros((0, 8, 12, 10), np.float32)).cuda())
    y = tt.nn.MaxPool2d(2, 2)(tt.nn.ReLU()(conv(x)))
    assert tuple(y.shape) == (0, 16, 3, 2)
    yo = O.conv2d_forward(np.zeros((0, 8, 12, 10), np.float32), conv.weight.data.get(), conv.bias.data.get(), 2, 1, 1)
    assert yo.shape == (0, 16, 6, 5)
    y.sum().backward()
    assert tuple(x.grad.shape) == (0, 8, 12, 10)
    np.testing.assert_array_equal(conv.weight.grad.get(), np.zeros((16, 8, 3, 3), np.float32))
    np.testing.assert_array_equal(conv.bias.grad.get(), np.zeros(16, np.float32))

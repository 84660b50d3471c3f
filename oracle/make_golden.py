"""Generate tests/golden/*.npz by running the REAL reference (tortto v1.3.4 numpy path, imported from
/root/reference via oracle/ref_import.py).  Run in this container only:

    python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  The fixtures are committed; the GPU box never needs /root/reference.
Every case stores its inputs (seeded `np.random.default_rng`) and the reference's outputs/gradients.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

tt = ref_import.import_reference()  # must come before anything that imports numpy submodules / torch
import numpy as np  # noqa: E402

nn = tt.nn
F = tt.nn.functional
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def f32(a):
    return np.array(np.asarray(a), dtype=np.float32, order='C', copy=True)


# ---------------------------------------------------------------------------------------------------------
# conv cases: (name, N, Cin, H, W, Cout, k, stride, padding, dilation, groups, bias)
# first entry = the reference's own known-answer configuration
# (examples/conv2d_result_speed_comparison.ipynb:21-24, 40-100) at a reduced batch
# ---------------------------------------------------------------------------------------------------------
CONV_CASES = [
    ("notebook_exotic", 2, 4, 64, 64, 16, (3, 2), (1, 3), (2, 3), (1, 2), 2, True),
    ("c3x3_s1_p1", 2, 8, 9, 11, 12, (3, 3), (1, 1), (1, 1), (1, 1), 1, False),
    ("c3x3_s2_p1", 2, 8, 10, 12, 16, (3, 3), (2, 2), (1, 1), (1, 1), 1, False),
    ("c3x3_s2_p1_odd", 2, 6, 9, 7, 4, (3, 3), (2, 2), (1, 1), (1, 1), 1, True),
    ("c1x1_s1", 3, 16, 5, 5, 8, (1, 1), (1, 1), (0, 0), (1, 1), 1, False),
    ("c1x1_s2", 3, 16, 8, 8, 32, (1, 1), (2, 2), (0, 0), (1, 1), 1, False),
    ("c7x7_s2_p3", 1, 3, 20, 20, 8, (7, 7), (2, 2), (3, 3), (1, 1), 1, False),
    ("c3x3_dil2", 2, 4, 12, 12, 6, (3, 3), (1, 1), (2, 2), (2, 2), 1, True),
    ("c3x3_groups4", 2, 8, 8, 8, 8, (3, 3), (1, 1), (1, 1), (1, 1), 4, True),
    ("depthwise_s2", 2, 6, 9, 9, 6, (3, 3), (2, 2), (1, 1), (1, 1), 6, False),
    ("c5x3_s32_p21", 2, 5, 13, 14, 7, (5, 3), (3, 2), (2, 1), (1, 1), 1, True),
    ("tc32_3x3_s1", 2, 32, 8, 8, 32, (3, 3), (1, 1), (1, 1), (1, 1), 1, False),
    ("tc64_3x3_s2", 2, 64, 8, 8, 128, (3, 3), (2, 2), (1, 1), (1, 1), 1, False),
    ("tc64_1x1_s2", 2, 64, 8, 8, 128, (1, 1), (2, 2), (0, 0), (1, 1), 1, False),
    ("batch1_1x1img", 1, 4, 1, 1, 4, (1, 1), (1, 1), (0, 0), (1, 1), 1, True),
]


def gen_conv():
    out = {}
    for seed, (name, n, ci, h, w, co, k, s, p, d, g, bias) in enumerate(CONV_CASES):
        rng = np.random.default_rng(100 + seed)
        x = f32(rng.standard_normal((n, ci, h, w)))
        wt = f32(rng.standard_normal((co, ci // g, *k)) * 0.2)
        b = f32(rng.standard_normal((co,))) if bias else None
        conv = nn.Conv2d(ci, co, k, stride=s, padding=p, dilation=d, groups=g, bias=bias)
        conv.weight.data[...] = wt
        if bias:
            conv.bias.data[...] = b
        xt = tt.tensor(x, requires_grad=True)
        y = conv(xt)
        dy = f32(rng.standard_normal(y.shape))
        y.backward(tt.tensor(dy))
        out[f"{name}/x"] = x
        out[f"{name}/w"] = wt
        if bias:
            out[f"{name}/b"] = b
            out[f"{name}/db"] = f32(conv.bias.grad)
        out[f"{name}/dy"] = dy
        out[f"{name}/y"] = f32(y.data)
        out[f"{name}/dx"] = f32(xt.grad)
        out[f"{name}/dw"] = f32(conv.weight.grad)
        out[f"{name}/cfg"] = np.array([n, ci, h, w, co, k[0], k[1], s[0], s[1], p[0], p[1], d[0], d[1], g, int(bias)])
    np.savez_compressed(os.path.join(OUT, "conv2d.npz"), **out)
    print("conv2d.npz", len(CONV_CASES), "cases")


# (name, N, Cin, H, W, Cout, k, stride, padding, output_padding, dilation, groups, bias)
CONVT_CASES = [
    ("unet_k2s2", 2, 8, 5, 6, 4, (2, 2), (2, 2), (0, 0), (0, 0), (1, 1), 1, True),
    ("k3s2p1op1", 2, 6, 5, 5, 4, (3, 3), (2, 2), (1, 1), (1, 1), (1, 1), 1, True),
    ("k3s1p1", 1, 4, 6, 7, 5, (3, 3), (1, 1), (1, 1), (0, 0), (1, 1), 1, False),
    ("k4s2p1_g2", 2, 4, 4, 4, 6, (4, 4), (2, 2), (1, 1), (0, 0), (1, 1), 2, False),
    ("k3s3_d2", 1, 3, 4, 5, 2, (3, 3), (3, 3), (0, 0), (1, 2), (2, 2), 1, True),
    ("tc64_k2s2", 2, 64, 4, 4, 32, (2, 2), (2, 2), (0, 0), (0, 0), (1, 1), 1, True),
]


def gen_convt():
    out = {}
    for seed, (name, n, ci, h, w, co, k, s, p, op, d, g, bias) in enumerate(CONVT_CASES):
        rng = np.random.default_rng(200 + seed)
        x = f32(rng.standard_normal((n, ci, h, w)))
        wt = f32(rng.standard_normal((ci, co // g, *k)) * 0.2)
        b = f32(rng.standard_normal((co,))) if bias else None
        m = nn.ConvTranspose2d(ci, co, k, stride=s, padding=p, output_padding=op, groups=g, bias=bias, dilation=d)
        m.weight.data[...] = wt
        if bias:
            m.bias.data[...] = b
        xt = tt.tensor(x, requires_grad=True)
        y = m(xt)
        dy = f32(rng.standard_normal(y.shape))
        y.backward(tt.tensor(dy))
        out[f"{name}/x"] = x
        out[f"{name}/w"] = wt
        if bias:
            out[f"{name}/b"] = b
            out[f"{name}/db"] = f32(m.bias.grad)
        out[f"{name}/dy"] = dy
        out[f"{name}/y"] = f32(y.data)
        out[f"{name}/dx"] = f32(xt.grad)
        out[f"{name}/dw"] = f32(m.weight.grad)
        out[f"{name}/cfg"] = np.array([n, ci, h, w, co, k[0], k[1], s[0], s[1], p[0], p[1], op[0], op[1],
                                       d[0], d[1], g, int(bias)])
    np.savez_compressed(os.path.join(OUT, "conv_transpose2d.npz"), **out)
    print("conv_transpose2d.npz", len(CONVT_CASES), "cases")


# (name, shape, affine, track, momentum, training, steps)
BN_CASES = [
    ("train_affine", (4, 6, 5, 7), True, True, 0.1, True, 2),
    ("train_noaffine", (3, 5, 4, 4), False, True, 0.1, True, 1),
    ("train_cumulative", (4, 3, 6, 6), True, True, None, True, 3),
    ("train_notrack", (2, 4, 3, 3), True, False, 0.1, True, 1),
    ("eval_running", (4, 6, 5, 7), True, True, 0.1, False, 1),
    ("train_c32", (8, 32, 8, 8), True, True, 0.1, True, 1),
    ("train_shifted", (16, 8, 8, 8), True, True, 0.1, True, 1),
]


def gen_bn():
    out = {}
    for seed, (name, shape, affine, track, mom, training, steps) in enumerate(BN_CASES):
        rng = np.random.default_rng(300 + seed)
        c = shape[1]
        bn = nn.BatchNorm2d(c, momentum=mom, affine=affine, track_running_stats=track)
        if affine:
            bn.weight.data[...] = f32(rng.standard_normal(c) * 0.5 + 1.0)
            bn.bias.data[...] = f32(rng.standard_normal(c) * 0.5)
        if track:
            bn.running_mean.data[...] = f32(rng.standard_normal(c) * 0.1)
            bn.running_var.data[...] = f32(rng.random(c) + 0.5)
            out[f"{name}/rm0"] = f32(bn.running_mean.data)
            out[f"{name}/rv0"] = f32(bn.running_var.data)
        if affine:
            out[f"{name}/gamma"] = f32(bn.weight.data)
            out[f"{name}/beta"] = f32(bn.bias.data)
        bn.train(training)
        for st in range(steps):
            x = f32(rng.standard_normal(shape) * (1.0 + 0.5 * st) + (3.0 if "shifted" in name else 0.3 * st))
            xt = tt.tensor(x, requires_grad=True)
            y = bn(xt)
            dy = f32(rng.standard_normal(shape))
            if affine:
                bn.weight.grad = None
                bn.bias.grad = None
            y.backward(tt.tensor(dy))
            out[f"{name}/x{st}"] = x
            out[f"{name}/dy{st}"] = dy
            out[f"{name}/y{st}"] = f32(y.data)
            out[f"{name}/dx{st}"] = f32(xt.grad)
            if affine:
                out[f"{name}/dgamma{st}"] = f32(bn.weight.grad)
                out[f"{name}/dbeta{st}"] = f32(bn.bias.grad)
            if track:
                out[f"{name}/rm{st + 1}"] = f32(bn.running_mean.data)
                out[f"{name}/rv{st + 1}"] = f32(bn.running_var.data)
                out[f"{name}/nbt{st + 1}"] = f32(bn.num_batches_tracked.data)
        out[f"{name}/cfg"] = np.array([int(affine), int(track), -1.0 if mom is None else mom, int(training), steps,
                                       bn.eps], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "batch_norm.npz"), **out)
    print("batch_norm.npz", len(BN_CASES), "cases")


def gen_bn_large_mean():
    """|mean| / sd ~ 1e3 per channel: the regime where one-pass sum(x^2) statistics in fp32 lose the
    variance.  Stores the REAL reference's fp32 results and a float64 evaluation of the same formulas
    (grad_nn.py:923-959, 977-988), so a test can hold an implementation to the reference's own distance from the truth."""
    out = {}
    rng = np.random.default_rng(777)
    shape = (16, 8, 16, 16)
    c = shape[1]
    offs = (1000.0 * (1.0 + 0.1 * np.arange(c)) * np.where(np.arange(c) % 2, -1.0, 1.0)).reshape(1, c, 1, 1)
    x = f32(rng.standard_normal(shape) + offs)
    dy = f32(rng.standard_normal(shape))
    gamma = f32(rng.standard_normal(c) * 0.5 + 1.0)
    beta = f32(rng.standard_normal(c) * 0.5)
    bn = nn.BatchNorm2d(c)
    bn.weight.data[...] = gamma
    bn.bias.data[...] = beta
    bn.train(True)
    xt = tt.tensor(x, requires_grad=True)
    y = bn(xt)
    y.backward(tt.tensor(dy))
    out.update(x=x, dy=dy, gamma=gamma, beta=beta, y=f32(y.data), dx=f32(xt.grad), dgamma=f32(bn.weight.grad),
               dbeta=f32(bn.bias.grad), rm=f32(bn.running_mean.data), rv=f32(bn.running_var.data))
    # float64 evaluation of the same definition
    x64, dy64, g64, b64 = x.astype(np.float64), dy.astype(np.float64), gamma.astype(np.float64), beta.astype(np.float64)
    n = x64.size // c
    mu = x64.mean((0, 2, 3), keepdims=True)
    var = x64.var((0, 2, 3), keepdims=True)
    sd = np.sqrt(var + bn.eps)
    xh = (x64 - mu) / sd
    out["y64"] = xh * g64.reshape(1, c, 1, 1) + b64.reshape(1, c, 1, 1)
    gg = dy64 * g64.reshape(1, c, 1, 1)
    out["dx64"] = (n * gg - gg.sum((0, 2, 3), keepdims=True) - xh * (gg * xh).sum((0, 2, 3), keepdims=True)) / n / sd
    out["dgamma64"] = (dy64 * xh).sum((0, 2, 3))
    out["dbeta64"] = dy64.sum((0, 2, 3))
    out["rv64"] = 0.9 * 1.0 + 0.1 * var.reshape(-1) * n / (n - 1)
    out["rm64"] = 0.1 * mu.reshape(-1)
    out["eps"] = np.array([bn.eps])
    np.savez_compressed(os.path.join(OUT, "batch_norm_large_mean.npz"), **out)
    print("batch_norm_large_mean.npz")


def gen_relu():
    out = {}
    rng = np.random.default_rng(400)
    x = f32(rng.standard_normal((3, 5, 6, 7)))
    x[0, 0, 0, :3] = [0.0, -0.0, np.nan]
    xt = tt.tensor(x, requires_grad=True)
    y = F.relu(xt)
    dy = f32(rng.standard_normal(x.shape))
    y.backward(tt.tensor(dy))
    out["x"], out["dy"], out["y"], out["dx"] = x, dy, f32(y.data), f32(xt.grad)
    # in-place variant on a non-leaf
    xt2 = tt.tensor(x, requires_grad=True)
    h = xt2 * 2.0
    v0 = h._version
    y2 = F.relu(h, inplace=True)
    out["inplace_version_bump"] = np.array([h._version - v0])
    y2.backward(tt.tensor(dy))
    out["y_inplace"], out["dx_inplace"] = f32(y2.data), f32(xt2.grad)
    np.savez_compressed(os.path.join(OUT, "relu.npz"), **out)
    print("relu.npz")


# (name, shape, kernel, stride, padding, dilation, ceil_mode)
POOL_CASES = [
    ("resnet_k3s2p1", (2, 4, 12, 12), (3, 3), (2, 2), (1, 1), (1, 1), False),
    ("unet_k2s2", (2, 3, 8, 10), (2, 2), (2, 2), (0, 0), (1, 1), False),
    ("k3s1_overlap", (1, 2, 7, 7), (3, 3), (1, 1), (1, 1), (1, 1), False),
    ("k3s2_ceil", (2, 3, 10, 11), (3, 3), (2, 2), (0, 0), (1, 1), True),
    ("k2s2_ceil_odd", (1, 2, 7, 9), (2, 2), (2, 2), (0, 0), (1, 1), True),
    ("k3s2p1_ceil", (1, 2, 8, 8), (3, 3), (2, 2), (1, 1), (1, 1), True),
    ("k2_dil2", (1, 3, 9, 9), (2, 2), (1, 1), (0, 0), (2, 2), False),
    ("k32_s21", (2, 2, 9, 8), (3, 2), (2, 1), (1, 0), (1, 1), False),
    ("ties_k3s2p1", (1, 2, 9, 9), (3, 3), (2, 2), (1, 1), (1, 1), False),
]


def gen_pool():
    out = {}
    for seed, (name, shape, k, s, p, d, ceil) in enumerate(POOL_CASES):
        rng = np.random.default_rng(500 + seed)
        if name.startswith("ties"):
            x = f32(rng.integers(0, 3, shape))  # many equal values: exercises first-max tie-break + overwrite order
        else:
            x = f32(rng.standard_normal(shape))
        xt = tt.tensor(x, requires_grad=True)
        y = F.max_pool2d(xt, k, s, p, d, ceil)
        dy = f32(rng.standard_normal(y.shape))
        y.backward(tt.tensor(dy))
        out[f"{name}/x"], out[f"{name}/dy"] = x, dy
        out[f"{name}/y"], out[f"{name}/dx"] = f32(y.data), f32(xt.grad)
        out[f"{name}/cfg"] = np.array([k[0], k[1], s[0], s[1], p[0], p[1], d[0], d[1], int(ceil)])
    np.savez_compressed(os.path.join(OUT, "max_pool2d.npz"), **out)
    print("max_pool2d.npz", len(POOL_CASES), "cases")


# ---------------------------------------------------------------------------------------------------------
# whole training step: PreactResNet (examples/resnet/preact_resnet18/preact_resnet18.ipynb cell 10), reduced
# depth/width so the fixture stays small: layers [1,1,1,1], channels [32,32,64,64], batch 8, one SGD step
# (lr=0.1, momentum=0.9, weight_decay=1e-4 as in cell 11) repeated twice so momentum buffers are exercised.
# ---------------------------------------------------------------------------------------------------------
def conv3x3(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=3, stride=stride, padding=1, bias=False)


def conv1x1(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=1, stride=stride, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_channels, channels, stride=1, downsample=None):
        super().__init__()
        self.act = nn.Sequential(nn.BatchNorm2d(in_channels), nn.ReLU())
        self.residual = nn.Sequential(conv3x3(in_channels, channels, stride), nn.BatchNorm2d(channels), nn.ReLU(),
                                      conv3x3(channels, channels))
        self.downsample = nn.Sequential() if downsample is None else downsample

    def forward(self, x):
        out = self.act(x)
        shortcut = self.downsample(x)
        out = self.residual(out)
        return out + shortcut


class PreactResNet(nn.Module):
    def __init__(self, layers, channels, num_classes=10):
        super().__init__()
        self.in_channels = channels[0]
        self.conv1 = conv3x3(3, self.in_channels)
        self.layer1 = self._make_layer(channels[0], layers[0])
        self.layer2 = self._make_layer(channels[1], layers[1], 2)
        self.layer3 = self._make_layer(channels[2], layers[2], 2)
        self.layer4 = self._make_layer(channels[3], layers[3], 2)
        self.bn = nn.BatchNorm2d(self.in_channels)
        self.relu = nn.ReLU()
        self.fc = nn.Sequential(nn.Linear(self.in_channels, num_classes), nn.LogSoftmax(dim=-1))

    def _make_layer(self, channels, blocks, stride=1):
        downsample = None
        if stride != 1 or self.in_channels != channels:
            downsample = nn.Sequential(conv1x1(self.in_channels, channels, stride))
        layers = [BasicBlock(self.in_channels, channels, stride, downsample)]
        self.in_channels = channels
        for _ in range(1, blocks):
            layers.append(BasicBlock(self.in_channels, channels))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.conv1(x)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = self.relu(self.bn(x))
        x = tt.mean(x, (-1, -2), True)
        x = tt.flatten(x, 1)
        return self.fc(x)


def gen_step():
    out = {}
    tt.manual_seed(7)
    net = PreactResNet([1, 1, 1, 1], [32, 32, 64, 64])
    rng = np.random.default_rng(7)
    names = [k for k, _ in net.named_parameters()]
    for k, p in net.named_parameters():
        out[f"init/{k}"] = f32(p.data)
    crit = nn.NLLLoss()
    opt = tt.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    net.train()
    for step in range(2):
        x = f32(rng.standard_normal((8, 3, 16, 16)))
        lab = rng.integers(0, 10, 8).astype(np.int64)
        opt.zero_grad()
        logp = net(tt.tensor(x))
        loss = crit(logp, tt.tensor(lab, dtype=np.int64))
        loss.backward()
        out[f"step{step}/x"], out[f"step{step}/labels"] = x, lab
        out[f"step{step}/logp"] = f32(logp.data)
        out[f"step{step}/loss"] = f32(loss.data)
        for k, p in net.named_parameters():
            out[f"step{step}/grad/{k}"] = f32(p.grad)
        opt.step()
        if step == 1:
            for k, p in net.named_parameters():
                out[f"step{step}/param/{k}"] = f32(p.data)
    sd = net.state_dict()
    for k, v in sd.items():
        if "running" in k or "num_batches" in k:
            out[f"final/{k}"] = f32(v)
    out["param_names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "preact_step.npz"), **out)
    print("preact_step.npz", len(names), "params")


def _example_models():
    """the SAME model source the B200 package uses (pytortto_b200/examples.py), instantiated on the reference"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ttb_examples", os.path.join(os.path.dirname(HERE), "pytortto_b200",
                                                                              "examples.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.make_models(tt)


def gen_models():
    """Whole-model forward/backward of the other BASELINE configurations at reduced size, from the real reference:
    UNet (conv+bias, BN, ReLU, MaxPool2d(2,2), ConvTranspose2d(k2,s2), cat, BCEWithLogits - config 4) and a
    bottleneck ResNet with the 7x7/s2 stem and the overlapping MaxPool2d(3,2,1) (config 3)."""
    M = _example_models()
    out = {}
    # ---- UNet ----
    tt.manual_seed(21)
    net = M["UNet"](3, 1, [32, 64])
    rng = np.random.default_rng(21)
    x = f32(rng.standard_normal((2, 3, 16, 16)))
    target = f32(rng.integers(0, 2, (2, 1, 16, 16)))
    net.train()
    logits = net(tt.tensor(x))
    loss = nn.BCEWithLogitsLoss()(logits, tt.tensor(target))
    loss.backward()
    out["unet/x"], out["unet/target"] = x, target
    out["unet/logits"], out["unet/loss"] = f32(logits.data), f32(loss.data)
    out["unet/param_names"] = np.array([k for k, _ in net.named_parameters()])
    for k, p in net.named_parameters():
        out[f"unet/grad/{k}"] = f32(p.grad)
    # ---- bottleneck ResNet with stem + overlapping max-pool ----
    tt.manual_seed(22)
    net = M["ResNet"](M["Bottleneck"], [1, 1, 1, 1], [8, 8, 16, 16])
    x = f32(rng.standard_normal((4, 3, 32, 32)))
    lab = rng.integers(0, 10, 4).astype(np.int64)
    net.train()
    logp = net(tt.tensor(x))
    loss = nn.NLLLoss()(logp, tt.tensor(lab, dtype=np.int64))
    loss.backward()
    out["resnet/x"], out["resnet/labels"] = x, lab
    out["resnet/logp"], out["resnet/loss"] = f32(logp.data), f32(loss.data)
    out["resnet/param_names"] = np.array([k for k, _ in net.named_parameters()])
    for k, p in net.named_parameters():
        out[f"resnet/grad/{k}"] = f32(p.grad)
    np.savez_compressed(os.path.join(OUT, "models.npz"), **out)
    print("models.npz")


if __name__ == "__main__":
    if "--only-bn-large-mean" in sys.argv:
        gen_bn_large_mean()
        sys.exit(0)
    gen_bn_large_mean()
    gen_models()
    gen_conv()
    gen_convt()
    gen_bn()
    gen_relu()
    gen_pool()
    gen_step()

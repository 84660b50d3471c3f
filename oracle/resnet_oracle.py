"""CPU oracle for one whole PreactResNet training step, composed from oracle/tortto_oracle.py.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).

Model = the reference's example network, examples/resnet/preact_resnet18/preact_resnet18.ipynb cell 10
(`PreactResNet(BasicBlock, layers, channels)`: conv3x3 stem, 4 stages of pre-activation BasicBlocks
[BN-ReLU-conv3x3-BN-ReLU-conv3x3 + (1x1 strided conv | identity) shortcut], BN-ReLU, global mean, Linear,
LogSoftmax) trained with NLLLoss + SGD(momentum, weight_decay) (cell 11).  Parameter names follow the
reference's `named_parameters()` (nn/modules/module.py:108-131) so fixtures can be loaded by name.
Pinned by tests/test_oracle_golden.py::test_preact_step against tests/golden/preact_step.npz.
"""
import math

import numpy as np

from . import tortto_oracle as O


def param_shapes(layers, channels, num_classes=10, in_ch=3):
    """Ordered {name: shape} exactly as the reference's named_parameters() enumerates them."""
    shapes = {"conv1.weight": (channels[0], in_ch, 3, 3)}
    cin = channels[0]
    for li, (nblk, ch) in enumerate(zip(layers, channels), start=1):
        for b in range(nblk):
            stride = 2 if (li > 1 and b == 0) else 1
            p = f"layer{li}.{b}"
            shapes[f"{p}.act.0.weight"] = (cin,)
            shapes[f"{p}.act.0.bias"] = (cin,)
            shapes[f"{p}.residual.0.weight"] = (ch, cin, 3, 3)
            shapes[f"{p}.residual.1.weight"] = (ch,)
            shapes[f"{p}.residual.1.bias"] = (ch,)
            shapes[f"{p}.residual.3.weight"] = (ch, ch, 3, 3)
            if stride != 1 or cin != ch:
                shapes[f"{p}.downsample.0.weight"] = (ch, cin, 1, 1)
            cin = ch
    shapes["bn.weight"] = (cin,)
    shapes["bn.bias"] = (cin,)
    shapes["fc.0.weight"] = (num_classes, cin)
    shapes["fc.0.bias"] = (num_classes,)
    return shapes


def init_params(layers, channels, num_classes=10, seed=0):
    """kaiming_uniform_(a=sqrt(5)) like nn/modules/conv.py:50-56 / linear.py (bound = 1/sqrt(fan_in)); BN gamma=1,
    beta=0 (batchnorm.py:43-47).  Uses its own RNG stream (not bit-identical to the reference's np.random one)."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, shp in param_shapes(layers, channels, num_classes).items():
        if len(shp) == 1 and not name.startswith("fc"):
            params[name] = np.ones(shp, np.float32) if name.endswith("weight") else np.zeros(shp, np.float32)
        else:
            fan_in = int(np.prod(shp[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            params[name] = rng.uniform(-bound, bound, shp).astype(np.float32)
    return params


def init_buffers(params):
    bufs = {}
    for name, p in params.items():
        if p.ndim == 1 and name.endswith(".weight") and not name.startswith("fc"):
            base = name[:-len("weight")]
            bufs[base + "running_mean"] = np.zeros_like(p)
            bufs[base + "running_var"] = np.ones_like(p)
            bufs[base + "num_batches_tracked"] = np.zeros((), np.float32)
    return bufs


class StepOracle:
    """forward + backward + SGD of the network on numpy arrays (NCHW float32)."""

    def __init__(self, layers, channels, params, buffers=None, momentum_bn=0.1, eps=1e-5, dtype=np.float32):
        # dtype: float32 = the reference's arithmetic; float64 (params / inputs given as float64 too) = the same
        # formulas evaluated in double, used by tests as the "truth" both fp32 implementations are measured against
        self.dtype = dtype
        self.layers, self.channels = layers, channels
        self.params = params
        self.buffers = init_buffers(params) if buffers is None else buffers
        self.momentum_bn, self.eps = momentum_bn, eps
        self.sgd_bufs = {k: None for k in params}

    # -- building blocks that record what backward needs on self.tape -----------------------------------
    def _bn_relu(self, x, prefix, tape):
        g, b = self.params[prefix + "weight"], self.params[prefix + "bias"]
        self.buffers[prefix + "num_batches_tracked"] = self.buffers[prefix + "num_batches_tracked"] + 1
        y, rm, rv, saved = O.batch_norm_forward(x, g, b, self.buffers[prefix + "running_mean"],
                                                self.buffers[prefix + "running_var"], True, self.momentum_bn,
                                                self.eps)
        self.buffers[prefix + "running_mean"], self.buffers[prefix + "running_var"] = rm, rv
        a = O.relu_forward(y)

        def back(da, grads):
            dy = O.relu_backward(da, a)
            dx, dg, db = O.batch_norm_backward(dy, x, g, saved)
            grads[prefix + "weight"] = dg
            grads[prefix + "bias"] = db
            return dx
        tape.append(back)
        return a

    def _conv(self, x, name, stride, pad, tape, need_dx=True):
        w = self.params[name]
        y = O.conv2d_forward(x, w, None, stride, pad)

        def back(dy, grads):
            dx, dw, _ = O.conv2d_backward(x, w, dy, stride, pad, need_dx=need_dx)
            grads[name] = dw
            return dx
        tape.append(back)
        return y

    def forward_backward(self, x, labels):
        grads = {}
        t_stem = []
        h = self._conv(x, "conv1.weight", 1, 1, t_stem, need_dx=False)
        blocks = []  # (tape_act, tape_residual, tape_shortcut)
        cin = self.channels[0]
        for li, (nblk, ch) in enumerate(zip(self.layers, self.channels), start=1):
            for b in range(nblk):
                stride = 2 if (li > 1 and b == 0) else 1
                p = f"layer{li}.{b}"
                t_act, t_res, t_sc = [], [], []
                a = self._bn_relu(h, f"{p}.act.0.", t_act)
                if stride != 1 or cin != ch:
                    sc = self._conv(h, f"{p}.downsample.0.weight", stride, 0, t_sc)
                else:
                    sc = h
                r = self._conv(a, f"{p}.residual.0.weight", stride, 1, t_res)
                r = self._bn_relu(r, f"{p}.residual.1.", t_res)
                r = self._conv(r, f"{p}.residual.3.weight", 1, 1, t_res)
                h = r + sc
                blocks.append((t_act, t_res, t_sc))
                cin = ch
        t_head = []
        a = self._bn_relu(h, "bn.", t_head)
        pooled = a.mean(axis=(-1, -2), keepdims=True, dtype=self.dtype)  # tt.mean(x,(-1,-2),True)
        flat = pooled.reshape(pooled.shape[0], -1)
        w, bfc = self.params["fc.0.weight"], self.params["fc.0.bias"]
        logits = O.linear_forward(flat, w, bfc)
        logp = O.log_softmax_forward(logits)
        loss = O.nll_loss_forward(logp, labels)

        # ---- backward ----
        dlogp = O.nll_loss_backward(self.dtype(1.0), logp, labels)
        dlogits = O.log_softmax_backward(dlogp, logp)
        dflat, dw, db = O.linear_backward(dlogits, flat, w)
        grads["fc.0.weight"], grads["fc.0.bias"] = dw.astype(self.dtype), db.astype(self.dtype)
        hw = a.shape[-1] * a.shape[-2]
        da = np.broadcast_to(dflat.reshape(pooled.shape) / self.dtype(hw), a.shape).astype(self.dtype)
        dh = t_head[0](da, grads)
        for t_act, t_res, t_sc in reversed(blocks):
            d = dh
            for fn in reversed(t_res):
                d = fn(d, grads)
            d_in = t_act[0](d, grads)
            d_sc = t_sc[0](dh, grads) if t_sc else dh
            dh = d_in + d_sc
        t_stem[0](dh, grads)
        return loss, logp, grads

    def sgd(self, grads, lr, momentum=0.9, weight_decay=1e-4):
        names = list(self.params)
        plist = [self.params[k] for k in names]
        glist = [grads[k] for k in names]
        blist = [self.sgd_bufs[k] for k in names]
        O.sgd_step(plist, glist, blist, lr, momentum, weight_decay)
        for k, b in zip(names, blist):
            self.sgd_bufs[k] = b

    def train_step(self, x, labels, lr=0.1, momentum=0.9, weight_decay=1e-4):
        loss, logp, grads = self.forward_backward(x, labels)
        self.sgd(grads, lr, momentum, weight_decay)
        return loss, logp, grads

"""Callers of the hot path inside a training step: ToCopy (device transfer), Add,
Mul, Sum, Mean, View, Transpose, Mm, Exp, Cat, LogSoftmax, NllLoss, BinaryCrossEntropyWithLogits.

These are NOT the accelerated path: they keep the reference's Function names and gradient formulas
(/root/reference/src/tortto/autograd/grad_fcn.py, grad_nn.py:287-392) and run as small device-array expressions
(activation-sized adds go through ttb_add).  They exist so that an unchanged tortto training script runs end to end
on the device array.
"""
import numpy as np
import torch

from .. import ops
from ..xparray import cparray
from .function import Function, grad_slot
from .grad_mode import is_grad_enabled
from .helper import build_links


def _scalar_like(x, value):
    return cparray(torch.full((), float(value), dtype=x.t.dtype, device=x.t.device))


class ToCopy(Function):
    """reference grad_fcn.py:849-878: .cuda() / .cpu(); the gradient travels the opposite way."""
    _keep_grad_slots = True

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, = inputs
        xd0 = xt0.data
        target = params['target_device']
        if target == 'cuda':
            yd0 = xd0 if xd0.__class__ is cparray else cparray.from_numpy(xd0)
        elif target == 'cpu':
            yd0 = xd0.get() if xd0.__class__ is cparray else xd0
        else:
            raise RuntimeError(f"unknown device {target}")
        if yd0 is xd0:
            return xt0
        return build_links(yd0, grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        if ctx.params['target_device'] == 'cuda':
            return gd0.get() if gd0.__class__ is cparray else gd0
        return cparray.from_numpy(gd0) if gd0.__class__ is not cparray else gd0


def _unbroadcast(g, shape):
    """sum a broadcast gradient back to `shape` (reference helper.reverse_broadcast, helper.py:32-42)."""
    if tuple(g.shape) == tuple(shape):
        return g
    t = g.t
    lead = t.dim() - len(shape)
    if lead > 0:
        t = t.sum(dim=tuple(range(lead)))
    dims = tuple(i for i, s in enumerate(shape) if s == 1 and t.shape[i] != 1)
    if dims:
        t = t.sum(dim=dims, keepdim=True)
    return cparray(t)


class Add(Function):
    _absorbs = ("conv", "bn")

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, xt1 = inputs
        xd0, xd1 = xt0.data, xt1.data
        if xd0.__class__ is not cparray or xd1.__class__ is not cparray:
            raise RuntimeError("add: both operands must be on the CUDA device")
        # (training forward of 4-D operands: the same pass emits the statistics of the sum for the BatchNorm that follows)
        want_stats = is_grad_enabled() and xd0.ndim == 4
        deferred = ops.resolve_pending(Add, (xt0, xt1))
        if deferred is not None:
            other = xd1 if deferred is xd0 else xd0
            if deferred._thunk.job.kind == "conv":  # conv + shortcut (+ statistics of the sum) is ONE kernel
                yd0 = ops.conv_add_fused(deferred, other, want_stats)
            else:  # a BatchNorm normalise pass: the sum stays deferred - a ReLU may follow (post-activation blocks)
                yd0 = ops.bn_add_deferred(deferred, other)
        else:
            yd0 = ops.add_arrays(xd0, xd1, stats=want_stats)
        ctx.params['shapes'] = (xd0.shape, xd1.shape)
        return build_links(yd0, grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        s0, s1 = ctx.params['shapes']
        g0 = _unbroadcast(gd0, s0) if ctx.needs_input_grad[0] else None
        g1 = _unbroadcast(gd0, s1) if ctx.needs_input_grad[1] else None
        return g0, g1


class Mul(Function):
    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, xt1 = inputs
        yd0 = cparray(xt0.data.t * xt1.data.t)
        ctx.save_for_backward(xt0, xt1)
        return build_links(yd0, grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        xd0, xd1 = ctx.saved_tensors
        g0 = _unbroadcast(cparray(gd0.t * xd1.t), xd0.shape) if ctx.needs_input_grad[0] else None
        g1 = _unbroadcast(cparray(gd0.t * xd0.t), xd1.shape) if ctx.needs_input_grad[1] else None
        return g0, g1


def _norm_dims(dim, ndim):
    if dim is None:
        return tuple(range(ndim))
    if isinstance(dim, int):
        dim = (dim,)
    return tuple(sorted(d % ndim for d in dim))


class Sum(Function):
    """reference grad_fcn.py:1033-1056"""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, = inputs
        xd0 = xt0.data
        dims = _norm_dims(params['dim'], xd0.ndim)
        yd0 = cparray(xd0.t.sum(dim=dims, keepdim=params['keepdim']))
        ctx.params['shape'] = xd0.shape
        ctx.params['dims'] = dims
        return build_links(yd0, grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        shape, dims = ctx.params['shape'], ctx.params['dims']
        g = gd0.t
        if not ctx.params['keepdim']:
            for d in dims:
                g = g.unsqueeze(d)
        return cparray(g.expand(shape).contiguous())


class Mean(Function):
    """reference grad_fcn.py:1058-1093 (global average pooling in the ResNets: tt.mean(x, (-1,-2), True))"""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, = inputs
        xd0 = xt0.data
        dims = _norm_dims(params['dim'], xd0.ndim)
        ctx.params['shape'] = xd0.shape
        ctx.params['dims'] = dims
        ctx.params['hw_kernel'] = xd0.ndim == 4 and dims == (2, 3) and xd0.t.dtype == torch.float32 and xd0.size > 0
        if ctx.params['hw_kernel']:  # global average pooling of an NHWC activation: one kernel
            yd0 = ops.mean_hw(xd0, params['keepdim'])
        else:
            yd0 = cparray(xd0.t.mean(dim=dims, keepdim=params['keepdim']))
        return build_links(yd0, grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        shape, dims = ctx.params['shape'], ctx.params['dims']
        if ctx.params['hw_kernel'] and gd0.t.dtype == torch.float32:
            return ops.mean_hw_bwd(gd0, shape)
        count = 1
        for d in dims:
            count *= shape[d]
        g = gd0.t
        if not ctx.params['keepdim']:
            for d in dims:
                g = g.unsqueeze(d)
        out = torch.empty(shape, dtype=g.dtype, device=g.device,
                          memory_format=torch.channels_last if len(shape) == 4 else torch.contiguous_format)
        torch.div(g.expand(shape), count, out=out)
        return cparray(out)


class View(Function):
    """reshape / flatten (reference grad_fcn.py View): logical NCHW element order, like the reference."""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, = inputs
        xd0 = xt0.data
        ctx.params['in_shape'] = xd0.shape
        return build_links(xd0.reshape(params['shape']), grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        return gd0.reshape(ctx.params['in_shape'])


class Transpose(Function):
    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, = inputs
        yd0 = cparray(xt0.data.t.transpose(params['dim0'], params['dim1']).contiguous())
        return build_links(yd0, grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        return cparray(gd0.t.transpose(ctx.params['dim0'], ctx.params['dim1']).contiguous())


class Mm(Function):
    """matmul of 2-D operands (the classifier head: x @ W.T, nn/functional.py:54-63).  fp32 (no TF32)."""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, xt1 = inputs
        yd0 = cparray(torch.matmul(xt0.data.t, xt1.data.t))
        ctx.save_for_backward(xt0, xt1)
        return build_links(yd0, grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        xd0, xd1 = ctx.saved_tensors
        g0 = cparray(torch.matmul(gd0.t, xd1.t.transpose(-1, -2))) if ctx.needs_input_grad[0] else None
        g1 = cparray(torch.matmul(xd0.t.transpose(-1, -2), gd0.t)) if ctx.needs_input_grad[1] else None
        return g0, g1


class Linear(Function):
    """y = x @ W^T (+ b) as ONE node (the reference builds Transpose + Mm + Add, nn/functional.py:54-63): forward and
    each of the three gradients is one fp32 CUDA-core GEMM / column-sum kernel (ttb_matmul, ttb_bias_grad).
    inputs = (x (M, K), weight (N, K), bias (N,) | None)."""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, xt1, xt2 = inputs
        xd0, xd1 = xt0.data, xt1.data
        m, k = xd0.shape
        n = xd1.shape[0]
        if xd1.shape[1] != k:
            raise RuntimeError(f'mat1 and mat2 shapes cannot be multiplied ({m}x{k} and {xd1.shape[1]}x{n})')
        yd0 = ops.matmul(xd0, xd1, None if xt2 is None else xt2.data, m, n, k, k, 1, 1, k)
        ctx.save_for_backward(xt0, xt1)
        return build_links(yd0, grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        xd0, xd1 = ctx.saved_tensors
        m, k = xd0.shape
        n = xd1.shape[0]
        g0 = ops.matmul(gd0, xd1, None, m, k, n, n, 1, k, 1) if ctx.needs_input_grad[0] else None      # g @ W
        g1 = ops.matmul(gd0, xd0, None, n, k, m, 1, n, k, 1, out=grad_slot(ctx, 1)) if ctx.needs_input_grad[1] else None  # g^T @ x
        g2 = ops.colsum(gd0, out=grad_slot(ctx, 2)) if ctx.needs_input_grad[2] else None
        return g0, g1, g2


class Exp(Function):
    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, = inputs
        yt0 = build_links(cparray(torch.exp(xt0.data.t)), grad_fn=ctx)
        ctx.save_for_backward(yt0)
        return yt0

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        yd0, = ctx.saved_tensors
        return cparray(gd0.t * yd0.t)


class Cat(Function):
    """reference grad_fcn.py:881-904 (UNet skip connections, channel axis)."""

    @staticmethod
    def forward(ctx, *inputs, **params):
        dim = params['dim']
        datas = [t.data for t in inputs]
        ctx.params['sizes'] = [a.shape[dim] for a in datas]
        nd = datas[0].ndim
        ctx.params['channel_kernel'] = (nd == 4 and dim % nd == 1 and all(a.__class__ is cparray and a.ndim == 4 and
                                        a.t.dtype == torch.float32 and a.shape[0] == datas[0].shape[0] and
                                        a.shape[2:] == datas[0].shape[2:] for a in datas) and datas[0].size > 0)
        if ctx.params['channel_kernel']:  # NHWC channel concatenation (UNet skip connections): one strided copy per input
            return build_links(ops.cat_channels(datas), grad_fn=ctx)
        out = torch.cat([a.t for a in datas], dim=dim)
        return build_links(cparray(out), grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        if ctx.params['channel_kernel'] and gd0.t.dtype == torch.float32:
            return tuple(ops.split_channels(gd0, ctx.params['sizes'], ctx.needs_input_grad))
        parts = torch.split(gd0.t, ctx.params['sizes'], dim=ctx.params['dim'])
        return tuple(cparray(p.contiguous(memory_format=torch.channels_last) if p.dim() == 4 else p.contiguous())
                     if need else None for p, need in zip(parts, ctx.needs_input_grad))


class LogSoftmax(Function):
    """reference grad_nn.py:373-392"""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, = inputs
        x = xt0.data.t
        dim = params['dim']
        ctx.params['row_kernel'] = x.dim() == 2 and dim % 2 == 1 and x.dtype == torch.float32 and x.numel() > 0
        if ctx.params['row_kernel']:
            yt0 = build_links(ops.log_softmax(xt0.data), grad_fn=ctx)
        else:
            aug = x - x.max(dim=dim, keepdim=True).values
            y = aug - torch.log(torch.sum(torch.exp(aug), dim=dim, keepdim=True))
            yt0 = build_links(cparray(y), grad_fn=ctx)
        ctx.save_for_backward(yt0)
        return yt0

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        yd0, = ctx.saved_tensors
        if ctx.params['row_kernel'] and gd0.t.dtype == torch.float32:
            return ops.log_softmax_bwd(gd0, yd0)
        g = gd0.t
        return cparray(g - g.sum(dim=ctx.params['dim'], keepdim=True) * torch.exp(yd0.t))


class NllLoss(Function):
    """reference grad_nn.py:287-349; input (N, C) log-probabilities, integer class targets."""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, = inputs
        target = params['target']
        reduction = params['reduction']
        if params.get('weight') is not None:
            raise NotImplementedError("nll_loss class weights are not on the accelerated path yet")
        x = xt0.data.t
        tgt = target.data.t if target.data.__class__ is cparray else torch.from_numpy(np.asarray(target.data)).to(x.device)
        tgt = tgt.long()
        ignore_index = params.get('ignore_index', -100)
        if reduction not in ('none', 'mean', 'sum'):
            raise ValueError(f'{reduction} is not a valid value for reduction')
        ctx.params['kernel'] = x.dim() == 2 and x.dtype == torch.float32 and x.numel() > 0
        if ctx.params['kernel']:  # one fixed-order reduction kernel; the row count of the mean stays on the device
            tgt = tgt.contiguous()
            y, count = ops.nll_loss(xt0.data, tgt, ignore_index, reduction)
            ctx.params.update(tgt=tgt, count=count, shape=tuple(x.shape), ignore_index=ignore_index)
            return build_links(y, grad_fn=ctx)
        picked = -x.gather(1, tgt.clamp(min=0).unsqueeze(1)).squeeze(1)
        w = None
        if ignore_index >= 0:
            w = (tgt != ignore_index)
            picked = picked * w
        if reduction == 'sum':
            y = picked.sum()
        elif reduction == 'mean':
            if w is None:
                n = picked.numel()
                y = picked.mean()
            else:
                n = w.sum()
                y = picked.sum() / n
            ctx.params['N'] = n
        elif reduction == 'none':
            y = picked
        else:
            raise ValueError(f'{reduction} is not a valid value for reduction')
        ctx.params['tgt'] = tgt
        ctx.params['w'] = w
        ctx.params['shape'] = tuple(x.shape)
        return build_links(cparray(y), grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        if ctx.params['kernel']:
            return ops.nll_loss_bwd(gd0, ctx.params['tgt'], ctx.params['shape'], ctx.params['ignore_index'],
                                    ctx.params['reduction'], ctx.params['count'])
        g = gd0.t
        if ctx.params['reduction'] == 'mean':
            g = g / ctx.params['N']
        tgt, w = ctx.params['tgt'], ctx.params['w']
        out = torch.zeros(ctx.params['shape'], dtype=g.dtype, device=g.device)
        val = (-g).expand(tgt.shape) if g.dim() == 0 else -g
        if w is not None:
            val = val * w
        out.scatter_(1, tgt.clamp(min=0).unsqueeze(1), val.unsqueeze(1).to(out.dtype))
        return cparray(out)


class BinaryCrossEntropyWithLogits(Function):
    """reference grad_nn.py:236-285 (UNet loss); mean / sum / none reduction, no weights."""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, xt1 = inputs
        if params.get('weight') is not None or params.get('pos_weight') is not None:
            raise NotImplementedError("bce_with_logits weights are not on the accelerated path yet")
        x, t = xt0.data.t, xt1.data.t
        red = params['reduction']
        if red not in ('none', 'mean', 'sum'):
            raise ValueError(f'{red} is not a valid value for reduction')
        if x.shape != t.shape:
            raise ValueError(f'Target size ({tuple(t.shape)}) must be the same as input size ({tuple(x.shape)})')
        ctx.save_for_backward(xt0, xt1)
        return build_links(ops.bce_logits(xt0.data, xt1.data, red), grad_fn=ctx)

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        xd0, xd1 = ctx.saved_tensors
        g0 = ops.bce_logits_bwd(xd0, xd1, gd0, ctx.params['reduction']) if ctx.needs_input_grad[0] else None
        return g0, None

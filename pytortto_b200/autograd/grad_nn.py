"""The hot path: Convolution, TransposedConvolution, BatchNorm, Relu, MaxPool2DWithIndices as autograd Functions
that call the sm_100a kernels through the C ABI.

Each class keeps the reference's name, `forward(ctx, *inputs, **params)` / `backward(ctx, *grad_outputs)`
signature, parameter names, saved state and error messages (/root/reference/src/tortto/autograd/grad_nn.py,
lines cited per class), so `nn.functional` and the modules above it are unchanged callers.  What differs is below
the boundary: no `x_padded` copy is kept for wgrad (the TMA im2col load zero-fills the halo), dgrad is a gather
(no zero-inserted `grad_dilated`), BN is two fused passes per direction.
"""
import os
import math

from .. import ops
from ..xparray import cparray
from .function import AccumulateGrad, Function, grad_slot
from .grad_mode import is_grad_enabled
from .helper import build_links, inplace_precheck, inplace_update

prod = math.prod


def _require_cuda(name, *arrays):
    for a in arrays:
        if a is not None and a.__class__ is not cparray:
            raise RuntimeError(f"{name}: pytortto_b200 implements the CUDA path only (got a host array); "
                               f"move tensors and modules with .cuda()")


class Relu(Function):
    """reference grad_nn.py:33-69: y = maximum(x, 0) (NaN-propagating); saves the OUTPUT; dx = dy * (y > 0)."""
    _absorbs = ("bn", "bn_add")

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, = inputs
        xd0 = xt0.data
        _require_cuda('relu', xd0)
        deferred = ops.resolve_pending(Relu, (xt0,))  # a deferred bn(x) [+ identity]: the ReLU joins its normalise pass
        if params['inplace']:
            inplace_precheck(xt0)
            if deferred is not None:
                ops.bn_relu_fused(deferred, out=xd0)
            else:
                ops.relu_fwd(xd0, inplace=True)
            yt0 = inplace_update(xt0, ctx)
        elif deferred is not None:
            yt0 = build_links(ops.bn_relu_fused(deferred), grad_fn=ctx)
        else:
            yt0 = build_links(ops.relu_fwd(xd0), grad_fn=ctx)
        ctx.save_for_backward(yt0)
        return yt0

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        yd0, = ctx.saved_tensors
        return ops.relu_bwd(gd0, yd0)


class Convolution(Function):
    """reference grad_nn.py:684-734.  inputs = (input, weight, bias|None); params stride/padding/dilation/groups."""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, xt1, xt2 = inputs
        xd0, xd1 = xt0.data, xt1.data
        xd2 = None if xt2 is None else xt2.data
        stride = tuple(params['stride'])
        padding = tuple(params['padding'])
        dilation = tuple(params['dilation'])
        groups = params['groups']
        if xd0.__class__ is not xd1.__class__:
            raise RuntimeError(f'Input type ({xd0.__class__.__name__}) and weight type ({xd1.__class__.__name__}) '
                               f'should be the same')
        if xd2 is not None and xd0.__class__ is not xd2.__class__:
            raise RuntimeError(f'Input type ({xd0.__class__.__name__}) and bias type ({xd2.__class__.__name__}) '
                               f'should be the same')
        if xd0.ndim != 4:
            raise RuntimeError(f'Expected 3D (unbatched) or 4D (batched) input to conv2d, '
                               f'but got input of size: {xd0.shape}')
        if groups * xd1.shape[-3] != xd0.shape[-3]:
            raise RuntimeError(f'Given groups={groups}, weight of size {xd1.shape}, '
                               f'expected input{xd0.shape} to have {groups * xd1.shape[-3]} channels, '
                               f'but got {xd0.shape[-3]} channels instead')
        _require_cuda('conv2d', xd0)
        d = ops.conv_desc(xd0.shape, xd1.shape, stride, padding, dilation, groups)
        # while a graph is being recorded (training forward) the epilogue also emits the per-channel sum / sum of squares
        # of the output: the BatchNorm that reads it next (grad_nn.py:923-924 of the reference) skips its statistics pass
        want_stats = is_grad_enabled() and ops.conv_fused_info(d)[0]
        # (the launch waits for the next operator: a residual Add is absorbed into the epilogue, ops.conv2d_fprop_deferred)
        yd0 = ops.conv2d_fprop_deferred(xd0, xd1, xd2, d, want_stats)
        yt0 = build_links(yd0, grad_fn=ctx)
        ctx.save_for_backward(xt0, xt1)
        ctx.params['desc'] = d
        if xt0.requires_grad:
            ops.register_dgrad_weight(xd1, d)  # its dgrad re-ordering joins the step's one multi-tensor launch
        return yt0

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        xd0, xd1 = ctx.saved_tensors
        d = ctx.params['desc']
        grad0, grad1, grad2 = None, None, None
        if ctx.needs_input_grad[2]:  # (on the wgrad stream when it feeds a leaf: see the weight gradient below)
            grad2 = ops.bias_grad(gd0, out=grad_slot(ctx, 2),
                                  overlap=ctx.next_functions[2][0].__class__ is AccumulateGrad)
        # wgrad forks to a second stream (joined at the end of backward), dgrad stays on the critical path.  Measured on
        # B200 (preact_resnet18, batch 256, graph replay): 3.81 ms/step in this order, 3.89 with dgrad queued first,
        # 3.92 without the fork - the two tensor-bound kernels cannot share an SM (shared memory) and an HBM-bound
        # BatchNorm pass gains nothing from running beside a wgrad (scripts/overlap_probe.py), so the gain is only
        # the overlap of each kernel's partial last wave / epilogue with the start of the next
        # The fork is only safe when dW goes straight to a leaf: AccumulateGrad, the DP bucket launch and the end of the
        # sweep join the wgrad stream before anything reads it.  A non-leaf weight (w * mask, weight standardisation, a
        # view of a parameter) hands dW to another node's backward on the main stream, so it is computed in order.
        if ctx.needs_input_grad[1]:
            to_leaf = ctx.next_functions[1][0].__class__ is AccumulateGrad
            grad1 = ops.conv2d_wgrad(xd0, gd0, d, overlap=to_leaf, out=grad_slot(ctx, 1))
        # `_accum0`: a gradient that already reached the input through another branch (a ResNet block's identity shortcut);
        # the engine hands it over (Tensor._sweep) and the dgrad epilogue adds it instead of a separate add kernel
        accum = ctx.params.pop('_accum0', None)
        if ctx.needs_input_grad[0]:
            # the input came out of a BatchNorm(+ReLU) node: its backward runs next on this gradient and can take the dgrad
            # launch over, with the epilogue that emits its two reductions (ops.conv2d_dgrad_for_batchnorm)
            fn0 = ctx.next_functions[0][0]
            if fn0 is not None and getattr(getattr(fn0, '_forward_cls', None), '_absorbs_dgrad', False):
                grad0 = ops.conv2d_dgrad_for_batchnorm(gd0, xd1, d, accum=accum)
            else:
                grad0 = ops.conv2d_dgrad(gd0, xd1, d, accum=accum)
        return grad0, grad1, grad2


Convolution._accumulates_input0 = os.environ.get("TORTTO_B200_FOLD_ACCUM", "1") != "0"


class TransposedConvolution(Function):
    """reference grad_nn.py:737-779: forward IS the dgrad of a convolution whose 'output' is this op's input
    (:755); backward: wgrad with roles swapped (:776), input gradient = plain convolution of dy (:778).
    weight is (Cin, Cout/groups, kh, kw)."""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, xt1, xt2 = inputs
        xd0, xd1 = xt0.data, xt1.data
        stride = tuple(params['stride'])
        padding = tuple(params['padding'])
        dilation = tuple(params['dilation'])
        groups = params['groups']
        output_padding = params['output_padding']
        if xd0.ndim != 4:
            raise RuntimeError(f'Expected 3D (unbatched) or 4D (batched) input to conv_transpose2d, '
                               f'but got input of size: {xd0.shape}')
        if xd1.shape[-4] != xd0.shape[-3]:
            raise RuntimeError(f'Given transposed=1, weight of size {xd1.shape}, '
                               f'expected input {xd0.shape} to have {xd1.shape[-4]} channels, '
                               f'but got {xd0.shape[-3]} channels instead')
        _require_cuda('conv_transpose2d', xd0, xd1)
        op = tuple(output_padding) if hasattr(output_padding, '__len__') else (output_padding,)
        hop = op[0]
        wop = hop if len(op) == 1 else op[1]  # grad_nn.py:675-676
        n, cin, hi, wi = xd0.shape
        _, cog, kh, kw = xd1.shape
        ho = (hi - 1) * stride[0] - 2 * padding[0] + dilation[0] * (kh - 1) + hop + 1
        wo = (wi - 1) * stride[1] - 2 * padding[1] + dilation[1] * (kw - 1) + wop + 1
        # the underlying convolution maps (n, cog*groups, ho, wo) -> (n, cin, hi, wi)
        d = ops.conv_desc((n, cog * groups, ho, wo), xd1.shape, stride, padding, dilation, groups, out_hw=(hi, wi))
        yd0 = ops.conv2d_dgrad(xd0, xd1, d)
        if xt2 is not None:
            ops.add_bias_(yd0, xt2.data)
        yt0 = build_links(yd0, grad_fn=ctx)
        ctx.save_for_backward(xt0, xt1)
        ctx.params['desc'] = d
        return yt0

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, = grad_outputs
        xd0, xd1 = ctx.saved_tensors
        d = ctx.params['desc']
        grad0, grad1, grad2 = None, None, None
        if ctx.needs_input_grad[2]:
            grad2 = ops.bias_grad(gd0, out=grad_slot(ctx, 2))
        if ctx.needs_input_grad[1]:
            grad1 = ops.conv2d_wgrad(gd0, xd0, d, out=grad_slot(ctx, 1))
        if ctx.needs_input_grad[0]:
            grad0 = ops.conv2d_fprop(gd0, xd1, None, d)
        return grad0, grad1, grad2


class MaxPool2DWithIndices(Function):
    """reference grad_nn.py:828-864.  The saved 'pos' is a byte index (r*kw+s) per output element instead of the
    reference's six broadcast index arrays; backward keeps the reference's last-writer-wins overlap semantics."""

    @staticmethod
    def forward(ctx, *inputs, **params):
        xt0, = inputs
        xd0 = xt0.data
        kernel_size = tuple(params['kernel_size'])
        stride = tuple(params['stride'])
        padding = tuple(params['padding'])
        dilation = tuple(params['dilation'])
        ceil_mode = params['ceil_mode']
        return_indices = params['return_indices']
        if padding[0] * 2 > kernel_size[0] or padding[1] * 2 > kernel_size[1]:
            raise RuntimeError(f'pad should be smaller than or equal to half of kernel size, '
                               f'but got padW = {padding[1]}, padH = {padding[0]}, kW = {kernel_size[1]}, '
                               f'kH = {kernel_size[0]}')
        if xd0.ndim != 3 and xd0.ndim != 4:
            raise RuntimeError('non-empty 3D or 4D (batch mode) tensor expected for input')
        _require_cuda('max_pool2d', xd0)
        low_dim = xd0.ndim == 3
        if low_dim:
            xd0 = xd0.reshape((1,) + tuple(xd0.shape))
        d = ops.pool_desc(xd0.shape, kernel_size, stride, padding, dilation, ceil_mode)
        yd0, pos = ops.maxpool2d_fwd(xd0, d)
        if low_dim:
            yd0 = yd0.reshape(tuple(yd0.shape)[1:])
        yt0 = build_links(yd0, grad_fn=ctx)
        ctx.save_for_backward(xt0)
        ctx.params['pos'] = pos
        ctx.params['desc'] = d
        ctx.params['low_dim'] = low_dim
        return (yt0, pos) if return_indices else yt0

    @staticmethod
    def backward(ctx, *grad_outputs):
        gd0, *_ = grad_outputs
        d = ctx.params['desc']
        if ctx.params['low_dim']:
            gd0 = gd0.reshape((1,) + tuple(gd0.shape))
        grad0 = ops.maxpool2d_bwd(gd0, ctx.params['pos'], d, accumulate=_POOL_ACCUMULATE[0])
        if ctx.params['low_dim']:
            grad0 = grad0.reshape(tuple(grad0.shape)[1:])
        return grad0


_POOL_ACCUMULATE = [False]  # reference semantics by default; set_maxpool_backward_accumulate(True) = PyTorch's


def set_maxpool_backward_accumulate(flag):
    _POOL_ACCUMULATE[0] = bool(flag)


def _verify_batch_size(size):
    size_prods = size[0]
    for i in range(len(size) - 2):
        size_prods *= size[i + 2]
    if size_prods == 1:
        raise ValueError(f'Expected more than 1 value per channel when training, got input size {size}')


class _BatchNormBase(Function):
    """shared implementation of BatchNorm and the fused BatchNorm+ReLU node"""
    _fuse_relu = False

    @classmethod
    def _forward(cls, ctx, xt0, xt1, xt2):
        xd0 = xt0.data
        running_mean = ctx.params['running_mean']
        running_var = ctx.params['running_var']
        training = ctx.params['training']
        momentum = ctx.params['momentum']
        eps = ctx.params['eps']
        _require_cuda('batch_norm', xd0)
        gamma = None if xt1 is None else xt1.data
        beta = None if xt2 is None else xt2.data
        from .. import distributed as dist
        # the SyncBN slot of a layer is keyed by the module's construction index (identical on every rank; an id() would
        # differ between ranks and can be reused after garbage collection); functional calls without a module tag take
        # the library all-reduce
        ident = xt1 if xt1 is not None else running_mean
        sid = getattr(ident, '_bn_sync_id', None)
        hook = dist.bn_forward_hook(None if sid is None else (sid, 'f')) if training else None
        relu = cls._fuse_relu
        if training:
            _verify_batch_size(xd0.shape)
            track = running_mean is not None and running_var is not None
            yd0, stats, count = ops.bn_forward_train(xd0, gamma, beta, running_mean.data if track else None,
                                                     running_var.data if track else None, momentum, eps, relu=relu,
                                                     reduce_hook=hook)
        elif running_mean is not None and running_var is not None:
            yd0, stats, count = ops.bn_forward_eval(xd0, gamma, beta, running_mean.data, running_var.data, eps, relu=relu)
        else:
            yd0, stats, count = ops.bn_forward_train(xd0, gamma, beta, None, None, None, eps, relu=relu)
        yt0 = build_links(yd0, grad_fn=ctx)
        if relu:
            ctx.save_for_backward(xt0, xt1, yt0)
        else:
            ctx.save_for_backward(xt0, xt1)
        ctx.params = {'stats': stats, 'count': count, 'synced': hook is not None,
                      'sync_key': None if sid is None else (sid, 'b'),
                      'affine_ids': (None if xt1 is None else id(xt1), None if xt2 is None else id(xt2))}
        return yt0

    @classmethod
    def _backward(cls, ctx, gd0):
        saved = ctx.saved_tensors
        xd0, xd1 = saved[0], saved[1]
        relu_out = saved[2] if cls._fuse_relu else None
        from .. import distributed as dist
        hook = dist.bn_backward_hook(ctx.params['sync_key']) if ctx.params['synced'] else None
        if hook is not None:  # dgamma / dbeta below come from all-reduced sums: the DP layer must not reduce them again
            dist.note_synced_bn_params(ctx.params['affine_ids'])
        # `_accum0`: a gradient that already reached the input through another branch (the residual shortcut); the
        # engine hands it over (Tensor._sweep) and the dx pass adds it instead of a separate add kernel
        accum = ctx.params.pop('_accum0', None)
        return ops.bn_backward(gd0, xd0, xd1, ctx.params['stats'], ctx.params['count'], relu_out=relu_out,
                               need_dx=ctx.needs_input_grad[0], need_dgamma=ctx.needs_input_grad[1],
                               need_dbeta=ctx.needs_input_grad[2], reduce_hook=hook, accum=accum,
                               fused_relu=cls._fuse_relu,
                               out_dgamma=grad_slot(ctx, 1) if ctx.needs_input_grad[1] else None,
                               out_dbeta=grad_slot(ctx, 2) if ctx.needs_input_grad[2] else None)


# its backward can fold the pending gradient of input 0 into dx (TORTTO_B200_FOLD_ACCUM=0: separate add kernel)
_BatchNormBase._accumulates_input0 = os.environ.get("TORTTO_B200_FOLD_ACCUM", "1") != "0"
# (see Convolution.backward.)  Off by default: on B200 the fused launch is a wash - the statistics pass it removes
# (preact_resnet18: 0.29 -> 0.11 ms per step) comes back as a longer dgrad epilogue (0.77 -> 0.95 ms), the step gains 0.75 %
# in TF32 mode and loses 1.3 % in bf16 mode (profiles/r2_dgrad_bn_ab.txt).  TORTTO_B200_DGRAD_BN=1 switches it on.
_BatchNormBase._absorbs_dgrad = os.environ.get("TORTTO_B200_DGRAD_BN", "0") != "0"


class BatchNormRelu(_BatchNormBase):
    """BatchNorm immediately followed by ReLU as ONE node (what `nn.Sequential(BatchNorm2d, ReLU)` lowers to):
    forward writes max(bn(x), 0) in the normalise pass; backward masks dy with (y > 0) inside the two BN backward
    kernels (Relu.backward, reference grad_nn.py:64-69, folded in).  Saves 1 read + 1 write forward and 2 reads +
    1 write backward per element against separate nodes; numerically identical to BatchNorm -> Relu."""
    _fuse_relu = True

    @staticmethod
    def forward(ctx, *inputs, **params):
        return BatchNormRelu._forward(ctx, *inputs)

    @staticmethod
    def backward(ctx, *grad_outputs):
        return BatchNormRelu._backward(ctx, grad_outputs[0])


class BatchNorm(_BatchNormBase):
    """reference grad_nn.py:907-989.  inputs = (input, weight|None, bias|None); params running_mean, running_var
    (Tensors or None), training, momentum, eps.  Training: batch mean / BIASED variance normalise, running stats get
    the UNBIASED variance (:923-930).  Saves input + weight and keeps mean, var+eps, sd for backward (:962-963);
    backward always uses those saved statistics, also in eval mode (:967-989).

    Under `pytortto_b200.distributed` with sync_bn the per-channel sums are all-reduced (forward: sum x, sum x^2;
    backward: sum dy, sum dy*(x-mean)), which makes a k-GPU run equal the single-process global batch."""

    @staticmethod
    def forward(ctx, *inputs, **params):
        return BatchNorm._forward(ctx, *inputs)

    @staticmethod
    def backward(ctx, *grad_outputs):
        return BatchNorm._backward(ctx, grad_outputs[0])

"""`tortto.nn.functional` for the conv-net path - the entry points kept verbatim from the reference
(/root/reference/src/tortto/nn/functional.py:6-11, 54-63, 80-121): same names, argument order, defaults."""
from ..autograd.grad_fcn import (BinaryCrossEntropyWithLogits, Linear, LogSoftmax, NllLoss, View)
from ..autograd.grad_nn import BatchNorm, BatchNormRelu, Convolution, MaxPool2DWithIndices, Relu, TransposedConvolution
from ..VariableFunctions import matmul


def relu(input, inplace=False):
    return Relu.apply(input, inplace=inplace)


def relu_(input):
    return Relu.apply(input, inplace=True)


def _unsqueeze0(t):
    return View.apply(t, shape=(1,) + tuple(t.shape))


def _squeeze0(t):
    return View.apply(t, shape=tuple(t.shape)[1:])


def conv2d(input, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1):
    low_dim = input.ndim == 3
    if low_dim:
        input = _unsqueeze0(input)
    result = Convolution.apply(input, weight, bias, stride=stride, padding=padding, dilation=dilation, groups=groups)
    return _squeeze0(result) if low_dim else result


def conv_transpose2d(input, weight, bias=None, stride=(1, 1), padding=(0, 0), output_padding=(0, 0), groups=1,
                     dilation=(1, 1)):
    low_dim = input.ndim == 3
    if low_dim:
        input = _unsqueeze0(input)
    result = TransposedConvolution.apply(input, weight, bias, stride=stride, padding=padding,
                                         output_padding=output_padding, dilation=dilation, groups=groups)
    return _squeeze0(result) if low_dim else result


def max_pool2d(input, kernel_size, stride=(1, 1), padding=(0, 0), dilation=(1, 1), ceil_mode=False,
               return_indices=False):
    return MaxPool2DWithIndices.apply(input, kernel_size=kernel_size, stride=stride, padding=padding,
                                      dilation=dilation, ceil_mode=ceil_mode, return_indices=return_indices)


def _conv_desc_of(input, conv):
    from .. import ops
    return ops.conv_desc(tuple(input.shape), tuple(conv.weight.shape), tuple(conv.stride), tuple(conv.padding),
                         tuple(conv.dilation), conv.groups)


def conv2d_epilogue_available(input, conv):
    """the convolution runs on the tensor path, whose epilogue can apply a folded BatchNorm / ReLU"""
    from .. import ops
    from ..xparray import cparray
    if input.data.__class__ is not cparray or conv.weight.data.__class__ is not cparray:
        return False
    if conv.groups * conv.weight.shape[1] != input.shape[1]:
        return False  # (let the plain path raise the reference's error)
    return ops.conv_fused_info(_conv_desc_of(input, conv))[0]


def conv2d_bn_eval(input, conv, bn, relu):
    """Inference form of Conv2d -> BatchNorm2d(eval) [-> ReLU]: one convolution kernel; the BatchNorm (running statistics,
    reference grad_nn.py:932-959) is folded to per-channel scale / shift applied, with the ReLU (:58), in its epilogue.
    Called by nn.Sequential under no_grad; no graph is recorded."""
    from .. import ops
    from ..tensor import Tensor
    d = _conv_desc_of(input, conv)
    y = ops.conv2d_bn_eval(input.data, conv.weight.data, None if conv.bias is None else conv.bias.data, d,
                           bn.running_mean.data, bn.running_var.data, bn.eps,
                           None if bn.weight is None else bn.weight.data, None if bn.bias is None else bn.bias.data, relu)
    return Tensor(y, copy=False, dtype=y.dtype)


def batch_norm(input, running_mean, running_var, weight=None, bias=None, training=False, momentum=0.1, eps=1e-5):
    return BatchNorm.apply(input, weight, bias, running_mean=running_mean, running_var=running_var,
                           training=training, momentum=momentum, eps=eps)


def batch_norm_relu(input, running_mean, running_var, weight=None, bias=None, training=False, momentum=0.1, eps=1e-5):
    """relu(batch_norm(...)) as one fused node; `nn.Sequential` calls this for a BatchNorm module followed by ReLU."""
    return BatchNormRelu.apply(input, weight, bias, running_mean=running_mean, running_var=running_var,
                               training=training, momentum=momentum, eps=eps)


def linear(input, weight, bias):
    if input.ndim == 2 and weight.ndim == 2 and input.is_cuda and weight.is_cuda and input.dtype == weight.dtype == 'float32':
        return Linear.apply(input, weight, bias)  # one fused node: x @ W^T + b
    output = matmul(input, weight.T)
    return output if bias is None else output + bias


def log_softmax(input, dim):
    return LogSoftmax.apply(input, dim=dim)


def nll_loss(input, target, weight=None, ignore_index=-100, reduction='mean'):
    return NllLoss.apply(input, target=target, weight=weight, ignore_index=ignore_index, reduction=reduction)


def binary_cross_entropy_with_logits(input, target, weight=None, pos_weight=None, reduction='mean'):
    return BinaryCrossEntropyWithLogits.apply(input, target, weight=weight, pos_weight=pos_weight, reduction=reduction)

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

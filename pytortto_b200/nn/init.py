"""Parameter initialisers with the reference's formulas and RNG source (nn/init.py:6-109): draws come from
numpy's global generator, so `manual_seed(s)` followed by module construction reproduces the reference's
initial weights bit for bit (modules are built on the host and moved with .cuda(), like the examples do)."""
import math

import numpy as np

from ..xparray import cparray


def _assign(tensor, values):
    d = tensor.data
    if d.__class__ is cparray:
        d.t.copy_(cparray.from_numpy(values.astype(d.dtype)).t)
        d._touched()
    else:
        d[...] = values.astype(d.dtype)
    return tensor


def uniform_(tensor, a=0, b=1):
    return _assign(tensor, np.random.uniform(low=a, high=b, size=tensor.shape))


def normal_(tensor, mean=0., std=1.):
    return _assign(tensor, np.random.normal(loc=mean, scale=std, size=tensor.shape))


def constant_(tensor, val):
    tensor.data.fill(val)
    return tensor


def ones_(tensor):
    return constant_(tensor, 1.)


def zeros_(tensor):
    return constant_(tensor, 0.)


def _calculate_fan_in_and_fan_out(tensor):
    if tensor.dim() < 2:
        raise ValueError("Fan in and fan out can not be computed for tensor with fewer than 2 dimensions")
    receptive = 1
    for s in tensor.shape[2:]:
        receptive *= s
    return tensor.shape[1] * receptive, tensor.shape[0] * receptive


def _calculate_correct_fan(tensor, mode):
    mode = mode.lower()
    if mode not in ('fan_in', 'fan_out'):
        raise ValueError("Mode {} not supported, please use one of {}".format(mode, ['fan_in', 'fan_out']))
    fan_in, fan_out = _calculate_fan_in_and_fan_out(tensor)
    return fan_in if mode == 'fan_in' else fan_out


def calculate_gain(nonlinearity, param=None):
    linear_fns = ['linear', 'conv1d', 'conv2d', 'conv3d', 'conv_transpose1d', 'conv_transpose2d', 'conv_transpose3d']
    if nonlinearity in linear_fns or nonlinearity == 'sigmoid':
        return 1
    if nonlinearity == 'tanh':
        return 5.0 / 3
    if nonlinearity == 'relu':
        return math.sqrt(2.0)
    if nonlinearity == 'leaky_relu':
        if param is None:
            negative_slope = 0.01
        elif not isinstance(param, bool) and isinstance(param, (int, float)):
            negative_slope = param
        else:
            raise ValueError("negative_slope {} not a valid number".format(param))
        return math.sqrt(2.0 / (1 + negative_slope ** 2))
    if nonlinearity == 'selu':
        return 3.0 / 4
    raise ValueError("Unsupported nonlinearity {}".format(nonlinearity))


def kaiming_uniform_(tensor, a=0, mode='fan_in', nonlinearity='leaky_relu'):
    fan = _calculate_correct_fan(tensor, mode)
    std = calculate_gain(nonlinearity, a) / math.sqrt(fan)
    bound = math.sqrt(3.0) * std
    return uniform_(tensor, -bound, bound)


def xavier_uniform_(tensor, gain=1.):
    fan_in, fan_out = _calculate_fan_in_and_fan_out(tensor)
    a = math.sqrt(3.0) * gain * math.sqrt(2.0 / float(fan_in + fan_out))
    return uniform_(tensor, -a, a)

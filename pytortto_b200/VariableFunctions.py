"""Free functions of the package namespace (`tt.mean`, `tt.flatten`, `tt.zeros`, `tt.manual_seed`, ...), the
subset of the reference's VariableFunctions.py a CNN training script touches."""
import numpy as np

from .autograd import grad_fcn as G
from .tensor import Tensor, float32, tensor
from .xparray import cparray


def manual_seed(seed):
    """reference VariableFunctions.py:8-9: seeds numpy's global RNG (parameter init draws from it)."""
    np.random.seed(seed)


def _as_tensor(x, like):
    if isinstance(x, Tensor):
        return x
    arr = np.asarray(x, dtype=like.dtype)
    t = Tensor(arr, dtype=like.dtype, copy=False)
    return t.cuda() if like.is_cuda else t


def add(a, b):
    if not isinstance(a, Tensor):
        a, b = b, a
    return G.Add.apply(a, _as_tensor(b, a))


def mul(a, b):
    if not isinstance(a, Tensor):
        a, b = b, a
    return G.Mul.apply(a, _as_tensor(b, a))


def sum(input, dim=None, keepdim=False):
    return G.Sum.apply(input, dim=dim, keepdim=keepdim)


def mean(input, dim=None, keepdim=False):
    return G.Mean.apply(input, dim=dim, keepdim=keepdim)


def exp(input):
    return G.Exp.apply(input)


def reshape(input, *shape):
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
        shape = tuple(shape[0])
    if -1 in shape:
        known = 1
        for s in shape:
            if s != -1:
                known *= s
        shape = tuple(input.numel() // known if s == -1 else s for s in shape)
    return G.View.apply(input, shape=tuple(shape))


def flatten(input, start_dim=0, end_dim=-1):
    shp = input.shape
    n = len(shp)
    s, e = start_dim % n, end_dim % n
    mid = 1
    for v in shp[s:e + 1]:
        mid *= v
    return G.View.apply(input, shape=tuple(shp[:s]) + (mid,) + tuple(shp[e + 1:]))


def transpose(input, dim0, dim1):
    return G.Transpose.apply(input, dim0=dim0, dim1=dim1)


def matmul(a, b):
    return G.Mm.apply(a, b)


def cat(tensors, dim=0):
    return G.Cat.apply(*tensors, dim=dim)


def _shape_arg(shape):
    return tuple(shape[0]) if len(shape) == 1 and hasattr(shape[0], '__iter__') else tuple(shape)


def zeros(*shape, dtype=None, requires_grad=False):
    return Tensor(np.zeros(_shape_arg(shape), dtype=dtype or float32), dtype=dtype or float32, copy=False,
                  requires_grad=requires_grad)


def ones(*shape, dtype=None, requires_grad=False):
    return Tensor(np.ones(_shape_arg(shape), dtype=dtype or float32), dtype=dtype or float32, copy=False,
                  requires_grad=requires_grad)


def empty(*shape, dtype=None, requires_grad=False):
    return Tensor(np.empty(_shape_arg(shape), dtype=dtype or float32), dtype=dtype or float32, copy=False,
                  requires_grad=requires_grad)


def randn(*shape, dtype=None, requires_grad=False):
    return Tensor(np.random.randn(*_shape_arg(shape)), dtype=dtype or float32, copy=False, requires_grad=requires_grad)


__all__ = ['manual_seed', 'add', 'mul', 'sum', 'mean', 'exp', 'reshape', 'flatten', 'transpose', 'matmul', 'cat',
           'zeros', 'ones', 'empty', 'randn', 'tensor', 'Tensor', 'cparray']

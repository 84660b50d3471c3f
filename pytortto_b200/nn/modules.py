"""Module system + the layers of the conv-net path, API-compatible with the reference's tortto.nn.modules
(/root/reference/src/tortto/nn/modules/*.py): `Module` (parameter / buffer / child registration through attribute
assignment, `parameters()`, `named_parameters()`, `state_dict()` as numpy arrays, `.cuda()` rebuilding Parameters,
train/eval), `Sequential`, `ModuleList`, `Identity`, `Conv2d`, `ConvTranspose2d`, `BatchNorm2d`, `ReLU`,
`MaxPool2d`, `Linear`, `LogSoftmax`, `NLLLoss`, `BCEWithLogitsLoss`.  These are thin callers of nn.functional; they
are mirrored (not accelerated) so an unchanged tortto model definition runs on the B200 path.
"""
import math
from collections import OrderedDict
from collections.abc import Iterable
from itertools import repeat

import numpy as np

from .. import VariableFunctions as V
from ..autograd.grad_mode import is_grad_enabled, no_grad
from ..tensor import Tensor
from ..xparray import cparray
from . import functional as F
from . import init
from .parameter import Parameter


def _ntuple(n):
    def parse(x):
        if isinstance(x, Iterable):
            return tuple(x)[:n]  # reference nn/modules/utils.py:5-10 (note: truncates, never pads)
        return tuple(repeat(x, n))
    return parse


_single = _ntuple(1)
_pair = _ntuple(2)


class Module:
    def __init__(self):
        self.training = True
        self._parameters = OrderedDict()
        self._modules = OrderedDict()
        self._buffers = OrderedDict()

    # ---- registration --------------------------------------------------------------------------------------
    def register_parameter(self, name, param):
        if '_parameters' not in self.__dict__:
            raise AttributeError("cannot assign parameter before Module.__init__() call")
        if param is not None and not isinstance(param, Parameter):
            raise TypeError(f"cannot assign '{type(param).__name__}' object to parameter '{name}' "
                            "(nn.Parameter or None required)")
        self._parameters[name] = param

    def register_buffer(self, name, tensor):
        if '_buffers' not in self.__dict__:
            raise AttributeError("cannot assign buffer before Module.__init__() call")
        if not isinstance(name, str):
            raise TypeError("buffer name should be a string. Got {}".format(type(name)))
        if '.' in name:
            raise KeyError("buffer name can't contain \".\"")
        if tensor is not None and not isinstance(tensor, Tensor):
            raise TypeError(f"cannot assign '{type(tensor).__name__}' object to buffer '{name}' "
                            "(Tensor or None required)")
        self._buffers[name] = tensor

    def add_module(self, name, module):
        if module is not None and not isinstance(module, Module):
            raise TypeError("{} is not a Module subclass".format(type(module).__name__))
        self._modules[name] = module

    def __getattr__(self, name):
        d = self.__dict__
        for store in ('_parameters', '_buffers', '_modules'):
            if store in d and name in d[store]:
                return d[store][name]
        raise AttributeError("'{}' object has no attribute '{}'".format(type(self).__name__, name))

    def __setattr__(self, name, value):
        d = self.__dict__
        if isinstance(value, Parameter):
            if '_parameters' not in d:
                raise AttributeError("cannot assign parameters before Module.__init__() call")
            for store in ('_buffers', '_modules'):
                d[store].pop(name, None)
            d.pop(name, None)
            self._parameters[name] = value
        elif '_parameters' in d and name in d['_parameters']:
            if value is not None:
                raise TypeError(f"cannot assign '{type(value).__name__}' as parameter '{name}' "
                                "(nn.Parameter or None expected)")
            self._parameters[name] = None
        elif isinstance(value, Module):
            if '_modules' not in d:
                raise AttributeError("cannot assign module before Module.__init__() call")
            d.pop(name, None)
            self._modules[name] = value
        elif '_modules' in d and name in d['_modules']:
            self._modules[name] = value
        elif '_buffers' in d and name in d['_buffers']:
            if value is not None and not isinstance(value, Tensor):
                raise TypeError(f"cannot assign '{type(value).__name__}' as buffer '{name}' (Tensor or None expected)")
            self._buffers[name] = value
        else:
            object.__setattr__(self, name, value)

    def __delattr__(self, name):
        for store in ('_parameters', '_buffers', '_modules'):
            if name in self.__dict__.get(store, ()):
                del self.__dict__[store][name]
                return
        object.__delattr__(self, name)

    # ---- calling ---------------------------------------------------------------------------------------------
    def forward(self, *input, **kwargs):
        raise NotImplementedError

    def __call__(self, *input, **kwargs):
        return self.forward(*input, **kwargs)

    # ---- traversal -------------------------------------------------------------------------------------------
    def named_modules(self, memo=None, prefix=''):
        if memo is None:
            memo = set()
        if self not in memo:
            memo.add(self)
            yield prefix, self
            for name, module in self._modules.items():
                if module is None:
                    continue
                yield from module.named_modules(memo, prefix + ('.' if prefix else '') + name)

    def modules(self):
        for _, m in self.named_modules():
            yield m

    def named_children(self):
        seen = set()
        for name, module in self._modules.items():
            if module is not None and module not in seen:
                seen.add(module)
                yield name, module

    def children(self):
        for _, m in self.named_children():
            yield m

    def _named_members(self, get_members_fn, prefix='', recurse=True):
        seen = set()
        mods = self.named_modules(prefix=prefix) if recurse else [(prefix, self)]
        for mprefix, module in mods:
            for k, v in get_members_fn(module):
                if v is None or v in seen:
                    continue
                seen.add(v)
                yield mprefix + ('.' if mprefix else '') + k, v

    def named_parameters(self, prefix='', recurse=True):
        yield from self._named_members(lambda m: m._parameters.items(), prefix, recurse)

    def parameters(self, recurse=True):
        for _, p in self.named_parameters(recurse=recurse):
            yield p

    def named_buffers(self, prefix='', recurse=True):
        yield from self._named_members(lambda m: m._buffers.items(), prefix, recurse)

    def buffers(self, recurse=True):
        for _, b in self.named_buffers(recurse=recurse):
            yield b

    # ---- modes / device ----------------------------------------------------------------------------------------
    def train(self, mode=True):
        self.training = mode
        for m in self.children():
            m.train(mode)
        return self

    def eval(self):
        return self.train(False)

    def requires_grad_(self, requires_grad=True):
        for p in self.parameters():
            p.requires_grad_(requires_grad)
        return self

    def zero_grad(self):
        for p in self.parameters():
            p.grad = None

    def apply(self, fn):
        for m in self.children():
            m.apply(fn)
        fn(self)
        return self

    def cuda(self):
        return self._apply(lambda t: t.cuda())

    def cpu(self):
        return self._apply(lambda t: t.cpu())

    def _apply(self, fn):
        """reference module.py:357-385: every parameter is REPLACED by a new Parameter object (build optimizers
        after .cuda()), buffers are replaced by the moved tensors."""
        for m in self.children():
            m._apply(fn)
        for key, param in self._parameters.items():
            if param is None:
                continue
            with no_grad():
                applied = fn(param)
            out = Parameter(applied, param.requires_grad)
            if param.grad is not None:
                g = param.grad
                if applied.is_cuda and g.__class__ is not cparray:
                    g = cparray.from_numpy(g)
                elif not applied.is_cuda and g.__class__ is cparray:
                    g = g.get()
                out.grad = g
            self._parameters[key] = out
        for key, buf in self._buffers.items():
            if buf is not None:
                self._buffers[key] = fn(buf)
        return self

    # ---- (de)serialisation: always numpy arrays keyed 'a.b.weight' (module.py:108-131) ---------------------------
    def state_dict(self, destination=None, prefix=''):
        if destination is None:
            destination = OrderedDict()
        for name, t in list(self._parameters.items()) + list(self._buffers.items()):
            if t is not None:
                d = t.data
                destination[prefix + name] = d.get() if d.__class__ is cparray else np.asarray(d)
        for name, m in self._modules.items():
            if m is not None:
                m.state_dict(destination, prefix + name + '.')
        return destination

    def load_state_dict(self, state_dict, strict=True):
        missing, unexpected, errors = [], [], []
        own = OrderedDict()
        for prefix, m in self.named_modules():
            for name, t in list(m._parameters.items()) + list(m._buffers.items()):
                if t is not None:
                    own[prefix + ('.' if prefix else '') + name] = t
        for key, t in own.items():
            if key not in state_dict:
                missing.append(key)
                continue
            src = state_dict[key]
            if tuple(np.shape(src)) != tuple(t.shape):
                errors.append('size mismatch for {}: copying a param with shape {} from checkpoint, '
                              'the shape in current model is {}.'.format(key, tuple(np.shape(src)), t.shape))
                continue
            with no_grad():
                t.copy_(src)
        for m in self.modules():
            if hasattr(m, '_nbt'):
                m._nbt = None
        for key in state_dict:
            if key not in own:
                unexpected.append(key)
        if strict:
            if unexpected:
                errors.insert(0, 'Unexpected key(s) in state_dict: {}. '.format(', '.join(f'"{k}"' for k in unexpected)))
            if missing:
                errors.insert(0, 'Missing key(s) in state_dict: {}. '.format(', '.join(f'"{k}"' for k in missing)))
        if errors:
            raise RuntimeError('Error(s) in loading state_dict for {}:\n\t{}'.format(self.__class__.__name__,
                                                                                     "\n\t".join(errors)))
        return missing, unexpected

    # ---- repr --------------------------------------------------------------------------------------------------
    def extra_repr(self):
        return ''

    def __repr__(self):
        lines = [f'({k}): ' + repr(m).replace('\n', '\n  ') for k, m in self._modules.items()]
        head = self.__class__.__name__ + '(' + self.extra_repr()
        return head + (('\n  ' + '\n  '.join(lines) + '\n') if lines else '') + ')'


class Sequential(Module):
    def __init__(self, *args):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for k, m in args[0].items():
                self.add_module(k, m)
        else:
            for i, m in enumerate(args):
                self.add_module(str(i), m)

    def __len__(self):
        return len(self._modules)

    def __iter__(self):
        return iter(self._modules.values())

    def __getitem__(self, idx):
        vals = list(self._modules.values())
        return Sequential(*vals[idx]) if isinstance(idx, slice) else vals[idx]

    def forward(self, x):
        mods = list(self._modules.values())
        i, n = 0, len(mods)
        while i < n:
            m = mods[i]
            # peephole: BatchNorm immediately followed by ReLU inside one Sequential is a single fused node
            # (same values, same gradients; the intermediate is not observable from outside the container)
            if i + 1 < n and isinstance(m, _BatchNorm) and type(mods[i + 1]) is ReLU and _FUSE_BN_RELU[0]:
                x = m(x, fuse_relu=True)
                i += 2
            elif i + 1 < n and type(m) is Conv2d and _conv_bn_eval_fusable(m, mods[i + 1], x):
                # inference peephole: Conv2d -> BatchNorm2d(eval) [-> ReLU] is one convolution whose epilogue applies the
                # folded per-channel scale / shift and the ReLU (no intermediate tensor)
                relu = i + 2 < n and type(mods[i + 2]) is ReLU
                x = F.conv2d_bn_eval(x, m, mods[i + 1], relu)
                i += 3 if relu else 2
            else:
                x = m(x)
                i += 1
        return x


_FUSE_BN_RELU = [True]


def _conv_bn_eval_fusable(conv, bn, x):
    return (_FUSE_BN_RELU[0] and type(bn) is BatchNorm2d and not bn.training and bn.running_mean is not None
            and bn.running_var is not None and not is_grad_enabled() and conv.padding_mode == 'zeros' and x.ndim == 4
            and F.conv2d_epilogue_available(x, conv))


def set_bn_relu_fusion(flag):
    """Enable / disable the Sequential(BatchNorm, ReLU) -> fused node peephole (on by default)."""
    _FUSE_BN_RELU[0] = bool(flag)


class ModuleList(Module):
    def __init__(self, modules=None):
        super().__init__()
        if modules is not None:
            for m in modules:
                self.append(m)

    def append(self, m):
        self.add_module(str(len(self._modules)), m)
        return self

    def __len__(self):
        return len(self._modules)

    def __iter__(self):
        return iter(self._modules.values())

    def __getitem__(self, idx):
        vals = list(self._modules.values())
        return ModuleList(vals[idx]) if isinstance(idx, slice) else vals[idx]


class Identity(Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x):
        return x


class ReLU(Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, x):
        return F.relu(x, inplace=self.inplace)

    def extra_repr(self):
        return 'inplace=True' if self.inplace else ''


class LogSoftmax(Module):
    def __init__(self, dim=None):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        return F.log_softmax(x, self.dim)


class NLLLoss(Module):
    def __init__(self, weight=None, ignore_index=-100, reduction='mean'):
        super().__init__()
        self.weight, self.ignore_index, self.reduction = weight, ignore_index, reduction

    def forward(self, input, target):
        return F.nll_loss(input, target, self.weight, self.ignore_index, self.reduction)


class BCEWithLogitsLoss(Module):
    def __init__(self, weight=None, reduction='mean', pos_weight=None):
        super().__init__()
        self.weight, self.reduction, self.pos_weight = weight, reduction, pos_weight

    def forward(self, input, target):
        return F.binary_cross_entropy_with_logits(input, target, self.weight, self.pos_weight, self.reduction)


class Linear(Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_feature = in_features
        self.out_feature = out_features
        self.weight = Parameter(V.zeros((out_features, in_features)))
        if bias:
            self.bias = Parameter(V.zeros(out_features))
        else:
            self.bias = None
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)

    def extra_repr(self):
        return f'in_feature={self.in_feature}, out_feature={self.out_feature}, bias={self.bias is not None}'


class _ConvNd(Module):
    """reference nn/modules/conv.py:8-72"""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, transposed, output_padding,
                 groups, bias, padding_mode):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError('in_channels must be divisible by groups')
        if out_channels % groups != 0:
            raise ValueError('out_channels must be divisible by groups')
        valid_padding_modes = {'zeros', 'reflect', 'replicate'}
        if padding_mode not in valid_padding_modes:
            raise ValueError("padding_mode must be one of {}, but got padding_mode='{}'".format(
                valid_padding_modes, padding_mode))
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation = kernel_size, stride, padding, dilation
        self.transposed, self.output_padding, self.groups, self.padding_mode = transposed, output_padding, groups, padding_mode
        if transposed:
            self.weight = Parameter(V.zeros((in_channels, out_channels // groups, *kernel_size)))
        else:
            self.weight = Parameter(V.zeros((out_channels, in_channels // groups, *kernel_size)))
        if bias:
            self.bias = Parameter(V.zeros(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            if fan_in != 0:
                bound = 1 / math.sqrt(fan_in)
                init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        s = f'{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}'
        if self.padding != (0,) * len(self.padding):
            s += f', padding={self.padding}'
        if self.dilation != (1,) * len(self.dilation):
            s += f', dilation={self.dilation}'
        if self.groups != 1:
            s += f', groups={self.groups}'
        if self.bias is None:
            s += ', bias=False'
        return s


class Conv2d(_ConvNd):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='zeros'):
        super().__init__(in_channels, out_channels, _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation),
                         False, _pair(0), groups, bias, padding_mode)

    def forward(self, inpt):
        if self.padding_mode != 'zeros':
            raise NotImplementedError("TODO: Currently only support mode='constant'")  # reference functional.py:74-77
        return F.conv2d(inpt, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


class ConvTranspose2d(_ConvNd):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, groups=1,
                 bias=True, dilation=1, padding_mode='zeros'):
        if padding_mode != 'zeros':
            raise ValueError('Only "zeros" padding mode is supported for {}'.format(self.__class__.__name__))
        super().__init__(in_channels, out_channels, _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation),
                         True, _pair(output_padding), groups, bias, padding_mode)

    def _output_padding(self, input, output_size, stride, padding, kernel_size, dilation=None):
        if output_size is None:
            return _single(self.output_padding)  # reference conv.py:129 - only the first element survives
        k = input.dim() - 2
        if len(output_size) == k + 2:
            output_size = output_size[2:]
        if len(output_size) != k:
            raise ValueError("output_size must have {} or {} elements (got {})".format(k, k + 2, len(output_size)))
        res = []
        for d in range(k):
            lo = ((input.shape[d + 2] - 1) * stride[d] - 2 * padding[d] +
                  (dilation[d] if dilation is not None else 1) * (kernel_size[d] - 1) + 1)
            hi = lo + stride[d] - 1
            if output_size[d] < lo or output_size[d] > hi:
                raise ValueError("requested an output size of {}, but valid sizes range from {} to {} (for an input "
                                 "of {})".format(output_size, lo, hi, input.shape[2:]))
            res.append(output_size[d] - lo)
        return res

    def forward(self, inpt, output_size=None):
        output_padding = self._output_padding(inpt, output_size, self.stride, self.padding, self.kernel_size,
                                              self.dilation)
        return F.conv_transpose2d(inpt, self.weight, self.bias, self.stride, self.padding, output_padding,
                                  self.groups, self.dilation)


class _MaxPoolNd(Module):
    def __init__(self, kernel_size, stride=None, padding=0, dilation=1, return_indices=False, ceil_mode=False):
        super().__init__()
        self.kernel_size = kernel_size
        self.stride = stride if (stride is not None) else kernel_size
        self.padding, self.dilation = padding, dilation
        self.return_indices, self.ceil_mode = return_indices, ceil_mode

    def extra_repr(self):
        return (f'kernel_size={self.kernel_size}, stride={self.stride}, padding={self.padding}, '
                f'dilation={self.dilation}, ceil_mode={self.ceil_mode}')


class MaxPool2d(_MaxPoolNd):
    def forward(self, input):
        return F.max_pool2d(input, _pair(self.kernel_size), _pair(self.stride), _pair(self.padding),
                            _pair(self.dilation), self.ceil_mode, self.return_indices)


class _BatchNorm(Module):
    """reference nn/modules/batchnorm.py:8-93"""

    _sync_counter = [0]  # construction order is the same on every rank (SPMD): a deterministic per-layer SyncBN identity

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        _BatchNorm._sync_counter[0] += 1
        self._sync_id = _BatchNorm._sync_counter[0]
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.track_running_stats = affine, track_running_stats
        if affine:
            self.weight = Parameter(V.ones(num_features))
            self.bias = Parameter(V.zeros(num_features))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)
        if track_running_stats:
            self.register_buffer('running_mean', V.zeros(num_features))
            self.register_buffer('running_var', V.ones(num_features))
            self.register_buffer('num_batches_tracked', Tensor(np.array(0)))  # float32 0-d, like the reference
        else:
            self.register_buffer('running_mean', None)
            self.register_buffer('running_var', None)
            self.register_buffer('num_batches_tracked', None)

    def reset_running_stats(self):
        if self.track_running_stats:
            init.zeros_(self.running_mean)
            init.ones_(self.running_var)
            init.zeros_(self.num_batches_tracked)

    def reset_parameters(self):
        self.reset_running_stats()
        if self.affine:
            init.ones_(self.weight)
            init.zeros_(self.bias)

    def _check_input_dim(self, input):
        raise NotImplementedError

    def _flush_counter(self):
        if getattr(self, '_nbt', None) is not None and self.num_batches_tracked is not None:
            self.num_batches_tracked.data.fill(self._nbt)

    def state_dict(self, destination=None, prefix=''):
        self._flush_counter()
        return super().state_dict(destination, prefix)

    def extra_repr(self):
        return (f'{self.num_features}, eps={self.eps}, momentum={self.momentum}, affine={self.affine}, '
                f'track_running_stats={self.track_running_stats}')

    def forward(self, inpt, fuse_relu=False):
        self._check_input_dim(inpt)
        if self.training and self.track_running_stats:
            # host-side step counter: the reference bumps a float32 device scalar and reads it back with .item()
            # when momentum is None (a device sync per layer per step); here the count lives on the host and
            # is written to the `num_batches_tracked` buffer when the state is exported (state_dict)
            if getattr(self, '_nbt', None) is None:
                self._nbt = float(self.num_batches_tracked.item())
            self._nbt += 1.0
            if self.momentum is None:
                # cumulative moving average: the factor 1/n changes every step and is a HOST scalar - a CUDA-graph replay
                # would keep applying the factor of the captured step, silently diverging from eager execution
                import torch
                if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("BatchNorm with momentum=None (cumulative moving average) cannot be captured in a "
                                       "CUDA graph: its averaging factor 1/num_batches_tracked changes every step")
                factor = 1.0 / self._nbt
            else:
                factor = self.momentum
        else:
            factor = None
        bn_training = True if self.training else (self.running_mean is None and self.running_var is None)
        ident = self.weight if self.weight is not None else self.running_mean
        if ident is not None:  # (Module._apply replaces Parameter objects on .cuda(): tag whatever object is current)
            ident._bn_sync_id = self._sync_id
        fn = F.batch_norm_relu if fuse_relu else F.batch_norm
        return fn(inpt,
                            self.running_mean if not self.training or self.track_running_stats else None,
                            self.running_var if not self.training or self.track_running_stats else None,
                            self.weight, self.bias, bn_training, factor, self.eps)


class BatchNorm2d(_BatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(input.dim()))


class BatchNorm1d(_BatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 2:
            raise ValueError('expected 2D input (got {}D input)'.format(input.dim()))


__all__ = ['Module', 'Sequential', 'ModuleList', 'Identity', 'ReLU', 'LogSoftmax', 'NLLLoss', 'BCEWithLogitsLoss',
           'Linear', 'Conv2d', 'ConvTranspose2d', 'MaxPool2d', 'BatchNorm2d', 'BatchNorm1d', 'Parameter',
           'set_bn_relu_fusion']

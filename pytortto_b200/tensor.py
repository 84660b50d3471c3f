"""`Tensor`: the user-facing array + autograd handle, API-compatible with the reference's tortto.Tensor for the
surface a CNN training step uses (/root/reference/src/tortto/tensor.py:19-613): `.data` (numpy array on the host,
`cparray` on the device), `.grad` (a RAW array, function.py:88-93), `.grad_fn`, `.requires_grad`, `.cuda()/.cpu()`,
`.backward()`, `.item()`, `.numpy()`, `.detach()`, arithmetic that builds graph nodes.

Host tensors are containers only: this package implements the CUDA path; the reference's numpy path is the oracle,
not a fallback.  Calling a hot-path op on a host tensor raises.
"""
import numpy as np

from .autograd.function import AccumulateGrad
from .autograd.helper import toposort
from .xparray import cparray

float16, float32, float64 = np.float16, np.float32, np.float64
int16, int32, int64 = np.int16, np.int32, np.int64


class Tensor:
    def __init__(self, data, requires_grad=False, dtype=None, copy=True, **kwargs):
        if dtype is None:
            dtype = float32  # the reference defaults everything to float32 (tensor.py:21-22)
        if data.__class__ is cparray:
            if data.dtype != np.dtype(dtype):
                data = data.astype(dtype)
            elif copy:
                data = data.copy()
            self.data = data
        else:
            self.data = np.array(data, dtype=dtype, copy=bool(copy)) if copy else np.asarray(data, dtype=dtype)
        self.grad = None
        self.grad_fn = kwargs.get('grad_fn')
        self._requires_grad = requires_grad
        self._output_idx = kwargs.get('_output_idx')

    # ---- properties ---------------------------------------------------------------------------------------
    @property
    def _version(self):
        d = self.data
        return d._version[0] if d.__class__ is cparray else 0

    @property
    def requires_grad(self):
        return self._requires_grad

    @requires_grad.setter
    def requires_grad(self, val):
        if val.__class__ is not bool:
            raise RuntimeError('requires_grad must be a bool')
        if val and not np.issubdtype(self.data.dtype, np.floating) and not np.issubdtype(self.data.dtype, np.complexfloating):
            raise RuntimeError('only Tensors of floating point and complex dtype can require gradients')
        self._requires_grad = val

    def requires_grad_(self, val=True):
        self.requires_grad = val
        return self

    @property
    def is_cuda(self):
        return self.data.__class__ is cparray

    @property
    def device(self):
        return self.data.device if self.data.__class__ is cparray else 'cpu'

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def ndim(self):
        return len(self.data.shape)

    @property
    def is_leaf(self):
        return self.grad_fn is None

    @property
    def T(self):
        from . import VariableFunctions as V
        if self.ndim != 2:
            raise RuntimeError(f"x.T expects a tensor with 2 dimensions, but self is {self.ndim}D")
        return V.transpose(self, 1, 0)

    def __len__(self):
        return self.shape[0]

    def __hash__(self):
        return id(self)

    def __repr__(self):
        host = self.data.get() if self.is_cuda else self.data
        dev = f", device='{self.device}'" if self.is_cuda else ''
        gf = f', grad_fn=<{self.grad_fn.__class__.__name__}>' if self.grad_fn else ''
        rg = ', requires_grad=True' if self.requires_grad and not self.grad_fn else ''
        return f'tensor({np.array2string(host, separator=", ", precision=4)}{dev}{gf}{rg})'

    # ---- shape helpers ------------------------------------------------------------------------------------
    def dim(self):
        return self.ndim

    def size(self, dim=None):
        if dim is None:
            return self.shape
        n = self.ndim
        if dim < -n or dim > n - 1:
            raise IndexError(f'Dimension out of range (expected to be in range of [-{n}, {n - 1}], but got {dim})')
        return self.shape[dim]

    def numel(self):
        return int(self.data.size)

    def item(self):
        return self.data.item()

    def item_async(self):
        """Extension (not in the reference): start the device->host copy of a one-element CUDA tensor into pinned
        memory on the current stream and return a handle; `handle.get()` waits for that copy only.  Lets a training
        loop read step i's loss after it has queued step i+1 instead of draining the GPU every step."""
        if not self.is_cuda:
            v = self.data.item()
            return _ReadyScalar(v)
        return _PendingScalar(self.data.t)

    def data_ptr(self):
        return self.data.data.ptr if self.is_cuda else self.data.ctypes.data

    def numpy(self):
        if self.is_cuda:
            raise RuntimeError("can't convert cuda tensor to numpy. Use Tensor.cpu() to copy the tensor to host memory first.")
        if self.requires_grad:
            raise RuntimeError("Can't call numpy() on Tensor that requires grad. Use tensor.detach().numpy() instead.")
        return np.asarray(self.data)

    def detach(self):
        return Tensor(self.data, requires_grad=False, dtype=self.data.dtype, copy=False)

    # ---- device movement ----------------------------------------------------------------------------------
    def cuda(self):
        from .autograd.grad_fcn import ToCopy
        return ToCopy.apply(self, target_device='cuda')

    def cpu(self):
        from .autograd.grad_fcn import ToCopy
        return ToCopy.apply(self, target_device='cpu')

    # ---- graph-building operators (only what the conv-net training step and its callers use) ------------------
    def __add__(self, other):
        from . import VariableFunctions as V
        return V.add(self, other)

    __radd__ = __add__

    def __mul__(self, other):
        from . import VariableFunctions as V
        return V.mul(self, other)

    __rmul__ = __mul__

    def __eq__(self, other):
        o = other.data if isinstance(other, Tensor) else other
        return Tensor(self.data == o, dtype=np.bool_, copy=False)

    def sum(self, dim=None, keepdim=False):
        from . import VariableFunctions as V
        return V.sum(self, dim, keepdim)

    def mean(self, dim=None, keepdim=False):
        from . import VariableFunctions as V
        return V.mean(self, dim, keepdim)

    def argmax(self, dim=None):
        d = self.data
        return Tensor(d.argmax(dim), dtype=np.int64, copy=False)

    def view(self, *shape):
        from . import VariableFunctions as V
        return V.reshape(self, *shape)

    reshape = view

    def flatten(self, start_dim=0, end_dim=-1):
        from . import VariableFunctions as V
        return V.flatten(self, start_dim, end_dim)

    def exp(self):
        from . import VariableFunctions as V
        return V.exp(self)

    def copy_(self, other):
        src = other.data if isinstance(other, Tensor) else other
        if self.is_cuda:
            if src.__class__ is not cparray:
                src = cparray.from_numpy(np.asarray(src), dtype=self.data.dtype)
            self.data.t.copy_(src.t.reshape(self.data.t.shape) if src.ndim != self.data.ndim else src.t)
            self.data._version[0] += 1
            self.data._touched()
        else:
            self.data[...] = src.get() if src.__class__ is cparray else np.asarray(src)
        return self

    # ---- backward -----------------------------------------------------------------------------------------
    def backward(self, gradient=None):
        """Reverse-mode sweep with the reference's engine semantics (tensor.py:568-613): seed, visit nodes in
        `toposort` order, hand each produced gradient to its consumer slot (first arrival stored, later ones
        summed), then drop the node's saved state."""
        if not self.requires_grad:
            raise RuntimeError('element 0 of tensors does not require grad and does not have a grad_fn')
        from . import ops
        if not self.is_cuda:
            raise RuntimeError('pytortto_b200 runs backward on the CUDA path only; move the tensor with .cuda()')
        if gradient is None:
            if self.data.size != 1:
                raise RuntimeError('grad can be implicitly created only for scalar outputs')
            gradient = cparray(self.data.t.new_ones(self.data.t.shape))
        elif isinstance(gradient, Tensor):
            if gradient.device != self.device:
                raise RuntimeError(f'invalid gradient at index 0 - expected device {self.device} but got {gradient.device}')
            gradient = gradient.data.copy()
        if tuple(self.data.shape) != tuple(gradient.shape):
            raise RuntimeError('grad can be implicitly created only for scalar outputs')
        if self.grad_fn is None:
            acc = AccumulateGrad()
            acc.variable = self
            acc.grad = [gradient]
            acc.apply(gradient)
            return
        self.grad_fn.grad[self._output_idx] = gradient
        ops.begin_backward_sweep()
        try:
            self._sweep(ops)
        finally:
            ops.end_backward_sweep()

    def _sweep(self, ops):
        for node in toposort(self.grad_fn):
            nf = node.next_functions
            fused0 = False
            fcls = node._forward_cls
            if fcls is not None and getattr(fcls, '_accumulates_input0', False) and node.needs_input_grad[0] \
                    and nf[0][0] is not None:
                # a gradient for input 0 has already arrived through another branch: let this op add it inside its own
                # dx pass (the `grad += new` of tensor.py:597-599 without a separate elementwise kernel)
                fn0, ind0 = nf[0]
                pending = fn0.grad[ind0]
                if pending is not None and pending.__class__ is cparray:
                    node.params['_accum0'] = pending
                    fused0 = True
            grads = node.apply(*node.grad)
            for i in range(len(nf)):
                g = grads[i]
                if g is not None:
                    fn, ind = nf[i]
                    slot = fn.grad
                    if slot[ind] is None or (i == 0 and fused0):
                        slot[ind] = g  # (fused0: g already contains the gradient that was pending in the slot)
                    else:
                        slot[ind] = ops.add_arrays(slot[ind], g)  # out of place: producers may share `g`
            node.clear()


class _ReadyScalar:
    def __init__(self, v):
        self._v = v

    def get(self):
        return self._v


class _PendingScalar:
    _pool = []  # recycled (pinned buffer, event) pairs

    def __init__(self, t):
        import torch
        if t.numel() != 1:
            raise ValueError("only one element tensors can be converted to Python scalars")
        self._buf, self._ev = _PendingScalar._pool.pop() if _PendingScalar._pool else (
            torch.empty(1, dtype=torch.float64).pin_memory(), torch.cuda.Event())
        self._host = self._buf.view(torch.uint8)[:t.element_size()].view(t.dtype)
        self._host.copy_(t.reshape(1), non_blocking=True)
        self._ev.record()

    def get(self):
        self._ev.synchronize()
        v = self._host.item()
        _PendingScalar._pool.append((self._buf, self._ev))
        return v


def tensor(data, requires_grad=False, dtype=None, copy=True, **kwargs):
    """tt.tensor(...) (reference VariableFunctions / tensor.py): dtype follows the data for numpy inputs
    (float64 -> float32 default; integer labels keep their dtype)."""
    if isinstance(data, Tensor):
        data = data.data
    if dtype is None:
        if data.__class__ is cparray or isinstance(data, np.ndarray):
            dtype = data.dtype
            if dtype == np.float64:
                dtype = float32
        else:
            probe = np.asarray(data)
            dtype = float32 if probe.dtype.kind == 'f' else probe.dtype
    return Tensor(data, requires_grad=requires_grad, dtype=dtype, copy=copy, **kwargs)

"""Device-memory layer: `cparray`, the B200 device array behind `Tensor.data` after `.cuda()`.

Mirrors the role of the reference's `cparray(cp.ndarray)` (/root/reference/src/tortto/xparray.py:44-78): a device
buffer with a logical numpy-style shape, a `_version` list shared by views (in-place bookkeeping, xparray.py:65-78),
`.data.ptr` (tensor.py:355-357) and `.get()` (module.py:113-121).  CuPy is not part of this stack: the memory
comes from torch's CUDA caching allocator (plumbing only - allocation, streams, H2D/D2H copies); all hot-path math
runs in libtortto_b200.so on the raw pointers.

HBM layout: 4-D arrays keep the logical (N, C, H, W) shape of the reference but are stored NHWC (channels innermost) -
the same bytes as a torch `channels_last` tensor - so that one pixel's channels are one contiguous TMA row.  Conv
weights (Cout, Cin/g, kh, kw) stored this way are exactly the [K][R][S][C] layout the kernels read.  Arrays of any
other rank are plain C-contiguous.
"""
import numpy as np
import torch

from . import _cabi

_TORCH_TO_NP = {torch.float32: np.dtype("float32"), torch.float64: np.dtype("float64"), torch.int64: np.dtype("int64"),
                torch.int32: np.dtype("int32"), torch.bool: np.dtype("bool"), torch.uint8: np.dtype("uint8"),
                torch.float16: np.dtype("float16")}
_NP_TO_TORCH = {v: k for k, v in _TORCH_TO_NP.items()}


def current_stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("pytortto_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def empty_device(shape, dtype=torch.float32):
    """Uninitialised device storage in the canonical layout for its rank (NHWC for 4-D)."""
    shape = tuple(int(s) for s in shape)
    if len(shape) == 4:
        return torch.empty(shape, dtype=dtype, device=_device(), memory_format=torch.channels_last)
    return torch.empty(shape, dtype=dtype, device=_device())


class _Mem:
    __slots__ = ("ptr",)

    def __init__(self, ptr):
        self.ptr = ptr


# Bumped by every in-place write to a 4-D device array that does not go through a kernel which maintains the array's
# bf16 shadow itself: ops.py re-derives the bf16 copies of the conv weights when it has moved (weights change once per
# optimizer step; anything else that writes a parameter - copy_, load_state_dict, broadcast - lands here too).
mutation_epoch = [0]


class cparray:
    """Device array.  `t` is the owning torch tensor (logical shape, canonical physical layout).
    `_h` / `_hver`: optional bf16 shadow of the same values (same physical layout) and the `_version` it was written
    at - co-written by the producing kernel in bf16 math mode so the next convolution reads 2-byte operands."""
    __slots__ = ("_t", "_thunk", "_version", "base", "_h", "_hver", "_bnstats", "__weakref__")

    # `t` is a property so that an array can be DEFERRED: its storage (`_t`) exists, but the kernel that fills it
    # (`_thunk`) has not been launched yet and runs on the first access to `t`.  ops.conv2d_fprop_deferred uses this to let
    # the residual `Add` that follows a convolution run inside the convolution's epilogue (ops.resolve_pending).
    @property
    def t(self):
        th = self._thunk
        if th is not None:
            self._thunk = None
            th(self)
        return self._t

    @t.setter
    def t(self, value):
        self._t = value

    def __init__(self, t, version=None, base=None):
        if t.__class__ is not torch.Tensor:
            raise TypeError("cparray wraps a torch CUDA tensor; use cparray.from_numpy / tensor(...).cuda()")
        if t.dim() == 4 and not t.is_contiguous(memory_format=torch.channels_last):
            t = t.contiguous(memory_format=torch.channels_last)
        elif t.dim() != 4 and not t.is_contiguous():
            t = t.contiguous()
        self._t = t
        self._thunk = None
        self._version = [0] if version is None else version
        self.base = base
        self._h = None
        self._hver = 0
        # (partials, chunks, version): per-channel sum / sum-of-squares partials of this array, emitted by the kernel that
        # produced it (a conv epilogue, the fused add) for the BatchNorm that reads it next; valid while _version matches
        self._bnstats = None

    def _touched(self):
        """an in-place write happened outside the shadow-maintaining kernels"""
        self._h = None
        self._bnstats = None
        if self._t.dim() == 4:
            mutation_epoch[0] += 1

    # ---- construction / transfer --------------------------------------------------------------------------
    @classmethod
    def from_numpy(cls, a, dtype=None):
        """Host -> device.  4-D float32 arrays are copied as NCHW and re-laid out to NHWC by ttb_nchw_to_nhwc."""
        a = np.asarray(a)
        if dtype is not None:
            a = a.astype(dtype, copy=False)
        a = np.require(a, requirements='C')  # keeps 0-d arrays 0-d (ascontiguousarray would make them 1-d)
        if a.dtype not in _NP_TO_TORCH:
            raise TypeError(f"unsupported dtype {a.dtype}")
        host = torch.from_numpy(a.view(np.ndarray))
        if a.ndim == 4 and a.dtype == np.float32 and a.size > 0:
            staged = host.to(_device(), non_blocking=True)
            out = empty_device(a.shape)
            n, c, h, w = a.shape
            _cabi.call("ttb_nchw_to_nhwc", staged.data_ptr(), out.data_ptr(), n, c, h, w, current_stream_ptr())
            return cls(out)
        return cls(host.to(_device(), non_blocking=host.is_pinned()))  # pinned source: truly asynchronous

    def get(self):
        """Device -> host numpy array in the logical (reference) index order."""
        t = self.t
        if t.dim() == 4 and t.dtype == torch.float32 and t.numel() > 0:
            n, c, h, w = t.shape
            nchw = torch.empty((n, c, h, w), dtype=t.dtype, device=t.device)
            _cabi.call("ttb_nhwc_to_nchw", t.data_ptr(), nchw.data_ptr(), n, c, h, w, current_stream_ptr())
            return nchw.cpu().numpy()
        return t.cpu().numpy() if t.dim() != 4 else t.contiguous().cpu().numpy()

    # ---- numpy-style attributes -----------------------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def ndim(self):
        return self._t.dim()

    @property
    def size(self):
        return self._t.numel()

    @property
    def dtype(self):
        return _TORCH_TO_NP[self._t.dtype]

    @property
    def itemsize(self):
        return self._t.element_size()

    @property
    def nbytes(self):
        return self._t.numel() * self._t.element_size()

    @property
    def strides(self):
        es = self._t.element_size()
        return tuple(s * es for s in self._t.stride())

    @property
    def data(self):
        return _Mem(self.t.data_ptr())

    @property
    def ptr(self):
        return self.t.data_ptr()

    @property
    def device(self):
        return self._t.device

    @property
    def flags(self):
        return {"C_CONTIGUOUS": self._t.is_contiguous(), "OWNDATA": self.base is None}

    def __len__(self):
        return self._t.shape[0]

    def __repr__(self):
        return f"cparray(shape={self.shape}, dtype={self.dtype}, device='{self.device}')"

    # ---- the few array operations host-side code (optimizers, schedulers, user scripts) performs ------------
    def _wrap(self, t, view=False):
        return cparray(t, self._version if view else None, self if view else None)

    def copy(self):
        return cparray(self.t.clone(memory_format=torch.preserve_format))

    def astype(self, dtype, copy=True):
        td = _NP_TO_TORCH[np.dtype(dtype)]
        if td == self.t.dtype and not copy:
            return self
        return cparray(self.t.to(td, copy=True))

    def view(self, cls=None):
        return self

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        src = self.t
        if src.dim() == 4:
            src = src.contiguous()  # logical NCHW order, as numpy's reshape of an NCHW array
            return cparray(src.reshape(shape))
        return cparray(src.reshape(shape), self._version, self)

    def item(self):
        return self.t.item()

    def fill(self, value):
        self.t.fill_(value)
        self._touched()

    def sum(self, axis=None, keepdims=False):
        return cparray(self.t.sum() if axis is None else self.t.sum(dim=axis, keepdim=keepdims))

    def mean(self, axis=None, keepdims=False):
        return cparray(self.t.mean() if axis is None else self.t.mean(dim=axis, keepdim=keepdims))

    def argmax(self, axis=None):
        return cparray(self.t.argmax() if axis is None else self.t.argmax(dim=axis))

    def __getitem__(self, key):
        if isinstance(key, cparray):
            key = key.t
        elif isinstance(key, tuple):
            key = tuple(k.t if isinstance(k, cparray) else k for k in key)
        out = self.t[key]
        return cparray(out) if out.dim() == 4 or not _shares(out, self.t) else cparray(out, self._version, self)

    def __setitem__(self, key, value):
        if isinstance(key, cparray):
            key = key.t
        if isinstance(value, cparray):
            value = value.t
        elif isinstance(value, np.ndarray):
            value = torch.from_numpy(np.ascontiguousarray(value)).to(self.t.device)
        self.t[key] = value
        self._touched()

    @staticmethod
    def _operand(o):
        return o.t if isinstance(o, cparray) else o

    def __add__(self, o):
        return cparray(self.t + self._operand(o))

    __radd__ = __add__

    def __sub__(self, o):
        return cparray(self.t - self._operand(o))

    def __rsub__(self, o):
        return cparray(self._operand(o) - self.t)

    def __mul__(self, o):
        return cparray(self.t * self._operand(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return cparray(self.t / self._operand(o))

    def __neg__(self):
        return cparray(-self.t)

    def __eq__(self, o):
        return cparray(self.t == self._operand(o))

    def __ne__(self, o):
        return cparray(self.t != self._operand(o))

    def __gt__(self, o):
        return cparray(self.t > self._operand(o))

    def __lt__(self, o):
        return cparray(self.t < self._operand(o))

    __hash__ = None

    def __iadd__(self, o):
        self.t.add_(self._operand(o))
        self._touched()
        return self

    def __isub__(self, o):
        self.t.sub_(self._operand(o))
        self._touched()
        return self

    def __imul__(self, o):
        self.t.mul_(self._operand(o))
        self._touched()
        return self

    def __itruediv__(self, o):
        self.t.div_(self._operand(o))
        self._touched()
        return self


def _shares(a, b):
    try:
        return a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr()
    except Exception:
        return False


def new_like(x, shape=None):
    """Fresh uninitialised float32 device array (canonical layout)."""
    return cparray(empty_device(x.shape if shape is None else shape, x.t.dtype if shape is None else torch.float32))


def new_f32(shape):
    return cparray(empty_device(shape, torch.float32))


def zeros_f32(shape):
    a = new_f32(shape)
    a.t.zero_()
    return a

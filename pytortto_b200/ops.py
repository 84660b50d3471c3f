"""Array-level operators: one Python function per C-ABI kernel family, taking and returning `cparray`s.

This is the layer the autograd Functions (autograd/grad_nn.py, grad_fcn.py) call; it owns output / workspace
allocation (the C ABI never allocates) and descriptor construction.  Every function enqueues on torch's current
CUDA stream and never synchronises.
"""
import ctypes
import math
import os
import weakref

import torch

from . import _cabi
from . import xparray as _xp
from .xparray import cparray, current_stream_ptr, empty_device, new_f32

_MATH_MODES = {"fp32": _cabi.TTB_MATH_FP32, "tf32": _cabi.TTB_MATH_TF32, "bf16": _cabi.TTB_MATH_BF16}
_math_mode = _MATH_MODES[os.environ.get("TORTTO_B200_MATH", "tf32").lower()]


def set_math_mode(mode):
    """'tf32' (default: tcgen05 kind::tf32, fp32 accumulate), 'fp32' (exact CUDA-core path) or 'bf16'."""
    global _math_mode
    _math_mode = _MATH_MODES[mode.lower()]


def get_math_mode():
    return {v: k for k, v in _MATH_MODES.items()}[_math_mode]


def _ptr(a):
    return None if a is None else a.t.data_ptr()


# ---------------------------------------------------------------------------------------------------------
# bf16 math mode: operands live in HBM as bf16 "shadows"
# ---------------------------------------------------------------------------------------------------------
# In bf16 mode the tensor-core kernels read 2-byte operands.  Converting x / w / dy inside every conv call costs three
# extra HBM passes per call (what round 1 did: bf16 was no faster than TF32).  Instead
#   * activations and gradients: the PRODUCING kernel (BatchNorm apply (+ReLU), BatchNorm backward apply, ReLU)
#     co-writes a bf16 copy next to its fp32 output (`cparray._h`); a conv whose operand has no valid shadow converts it
#     once with ttb_to_bf16 and caches the copy on the array (fprop's conversion of x is reused by wgrad);
#   * weights: every conv weight seen so far is re-packed ([K][R][S][C] for fprop, [C][R][S][K] for dgrad, both bf16) by
#     ONE multi-tensor launch when the first conv after a weight change runs (`xparray.mutation_epoch`: optimizer
#     steps, copy_, load_state_dict, broadcast).
# fp32 tensors stay what the API shows; accumulation and every output stay fp32.
def _bf16_mode():
    return _math_mode == _cabi.TTB_MATH_BF16


def _shadow_wanted(x):
    return _math_mode == _cabi.TTB_MATH_BF16 and x.ndim == 4 and x.shape[1] % 64 == 0 and x.t.dtype == torch.float32


def _new_shadow(a):
    h = torch.empty_like(a.t, dtype=torch.bfloat16)
    a._h, a._hver = h, a._version[0]
    return h


def bf16_of(a):
    """bf16 copy of a device array in the same physical layout: the producer's shadow, else converted now and cached."""
    h = a._h
    if h is not None and a._hver == a._version[0]:
        return h
    h = _new_shadow(a)
    _cabi.call("ttb_to_bf16", a.t.data_ptr(), h.data_ptr(), a.size, current_stream_ptr())
    return h


_bf16_ok = {}


def conv_bf16_supported(d, pass_):
    key = (id(d), pass_)
    ok = _bf16_ok.get(key)
    if ok is None:
        ok = _bf16_ok[key] = (d, bool(_cabi.load().ttb_conv2d_bf16_supported(ctypes.byref(d), pass_)))  # keeps d alive
    return ok[1]


_tf32_twins = {}


def _tf32_twin(d):
    """the same problem in TF32 mode (bf16-mode layers whose channel counts do not fit 64-channel K-blocks, e.g. the
    3-channel stem, run on the TF32 tensor path instead of a 20x zero-padded bf16 one)"""
    t = _tf32_twins.get(id(d))
    if t is None:
        twin = _cabi.ConvDesc.from_buffer_copy(d)
        twin.math_mode = _cabi.TTB_MATH_TF32
        t = _tf32_twins[id(d)] = (d, twin)
    return t[1]


class _WeightPack:
    __slots__ = ("ref", "desc", "wh", "wth", "epoch")


_wpack = {}  # weight data_ptr -> _WeightPack


def weights_changed():
    """Call after writing conv weights through anything but cparray / Tensor methods (optimizer kernels do)."""
    _xp.mutation_epoch[0] += 1


def _weight_bf16(w, transposed):
    epoch = _xp.mutation_epoch[0]
    ptr = w.t.data_ptr()
    ent = _wpack.get(ptr)
    if ent is None or ent.ref() is not w or ent.wh.numel() != w.size:
        ent = _WeightPack()
        ent.ref = weakref.ref(w)
        k, c, r, s_ = w.shape
        ent.desc = _cabi.ConvDesc(1, c, r, s_, k, r, s_, 1, 1, 0, 0, 1, 1, 1, 1, 1, _cabi.TTB_MATH_BF16)
        ent.wh = torch.empty(w.size, dtype=torch.bfloat16, device=w.t.device)
        ent.wth = torch.empty(w.size, dtype=torch.bfloat16, device=w.t.device)
        ent.epoch = -1
        _wpack[ptr] = ent
    if ent.epoch != epoch:  # re-pack every stale weight that is still alive in one launch
        stale = []
        for p_, e in list(_wpack.items()):
            if e.ref() is None:
                del _wpack[p_]
            elif e.epoch != epoch:
                stale.append((p_, e))
        n = len(stale)
        descs = (ctypes.POINTER(_cabi.ConvDesc) * n)()
        src = (ctypes.c_void_p * n)()
        dst = (ctypes.c_void_p * n)()
        dstt = (ctypes.c_void_p * n)()
        for i, (p_, e) in enumerate(stale):
            descs[i] = ctypes.pointer(e.desc)
            src[i] = p_
            dst[i] = e.wh.data_ptr()
            dstt[i] = e.wth.data_ptr()
            e.epoch = epoch
        _cabi.call("ttb_conv2d_pack_weights_bf16", n, descs, src, dst, dstt, current_stream_ptr())
    return ent.wth if transposed else ent.wh


def _workspace(nbytes):
    if nbytes <= 0:
        return None, 0
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
    return ws, int(nbytes)


# ---------------------------------------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------------------------------------
def conv_out_hw(h, w, kh, kw, stride, padding, dilation):
    """reference autograd/grad_nn.py:541-542"""
    p = math.floor((h + 2 * padding[0] - dilation[0] * (kh - 1) - 1) / stride[0] + 1)
    q = math.floor((w + 2 * padding[1] - dilation[1] * (kw - 1) - 1) / stride[1] + 1)
    return p, q


_desc_cache = {}


def conv_desc(x_shape, w_shape, stride, padding, dilation, groups, out_hw=None):
    key = (x_shape, w_shape, stride, padding, dilation, groups, out_hw, _math_mode)
    d = _desc_cache.get(key)
    if d is None:
        n, c, h, w = x_shape
        k, _, r, s = w_shape
        p, q = conv_out_hw(h, w, r, s, stride, padding, dilation) if out_hw is None else out_hw
        d = _cabi.ConvDesc(n, c, h, w, k, r, s, stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1],
                           groups, p, q, _math_mode)
        _desc_cache[key] = d
    return d


# Fused conv epilogues (tensor path): per-channel scale / bias, a residual add, ReLU, and the per-channel sum / sum of
# squares of the stored output for the BatchNorm that reads it next (`cparray._bnstats`).  TORTTO_B200_EPILOGUE_STATS=0
# switches the statistics off (every BatchNorm then runs its own statistics pass).
_EPILOGUE_STATS = os.environ.get("TORTTO_B200_EPILOGUE_STATS", "1") != "0"
_fused_info = {}


def conv_fused_info(d):
    """(fused epilogue available, statistics chunks) of fprop of this problem"""
    ent = _fused_info.get(id(d))
    if ent is None:
        lib = _cabi.load()
        ok = bool(lib.ttb_conv2d_fused_epilogue_supported(ctypes.byref(d)))
        ent = _fused_info[id(d)] = (d, ok, int(lib.ttb_conv2d_fprop_stats_chunks(ctypes.byref(d))) if ok else 0)  # keeps d alive
    return ent[1], ent[2]


def conv2d_fprop(x, w, bias, d, scale=None, residual=None, relu=False, stats=False, out=None):
    """y = conv(x, w) [* scale[k]] [+ bias[k]] [+ residual] [max(., 0)]; `stats`: also emit the BatchNorm statistics
    partials of y (y._bnstats).  scale / residual / relu / stats need conv_fused_info(d)[0].  `out`: write into this array."""
    y = new_f32((d.n, d.k, d.p, d.q)) if out is None else out
    if y.size == 0:
        return y
    bf16_direct = d.math_mode == _cabi.TTB_MATH_BF16 and conv_bf16_supported(d, 0) and w.ndim == 4
    if d.math_mode == _cabi.TTB_MATH_BF16 and not bf16_direct:
        d = _tf32_twin(d)
    fused = scale is not None or residual is not None or relu or stats
    ep = None
    if fused:
        ok, chunks = conv_fused_info(d)
        if not ok:
            raise RuntimeError("conv2d_fprop: this problem has no fused epilogue (exact fp32 / grouped path)")
        partials = None
        if stats and chunks > 0 and _EPILOGUE_STATS:
            partials = torch.empty((chunks, 2, d.k), dtype=torch.float64, device=y.t.device)
            y._bnstats = (partials, chunks, y._version[0])
        ep = _cabi.ConvEpilogue(_ptr(scale), _ptr(bias), _ptr(residual), int(bool(relu)),
                                None if partials is None else partials.data_ptr())
    if bf16_direct:
        if ep is None:
            ep = _cabi.ConvEpilogue(None, _ptr(bias), None, 0, None)
        _cabi.call("ttb_conv2d_fprop_bf16", ctypes.byref(d), bf16_of(x).data_ptr(), _weight_bf16(w, False).data_ptr(),
                   ctypes.byref(ep), _ptr(y), current_stream_ptr())
        return y
    ws, nb = _workspace(_cabi.load().ttb_conv2d_workspace_size(ctypes.byref(d), 0))
    if ep is not None:
        _cabi.call("ttb_conv2d_fprop_fused", ctypes.byref(d), _ptr(x), _ptr(w), ctypes.byref(ep), _ptr(y),
                   None if ws is None else ws.data_ptr(), nb, current_stream_ptr())
    else:
        _cabi.call("ttb_conv2d_fprop", ctypes.byref(d), _ptr(x), _ptr(w), _ptr(bias), _ptr(y),
                   None if ws is None else ws.data_ptr(), nb, current_stream_ptr())
    return y


# Deferred producers ("the next operator decides the epilogue").
# A residual block ends with `conv(...) + shortcut` (pre-activation: reference examples' BasicBlock.forward) or with
# `relu(bn(conv(...)) + identity)` (post-activation: Bottleneck.forward of the ResNet-50 notebook).  The conv epilogue can
# add the shortcut (and emit the next BatchNorm's statistics of the SUM), and BatchNorm's normalise pass can add the
# identity and apply the ReLU - each fusion removes full HBM passes - but an operator's forward does not know what follows
# it.  So such a producer is DEFERRED: its output array exists, the launch is a thunk on the array (`cparray._thunk`, run by
# the first access to its storage) and `_pending` remembers it.  The very next operator (autograd Function.apply ->
# resolve_pending) either ABSORBS it -
#     Add  <- deferred convolution      : ONE launch conv + shortcut (+ statistics) into the Add's output
#     Add  <- deferred BatchNorm apply  : the sum is itself deferred ("bn_add": y = bn(x) + other)
#     ReLU <- deferred bn_add / bn      : ONE normalise pass y = max(bn(x) [+ other], 0) into the ReLU's output
# and the absorbed array stays unmaterialised (anything that still touches it later gets the plain producer) - or it is
# anything else, and the producer is launched first, in program order.  Nothing can run between the deferral and that
# decision except direct array accesses, which materialise.  Gradients are unaffected: none of the absorbed intermediate
# values is saved for backward (Convolution saves its input and weight, BatchNorm its input, Add nothing, ReLU its output).
_pending = [None]
_DEFER = os.environ.get("TORTTO_B200_DEFER", "1") != "0"  # (TORTTO_B200_DEFER=0: every producer launches immediately)


class _Job:
    """what a deferred launch needs; `versions`: (array, version) pairs that must be unchanged when it finally runs"""
    __slots__ = ("kind", "args", "versions")

    def check(self):
        for arr, ver in self.versions:
            if arr._version[0] != ver:
                raise RuntimeError("an operand of a deferred convolution / batch-norm was modified in place before its output was read")


def _defer(y, kind, args, operands, run):
    """mark `y` as produced by `run(y, *args)` later; returns y"""
    resolve_pending(None, ())
    job = _Job()
    job.kind, job.args = kind, args
    job.versions = tuple((a, a._version[0]) for a in operands if a is not None)

    def launch(arr, job=job, run=run):
        if _pending[0] is not None and _pending[0]() is arr:
            _pending[0] = None
        job.check()
        run(arr, *job.args)

    launch.job = job
    y._thunk = launch
    _pending[0] = weakref.ref(y)
    return y


def pending_array():
    ref = _pending[0]
    if ref is None:
        return None
    arr = ref()
    if arr is None or arr._thunk is None:
        _pending[0] = None
        return None
    return arr


def resolve_pending(op_cls, inputs):
    """Called before every operator's forward (and at the start of backward).  Returns the deferred array if `op_cls`
    (a Function class with `_absorbs` = kinds it can take) consumes it as one of `inputs` and can absorb it; otherwise
    launches the deferred producer (program order) and returns None."""
    arr = pending_array()
    if arr is None:
        return None
    kinds = getattr(op_cls, "_absorbs", ()) if op_cls is not None else ()
    if arr._thunk.job.kind in kinds:
        datas = [i.data for i in inputs if i is not None]
        if any(d is arr for d in datas):
            others = [d for d in datas if d is not arr]
            if all(o.__class__ is cparray and o._thunk is None and o.shape == arr.shape and o._t.dtype == torch.float32
                   for o in others):
                return arr
    arr.t  # noqa: B018 - materialises (program order)
    return None


def conv2d_fprop_deferred(x, w, bias, d, stats):
    """conv2d_fprop whose launch waits for the next operator (see above); an immediate launch when the problem has no fused
    epilogue."""
    if not _DEFER or not conv_fused_info(d)[0] or d.n * d.k * d.p * d.q == 0:
        return conv2d_fprop(x, w, bias, d, stats=stats)
    y = new_f32((d.n, d.k, d.p, d.q))
    return _defer(y, "conv", (x, w, bias, d, stats), (x, w, bias),
                  lambda arr, x, w, bias, d, stats: conv2d_fprop(x, w, bias, d, stats=stats, out=arr))


def conv_add_fused(deferred, other, stats):
    """conv(x, w) (+ bias) + other, the sum's BatchNorm statistics emitted by the same epilogue; `deferred` stays deferred"""
    job = deferred._thunk.job
    _pending[0] = None
    job.check()
    x, w, bias, d, _ = job.args
    return conv2d_fprop(x, w, bias, d, residual=other, stats=stats)


def _bn_apply_launch(y, x, other, m, c, stats_rows, relu):
    """y = bn(x) [+ other] [max(., 0)] with the statistics block `stats_rows` ([5][C]: mean, var+eps, sd, scale, shift)"""
    base = stats_rows.t.data_ptr()
    row = c * 4
    yh = _new_shadow(y).data_ptr() if _shadow_wanted(y) else None
    if other is None:
        _cabi.call("ttb_bn_apply", _ptr(x), _ptr(y), m, c, base, base + 3 * row, base + 4 * row, int(relu), yh,
                   current_stream_ptr())
    else:
        _cabi.call("ttb_bn_apply_add", _ptr(x), _ptr(other), _ptr(y), m, c, base, base + 3 * row, base + 4 * row, int(relu),
                   yh, current_stream_ptr())


def bn_apply_deferred(x, m, c, stats_rows, relu):
    """the normalise pass of a BatchNorm: deferred when a following Add / ReLU could absorb it (4-D, no fused ReLU)"""
    y = cparray(empty_device(x.shape))
    if relu or not _DEFER or x.ndim != 4 or c % 4 != 0 or x.size == 0:
        _bn_apply_launch(y, x, None, m, c, stats_rows, relu)
        return y
    return _defer(y, "bn", (x, None, m, c, stats_rows, False), (x,), _bn_apply_launch)


def bn_add_deferred(deferred, other):
    """Add absorbing a deferred BatchNorm apply: the sum bn(x) + other, itself deferred (a ReLU may follow)"""
    job = deferred._thunk.job
    _pending[0] = None
    job.check()
    x, _, m, c, stats_rows, _ = job.args
    y = cparray(empty_device(x.shape))
    return _defer(y, "bn_add", (x, other, m, c, stats_rows, False), (x, other), _bn_apply_launch)


def bn_relu_fused(deferred, out=None):
    """ReLU absorbing a deferred bn / bn_add: y = max(bn(x) [+ other], 0) in one pass.  `out`: write in place into the
    deferred array itself (ReLU(inplace=True))."""
    job = deferred._thunk.job
    _pending[0] = None
    job.check()
    x, other, m, c, stats_rows, _ = job.args
    if out is not None:
        out._thunk = None
        y = out
        y._h = None
    else:
        y = cparray(empty_device(x.shape))
    _bn_apply_launch(y, x, other, m, c, stats_rows, True)
    return y


def conv2d_bn_eval(x, w, conv_bias, d, mean, var, eps, gamma, beta, relu):
    """relu?(batch_norm_eval(conv(x, w) + conv_bias)) as ONE kernel: the BatchNorm is folded to the epilogue's per-channel
    scale / bias (inference; needs conv_fused_info(d)[0])"""
    c = d.k
    sb = new_f32((2, c))
    base = sb.t.data_ptr()
    _cabi.call("ttb_bn_fold_eval", _ptr(mean), _ptr(var), c, float(eps), _ptr(gamma), _ptr(beta), _ptr(conv_bias), base,
               base + 4 * c, current_stream_ptr())
    scale, bias = cparray(sb.t[0]), cparray(sb.t[1])
    return conv2d_fprop(x, w, bias, d, scale=scale, relu=relu)


def bn_stats_of(x):
    """(partials tensor, chunks) a producer attached to x, if still valid for its current contents"""
    bs = x._bnstats
    if bs is not None and bs[2] == x._version[0]:
        return bs[0], bs[1]
    return None


# dgrad needs the filters as [C][R][S][K]; the [K][R][S][C] -> [C][R][S][K] re-ordering of ALL conv layers of a step is
# one multi-tensor launch: Convolution.forward registers (weight, descriptor), the first dgrad of a backward sweep
# packs everything registered since the last sweep, and the packed copies are trusted for that sweep only (weights do
# not change inside a backward sweep; nothing else is assumed about when they change).  Measured on one B200, same
# box: 66.86 k images/s with the batched re-ordering vs 66.07 k with one small launch per layer.
_dgrad_pack = {"registry": {}, "packed": {}, "sweep": 0, "in_sweep": False,
               "enabled": os.environ.get("TORTTO_B200_BATCHED_HELPERS", "1") != "0"}
# (issuing the re-ordering launch from the wgrad stream at the start of the sweep, to run under the head's backward kernels,
# measured no gain: 3.340 vs 3.331 ms per step - it is not on the critical path)
_prepack_ok = {}


def begin_backward_sweep():
    resolve_pending(None, ())
    _dgrad_pack["sweep"] += 1
    _dgrad_pack["in_sweep"] = True


def end_backward_sweep():
    resolve_pending(None, ())  # a dgrad whose launch was deferred for a BatchNorm backward that never came
    _dgrad_pack["in_sweep"] = False
    _dgrad_pack["registry"].clear()  # (weights registered by a forward whose dgrad never ran must not be kept alive)
    packed = _dgrad_pack["packed"]
    if len(packed) > 64:  # drop the packed copies of weights that no longer exist (models rebuilt in one process)
        for ptr in [p_ for p_, e in packed.items() if e[2]() is None]:
            del packed[ptr]
    join_wgrad()  # weight gradients computed on the second stream are complete for whoever runs next


def register_dgrad_weight(w, d):
    if not _dgrad_pack["enabled"]:
        return
    ok = _prepack_ok.get(id(d))
    if ok is None:
        ok = _prepack_ok[id(d)] = bool(_cabi.load().ttb_conv2d_dgrad_prepacked_supported(ctypes.byref(d)))
    if ok:
        _dgrad_pack["registry"][w.t.data_ptr()] = (w, d)


def _flush_dgrad_pack():
    reg = _dgrad_pack["registry"]
    if not reg:
        return
    items = list(reg.values())
    reg.clear()
    n = len(items)
    descs = (ctypes.POINTER(_cabi.ConvDesc) * n)()
    src = (ctypes.c_void_p * n)()
    dst = (ctypes.c_void_p * n)()
    for i, (w, d) in enumerate(items):
        ptr = w.t.data_ptr()
        ent = _dgrad_pack["packed"].get(ptr)
        if ent is None or ent[0].numel() != w.t.numel() or ent[2]() is not w:
            ent = [torch.empty(w.t.numel(), dtype=torch.float32, device=w.t.device), -1, weakref.ref(w)]
            _dgrad_pack["packed"][ptr] = ent
        ent[1] = _dgrad_pack["sweep"]
        descs[i] = ctypes.pointer(d)
        src[i] = ptr
        dst[i] = ent[0].data_ptr()
    _cabi.call("ttb_conv2d_dgrad_pack_weights", n, descs, src, dst, current_stream_ptr())


_dgrad_bn_chunks = {}  # id(descriptor) -> rows of the partial buffer a dgrad + BatchNorm-backward-statistics launch writes
# Measured on B200 (profiles/r2_dgrad_bn_ab.txt): the fused launch saves the statistics pass (preact_resnet18: 0.29 -> 0.11
# ms of ttb_bn_bwd_reduce per step) but its longer epilogue costs the dgrad kernels about as much (0.77 -> 0.95 ms); net
# 3.098 -> 3.075 ms per step in TF32 mode (fewer launches), 2.636 -> 2.670 ms in bf16 mode, where the shorter main loop no
# longer hides the epilogue.  The whole fusion is therefore opt-in (TORTTO_B200_DGRAD_BN=1, grad_nn._BatchNormBase), and
# bf16 problems additionally need TORTTO_B200_DGRAD_BN_BF16=1.
_DGRAD_BN_BF16 = [os.environ.get("TORTTO_B200_DGRAD_BN_BF16", "0") != "0"]


def _dgrad_tensor_operands(dy, w, d):
    """(dy pointer, packed weight pointer) when this dgrad runs on the tensor path from ready-made operands - bf16 shadows,
    or fp32 with the weights pre-packed for this backward sweep - else None"""
    if d.math_mode == _cabi.TTB_MATH_BF16:
        if conv_bf16_supported(d, 1) and w.ndim == 4:
            return bf16_of(dy).data_ptr(), _weight_bf16(w, True).data_ptr()
        return None
    if _dgrad_pack["enabled"] and _dgrad_pack["in_sweep"]:  # (a dgrad outside backward = ConvTranspose2d forward)
        _flush_dgrad_pack()
        ent = _dgrad_pack["packed"].get(w.t.data_ptr())
        if ent is not None and ent[1] == _dgrad_pack["sweep"] and _prepack_ok.get(id(d)):
            return _ptr(dy), ent[0].data_ptr()
    return None


def conv2d_dgrad_for_batchnorm(dy, w, d, accum=None):
    """conv2d_dgrad whose consumer is a BatchNorm(+ReLU) backward node: the launch is deferred (see `_defer`) so that
    `bn_backward` can run it with the epilogue that also emits the two sums of the BatchNorm backward over dx
    (ttb_conv2d_dgrad_bn) - the statistics pass (two full reads) disappears.  Anything else that touches the array first
    gets the plain dgrad."""
    dx_shape = (d.n, d.c, d.h, d.w)
    if not _DEFER or not _dgrad_pack["in_sweep"] or d.n * d.c * d.h * d.w == 0 or \
            (d.math_mode == _cabi.TTB_MATH_BF16 and not _DGRAD_BN_BF16[0]) or \
            (accum is not None and (accum.shape != dx_shape or accum.t.dtype != torch.float32 or accum._thunk is not None)):
        return conv2d_dgrad(dy, w, d, accum)
    chunks = _dgrad_bn_chunks.get(id(d))
    if chunks is None:
        chunks = _dgrad_bn_chunks[id(d)] = int(_cabi.load().ttb_conv2d_dgrad_bn_stats_chunks(ctypes.byref(d)))
    if chunks <= 0 or (d.math_mode != _cabi.TTB_MATH_BF16 and not (_dgrad_pack["enabled"] and _prepack_ok.get(id(d)))):
        return conv2d_dgrad(dy, w, d, accum)
    if _dgrad_tensor_operands(dy, w, d) is None:  # (the operands the fused launch needs exist now, hence also later)
        return conv2d_dgrad(dy, w, d, accum)
    dx = new_f32(dx_shape)
    return _defer(dx, "dgrad", (dy, w, d, accum), (dy, w, accum),
                  lambda arr, dy, w, d, accum: conv2d_dgrad(dy, w, d, accum, out=arr))


def dgrad_bn_fused(deferred, x, mean_ptr, rscale_ptr, rshift_ptr):
    """BatchNorm backward absorbing the deferred dgrad that produces its incoming gradient: ONE conv launch writes
    `deferred` and the [chunks][2][C] partial sums; returns (partials, chunks) or None (then the caller materialises)."""
    job = deferred._thunk.job
    dy, w, d, accum = job.args
    if tuple(x.shape) != tuple(deferred.shape) or x.__class__ is not cparray or x.t.dtype != torch.float32:
        return None
    ops_ptrs = _dgrad_tensor_operands(dy, w, d)
    if ops_ptrs is None:
        return None
    _pending[0] = None
    job.check()
    chunks = _dgrad_bn_chunks[id(d)]
    partials = torch.empty((chunks, 2, d.c), dtype=torch.float64, device=x.t.device)
    bn = _cabi.DgradBnStats(_ptr(x), mean_ptr, rscale_ptr, rshift_ptr, partials.data_ptr())
    deferred._thunk = None
    deferred._h = None
    _cabi.call("ttb_conv2d_dgrad_bn", ctypes.byref(d), ops_ptrs[0], ops_ptrs[1], _ptr(accum), deferred._t.data_ptr(),
               ctypes.byref(bn), current_stream_ptr())
    return partials, chunks


def conv2d_dgrad(dy, w, d, accum=None, out=None):
    """`accum`: a gradient already pending for the same tensor (the engine's `grad += new`): added in the dgrad epilogue
    where the kernel can (pre-packed TF32 / bf16 tensor path), by a separate add otherwise.  `out`: the array to fill (a
    deferred launch finally running)."""
    dx = out if out is not None else new_f32((d.n, d.c, d.h, d.w))
    if dx.size == 0:
        return dx if accum is None else accum
    if accum is not None and (accum.shape != dx.shape or accum.t.dtype != torch.float32):
        return add_arrays(conv2d_dgrad(dy, w, d), accum)
    ops_ptrs = _dgrad_tensor_operands(dy, w, d)
    if ops_ptrs is not None:
        _cabi.call("ttb_conv2d_dgrad_bf16" if d.math_mode == _cabi.TTB_MATH_BF16 else "ttb_conv2d_dgrad_prepacked",
                   ctypes.byref(d), ops_ptrs[0], ops_ptrs[1], _ptr(accum), _ptr(dx), current_stream_ptr())
        return dx
    if d.math_mode == _cabi.TTB_MATH_BF16:
        d = _tf32_twin(d)
        ops_ptrs = _dgrad_tensor_operands(dy, w, d)
        if ops_ptrs is not None:
            _cabi.call("ttb_conv2d_dgrad_prepacked", ctypes.byref(d), ops_ptrs[0], ops_ptrs[1], _ptr(accum), _ptr(dx),
                       current_stream_ptr())
            return dx
    ws, nb = _workspace(_cabi.load().ttb_conv2d_workspace_size(ctypes.byref(d), 1))
    _cabi.call("ttb_conv2d_dgrad", ctypes.byref(d), _ptr(dy), _ptr(w), _ptr(dx),
               None if ws is None else ws.data_ptr(), nb, current_stream_ptr())
    if accum is None:
        return dx
    res = add_arrays(dx, accum)
    if out is not None:  # (a deferred launch must fill the array it was deferred on)
        out.t.copy_(res.t)
        return out
    return res


# Weight gradients are leaves of the backward pass: nothing downstream of a conv's backward needs dW before the
# optimizer (or the gradient all-reduce), while dX is on the critical path.  With `overlap=True` the wgrad kernels go to
# a second stream that forks from the current one, so the tensor-bound wgrad runs under the HBM-bound BatchNorm / ReLU
# backward kernels of the next layers; `join_wgrad()` (end of Tensor.backward, or whoever reads a gradient earlier)
# makes the current stream wait for it.  Operands are kept referenced until the join so the caching allocator cannot
# hand their memory out again; the fork / join also works under CUDA-graph capture (a parallel branch of the graph).
_overlap = {"enabled": os.environ.get("TORTTO_B200_WGRAD_OVERLAP", "1") != "0", "stream": {}, "keep": [], "dirty": False,
            # one multi-tensor split reduction at the join instead of one per layer: measured SLOWER (65.5 k vs 66.1 k
            # images/s) - by then the partial sums have left the L2 - so it stays a switch, off by default
            "defer_sums": os.environ.get("TORTTO_B200_DEFER_SPLIT_SUMS", "0") != "0",
            "sums": []}  # sums: (partials ptr, splits, size, dw ptr) of wgrads whose split reduction is still owed


def set_wgrad_overlap(flag):
    join_wgrad()
    _overlap["enabled"] = bool(flag)


def _side_stream():
    dev = torch.cuda.current_device()
    st = _overlap["stream"].get(dev)
    if st is None:
        st = _overlap["stream"][dev] = torch.cuda.Stream(device=dev)
    return st


def mark_side_stream_used():
    """work other than a wgrad was queued on the side stream (the DP layer issues bucket all-reduces from it)"""
    _overlap["dirty"] = True


def join_wgrad():
    if _overlap["dirty"]:
        side = _side_stream()
        sums = _overlap["sums"]
        if sums:  # the split reductions of every wgrad since the last join: one multi-tensor launch on the wgrad stream
            n = len(sums)
            part = (ctypes.c_void_p * n)(*[t[0] for t in sums])
            spl = (ctypes.c_int * n)(*[t[1] for t in sums])
            siz = (ctypes.c_int64 * n)(*[t[2] for t in sums])
            out = (ctypes.c_void_p * n)(*[t[3] for t in sums])
            _cabi.call("ttb_sum_splits_multi", n, part, spl, siz, out, side.cuda_stream)
            sums.clear()
        torch.cuda.current_stream().wait_stream(side)
        _overlap["keep"].clear()
        _overlap["dirty"] = False


def conv2d_wgrad(x, dy, d, overlap=False, out=None):
    """`out`: where the caller wants dW (a parameter's slot in a flat gradient bucket of the data-parallel layer)"""
    dw = out if out is not None else new_f32((d.k, d.c // d.groups, d.r, d.s))
    if dw.size == 0:
        return dw
    if d.math_mode == _cabi.TTB_MATH_BF16:
        if conv_bf16_supported(d, 2):
            # operand copies are made (when the producers did not co-write them) on the CURRENT stream, before the fork:
            # the dgrad that follows on this stream reuses dy's
            xh, dyh = bf16_of(x), bf16_of(dy)
            ws, nb = _workspace(_cabi.load().ttb_conv2d_workspace_size_bf16(ctypes.byref(d), 2))
            st = current_stream_ptr()
            if overlap and _overlap["enabled"]:
                side = _side_stream()
                side.wait_stream(torch.cuda.current_stream())
                st = side.cuda_stream
                _overlap["keep"].append((x, dy, dw, ws, xh, dyh))
                _overlap["dirty"] = True
            _cabi.call("ttb_conv2d_wgrad_bf16", ctypes.byref(d), xh.data_ptr(), dyh.data_ptr(), _ptr(dw),
                       None if ws is None else ws.data_ptr(), nb, st)
            return dw
        d = _tf32_twin(d)
    ws, nb = _workspace(_cabi.load().ttb_conv2d_workspace_size(ctypes.byref(d), 2))
    if overlap and _overlap["enabled"]:
        side = _side_stream()
        side.wait_stream(torch.cuda.current_stream())
        if _overlap["defer_sums"]:
            splits, partials = ctypes.c_int(0), ctypes.c_void_p(0)
            _cabi.call("ttb_conv2d_wgrad_partial", ctypes.byref(d), _ptr(x), _ptr(dy), _ptr(dw),
                       None if ws is None else ws.data_ptr(), nb, ctypes.byref(splits), ctypes.byref(partials), side.cuda_stream)
            if splits.value > 1:
                _overlap["sums"].append((partials.value, splits.value, dw.size, dw.t.data_ptr()))
        else:
            _cabi.call("ttb_conv2d_wgrad", ctypes.byref(d), _ptr(x), _ptr(dy), _ptr(dw),
                       None if ws is None else ws.data_ptr(), nb, side.cuda_stream)
        _overlap["keep"].append((x, dy, dw, ws))
        _overlap["dirty"] = True
        return dw
    _cabi.call("ttb_conv2d_wgrad", ctypes.byref(d), _ptr(x), _ptr(dy), _ptr(dw),
               None if ws is None else ws.data_ptr(), nb, current_stream_ptr())
    return dw


def bias_grad(dy, out=None, overlap=False):
    """dy (N,K,P,Q) -> (K,)   (gd0.sum((0,2,3)), reference grad_nn.py:727-728).  `overlap`: the result goes straight to a
    leaf, so - like the weight gradient - this full HBM read of dy leaves the critical path and runs on the wgrad stream
    (UNet at 8 x 512 x 512: 23 biased convolutions, 0.8 ms of bias gradients per step)."""
    n, k, p, q = dy.shape
    db = out if out is not None else new_f32((k,))
    st = current_stream_ptr()
    if overlap and _overlap["enabled"] and db.size:
        side = _side_stream()
        side.wait_stream(torch.cuda.current_stream())
        st = side.cuda_stream
        _overlap["keep"].append((dy, db))
        _overlap["dirty"] = True
    _cabi.call("ttb_bias_grad", _ptr(dy), _ptr(db), n * p * q, k, st)
    return db


def add_bias_(y, bias):
    """y (N,K,P,Q) += bias[:, None, None] in place (used where the bias cannot ride the conv epilogue)."""
    n, k, p, q = y.shape
    _cabi.call("ttb_add_bias", _ptr(y), _ptr(bias), n * p * q, k, current_stream_ptr())
    y._h = None
    return y


# ---------------------------------------------------------------------------------------------------------
# head ops: global mean, Linear, LogSoftmax, NLL, BCE-with-logits, channel cat / split
# ---------------------------------------------------------------------------------------------------------
def _f32_dense(*arrays):
    return all(a is not None and a.__class__ is cparray and a.t.dtype == torch.float32 for a in arrays)


def mean_hw(x, keepdim):
    n, c, h, w = x.shape
    y = new_f32((n, c, 1, 1) if keepdim else (n, c))
    _cabi.call("ttb_mean_hw_fwd", _ptr(x), _ptr(y), n, h * w, c, current_stream_ptr())
    return y


def mean_hw_bwd(dy, shape):
    n, c, h, w = shape
    dx = new_f32(shape)
    _cabi.call("ttb_mean_hw_bwd", _ptr(dy), _ptr(dx), n, h * w, c, current_stream_ptr())
    return dx


def matmul(a, b, bias, m, n, k, sam, sak, sbk, sbn, out=None):
    c = out if out is not None else new_f32((m, n))
    _cabi.call("ttb_matmul", _ptr(a), _ptr(b), _ptr(bias), _ptr(c), m, n, k, sam, sak, sbk, sbn, current_stream_ptr())
    return c


def colsum(g, out=None):
    """(M, N) -> (N,)"""
    m, n = g.shape
    out = out if out is not None else new_f32((n,))
    _cabi.call("ttb_bias_grad", _ptr(g), _ptr(out), m, n, current_stream_ptr())
    return out


def log_softmax(x):
    rows, cols = x.shape
    y = new_f32((rows, cols))
    _cabi.call("ttb_log_softmax_fwd", _ptr(x), _ptr(y), rows, cols, current_stream_ptr())
    return y


def log_softmax_bwd(dy, y):
    rows, cols = y.shape
    dx = new_f32((rows, cols))
    _cabi.call("ttb_log_softmax_bwd", _ptr(dy), _ptr(y), _ptr(dx), rows, cols, current_stream_ptr())
    return dx


_REDUCTIONS = {"none": 0, "mean": 1, "sum": 2}


def nll_loss(logp, target, ignore_index, reduction):
    rows, cols = logp.shape
    out = new_f32((rows,) if reduction == "none" else ())
    count = new_f32(())
    _cabi.call("ttb_nll_loss_fwd", _ptr(logp), target.data_ptr(), rows, cols, int(ignore_index), _REDUCTIONS[reduction],
               _ptr(out), _ptr(count), current_stream_ptr())
    return out, count


def nll_loss_bwd(g, target, shape, ignore_index, reduction, count):
    rows, cols = shape
    dx = new_f32((rows, cols))
    _cabi.call("ttb_nll_loss_bwd", _ptr(g), target.data_ptr(), rows, cols, int(ignore_index), _REDUCTIONS[reduction],
               _ptr(count), _ptr(dx), current_stream_ptr())
    return dx


def bce_logits(x, t, reduction):
    n = x.size
    out = cparray(empty_device(x.shape)) if reduction == "none" else new_f32(())
    ws = torch.empty(int(_cabi.load().ttb_bce_logits_workspace_size()), dtype=torch.uint8, device=x.t.device)
    _cabi.call("ttb_bce_logits_fwd", _ptr(x), _ptr(t), n, _REDUCTIONS[reduction], _ptr(out), ws.data_ptr(), current_stream_ptr())
    return out


def bce_logits_bwd(x, t, g, reduction):
    n = x.size
    dx = cparray(empty_device(x.shape))
    _cabi.call("ttb_bce_logits_bwd", _ptr(x), _ptr(t), _ptr(g), int(reduction == "none"),
               1.0 / n if reduction == "mean" else 1.0, n, _ptr(dx), current_stream_ptr())
    return dx


def cat_channels(arrays):
    """NHWC arrays of equal (N, H, W) -> one array with the channels concatenated (logical dim 1)"""
    n, _, h, w = arrays[0].shape
    ctot = sum(a.shape[1] for a in arrays)
    out = new_f32((n, ctot, h, w))
    off = 0
    st = current_stream_ptr()
    for a in arrays:
        c = a.shape[1]
        _cabi.call("ttb_copy_channels", _ptr(a), _ptr(out), n * h * w, c, ctot, 0, off, c, st)
        off += c
    return out


def split_channels(g, sizes, needed):
    n, ctot, h, w = g.shape
    outs, off = [], 0
    st = current_stream_ptr()
    for c, need in zip(sizes, needed):
        if need:
            o = new_f32((n, c, h, w))
            _cabi.call("ttb_copy_channels", _ptr(g), _ptr(o), n * h * w, ctot, c, off, 0, c, st)
            outs.append(o)
        else:
            outs.append(None)
        off += c
    return outs


# ---------------------------------------------------------------------------------------------------------
# batch norm
# ---------------------------------------------------------------------------------------------------------
def _rows_channels(x):
    shp = x.shape
    if x.ndim == 4:
        return shp[0] * shp[2] * shp[3], shp[1]
    if x.ndim == 2:
        return shp[0], shp[1]
    raise RuntimeError(f"batch_norm on the B200 path supports (N,C,H,W) and (N,C) inputs, got shape {shp}")


def _bn_partials(x, m, c):
    """statistics partials [chunks][2][C] of x: the producer's (conv epilogue / fused add) when still valid, else a
    statistics pass over x"""
    got = bn_stats_of(x)
    if got is not None:
        return got
    chunks = _cabi.load().ttb_bn_num_chunks(m, c)
    partials = torch.empty((chunks, 2, c), dtype=torch.float64, device=x.t.device)
    _cabi.call("ttb_bn_stats", _ptr(x), m, c, partials.data_ptr(), chunks, current_stream_ptr())
    return partials, chunks


def bn_sums(x, reduce_hook=None):
    """per-channel sum(x), sum(x^2) as double partials [chunks][2][C] -> (buffer, chunks, count).  Without a hook
    the per-chunk partials go straight to the finalize kernel (which sums them in fixed order); with a hook
    (SyncBN) `reduce_hook(partials, chunks, 2C, count)` returns the cross-rank [2][C] sums and the global count."""
    m, c = _rows_channels(x)
    partials, chunks = _bn_partials(x, m, c)
    if reduce_hook is None:
        return partials, chunks, m
    sums, m = reduce_hook(partials, chunks, 2 * c, m)
    return sums, 1, m


def bn_forward_train(x, gamma, beta, running_mean, running_var, momentum, eps, relu=False, reduce_hook=None):
    m, c = _rows_channels(x)
    stats = new_f32((5, c))  # rows: mean, var+eps, sd, scale (= gamma/sd), shift (= beta): y = (x - mean)*scale + shift
    base = stats.t.data_ptr()
    row = c * 4
    st = current_stream_ptr()
    fused = getattr(reduce_hook, "fused", None)  # SyncBN over NVLink peer memory: exchange + finalize are one kernel
    if fused is not None and x.t.is_cuda and fused.fits(2 * c, reduce_hook.key):
        partials, chunks = _bn_partials(x, m, c)
        count = m * fused.world  # equal shards by construction (distributed.shard_batch)
        fused.call("ttb_comm_bn_finalize", reduce_hook.key, partials, chunks, count, c, eps,
                   0.0 if momentum is None else momentum, _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
                   base, base + row, base + 2 * row, base + 3 * row, base + 4 * row, st)
    else:
        sums, nchunks, count = bn_sums(x, reduce_hook)
        _cabi.call("ttb_bn_finalize", sums.data_ptr(), nchunks, count, c, eps, 0.0 if momentum is None else momentum,
                   _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), base, base + row, base + 2 * row,
                   base + 3 * row, base + 4 * row, st)
    y = bn_apply_deferred(x, m, c, stats, relu)
    return y, stats, count


def bn_forward_eval(x, gamma, beta, mean, var, eps, relu=False):
    m, c = _rows_channels(x)
    stats = new_f32((5, c))
    base = stats.t.data_ptr()
    row = c * 4
    st = current_stream_ptr()
    _cabi.call("ttb_bn_prepare_eval", _ptr(mean), _ptr(var), c, eps, _ptr(gamma), _ptr(beta), base, base + row,
               base + 2 * row, base + 3 * row, base + 4 * row, st)
    y = bn_apply_deferred(x, m, c, stats, relu)
    return y, stats, m


_RECOMPUTE_RELU_MASK = os.environ.get("TORTTO_B200_RECOMPUTE_RELU_MASK", "1") != "0"


def bn_backward(dy, x, gamma, stats, count, relu_out=None, need_dx=True, need_dgamma=True, need_dbeta=True,
                reduce_hook=None, accum=None, fused_relu=False, out_dgamma=None, out_dbeta=None):
    """-> (dx, dgamma, dbeta).  `count` is the (global) number of elements per channel used in forward.
    `accum`: an array of x's shape that is added to dx inside the apply pass (never written).
    `fused_relu`: the node is BatchNorm+ReLU; the mask (y > 0) is recomputed from x with the scale / shift rows of
    `stats` (what forward normalised with, so bit-identical) instead of reading `relu_out` - one read less per pass."""
    m, c = _rows_channels(x)
    base = stats.t.data_ptr()
    row = c * 4
    st = current_stream_ptr()
    rsc = rsh = None
    if fused_relu and _RECOMPUTE_RELU_MASK:
        relu_out, rsc, rsh = None, base + 3 * row, base + 4 * row
    partials = None
    if relu_out is None and dy.__class__ is cparray and dy._thunk is not None and x.ndim == 4 and \
            getattr(getattr(dy._thunk, "job", None), "kind", None) == "dgrad":
        # the gradient is a dgrad whose launch is still deferred: ONE conv launch writes it AND the two sums over it
        fusedp = dgrad_bn_fused(dy, x, base, rsc, rsh)
        if fusedp is not None:
            partials, chunks = fusedp
    if partials is None:
        chunks = _cabi.load().ttb_bn_num_chunks(m, c)
        partials = torch.empty((chunks, 2, c), dtype=torch.float64, device=x.t.device)
        _cabi.call("ttb_bn_bwd_reduce", _ptr(dy), _ptr(x), base, _ptr(relu_out), rsc, rsh, m, c, partials.data_ptr(), chunks, st)
    dgamma = (out_dgamma if out_dgamma is not None else new_f32((c,))) if need_dgamma else None
    dbeta = (out_dbeta if out_dbeta is not None else new_f32((c,))) if need_dbeta else None
    coef = new_f32((3, c))
    fused = getattr(reduce_hook, "fused", None)
    if fused is not None and x.t.is_cuda and fused.fits(2 * c, reduce_hook.key):
        fused.call("ttb_comm_bn_bwd_finalize", reduce_hook.key, partials, chunks, count, c, _ptr(gamma), base + row,
                   base + 2 * row, _ptr(dgamma), _ptr(dbeta), _ptr(coef), st)
    else:
        if reduce_hook is None:
            sums, nchunks = partials, chunks
        else:
            (sums, _), nchunks = reduce_hook(partials, chunks, 2 * c, m), 1
        _cabi.call("ttb_bn_bwd_finalize", sums.data_ptr(), nchunks, count, c, _ptr(gamma), base + row, base + 2 * row,
                   _ptr(dgamma), _ptr(dbeta), _ptr(coef), st)
    dx = None
    if need_dx:
        dx = cparray(empty_device(x.shape))
        dxh = _new_shadow(dx).data_ptr() if _shadow_wanted(x) else None
        _cabi.call("ttb_bn_bwd_apply", _ptr(dy), _ptr(x), base, _ptr(relu_out), rsc, rsh, _ptr(coef), _ptr(accum), _ptr(dx), m, c,
                   dxh, st)
    return dx, dgamma, dbeta


# ---------------------------------------------------------------------------------------------------------
# relu / elementwise
# ---------------------------------------------------------------------------------------------------------
def relu_fwd(x, inplace=False):
    y = x if inplace else cparray(empty_device(x.shape))
    y._h = None  # (in place: the old shadow is stale)
    yh = _new_shadow(y).data_ptr() if (_shadow_wanted(y) and not inplace) else None
    _cabi.call("ttb_relu_fwd", _ptr(x), _ptr(y), x.size, yh, current_stream_ptr())
    return y


def relu_bwd(dy, y):
    dx = cparray(empty_device(dy.shape))
    _cabi.call("ttb_relu_bwd", _ptr(dy), _ptr(y), _ptr(dx), dy.size, current_stream_ptr())
    return dx


def add_arrays(a, b, stats=False):
    """a + b into a fresh array (same shape, float32, same canonical layout).  `stats`: for 4-D operands the same pass also
    emits the BatchNorm statistics partials of the sum (out._bnstats) - the residual add of a pre-activation block feeds
    the next block's BatchNorm."""
    if a.shape != b.shape or a.t.dtype != torch.float32 or b.t.dtype != torch.float32:
        return cparray(a.t + b.t)
    out = cparray(empty_device(a.shape))
    if stats and _EPILOGUE_STATS and a.ndim == 4 and a.size > 0:
        m, c = _rows_channels(a)
        chunks = _cabi.load().ttb_bn_num_chunks(m, c)
        partials = torch.empty((chunks, 2, c), dtype=torch.float64, device=a.t.device)
        _cabi.call("ttb_add_bn_stats", _ptr(a), _ptr(b), _ptr(out), m, c, partials.data_ptr(), chunks, current_stream_ptr())
        out._bnstats = (partials, chunks, out._version[0])
        return out
    _cabi.call("ttb_add", _ptr(a), _ptr(b), _ptr(out), a.size, current_stream_ptr())
    return out


def axpy_(alpha, x, y):
    _cabi.call("ttb_axpy", float(alpha), _ptr(x), _ptr(y), x.size, current_stream_ptr())
    y._touched()
    return y


def sgd_step_(param, grad, buf, lr, momentum, dampening, weight_decay, nesterov, first_step):
    _cabi.call("ttb_sgd_step", _ptr(param), _ptr(grad), _ptr(buf), param.size, float(lr), float(momentum),
               float(dampening), float(weight_decay), int(bool(nesterov)), int(bool(first_step)), current_stream_ptr())
    weights_changed()


def sgd_step_multi_(params, grads, bufs, first_flags, lr, momentum, dampening, weight_decay, nesterov):
    """Fused update of many parameter tensors (lists of cparrays; bufs entries may be None when momentum == 0)."""
    n = len(params)
    if n == 0:
        return
    P = (ctypes.c_void_p * n)(*[p.t.data_ptr() for p in params])
    G = (ctypes.c_void_p * n)(*[g.t.data_ptr() for g in grads])
    B = (ctypes.c_void_p * n)(*[None if b is None else b.t.data_ptr() for b in bufs])
    S = (ctypes.c_int64 * n)(*[p.size for p in params])
    F = (ctypes.c_ubyte * n)(*[1 if f else 0 for f in first_flags])
    _cabi.call("ttb_sgd_step_multi", n, P, G, B, S, F, float(lr), float(momentum), float(dampening),
               float(weight_decay), int(bool(nesterov)), current_stream_ptr())
    weights_changed()


# ---------------------------------------------------------------------------------------------------------
# max pool
# ---------------------------------------------------------------------------------------------------------
def pool_geometry(h, w, kernel_size, stride, padding, dilation, ceil_mode):
    """Output size under the reference's ceil_mode rules (autograd/grad_nn.py:557-580)."""
    kh, kw = kernel_size
    sh, sw = stride
    ph, pw = padding
    dh, dw = dilation
    rnd = math.ceil if ceil_mode else math.floor
    p = rnd((h + 2 * ph - dh * (kh - 1) - 1) / sh + 1)
    q = rnd((w + 2 * pw - dw * (kw - 1) - 1) / sw + 1)
    if ceil_mode:
        eh = (p - 1) * sh - (h + 2 * ph - dh * (kh - 1) - 1)
        ew = (q - 1) * sw - (w + 2 * pw - dw * (kw - 1) - 1)
        if eh + ph >= (kh - 1) * dh + 1:
            p -= 1
        if ew + pw >= (kw - 1) * dw + 1:
            q -= 1
    return p, q


def pool_desc(x_shape, kernel_size, stride, padding, dilation, ceil_mode):
    n, c, h, w = x_shape
    p, q = pool_geometry(h, w, kernel_size, stride, padding, dilation, ceil_mode)
    return _cabi.PoolDesc(n, c, h, w, kernel_size[0], kernel_size[1], stride[0], stride[1], padding[0], padding[1],
                          dilation[0], dilation[1], p, q)


def maxpool2d_fwd(x, d):
    y = new_f32((d.n, d.c, d.p, d.q))
    idx = cparray(empty_device((d.n, d.c, d.p, d.q), torch.uint8))
    _cabi.call("ttb_maxpool2d_fwd", ctypes.byref(d), _ptr(x), _ptr(y), _ptr(idx), current_stream_ptr())
    return y, idx


def maxpool2d_bwd(dy, idx, d, accumulate=False):
    dx = new_f32((d.n, d.c, d.h, d.w))
    _cabi.call("ttb_maxpool2d_bwd", ctypes.byref(d), _ptr(dy), _ptr(idx), _ptr(dx), int(accumulate),
               current_stream_ptr())
    return dx

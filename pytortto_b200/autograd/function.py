"""The operator protocol (the drop-in boundary): `Function` subclasses with static
`forward(ctx, *inputs, **params)` / `backward(ctx, *grad_outputs)`, invoked through `Cls.apply(...)`.

Same contract as the reference's autograd/function.py:10-179: forward receives Tensors (or None), returns
`build_links(array, grad_fn=ctx)`; backward receives raw arrays and returns one raw array / None per positional
input (arity checked); `ctx.save_for_backward` version-stamps tensors, `ctx.saved_tensors` re-checks them;
`ctx.needs_input_grad[i]`, `ctx.params`, `ctx.next_functions`, `AccumulateGrad` leaves.  One deliberate
difference: the per-op `XBackward` class is created once per Function subclass and cached, not re-created with
`type()` on every call (function.py:116) - SURVEY.md §8(f) rank 2 (Python dispatch overhead).
"""
from .helper import get_data
from .. import ops as _ops


class FunctionBase:
    _forward_cls = None
    __slots__ = ('variable', 'to_save', 'next_functions', 'prev_function_counts', 'needs_input_grad', 'grad', 'params',
                 'requires_grad', 'xp')

    def __init__(self):
        self.variable = None
        self.to_save = None
        self.next_functions = None
        self.prev_function_counts = 0
        self.needs_input_grad = None
        self.requires_grad = False
        self.grad = None
        self.params = None
        self.xp = None

    def save_for_backward(self, *tensors):
        self.to_save = tuple(None if t is None else (t, t._version) for t in tensors)

    @property
    def saved_tensors(self):
        return tuple(get_data(pair) for pair in self.to_save)

    def clear(self):
        self.to_save = None
        if self.__class__ is AccumulateGrad or getattr(self._forward_cls, '_keep_grad_slots', False):
            self.grad = [None]
        else:
            self.grad = None
            self.params = None


class BackwardFunction(FunctionBase):
    __slots__ = ()

    def apply(self, *args):
        out = self._forward_cls.backward(self, *args)
        if out.__class__ is not tuple and out.__class__ is not list:
            out = (out,)
        if len(out) != len(self.needs_input_grad):
            raise RuntimeError(f'function {self.__class__.__name__} returned an incorrect number of gradients'
                               f' (expected {len(self.needs_input_grad)}, got {len(out)})')
        return out


class AccumulateGrad(BackwardFunction):
    """Leaf node: stores / accumulates the gradient on `variable.grad` as a RAW array (function.py:70-93).
    Post-accumulation hooks (used by the data-parallel layer to start all-reduces while backward is still
    running) are called with the variable once its gradient for this backward pass is complete."""
    __slots__ = ()
    post_hooks = []

    def apply(self, *args):
        from .. import ops
        var = self.variable
        g = self.grad[0]
        if var.grad is None:
            var.grad = g
        else:
            ops.join_wgrad()  # `g` may still be in flight on the wgrad stream
            var.grad = ops.add_arrays(var.grad, g)
        for hook in AccumulateGrad.post_hooks:  # a hook that READS var.grad on the device calls ops.join_wgrad() first
            hook(var)


def grad_slot(ctx, i):
    """Where the gradient of input `i` should be written, if anyone cares: the data-parallel layer gives every parameter a
    view into a flat gradient bucket (`Parameter._grad_slot`); a backward that produces the gradient of a LEAF parameter
    which has no gradient yet writes straight into it (AccumulateGrad then just stores that view)."""
    fn = ctx.next_functions[i][0]
    if fn.__class__ is AccumulateGrad:
        var = fn.variable
        if var.grad is None:
            return getattr(var, '_grad_slot', None)
    return None


class Function(FunctionBase):
    __slots__ = ()
    _backward_cls = None

    def __init__(self, *args, **kwargs):
        raise RuntimeError(f"{self.__class__} should not be instantiated. Methods on autograd functions"
                           "are all static, so you should invoke them on the class itself.")

    @staticmethod
    def forward(ctx, *inputs, **params):
        raise NotImplementedError("You must implement the forward function for custom autograd.Function.")

    @staticmethod
    def backward(ctx, *grad_outputs):
        raise NotImplementedError("You must implement the backward method for your custom autograd.Function "
                                  "to use it with backward mode automatic differentiation.")

    @classmethod
    def apply(cls, *inputs, **params):
        bcls = cls.__dict__.get('_backward_cls')
        if bcls is None:
            bcls = type(cls.__name__ + 'Backward', (BackwardFunction,), {'_forward_cls': cls, '__slots__': ()})
            cls._backward_cls = bcls
        grad_fn = bcls()
        grad_fn.params = params

        next_functions = []
        needs_input_grad = []
        requires_grad = False
        for i in inputs:
            i_requires_grad = False if i is None else i.requires_grad
            needs_input_grad.append(i_requires_grad)
            if i_requires_grad:
                requires_grad = True
                fn = i.grad_fn
                if fn is None:
                    fn = AccumulateGrad()
                    fn.variable = i
                    fn.prev_function_counts = 1
                    fn.grad = [None]
                    fn.next_functions = ()
                    next_functions.append((fn, 0))
                else:
                    fn.prev_function_counts += 1
                    next_functions.append((fn, i._output_idx))
            else:
                next_functions.append((None, 0))
        grad_fn.needs_input_grad = tuple(needs_input_grad)
        grad_fn.next_functions = tuple(next_functions)
        grad_fn.requires_grad = requires_grad
        grad_fn.xp = inputs[0].data.__class__ if inputs and inputs[0] is not None else None

        # a deferred producer (ops._defer: a convolution / BatchNorm normalise pass) is launched now, in program order, unless this operator is the
        # one that can absorb it (Add, ReLU: they decide in their own forward)
        if _ops._pending[0] is not None and not cls.__dict__.get('_absorbs'):
            _ops.resolve_pending(None, ())
        results = cls.forward(grad_fn, *inputs, **params)
        n_out = 1 if results.__class__ is not tuple else len(results)
        grad_fn.grad = [None] * n_out
        return results

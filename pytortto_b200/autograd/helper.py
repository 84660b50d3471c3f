"""Graph helpers with the reference's names and semantics (autograd/helper.py:10-82): in-place bookkeeping,
version-checked saved tensors, `build_links`, and the reference-counting topological order used by backward."""
from .grad_mode import is_grad_enabled


def inplace_precheck(xt):
    if is_grad_enabled() and xt.requires_grad and xt.grad_fn is None:
        raise RuntimeError("a leaf Variable that requires grad is being used in an in-place operation.")


def inplace_update(tensor, grad_fn):
    tensor.data._version[0] += 1
    tensor._output_idx = 0
    tensor.requires_grad = grad_fn.requires_grad
    if grad_fn.requires_grad:
        tensor.grad_fn = grad_fn
    return tensor


def get_data(pair):
    if pair is None:
        return None
    tensor, version = pair
    if tensor._version == version:
        return tensor.data
    msg = '' if tensor.grad_fn is None else f', which is the output of {tensor.grad_fn.__class__.__name__},'
    raise RuntimeError(f'one of the variables needed for gradient computation has been modified '
                       f'by an inplace operation: [shape: {tensor.shape}]{msg} is at version '
                       f'{tensor._version}; expected version {version} instead.')


def build_links(data, grad_fn, copy=False, _output_idx=0):
    from ..tensor import Tensor
    if grad_fn.requires_grad and is_grad_enabled():
        return Tensor(data, requires_grad=True, grad_fn=grad_fn, copy=copy, _output_idx=_output_idx, dtype=data.dtype)
    return Tensor(data, copy=copy, dtype=data.dtype)


def toposort(end_node):
    """Yield nodes so that a node is visited only after every consumer of its outputs that is reachable has been
    (reference helper.py:53-82: ready list first, otherwise any pending candidate)."""
    ready = [end_node]
    pending = set()
    while ready or pending:
        node = ready.pop() if ready else pending.pop()
        yield node
        if node.next_functions is None:
            continue
        for fn, _ in node.next_functions:
            if fn is None:
                continue
            fn.prev_function_counts -= 1
            if fn.prev_function_counts == 0:
                ready.append(fn)
                pending.discard(fn)
            else:
                pending.add(fn)

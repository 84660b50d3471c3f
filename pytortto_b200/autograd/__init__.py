from .function import Function, BackwardFunction, AccumulateGrad, FunctionBase
from .grad_mode import no_grad, enable_grad, set_grad_enabled, is_grad_enabled
from .helper import build_links, inplace_precheck, inplace_update, toposort, get_data

__all__ = ['Function', 'BackwardFunction', 'AccumulateGrad', 'FunctionBase', 'no_grad', 'enable_grad',
           'set_grad_enabled', 'is_grad_enabled', 'build_links', 'inplace_precheck', 'inplace_update', 'toposort',
           'get_data']

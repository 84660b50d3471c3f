"""pytortto_b200 - the B200 (sm_100a) implementation of tortto's conv / BatchNorm / ReLU / MaxPool hot path behind
tortto's own API:  `import pytortto_b200 as tt`  then  `tt.nn.Conv2d(...).cuda()`, `loss.backward()`,
`tt.optim.SGD(...)` exactly as with the reference (samrere/pytortto v1.3.4).

Only the CUDA path exists here: host tensors are containers for `.cuda()` / `.cpu()`; the hot-path operators raise
on host arrays and there is no CPU or library fallback (the reference's numpy path is the oracle under oracle/).
"""
__version__ = '0.1.0'

from .xparray import cparray
from .tensor import Tensor, tensor, float16, float32, float64, int16, int32, int64
from .autograd.grad_mode import no_grad, enable_grad, set_grad_enabled, is_grad_enabled
from .VariableFunctions import (manual_seed, add, mul, sum, mean, exp, reshape, flatten, transpose, matmul, cat, zeros,
                                ones, empty, randn)
from .ops import set_math_mode, get_math_mode, set_wgrad_overlap
from .autograd.grad_nn import set_maxpool_backward_accumulate
from . import nn
from . import optim
from . import distributed
from . import cuda_graph
from .serialization import save, load


def cuda_is_available():
    import torch
    return torch.cuda.is_available()
from . import prefetch

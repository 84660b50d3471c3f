"""CUDA-graph capture of a whole training step (Python dispatch overhead).

The reference dispatches ~140 autograd nodes per ResNet-18 step from Python (~50 us each); on a
B200 the device needs only a few milliseconds for the whole step, so the interpreter becomes the bottleneck as soon
as the host has to wait for a result every step (`loss.item()`).  `GraphedStep` keeps the user's step function and
the public API unchanged, runs it a few times eagerly (allocations, momentum buffers, function attributes), then
records ONE more execution into a CUDA graph.  Replays launch the ~270 kernels of the step with a single driver
call; inputs are copied into the captured input buffers first, outputs are read from the captured output tensors.

    step = tt.cuda_graph.GraphedStep(train_step, (x, y), modules=[net])   # x, y: device Tensors of the step's shapes
    loss = step(x_new, y_new)          # -> the captured output Tensor(s), refreshed in place
    loss.item()

Constraints (checked or documented): fixed shapes; no host synchronisation inside the step (`.item()`, `.get()`);
host-side hyper-parameters are baked at capture time (call `recapture()` after changing the learning rate; Adam's step
count and bias corrections live on the device and advance inside the graph; BatchNorm with `momentum=None` - a host-side
1/n averaging factor - refuses to be captured); BatchNorm's host-side `num_batches_tracked` counter is advanced on every
replay for the modules passed in `modules`.  Side effect to know about: the `warmup` eager executions are REAL training
steps on the example inputs (parameters, momentum buffers and running statistics move); the recorded execution itself does
not run.
"""
import torch

from . import ops as _ops
from .tensor import Tensor


def _flatten_outputs(out):
    if isinstance(out, Tensor):
        return [out], False
    return list(out), True


class GraphedStep:
    def __init__(self, fn, example_inputs, modules=(), warmup=3):
        self.fn = fn
        self.modules = list(modules)
        self.warmup = warmup
        self.static_inputs = [Tensor(t.data.copy(), dtype=t.data.dtype, copy=False) for t in example_inputs]
        self.graph = None
        self.outputs = None
        self._multi = False
        self._capture()

    def _bn_modules(self):
        from .nn.modules import _BatchNorm
        for m in self.modules:
            for sub in m.modules():
                if isinstance(sub, _BatchNorm) and sub.training and sub.track_running_stats:
                    yield sub

    def _capture(self):
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # eager warm-up on a side stream, as CUDA-graph capture requires
            for _ in range(self.warmup):
                self.fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = self.fn(*self.static_inputs)
            _ops.resolve_pending(None, ())  # a producer whose launch is still deferred (ops._defer) belongs to the graph
        self.outputs, self._multi = _flatten_outputs(out)
        for bn in self._bn_modules():  # the recorded execution did not run: take back its host-side step count
            if getattr(bn, "_nbt", None) is not None:
                bn._nbt -= 1.0
        torch.cuda.synchronize()

    def recapture(self):
        """Re-record after a hyper-parameter change (e.g. the scheduler changed the learning rate)."""
        self._capture()

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.data.t.copy_(src.data.t, non_blocking=True)
        self.graph.replay()
        for bn in self._bn_modules():  # host-side step counter (nn.modules._BatchNorm.forward) is not in the graph
            if getattr(bn, "_nbt", None) is not None:
                bn._nbt += 1.0
        return tuple(self.outputs) if self._multi else self.outputs[0]

"""Adam / AdamW (reference optim/adam.py, optim/adamw.py + _functional.py:25-115) as ONE fused multi-tensor launch per
48 parameters (ttb_adam_step_multi) instead of ~10 array expressions per parameter.

The step count and the two bias corrections live in a 3-float DEVICE buffer per distinct step value (normally one for
the whole optimizer), advanced by a one-thread kernel per `step()`: a training step captured into a CUDA graph
(cuda_graph.GraphedStep) therefore replays with the right 1 - beta^t factors.  `state[p]['step']` keeps the
reference's host-side count for state_dict().
"""
import ctypes

import torch

from .. import _cabi, ops
from ..xparray import cparray, current_stream_ptr
from .optimizer import Optimizer


class Adam(Optimizer):
    _decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if not 0.0 <= weight_decay:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self._dev_steps = {}  # (group index, host step before this update) -> device [step, 1-b1^t, 1-b2^t]

    def _device_state(self, gi, step_before, betas, device):
        """The device counter that currently stands at `step_before` for group `gi` (created on first use)."""
        key = (gi, step_before)
        buf = self._dev_steps.pop(key, None)
        if buf is None:
            buf = torch.tensor([float(step_before), 1.0 - betas[0] ** step_before, 1.0 - betas[1] ** step_before],
                               dtype=torch.float32, device=device)
        self._dev_steps[(gi, step_before + 1)] = buf
        return buf

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._dev_steps = {}  # device counters are re-created from the loaded host-side step counts

    def step(self):
        st_ptr = current_stream_ptr()
        for gi, group in enumerate(self.param_groups):
            beta1, beta2 = group['betas']
            by_step = {}
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.data.__class__ is not cparray or p.grad.__class__ is not cparray:
                    raise RuntimeError(f"{self.__class__.__name__}.step: parameters and gradients must live on the CUDA device")
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = cparray(torch.zeros_like(p.data.t))
                    state['exp_avg_sq'] = cparray(torch.zeros_like(p.data.t))
                    if group['amsgrad']:
                        state['max_exp_avg_sq'] = cparray(torch.zeros_like(p.data.t))
                by_step.setdefault(state['step'], []).append((p, state))
                state['step'] += 1
            for step_before, items in by_step.items():
                dev = self._device_state(gi, step_before, (beta1, beta2), items[0][0].data.t.device)
                _cabi.call("ttb_adam_advance", dev.data_ptr(), float(beta1), float(beta2), st_ptr)
                n = len(items)
                P = (ctypes.c_void_p * n)(*[p.data.t.data_ptr() for p, _ in items])
                G = (ctypes.c_void_p * n)(*[p.grad.t.data_ptr() for p, _ in items])
                M = (ctypes.c_void_p * n)(*[s['exp_avg'].t.data_ptr() for _, s in items])
                V = (ctypes.c_void_p * n)(*[s['exp_avg_sq'].t.data_ptr() for _, s in items])
                X = (ctypes.c_void_p * n)(*[s['max_exp_avg_sq'].t.data_ptr() for _, s in items]) if group['amsgrad'] else None
                S = (ctypes.c_int64 * n)(*[p.data.size for p, _ in items])
                _cabi.call("ttb_adam_step_multi", n, P, G, M, V, X, S, dev.data_ptr(), float(group['lr']), float(beta1),
                           float(beta2), float(group['eps']), float(group['weight_decay']), int(self._decoupled), st_ptr)
        ops.weights_changed()


class AdamW(Adam):
    """reference optim/adamw.py: decoupled weight decay (p *= 1 - lr*wd before the Adam update), default wd 1e-2"""
    _decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)

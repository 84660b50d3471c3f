"""Adam (reference optim/adam.py + _functional.py:25-68) for the UNet configuration.  Caller-side arithmetic on
device arrays (not part of the accelerated hot path)."""
import math

import torch

from ..xparray import cparray
from .optimizer import Optimizer


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("amsgrad is not supported on the B200 path")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))

    def step(self):
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                g = p.grad.t
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = cparray(torch.zeros_like(p.data.t))
                    st['exp_avg_sq'] = cparray(torch.zeros_like(p.data.t))
                st['step'] += 1
                step = st['step']
                if group['weight_decay'] != 0:
                    g = g + p.data.t * group['weight_decay']
                m, v = st['exp_avg'].t, st['exp_avg_sq'].t
                m.mul_(beta1).add_(g, alpha=1 - beta1)
                v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
                bc1 = 1 - beta1 ** step
                bc2 = 1 - beta2 ** step
                denom = (v.sqrt() / math.sqrt(bc2)).add_(group['eps'])
                p.data.t.addcdiv_(m, denom, value=-group['lr'] / bc1)

"""SGD with momentum / dampening / weight decay / Nesterov (reference optim/sgd.py:5-59 + _functional.py:4-22).
The per-parameter chain of 3-5 array expressions becomes ONE fused multi-tensor launch per 48 parameters
(ttb_sgd_step_multi):
d_p = g + wd*p; buf = d_p (first step) | mom*buf + (1-damp)*d_p; p += -lr * (nesterov ? d_p + mom*buf : buf)."""
from .. import ops
from ..xparray import cparray, new_f32
from .optimizer import Optimizer, required


class SGD(Optimizer):
    def __init__(self, params, lr=required, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        if lr is not required and lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if momentum < 0.0:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))

    def step(self):
        for group in self.param_groups:
            lr, momentum = group['lr'], group['momentum']
            params, grads, bufs, firsts = [], [], [], []
            for p in group['params']:
                g = p.grad
                if g is None:
                    continue
                if p.data.__class__ is not cparray or g.__class__ is not cparray:
                    raise RuntimeError("SGD.step: parameters and gradients must live on the CUDA device")
                buf, first = None, False
                if momentum != 0:
                    st = self.state[p]
                    buf = st.get('momentum_buffer')
                    if buf is None:
                        buf = new_f32(p.data.shape)
                        st['momentum_buffer'] = buf
                        first = True
                params.append(p.data)
                grads.append(g)
                bufs.append(buf)
                firsts.append(first)
            ops.sgd_step_multi_(params, grads, bufs, firsts, lr, momentum, group['dampening'], group['weight_decay'],
                                group['nesterov'])

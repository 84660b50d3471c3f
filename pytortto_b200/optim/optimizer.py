"""Optimizer base (reference optim/optimizer.py:1-188): param_groups, per-parameter state, zero_grad()
(sets p.grad = None, :143-146), state_dict()/load_state_dict() with numpy payloads."""
from collections import defaultdict

import numpy as np

from ..xparray import cparray

required = object()


class Optimizer:
    def __init__(self, params, defaults):
        self.defaults = defaults
        self.state = defaultdict(dict)
        self.param_groups = []
        params = list(params)
        if len(params) == 0:
            raise ValueError("optimizer got an empty parameter list")
        if not isinstance(params[0], dict):
            params = [{'params': params}]
        for g in params:
            self.add_param_group(g)

    def add_param_group(self, group):
        group = dict(group)
        group['params'] = list(group['params'])
        for k, v in self.defaults.items():
            if v is required and k not in group:
                raise ValueError(f"parameter group didn't specify a value of required optimization parameter {k}")
            group.setdefault(k, v)
        self.param_groups.append(group)

    def zero_grad(self):
        for group in self.param_groups:
            for p in group['params']:
                p.grad = None

    def state_dict(self):
        index = {}
        packed_groups = []
        for g in self.param_groups:
            pg = {k: v for k, v in g.items() if k != 'params'}
            ids = []
            for p in g['params']:
                index.setdefault(id(p), len(index))
                ids.append(index[id(p)])
            pg['params'] = ids
            packed_groups.append(pg)
        packed_state = {}
        for p, st in self.state.items():
            packed_state[index[id(p)]] = {k: (v.get() if v.__class__ is cparray else v) for k, v in st.items()}
        return {'state': packed_state, 'param_groups': packed_groups}

    def load_state_dict(self, state_dict):
        groups = state_dict['param_groups']
        if len(groups) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        params = [p for g in self.param_groups for p in g['params']]
        for g, sg in zip(self.param_groups, groups):
            for k, v in sg.items():
                if k != 'params':
                    g[k] = v
        self.state = defaultdict(dict)
        for idx, st in state_dict['state'].items():
            p = params[int(idx)]
            self.state[p] = {k: (cparray.from_numpy(v) if isinstance(v, np.ndarray) and p.is_cuda and v.ndim > 0 else v)
                             for k, v in st.items()}

    def step(self):
        raise NotImplementedError

"""tt.save / tt.load: `np.save` of a pickled {'data': obj} like the reference (serialization.py:4-12), so
checkpoints written by either implementation load in the other."""
import numpy as np


def save(obj, f):
    np.save(f, {'data': obj}, allow_pickle=True)


def load(f):
    return np.load(f, allow_pickle=True).item()['data']

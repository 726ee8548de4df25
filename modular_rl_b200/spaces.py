"""Minimal stand-ins for gym.spaces.Box / Discrete (gym is not installed in this image).
agentzoo.make_mlps only reads `.shape` / `.n` and uses isinstance (agentzoo.py:25-33); if gym is
importable its classes are accepted as well."""
import numpy as np


class Box(object):
    def __init__(self, low, high, shape=None, dtype=np.float64):
        if shape is None:
            low, high = np.asarray(low, dtype), np.asarray(high, dtype)
            shape = low.shape
        else:
            low = np.full(shape, low, dtype)
            high = np.full(shape, high, dtype)
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self):
        return np.random.uniform(np.maximum(self.low, -1e3), np.minimum(self.high, 1e3)).astype(self.dtype)

    def __repr__(self):
        return "Box%s" % (self.shape,)


class Discrete(object):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()

    def sample(self):
        return int(np.random.randint(self.n))

    def __repr__(self):
        return "Discrete(%d)" % self.n


def is_box(space):
    return isinstance(space, Box) or type(space).__name__ == "Box"


def is_discrete(space):
    return isinstance(space, Discrete) or type(space).__name__ == "Discrete"

"""Minimal gymnasium.spaces stand-ins (shape/dtype/bounds containers)."""
import numpy as np


class Space:
    shape = None
    dtype = None


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = start
        self.shape = ()
        self.dtype = np.int64


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)


class Dict(Space):
    """Mapping of sub-spaces; like upstream, plain-dict input is key-sorted."""

    def __init__(self, spaces=None, **kw):
        spaces = dict(spaces or {})
        spaces.update(kw)
        self.spaces = {k: spaces[k] for k in sorted(spaces)}

    def __getitem__(self, key):
        return self.spaces[key]

    def __setitem__(self, key, value):
        self.spaces[key] = value

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

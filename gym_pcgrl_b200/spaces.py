"""Minimal gym.spaces look-alikes (gym is not a dependency): Discrete, MultiDiscrete, Box, Dict.
Only the attributes the reference and its wrappers read are provided."""
import numpy as np


class Space:
    shape = ()
    dtype = None

    def sample(self, rng=None):
        raise NotImplementedError


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self, rng=None):
        return int((rng or np.random).randint(self.n))

    def __repr__(self):
        return "Discrete(%d)" % self.n


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape
        self.dtype = np.dtype(np.int64)

    def sample(self, rng=None):
        rng = rng or np.random
        return np.asarray([rng.randint(n) for n in self.nvec], dtype=np.int64)

    def __repr__(self):
        return "MultiDiscrete(%s)" % self.nvec.tolist()


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low), self.shape).astype(np.float64)
        self.high = np.broadcast_to(np.asarray(high), self.shape).astype(np.float64)

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


class Dict(Space):
    def __init__(self, spaces=None):
        self.spaces = dict(spaces or {})

    def __getitem__(self, key):
        return self.spaces[key]

    def keys(self):
        return self.spaces.keys()

    def __repr__(self):
        return "Dict(%s)" % ", ".join("%s: %r" % kv for kv in self.spaces.items())

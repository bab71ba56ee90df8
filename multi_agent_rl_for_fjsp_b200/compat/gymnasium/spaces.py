"""Shape-only stand-ins for the four `gymnasium.spaces` classes the reference declares
(`/root/reference/utils/ObservationSpaces.py:14-105`, `utils/ActionSpaces.py:10-56`) and the
attributes its callers read: `.spaces`, `.n`, `.shape`, `.sample()` (`a2c.py:118-135,80`,
`train.py:268`)."""
from collections import OrderedDict

import numpy as np


class Space:
    shape = None
    dtype = None

    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = int(start)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.start, self.start + self.n))

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n

    def __repr__(self):
        return "Discrete(%d)" % self.n


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64):
        self.nvec = np.asarray(nvec, dtype=dtype)
        self.shape = self.nvec.shape
        self.dtype = np.dtype(dtype)

    def sample(self):
        return (np.random.random_sample(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

    def __repr__(self):
        return "MultiDiscrete(%s)" % (self.nvec.tolist(),)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def sample(self):
        u = np.random.random_sample(self.shape)
        return (self.low + u * (self.high - self.low)).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s)" % (self.low.min(), self.high.max(), self.shape)


class Dict(Space):
    def __init__(self, spaces=None, **kwargs):
        items = list((spaces or {}).items()) + list(kwargs.items())
        # gymnasium orders plain-dict keys alphabetically; the reference relies on that order when it
        # flattens observations (a2c.py:137-151 sorts keys itself, a2c.py:118-135 only sums sizes).
        if not isinstance(spaces, OrderedDict):
            items = sorted(items, key=lambda kv: kv[0])
        self.spaces = OrderedDict(items)

    def sample(self):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def contains(self, x):
        return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

    def __getitem__(self, key):
        return self.spaces[key]

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def values(self):
        return self.spaces.values()

    def __repr__(self):
        return "Dict(%s)" % ", ".join("%r: %r" % kv for kv in self.spaces.items())

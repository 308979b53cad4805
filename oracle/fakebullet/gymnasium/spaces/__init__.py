"""TEST INFRASTRUCTURE — minimal gymnasium.spaces (Box / Dict / Sequence / Discrete)."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = dtype
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def contains(self, x):
        return True

    def __contains__(self, x):
        return self.contains(x)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        super().__init__(shape, dtype, seed)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        super().__init__((), np.int64, seed)
        self.n = n
        self.start = start

    def sample(self):
        return int(self._rng.integers(self.start, self.start + self.n))


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kw):
        super().__init__(None, None, seed)
        self.spaces = dict(spaces or {}, **kw)

    def __getitem__(self, k):
        return self.spaces[k]

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}


class Sequence(Space):
    def __init__(self, space, seed=None, stack=False):
        super().__init__(None, None, seed)
        self.feature_space = space
        self.stack = stack

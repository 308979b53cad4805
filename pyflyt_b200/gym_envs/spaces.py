"""Observation / action spaces.  With ``gymnasium`` installed these ARE ``gymnasium.spaces.Box`` / ``Dict`` / ``Sequence``;
without it (the build image has no gymnasium wheel) a duck-typed stand-in with the attributes trainers read — ``shape``,
``dtype``, ``low``, ``high``, ``sample()``, ``contains()`` — keeps the env surface the same."""

from __future__ import annotations

import numpy as np

try:  # pragma: no cover - not installable in the build image
    from gymnasium import spaces as _gs

    Box, Dict, Sequence = _gs.Box, _gs.Dict, _gs.Sequence
    HAVE_GYMNASIUM = True
except Exception:
    HAVE_GYMNASIUM = False

    class Box:  # noqa: D101 - gymnasium.spaces.Box stand-in
        def __init__(self, low, high, shape=None, dtype=np.float64, seed=None):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape = tuple(int(s) for s in shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict(dict):  # noqa: D101 - gymnasium.spaces.Dict stand-in
        def __init__(self, spaces=None, **kw):
            super().__init__(spaces or {}, **kw)
            self.spaces = self

        def sample(self):
            return {k: v.sample() for k, v in self.items()}

        def contains(self, x) -> bool:
            return isinstance(x, dict) and set(x) == set(self) and all(self[k].contains(v) for k, v in x.items())

    class Sequence:  # noqa: D101 - gymnasium.spaces.Sequence stand-in
        def __init__(self, space, stack=False):
            self.feature_space, self.stack = space, stack

        def sample(self):
            return np.stack([self.feature_space.sample() for _ in range(2)])

        def contains(self, x) -> bool:
            return all(self.feature_space.contains(r) for r in x)


def batch_box(single: Box, n: int) -> Box:
    """gymnasium.vector.utils.batch_space for a Box"""
    return Box(low=np.broadcast_to(single.low, (n,) + single.shape).copy(), high=np.broadcast_to(single.high, (n,) + single.shape).copy(),
               dtype=single.dtype)

"""TEST INFRASTRUCTURE — the sliver of the gymnasium API the reference envs touch, so that
``/root/reference/PyFlyt/gym_envs`` can be imported for golden-vector generation.
Seeding follows gymnasium.utils.seeding.np_random: Generator(PCG64(SeedSequence(seed)))."""
import numpy as np

from . import spaces  # noqa: F401
from .spaces import Space  # noqa: F401


class Env:
    metadata: dict = {}
    render_mode = None
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env


class ObservationWrapper(Wrapper):
    pass


def make(*a, **k):
    raise NotImplementedError("stub gymnasium has no registry")

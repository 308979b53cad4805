"""pyflyt_b200 — a B200-native batched UAV physics stepper behind the PyFlyt API.

Only the hot path of the reference is rebuilt here: ``Aviary.step()`` and the per-drone
control / physics / state loop, as hand-written sm_100a CUDA kernels behind a C-ABI
(include/pyflyt_b200.h).  ``import pyflyt_b200`` does not need a GPU; constructing a stepper does.
"""

__version__ = "0.1.0"

from . import models  # noqa: F401


def __getattr__(name):
    # torch-dependent pieces load lazily so that model tables stay importable anywhere
    if name in ("BatchedAviary", "AviaryInitException"):
        from .core import aviary

        return getattr(aviary, name)
    if name in ("QuadXHoverVecEnv", "QuadXHoverEnv"):
        from .gym_envs import quadx_hover_env

        return getattr(quadx_hover_env, name)
    raise AttributeError(name)

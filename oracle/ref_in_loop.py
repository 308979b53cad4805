"""TEST INFRASTRUCTURE — run the UNMODIFIED reference (``/root/reference/PyFlyt``) on the fake Bullet.

``install()`` puts ``oracle/fakebullet`` (stand-ins for pybullet, pybullet_data, pybullet_utils,
gymnasium, pettingzoo) and ``/root/reference`` on ``sys.path`` so ``import PyFlyt`` resolves to the
reference's own, untouched source.  Used ONLY by golden-vector generators under ``tools/`` and by
tests that are skipped when ``/root/reference`` is absent (it does not exist on the GPU box).

Parity unpinned: the engine under the reference code is our restatement (fakebullet/pybullet.py),
not PyBullet; see DESIGN.md §oracle.
"""

from __future__ import annotations

import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("PYFLYT_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.realpath(__file__))
FAKE_ROOT = os.path.join(_HERE, "fakebullet")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "PyFlyt"))


def install() -> None:
    """Make ``import PyFlyt`` (reference) and ``import pybullet`` (fake) work in this process."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, FAKE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pybullet  # noqa: F401  (the fake)

    assert os.path.realpath(pybullet.__file__).startswith(os.path.realpath(FAKE_ROOT)), (
        "a real pybullet shadows the fake one; golden vectors must be generated with "
        "tools/dump_pybullet_trajectory.py instead"
    )


class ScriptedNoise:
    """A numpy Generator wrapper that records every ``normal`` draw the reference makes
    (motors.py:134-138 / boosters.py:241-245: ONE scalar per component per physics step)."""

    def __init__(self, seed: int):
        self._rng = np.random.default_rng(seed)
        self.normal_log: list[float] = []

    def normal(self, *args, **kwargs):
        v = self._rng.normal(*args, **kwargs)
        self.normal_log.append(float(v))
        return v

    def __getattr__(self, name):
        return getattr(self._rng, name)

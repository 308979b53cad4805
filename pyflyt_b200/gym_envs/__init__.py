"""Vectorised mirrors of ``PyFlyt.gym_envs`` (hot-path rows of SURVEY.md §8 only).

If ``gymnasium`` is importable the single-env adaptors are registered under the reference's ids
(PyFlyt/gym_envs/__init__.py:8-43) plus the ``-v2`` aliases BASELINE.json uses."""

from .fixedwing_waypoints_env import FixedwingWaypointsVecEnv  # noqa: F401
from .quadx_hover_env import QuadXHoverEnv, QuadXHoverVecEnv  # noqa: F401
from .quadx_waypoints_env import QuadXWaypointsVecEnv  # noqa: F401
from .rocket_landing_env import RocketLandingVecEnv  # noqa: F401
from .single_env import FixedwingWaypointsEnv, QuadXWaypointsEnv, RocketLandingEnv  # noqa: F401

try:  # pragma: no cover - gymnasium is not installed in the build image
    from gymnasium.envs.registration import register

    for _ver in ("v4", "v2"):
        register(id=f"PyFlyt/QuadX-Hover-{_ver}", entry_point="pyflyt_b200.gym_envs.quadx_hover_env:QuadXHoverEnv")
except Exception:
    pass

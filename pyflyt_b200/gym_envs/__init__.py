"""Vectorised mirrors of ``PyFlyt.gym_envs`` (hot-path rows of SURVEY.md §8 only).

If ``gymnasium`` is importable every in-scope env is registered under the reference's id
(PyFlyt/gym_envs/__init__.py:8-43) plus the ``-v2`` alias BASELINE.json uses, with a single-env entry point and a vector
entry point; without gymnasium ``pyflyt_b200.gym_envs.make`` / ``make_vec`` resolve the same ids."""

from .fixedwing_waypoints_env import FixedwingWaypointsVecEnv  # noqa: F401
from .quadx_hover_env import QuadXHoverEnv, QuadXHoverVecEnv  # noqa: F401
from .quadx_waypoints_env import QuadXWaypointsVecEnv  # noqa: F401
from .rocket_landing_env import RocketLandingVecEnv  # noqa: F401
from .single_env import FixedwingWaypointsEnv, QuadXWaypointsEnv, RocketLandingEnv  # noqa: F401

from .vector import PyFlytVectorEnv, env_ids, make, make_vec, register_all  # noqa: F401

REGISTERED = register_all()  # gymnasium ids (reference names + the -v2 aliases) when gymnasium is importable
